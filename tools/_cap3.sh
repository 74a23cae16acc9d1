O=gpurun_out; T=/tmp/ncu3; mkdir -p $T
python tools/prof_encode.py 16384 4 2 1 2>&1 | tail -1
python tools/prof_encode.py 16384 4 2 2 2>&1 | tail -1
python tools/prof_encode.py 4736 4 2 1 2>&1 | tail -1
python tools/prof_encode.py 2368 3 2 1 > $O/r02i_plain_enc_warp.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^ob_k_encode$' -s 1 -c 1 -o $T/enc_warp python tools/prof_encode.py 2368 3 2 1 > $O/r02i_ncu_enc_warp.log 2>&1
python tools/ncu_summary.py $T/enc_warp.ncu-rep > $O/r02_ncu_encoder_warp_kernel.txt 2>&1
python tools/ncu_hot_lines.py $T/enc_warp.ncu-rep 'ob_k_encodeP' opus_codec_b200/libopus_b200.so 60 '^ob_k_encode$' > $O/r02_hot_lines_encoder_warp.txt 2>&1
python tools/prof_encode.py 16384 1 2 2 > $O/r02i_plain_enc_thread.log 2>&1 && ncu --set full --clock-control none -k 'regex:ob_k_(encode_thread|analysis)' -s 2 -c 2 -o $T/enc_thread python tools/prof_encode.py 16384 1 2 2 > $O/r02i_ncu_enc_thread.log 2>&1
python tools/ncu_summary.py $T/enc_thread.ncu-rep > $O/r02_ncu_encoder_thread_kernel.txt 2>&1
grep -E "duration|DRAM read|DRAM write|stall" $O/r02_ncu_encoder_warp_kernel.txt $O/r02_ncu_encoder_thread_kernel.txt
