#!/bin/bash
# Run on the GPU box (under gpurun): every ncu capture DESIGN.md / bench.py quote, each after its own plain run has exited 0.
# The reports are condensed ON the box (tools/ncu_summary.py, tools/ncu_hot_lines.py) and deleted: only text comes back (gpurun_out <= 64 MiB).
TAG=${1:-r02}
O=gpurun_out
T=/tmp/ncu_$TAG; mkdir -p $T
NCU="ncu --set full --clock-control none"
python tools/prof_decode.py 4096 200 1 > $O/${TAG}_plain_dec.log 2>&1 &&
  $NCU -k 'regex:ob_k_(symbols|bands|synth)' -c 24 -o $T/dec python tools/prof_decode.py 4096 200 1 > $O/${TAG}_ncu_dec.log 2>&1
python tools/ncu_summary.py $T/dec.ncu-rep > $O/${TAG}_ncu_decoder_kernels_all.txt 2>&1
python tools/prof_decode_stereo.py 16384 10 1 2 > $O/${TAG}_plain_dec_stereo.log 2>&1 &&
  $NCU -k 'regex:ob_k_(symbols|bands|synth)' -c 24 -o $T/dec_stereo python tools/prof_decode_stereo.py 16384 10 1 2 > $O/${TAG}_ncu_dec_stereo.log 2>&1
python tools/ncu_summary.py $T/dec_stereo.ncu-rep > $O/${TAG}_ncu_decoder_kernels_stereo_all.txt 2>&1
python tools/prof_encode.py 2368 3 2 1 > $O/${TAG}_plain_enc_warp.log 2>&1 &&
  $NCU --import-source on -k 'regex:ob_k_encodeP' -s 1 -c 1 -o $T/enc_warp python tools/prof_encode.py 2368 3 2 1 > $O/${TAG}_ncu_enc_warp.log 2>&1
python tools/ncu_summary.py $T/enc_warp.ncu-rep > $O/${TAG}_ncu_encoder_warp_kernel.txt 2>&1
python tools/ncu_hot_lines.py $T/enc_warp.ncu-rep ob_k_encodeP opus_codec_b200/libopus_b200.so 60 > $O/${TAG}_hot_lines_encoder_warp.txt 2>&1
python tools/prof_encode.py 16384 1 2 2 > $O/${TAG}_plain_enc_thread.log 2>&1 &&
  $NCU -k 'regex:ob_k_(encode_thread|analysis)' -s 2 -c 2 -o $T/enc_thread python tools/prof_encode.py 16384 1 2 2 > $O/${TAG}_ncu_enc_thread.log 2>&1
python tools/ncu_summary.py $T/enc_thread.ncu-rep > $O/${TAG}_ncu_encoder_thread_kernel.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-live > $O/${TAG}_plain_bench.json 2> $O/${TAG}_plain_bench.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-live > $O/${TAG}_ncu_bench.log 2>&1
for f in dec dec_stereo enc_warp enc_thread; do tail -n 2 $O/${TAG}_plain_$f.log; done
du -sh $O; ls -la $O | grep $TAG
