#!/bin/bash
# Run on the GPU box (under gpurun): every ncu capture DESIGN.md / bench.py quote, each after its own plain run has exited 0.
# The reports are condensed ON the box (tools/ncu_summary.py, tools/ncu_hot_lines.py) and deleted: only text comes back (gpurun_out <= 64 MiB).
# usage: bash tools/capture_profiles.sh [tag]      -> gpurun_out/<tag>_*.txt / .json / .csv  (copy what is to be kept into profiles/)
TAG=${1:-r02}
O=gpurun_out
T=/tmp/ncu_$TAG; mkdir -p $T $O
NCU="ncu --set full --clock-control none --import-source on"
SO=opus_codec_b200/libopus_b200.so
# decoder, bench shape (4096 mono streams x 200 frames): the last of four steps
python tools/prof_decode.py 4096 200 1 > $O/${TAG}_plain_dec.log 2>&1 &&
  $NCU -k 'regex:ob_k_(symbols|bands|synth)' -s 15 -c 5 -o $T/dec python tools/prof_decode.py 4096 200 1 > $O/${TAG}_ncu_dec.log 2>&1
python tools/ncu_summary.py $T/dec.ncu-rep > $O/${TAG}_ncu_decoder_kernels.txt 2>&1
python tools/ncu_hot_lines.py $T/dec.ncu-rep ob_k_symbolsILi8 $SO 60 'ob_k_symbols$' > $O/${TAG}_hot_lines_symbols.txt 2>&1
python tools/ncu_hot_lines.py $T/dec.ncu-rep ob_k_bandsILi1 $SO 60 'ob_k_bands$' > $O/${TAG}_hot_lines_bands.txt 2>&1
python tools/ncu_hot_lines.py $T/dec.ncu-rep ob_k_synthILi1 $SO 60 'ob_k_synth$' > $O/${TAG}_hot_lines_synth.txt 2>&1
# decoder, stereo (16 384 streams x 10 frames)
python tools/prof_decode_stereo.py 16384 10 1 2 > $O/${TAG}_plain_dec_stereo.log 2>&1 &&
  $NCU -k 'regex:ob_k_(symbols|bands|synth)' -s 6 -c 3 -o $T/dec_stereo python tools/prof_decode_stereo.py 16384 10 1 2 > $O/${TAG}_ncu_dec_stereo.log 2>&1
python tools/ncu_summary.py $T/dec_stereo.ncu-rep > $O/${TAG}_ncu_decoder_kernels_stereo.txt 2>&1
# encoder, one warp per stream (one full wave: 2368 streams x 3 frames), and one lane per stream (16 384 streams x 1 frame) with its analysis kernel
python tools/prof_encode.py 2368 3 2 1 > $O/${TAG}_plain_enc_warp.log 2>&1 &&
  $NCU -k 'regex:^ob_k_encode$' -s 1 -c 1 -o $T/enc_warp python tools/prof_encode.py 2368 3 2 1 > $O/${TAG}_ncu_enc_warp.log 2>&1
python tools/ncu_summary.py $T/enc_warp.ncu-rep > $O/${TAG}_ncu_encoder_warp_kernel.txt 2>&1
python tools/ncu_hot_lines.py $T/enc_warp.ncu-rep ob_k_encodeP $SO 60 '^ob_k_encode$' > $O/${TAG}_hot_lines_encoder_warp.txt 2>&1
python tools/prof_encode.py 16384 1 2 2 > $O/${TAG}_plain_enc_thread.log 2>&1 &&
  $NCU -k 'regex:ob_k_(encode_thread|analysis)' -s 2 -c 2 -o $T/enc_thread python tools/prof_encode.py 16384 1 2 2 > $O/${TAG}_ncu_enc_thread.log 2>&1
python tools/ncu_summary.py $T/enc_thread.ncu-rep > $O/${TAG}_ncu_encoder_thread_kernel.txt 2>&1
python tools/make_dram_traffic.py $O/${TAG}_ncu_decoder_kernels.txt $O/${TAG}_ncu_encoder_thread_kernel.txt $O/${TAG}_ncu_encoder_warp_kernel.txt > $O/${TAG}_dram_traffic.json 2> $O/${TAG}_dram_traffic.err
python tools/sass_sizes.py $SO > $O/${TAG}_sass_sizes.txt 2>&1
# launch list of the bench command itself (per-launch times are cold-cache and serialised: the kernels' SHARES are what must agree with the bench line)
python bench.py --steps 2 --warmup 1 --no-live > $O/${TAG}_plain_bench.json 2> $O/${TAG}_plain_bench.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-live > $O/${TAG}_ncu_bench.log 2>&1
for f in dec dec_stereo enc_warp enc_thread; do tail -n 2 $O/${TAG}_plain_$f.log; done
rm -rf $T
du -sh $O; ls -la $O | grep $TAG
