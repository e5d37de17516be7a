#!/bin/bash
# D = 256 generic filter: 128-row CTAs (8-stage codebook ring) vs 256-row CTAs (2 or 4 stages), same box.
out=gpurun_out/exp_rows128.txt
: > $out
L=attention-models_b200/lib
timeout 300 python tools/tc_check.py >> $out 2>&1
for rep in 1 2; do
  for T in 16384 8192; do
    D=256 T=$T timeout 120 python tools/tc_time.py >> $out 2>&1
    VQ_TC_ROWS128=0 D=256 T=$T timeout 120 python tools/tc_time.py >> $out 2>&1
    echo "bs4:" >> $out; VQ_B200_LIB=$L/libvq_b200_bs4.so VQ_TC_ROWS128=0 D=256 T=$T timeout 120 python tools/tc_time.py >> $out 2>&1
  done
done
echo "== instrumented rows128" >> $out
VQ_B200_LIB=$L/libvq_b200_instr.so D=256 T=16384 timeout 120 python tools/tc_time.py 2>&1 | tail -3 >> $out
echo "== instrumented rows256" >> $out
VQ_B200_LIB=$L/libvq_b200_instr.so VQ_TC_ROWS128=0 D=256 T=16384 timeout 120 python tools/tc_time.py 2>&1 | tail -3 >> $out
echo "== 1M tokens: bs2 vs bs4" >> $out
D=256 T=1048576 timeout 120 python tools/tc_time.py >> $out 2>&1
VQ_B200_LIB=$L/libvq_b200_bs4.so D=256 T=1048576 timeout 120 python tools/tc_time.py >> $out 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 >> $out
cat $out
