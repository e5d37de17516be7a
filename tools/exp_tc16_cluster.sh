#!/bin/bash
# cfg3 filter: clusters of two 128-row CTAs (four accumulator stages per row half, codebook tiles multicast) vs 256-row CTAs
out=gpurun_out/exp_tc16_cluster.txt
: > $out
L=attention-models_b200/lib
echo "== correctness (clusters)" >> $out
timeout 120 python tools/tc_check.py 2 >> $out 2>&1 || echo "FAILED rc=$?" >> $out
for rep in 1 2; do
  echo "== clusters" >> $out
  timeout 120 python tools/tc_time.py >> $out 2>&1
  echo "== 256-row CTAs" >> $out
  VQ_TC16_CLUSTER=0 timeout 120 python tools/tc_time.py >> $out 2>&1
done
echo "== instrumented: clusters / 256-row" >> $out
VQ_B200_LIB=$L/libvq_b200_instr.so timeout 120 python tools/tc_time.py 2>&1 | grep "instrument" | tail -2 >> $out
VQ_TC16_CLUSTER=0 VQ_B200_LIB=$L/libvq_b200_instr.so timeout 120 python tools/tc_time.py 2>&1 | grep "instrument" | tail -2 >> $out
echo "== other sizes" >> $out
for T in 300 1000 4096 33000 100000; do D=32 T=$T timeout 90 python tools/tc_mismatch.py >> $out 2>&1; done
K=16384 D=32 T=65536 timeout 90 python tools/tc_mismatch.py >> $out 2>&1
K=512 D=32 T=5000 timeout 90 python tools/tc_mismatch.py >> $out 2>&1
echo "== parity tests" >> $out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 >> $out
cat $out
