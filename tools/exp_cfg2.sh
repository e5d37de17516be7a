#!/bin/bash
# cfg 2 (D = 256) encode: clustered 128-row filter (TMA multicast) vs 256-row CTAs, A/B on one box.
out=gpurun_out/exp_cfg2.txt
: > $out
L=attention-models_b200/lib
echo "== timeline: clusters" >> $out
timeout 90 python tools/encode_timeline.py >> $out 2>&1 || { echo "FAILED/timeout rc=$?" >> $out; }
echo "== timeline: rows256 (4 stages)" >> $out
VQ_TC_CLUSTER=0 timeout 90 python tools/encode_timeline.py >> $out 2>&1
echo "== timeline: clusters, again" >> $out
timeout 90 python tools/encode_timeline.py >> $out 2>&1
echo "== 8192 / 4096 tokens, clusters vs rows256" >> $out
for B in 32 16; do
B=$B timeout 90 python tools/encode_timeline.py 2>&1 | grep "mismatch\|k_dist_tc\|k_rescore" >> $out
B=$B VQ_TC_CLUSTER=0 timeout 90 python tools/encode_timeline.py 2>&1 | grep "mismatch\|k_dist_tc\|k_rescore" >> $out
done
echo "== instrumented: clusters / rows256" >> $out
VQ_B200_LIB=$L/libvq_b200_instr.so D=256 T=16384 timeout 120 python tools/tc_time.py 2>&1 | tail -2 >> $out
VQ_TC_CLUSTER=0 VQ_B200_LIB=$L/libvq_b200_instr.so D=256 T=16384 timeout 120 python tools/tc_time.py 2>&1 | tail -2 >> $out
echo "== 1M tokens: clusters, rows256" >> $out
D=256 T=1048576 timeout 120 python tools/tc_time.py >> $out 2>&1
VQ_TC_CLUSTER=0 D=256 T=1048576 timeout 120 python tools/tc_time.py >> $out 2>&1
echo "== odd sizes" >> $out
for T in 300 1000 18900 33000; do D=256 T=$T timeout 90 python tools/tc_mismatch.py >> $out 2>&1; done
K=768 D=256 T=5000 timeout 90 python tools/tc_mismatch.py >> $out 2>&1
K=1536 D=256 T=40000 timeout 90 python tools/tc_mismatch.py >> $out 2>&1
echo "== correctness (all shapes)" >> $out
timeout 300 python tools/tc_check.py >> $out 2>&1
echo "== parity tests" >> $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -3 >> $out
