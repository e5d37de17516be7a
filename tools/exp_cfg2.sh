#!/bin/bash
# cfg 2 (D = 256) encode: fp16-screened rescoring vs plain, A/B on one box.
out=gpurun_out/exp_cfg2.txt
: > $out
L=attention-models_b200/lib
for v in "" noscreen "" noscreen; do
  echo "== timeline: rescoring variant '${v:-screened}'" >> $out
  if [ -n "$v" ]; then export VQ_B200_LIB=$L/libvq_b200_$v.so; else unset VQ_B200_LIB; fi
  timeout 90 python tools/encode_timeline.py 2>&1 | grep -v "arn" >> $out
done
for v in "" noscreen; do
  echo "== 8192 tokens '${v:-screened}'" >> $out
  if [ -n "$v" ]; then export VQ_B200_LIB=$L/libvq_b200_$v.so; else unset VQ_B200_LIB; fi
  B=32 timeout 90 python tools/encode_timeline.py 2>&1 | grep "mism\|rescore" >> $out
  echo "== 1M tokens" >> $out
  D=256 T=1048576 timeout 120 python tools/step_timeline.py 3 2>&1 | grep "step\|rescore\|dist_tc\|scan" >> $out
done
unset VQ_B200_LIB
echo "== correctness (all shapes)" >> $out
timeout 300 python tools/tc_check.py >> $out 2>&1
echo "== parity tests" >> $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -3 >> $out
