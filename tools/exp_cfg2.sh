#!/bin/bash
# cfg 2 (D = 256): rescoring with fp16 screening over coalesced fp16 cell copies vs plain, A/B on one box.
out=gpurun_out/exp_cfg2.txt
: > $out
L=attention-models_b200/lib
for v in "" noscreen "" noscreen; do
  echo "== timeline: rescoring variant '${v:-screened}'" >> $out
  if [ -n "$v" ]; then export VQ_B200_LIB=$L/libvq_b200_$v.so; else unset VQ_B200_LIB; fi
  timeout 90 python tools/encode_timeline.py 2>&1 | grep -v "arn" >> $out
done
for v in "" noscreen; do
  if [ -n "$v" ]; then export VQ_B200_LIB=$L/libvq_b200_$v.so; else unset VQ_B200_LIB; fi
  echo "== 1M tokens '${v:-screened}'" >> $out
  D=256 T=1048576 timeout 120 python tools/step_timeline.py 3 2>&1 | grep "step\|rescore\|prep_rows\|scan" >> $out
  echo "== cfg2 fwd+bwd step '${v:-screened}'" >> $out
  FORM=vqgan D=256 T=16384 timeout 120 python tools/step_timeline.py 10 2>&1 | grep "step\|rescore\|prep_rows" >> $out
done
unset VQ_B200_LIB
echo "== correctness (all shapes)" >> $out
timeout 300 python tools/tc_check.py >> $out 2>&1
echo "== parity tests" >> $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -3 >> $out
cat $out
