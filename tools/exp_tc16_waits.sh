#!/bin/bash
# cfg3 filter: how the MMA issuers wait (serial waits / two probes in flight / probes sent behind the previous tile)
out=gpurun_out/exp_tc16_waits.txt
: > $out
L=attention-models_b200/lib
for rep in 1 2; do
for v in "" w1 w2; do
  if [ -n "$v" ]; then export VQ_B200_LIB=$L/libvq_b200_$v.so; else unset VQ_B200_LIB; fi
  echo "== variant '${v:-w0}'" >> $out
  timeout 120 python tools/tc_time.py >> $out 2>&1
done
done
for v in w1 w2; do
  export VQ_B200_LIB=$L/libvq_b200_$v.so
  echo "== correctness $v" >> $out
  timeout 200 python tools/tc_check.py 2 >> $out 2>&1
done
echo "== cfg1 A/B: old library (before today's commits) vs new" >> $out
for rep in 1 2; do
VQ_B200_LIB=$L/libvq_b200_old.so timeout 200 python bench.py --config cfg1 --steps 20 --warmup 5 --skip-sustained --skip-module --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('old', d['ms_per_step'], d['kernel_us_cupti'])" >> $out 2>&1
unset VQ_B200_LIB
timeout 200 python bench.py --config cfg1 --steps 20 --warmup 5 --skip-sustained --skip-module --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new', d['ms_per_step'], d['kernel_us_cupti'])" >> $out 2>&1
done
cat $out
