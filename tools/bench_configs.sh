#!/bin/bash
# bench.py over BASELINE.json's configs on one GPU (or: N=$1 ranks through torchrun): one JSON line per config
N=${1:-1}; shift
OUT=${OUT:-gpurun_out/bench_configs_n$N.jsonl}
: > $OUT
for cfg in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 300 python bench.py --config ${cfg%%@*} --steps 20 --warmup 5 $( [ "${cfg##*@}" != "$cfg" ] && echo --scaling ${cfg##*@} ) >> $OUT 2>> $OUT.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --config ${cfg%%@*} --steps 20 --warmup 5 $( [ "${cfg##*@}" != "$cfg" ] && echo --scaling ${cfg##*@} ) >> $OUT 2>> $OUT.err
  fi
  echo "rc=$? $cfg" >> $OUT.err
done
