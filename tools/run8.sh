#!/bin/bash
# one 8-GPU session: the 4- and 8-GPU parity tests, then bench.py at N = 8 / 4 (weak and strong)
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "4 or 8" > gpurun_out/multi8_pytest.log 2>&1; tail -3 gpurun_out/multi8_pytest.log
OUT=gpurun_out/bench_n8.jsonl tools/bench_configs.sh 8 cfg3 cfg3@strong cfg4@strong; grep rc= gpurun_out/bench_n8.jsonl.err
OUT=gpurun_out/bench_n4.jsonl tools/bench_configs.sh 4 cfg3 cfg3@strong cfg4@strong; grep rc= gpurun_out/bench_n4.jsonl.err
OUT=gpurun_out/bench_n2s.jsonl tools/bench_configs.sh 2 cfg3@strong cfg4@strong; grep rc= gpurun_out/bench_n2s.jsonl.err
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
