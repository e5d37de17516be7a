#!/bin/bash
# One 1-GPU session for the round's profiles: bench over the configs, the reference arm, the ncu launch list and one
# `ncu --set full` capture of the step's kernels (each only after the plain command exited 0), the config sweep.
R=${R:-r02}
OUT=gpurun_out/${R}_bench_configs.jsonl tools/bench_configs.sh 1 cfg3 cfg1 cfg2 cfg2fwd cfg4 sweep:1048576,16384,32 sweep:1048576,8192,256
grep rc= gpurun_out/${R}_bench_configs.jsonl.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_ref.err
timeout 120 python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_prep_rows_fused|k_dist_tc16|k_exact_finish16|k_backward_fused' -s 8 -c 4 -o gpurun_out/${R}_full -f \
    python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_ncufull.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_prep_nchw_fused|k_dist_tc<|k_rescore_g|k_finish' -s 12 -c 4 -o gpurun_out/${R}_full_cfg2 -f \
    python bench.py --config cfg2 --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_ncufull_cfg2.log 2>&1
timeout 300 python tools/sweep.py > gpurun_out/${R}_sweep_configs.txt 2>&1
timeout 60 python tools/exact_modes.py > gpurun_out/${R}_exact_modes.txt 2>&1
ls -la gpurun_out/${R}_*
