#!/bin/bash
# refresh the cfg3 artefacts after a kernel change: default bench line, reference arm, launch list, ncu --set full of the step
R=${R:-r02}
timeout 300 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --config cfg1 --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg1.json 2>/dev/null
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg4.json 2>/dev/null
timeout 120 python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_prep_rows_fused|k_dist_tc16|k_exact_finish16|k_backward_fused' -s 8 -c 4 -o gpurun_out/${R}_full -f \
    python bench.py --steps 6 --warmup 3 --no-graphs --skip-cpu --skip-e2e --skip-sustained --skip-module > gpurun_out/${R}_ncufull.log 2>&1
timeout 60 python tools/exact_modes.py > gpurun_out/${R}_exact_modes.txt 2>&1
python tools/bench_table.py gpurun_out/${R}_bench_default.json gpurun_out/${R}_bench_cfg1.json gpurun_out/${R}_bench_cfg4.json | cut -c1-220
