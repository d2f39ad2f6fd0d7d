#!/bin/bash
# Captures, for the CURRENT kernel, the ncu evidence bench.py and profiles/ quote (run on the GPU box through gpurun, after
# `python bench.py` has exited 0 without ncu):
#   1. launch list of the bench command (kernel shares of the step)          -> gpurun_out/${R}_ncu_launches_bench.csv
#   2. DRAM bytes + duration of one timed mh_sweep_kernel launch of the bench -> gpurun_out/${R}_ncu_sweep_dram_bench.csv
# Copy both into profiles/ afterwards (bench.py reads the second for roofline.traffic).
# Launch numbering: 5000 tuning sweeps = 25 launches of 200, then 3 warm-up steps; launch 29 is the first timed step.
R=${1:-r02}
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/${R}_ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mh_sweep \
    --launch-skip 28 --launch-count 1 --csv --log-file gpurun_out/${R}_ncu_sweep_dram_bench.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/${R}_ncu_dram.log 2>&1
tail -3 gpurun_out/${R}_ncu_sweep_dram_bench.csv
