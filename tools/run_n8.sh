#!/bin/bash
# One 8-GPU session (gpurun --gpus 8): multi-GPU equality check, BASELINE configs[2] through the mcmc entry point,
# the chain-sharded entry point at the reference's length, and the bench at N = 8.  Logs -> gpurun_out/r02_*_n8.*
cd /root/repo
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; grep -E "^world|Error|Traceback" gpurun_out/r02_dist_check_n$N.log | head -20
python tools/run_config2.py prepare /tmp/c2 > gpurun_out/r02_config2_n$N.log 2>&1
$TR tools/run_config2.py run /tmp/c2 --chains 256 --iter 20000 --burn 40000 --thin 100 --samples 0 100 >> gpurun_out/r02_config2_n$N.log 2>&1
$TR tools/run_config2.py run /tmp/c2 --chains 64 --iter 20000 --burn 40000 --thin 10 --samples 0 1 --seed 5 >> gpurun_out/r02_config2_n$N.log 2>&1
grep -E "^configs|^rank|^elapsed|Error|Traceback" gpurun_out/r02_config2_n$N.log | head -40
$TR tools/run_config3.py > gpurun_out/r02_config3_n$N.log 2>&1; grep -E "^configs|Error|Traceback" gpurun_out/r02_config3_n$N.log
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 1200 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
