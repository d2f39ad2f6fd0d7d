cd /root/repo
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
tail -c 400 gpurun_out/r02_bench_n1.json
bash tools/capture_bench_profiles.sh r02 2>&1 | tail -4
ncu --set full --import-source on --clock-control none -k regex:mh_sweep --launch-skip 25 --launch-count 1 -o gpurun_out/r02_sweep_final python tools/perf_probe.py 1184 16 20 5000 > gpurun_out/r02_ncu_final.log 2>&1
tail -2 gpurun_out/r02_ncu_final.log
