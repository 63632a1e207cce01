# one gpurun call (~6 min): ncu launch list of a short default bench run (kernel shares of the step) and ncu --set full
# captures of the round-2 kernels (staged Warp, 4-step-loop solver pass) for profiles/.
D=gpurun_out/capture
mkdir -p $D
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_default.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --sustained 0 > $D/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"warp_nchw_staged" -c 4 -o $D/warp_staged python profiles/prof_driver.py warp > $D/ncu_warp.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_4k_dense.csv python bench.py --workload 4k-dense --steps 4 --warmup 3 --no-cpu-baseline --no-extras --sustained 0 > $D/bench4k_under_ncu.log 2>&1
ls -la $D
