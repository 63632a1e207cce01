# one gpurun call (~4 min): ncu launch lists of short bench runs (kernel shares of the step) for profiles/.
D=gpurun_out/capture
mkdir -p $D
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_default.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --sustained 0 > $D/bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_4k_dense.csv python bench.py --workload 4k-dense --steps 4 --warmup 3 --no-cpu-baseline --no-extras --sustained 0 > $D/bench4k_under_ncu.log 2>&1
ls -la $D
