# one gpurun call: blocked-solver bit-identity, A/B timing
D=gpurun_out/solver_r2
mkdir -p $D
( timeout 600 python -m pytest tests/test_stab_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x -k "blocked or frame_stabilize or sequence or solver" ) > $D/pytest_default.log 2>&1; tail -3 $D/pytest_default.log
timeout 400 python profiles/sweep_solver_r2.py rolled 0,0:8,4:16,8 > $D/sweep5.txt 2>&1
cat $D/sweep5.txt
python bench.py --no-cpu-baseline > $D/bench_1080p.json 2> $D/bench_1080p.err; cut -c1-200 $D/bench_1080p.json
