# one gpurun call (~4 min): GPU test suite, smoke, the three bench lines, the reference arm, drop-in path timing.
D=gpurun_out/verify
mkdir -p $D
( time timeout 1200 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -4 $D/pytest_gpu.log
python __graft_entry__.py smoke > $D/smoke.log 2>&1; tail -1 $D/smoke.log
python bench.py > $D/bench_1080p.json 2> $D/bench_1080p.err
python bench.py --workload 4k-stab --no-cpu-baseline > $D/bench_4k_stab.json 2> $D/bench_4k.err
python bench.py --workload 4k-dense --no-cpu-baseline > $D/bench_4k_dense.json 2>> $D/bench_4k.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_ref.err
timeout 300 python profiles/time_dropin_paths.py > $D/time_dropin_paths.txt 2>&1; cat $D/time_dropin_paths.txt
cut -c1-300 $D/bench_1080p.json
