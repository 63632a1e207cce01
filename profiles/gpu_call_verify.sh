# one gpurun call (~2 min): GPU test suite, smoke, the three bench lines and the reference arm.
# The heavier captures (ncu --set full of the stage-A kernels, launch list of the bench) are in gpu_call_capture.sh.
D=gpurun_out/verify
mkdir -p $D
( time timeout 900 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -4 $D/pytest_gpu.log
python __graft_entry__.py smoke > $D/smoke.log 2>&1; tail -1 $D/smoke.log
python bench.py > $D/bench_1080p.json 2> $D/bench_1080p.err
python bench.py --workload 4k-stab --no-cpu-baseline > $D/bench_4k_stab.json 2> $D/bench_4k.err
python bench.py --workload 4k-dense --no-cpu-baseline > $D/bench_4k_dense.json 2>> $D/bench_4k.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_ref.err
cut -c1-300 $D/bench_1080p.json
