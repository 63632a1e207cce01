# one gpurun call (~8 min): what gpu_call_verify.sh does, plus the ncu --set full capture of the stage-A kernels
# (profiles/prof_driver.py) and the ncu launch list of a short default bench run.
D=gpurun_out/v1
mkdir -p $D
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $D/pytest_gpu.log 2>&1
tail -4 $D/pytest_gpu.log
python bench.py > $D/bench_1080p.json 2> $D/bench_1080p.err
python bench.py --workload 4k-stab --no-cpu-baseline > $D/bench_4k_stab.json 2> $D/bench_4k.err
python bench.py --workload 4k-dense --no-cpu-baseline > $D/bench_4k_dense.json 2>> $D/bench_4k.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_ref.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_a -o $D/stage_a python profiles/prof_driver.py stage_a stage_a_prep > $D/ncu_stage_a.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_default.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $D/bench_under_ncu.log 2>&1
python __graft_entry__.py smoke > $D/smoke.log 2>&1
cat $D/smoke.log | tail -2
cut -c1-600 $D/bench_1080p.json
