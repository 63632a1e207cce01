D=gpurun_out/qg
mkdir -p $D
( timeout 900 python -m pytest tests/test_stab_gpu.py tests/test_fuzz_gpu.py tests/test_config0_video.py tests/test_baseline_sizes_gpu.py tests/test_host_shims.py -m gpu -q -x ) > $D/pytest.log 2>&1; tail -3 $D/pytest.log
python bench.py --no-cpu-baseline > $D/bench_1080p.json 2> $D/b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/qg/bench_1080p.json').read().strip().splitlines()[-1])
print(round(d['value'],1),round(d['e2e']['value'],1),round(d['sustained']['value'],1), d['roofline']['us_per_launch'], [(e['workload'], round(e['value'],1), round(e['e2e']['value'],1)) for e in d.get('extra',[])])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"solver_rolled_kernel<8" -c 4 -o $D/rolled8_qg python profiles/prof_driver.py solver > $D/ncu.log 2>&1; tail -2 $D/ncu.log
