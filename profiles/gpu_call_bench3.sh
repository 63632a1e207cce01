D=gpurun_out/verify
mkdir -p $D
( timeout 900 python -m pytest tests/test_stab_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x ) > $D/pytest_stab.log 2>&1; tail -2 $D/pytest_stab.log
python bench.py --workload 4k-stab --no-cpu-baseline --no-extras > $D/bench_4k_stab.json 2> $D/bench_4k.err
python bench.py --workload 4k-dense --no-cpu-baseline --no-extras > $D/bench_4k_dense.json 2>> $D/bench_4k.err
python bench.py > $D/bench_1080p.json 2> $D/bench_1080p.err
for f in bench_1080p bench_4k_stab bench_4k_dense; do python -c "
import json;d=json.loads(open('$D/$f.json').read().strip().splitlines()[-1]);r=d['roofline'];print('$f',round(d['value'],1),round(d['e2e']['value'],1),r['kernel'][:28],round(r['us_per_launch'],1),round(r['frac'],2),round(r['dram_frac'],3),round(r['fused_stage_a']['frac'],3), (d.get('sustained') or {}).get('value'))"; done
