# odd-depth sweep plan: solver parity tests, then bench A/B (classic plan 0x0800 vs the new default)
D=gpurun_out/plan
mkdir -p $D
( timeout 1200 python -m pytest tests/test_stab_gpu.py tests/test_fuzz_gpu.py tests/test_config0_video.py -m gpu -q -x ) > $D/pytest_stab.log 2>&1; tail -3 $D/pytest_stab.log
for i in 1 2; do
python bench.py --no-cpu-baseline --no-extras --solver-mode 0x0800 > $D/b_classic$i.json 2>> $D/b.err
python bench.py --no-cpu-baseline --no-extras > $D/b_new$i.json 2>> $D/b.err
done
python bench.py --no-cpu-baseline --no-extras --workload 4k-stab --solver-mode 0x0800 > $D/b4k_classic.json 2>> $D/b.err
python bench.py --no-cpu-baseline --no-extras --workload 4k-stab > $D/b4k_new.json 2>> $D/b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/plan/b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]);print(f, round(d['value'],1), round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $D/b.err
