D=gpurun_out/verify
mkdir -p $D
( timeout 900 python -m pytest tests/test_stab_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x ) > $D/pytest_stab.log 2>&1; tail -2 $D/pytest_stab.log
python bench.py --no-cpu-baseline --no-extras > $D/b_merge.json 2> $D/b.err
python bench.py --no-cpu-baseline --no-extras --solver-mode 0x800 > $D/b_nomerge.json 2>> $D/b.err
for f in b_merge b_nomerge; do python -c "
import json;d=json.loads(open('$D/$f.json').read().strip().splitlines()[-1]);print('$f',round(d['value'],1),round(d['e2e']['value'],1),round(d['sustained']['value'],1))"; done
