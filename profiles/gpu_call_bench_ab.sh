D=gpurun_out/verify
mkdir -p $D
( timeout 900 python -m pytest tests/test_stab_gpu.py tests/test_fuzz_gpu.py tests/test_host_shims.py tests/test_config0_video.py -m gpu -q -x ) > $D/pytest_stab.log 2>&1; tail -2 $D/pytest_stab.log
python bench.py --no-cpu-baseline > $D/b_new.json 2> $D/b.err
python -c "
import json;d=json.loads(open('$D/b_new.json').read().strip().splitlines()[-1]);print(round(d['value'],1),round(d['e2e']['value'],1),round(d['sustained']['value'],1), [(e['workload'], round(e['value'],1), round(e['e2e']['value'],1)) for e in d.get('extra',[])])"
