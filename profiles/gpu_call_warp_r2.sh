D=gpurun_out/warp_r2
mkdir -p $D
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "warp" ) > $D/pytest_warp.log 2>&1; tail -5 $D/pytest_warp.log
timeout 300 python profiles/time_warp_r2.py > $D/time_warp.txt 2>&1; cat $D/time_warp.txt
python bench.py --workload 4k-dense --no-cpu-baseline > $D/bench_4k_dense.json 2> $D/bench_4k_dense.err; python -c "
import json;d=json.loads(open(\"$D/bench_4k_dense.json\").read().strip().splitlines()[-1]);print(d[\"value\"],d[\"e2e\"][\"value\"]);[print(o[\"kernel\"],o[\"shape\"],round(o[\"us_per_launch\"],1),round(o[\"frac\"],3)) for o in d[\"roofline\"][\"custom_ops\"]]"
