D=gpurun_out/corr_share
mkdir -p $D
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "corr" ) > $D/pytest_corr.log 2>&1; tail -3 $D/pytest_corr.log
: > $D/time_corr.txt
for m in 3 5 6 5 6; do VSC_CORR_MODE=$m timeout 200 python profiles/time_corr_r2.py mode$m >> $D/time_corr.txt 2>&1; done
cat $D/time_corr.txt
