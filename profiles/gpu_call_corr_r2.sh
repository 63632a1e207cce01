D=gpurun_out/corr_r2
mkdir -p $D
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "corr" ) > $D/pytest_corr.log 2>&1; tail -3 $D/pytest_corr.log
: > $D/time_corr.txt
timeout 200 python profiles/time_corr_r2.py unroll8 >> $D/time_corr.txt 2>&1
for v in cu1 cu2 cu4; do f=$PWD/video-stream-consistency_b200/lib/libvsc_b200_$v.so; [ -f $f ] && VSC_B200_LIB=$f timeout 200 python profiles/time_corr_r2.py $v >> $D/time_corr.txt 2>&1; done
cat $D/time_corr.txt
