D=gpurun_out/corr_share
mkdir -p $D
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "corr" ) > $D/pytest_corr.log 2>&1; tail -3 $D/pytest_corr.log
: > $D/time_corr.txt
VSC_CORR_MODE=6 timeout 200 python profiles/time_corr_r2.py kc8s3 >> $D/time_corr.txt 2>&1
for v in kc4s6 kc4s4 kc2s12; do f=$PWD/video-stream-consistency_b200/lib/libvsc_b200_$v.so; [ -f $f ] && VSC_CORR_MODE=6 VSC_B200_LIB=$f timeout 200 python profiles/time_corr_r2.py $v >> $D/time_corr.txt 2>&1; done
for v in kc4s6; do f=$PWD/video-stream-consistency_b200/lib/libvsc_b200_$v.so; [ -f $f ] && VSC_CORR_MODE=5 VSC_B200_LIB=$f timeout 200 python profiles/time_corr_r2.py ${v}_w64 >> $D/time_corr.txt 2>&1; done
cat $D/time_corr.txt
