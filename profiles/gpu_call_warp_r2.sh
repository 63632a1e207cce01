D=gpurun_out/warp_r2
mkdir -p $D
( timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x -k "warp" ) > $D/pytest_warp.log 2>&1; tail -3 $D/pytest_warp.log
timeout 300 python profiles/time_warp_r2.py > $D/time_warp.txt 2>&1; head -8 $D/time_warp.txt
for v in p4; do f=$PWD/video-stream-consistency_b200/lib/libvsc_b200_$v.so; [ -f $f ] && { VSC_B200_LIB=$f timeout 300 python profiles/time_warp_r2.py > $D/time_warp_$v.txt 2>&1; echo "== $v"; head -8 $D/time_warp_$v.txt; }; done
