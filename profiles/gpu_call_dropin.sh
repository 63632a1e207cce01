D=gpurun_out/verify
mkdir -p $D
( timeout 600 python -m pytest tests/test_host_shims.py tests/test_config0_video.py -m gpu -q -x ) > $D/pytest_shims.log 2>&1; tail -3 $D/pytest_shims.log
timeout 300 python profiles/time_dropin_paths.py > $D/time_dropin_paths.txt 2>&1; cat $D/time_dropin_paths.txt
