D=gpurun_out/qg
mkdir -p $D
( timeout 600 python -m pytest tests/test_stab_gpu.py -m gpu -q -x -k "blocked" ) > $D/pytest.log 2>&1; tail -2 $D/pytest.log
timeout 400 python profiles/sweep_solver_qg.py 28 > $D/sweep_tail.txt 2>&1
cat $D/sweep_tail.txt
