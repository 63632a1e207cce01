D=gpurun_out/rows
mkdir -p $D
timeout 200 python -m pytest tests/test_stab_gpu.py -m gpu -x -q -k "stage_a" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"stage_a" --csv --log-file $D/rows_ncu.csv python profiles/time_stage_a.py --rows --ncu > $D/ncu.log 2>&1
python profiles/time_stage_a.py --rows > $D/rows_events.txt 2>&1
tail -45 $D/rows_events.txt
