# one gpurun call: stage-A / Warp kernel-selection timings (events + ncu launch list), then the GPU test suite
mkdir -p gpurun_out/s2b
python profiles/time_stage_a.py > gpurun_out/s2b/time_stage_a.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:"stage_a|warp_nchw" --csv --log-file gpurun_out/s2b/modes_ncu.csv python profiles/time_stage_a.py --ncu > gpurun_out/s2b/ncu.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2b/pytest_gpu.log 2>&1
tail -4 gpurun_out/s2b/pytest_gpu.log
cat gpurun_out/s2b/time_stage_a.txt
