D=gpurun_out/w1
mkdir -p $D
timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q 2>&1 | tail -4
python profiles/time_stage_a.py --warp > $D/time_warp.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"warp_nchw" --csv --log-file $D/warp_ncu.csv python profiles/time_stage_a.py --warp --ncu > $D/ncu.log 2>&1
cat $D/time_warp.txt
