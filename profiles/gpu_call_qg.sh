D=gpurun_out/qg
mkdir -p $D
VSC_B200_LIB=$PWD/video-stream-consistency_b200/lib/libvsc_b200_pfnb.so timeout 400 python profiles/sweep_solver_qg.py 28 > $D/sweep_pfnb.txt 2>&1
cat $D/sweep_pfnb.txt
