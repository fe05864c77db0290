# development aid: how the OSQP-flavour kernel scales with the batch, and one ncu capture of it
cd $GRAFT_REPO_ROOT
for b in 32 1024 8192; do C5_BATCH=$b timeout 300 python tools/gpu_osqp_check.py c5 2>&1 | grep "C5 dense"; done
C2_BATCH=4096 timeout 300 python tools/gpu_osqp_check.py c2 2>&1 | grep "C2 circle"
C4_BATCH=256 timeout 300 python tools/gpu_osqp_check.py c4 2>&1 | grep "C4 sparse"
C5_BATCH=2048 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lcqp_osqp_kernel -s 1 -c 1 -f -o gpurun_out/${TAG:-osqp}_osqp python tools/gpu_osqp_check.py c5 > gpurun_out/${TAG:-osqp}_ncu.log 2>&1
tail -3 gpurun_out/${TAG:-osqp}_ncu.log
