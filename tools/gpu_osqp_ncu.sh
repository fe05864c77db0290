# development aid: ncu capture of the OSQP-flavour kernel on a shortened run (MAXIT outer iterations)
cd $GRAFT_REPO_ROOT
export MAXIT=${MAXIT:-3}
C5_BATCH=${C5_BATCH:-28416} timeout 120 python tools/gpu_osqp_check.py c5 > gpurun_out/${TAG}_short.log 2>&1
cat gpurun_out/${TAG}_short.log
C5_BATCH=${C5_BATCH:-28416} timeout 800 ncu --section SpeedOfLight --section WarpStateStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section InstructionStats --clock-control none -k regex:lcqp_osqp_kernel -c 1 -f -o gpurun_out/${TAG}_osqp python tools/gpu_osqp_check.py c5 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
