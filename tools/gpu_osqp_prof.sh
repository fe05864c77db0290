# development aid: GPU tests of the OSQP flavour + throughput of its kernels
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_osqp_flavour.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG:-osqp}_pytest.log; cat gpurun_out/${TAG:-osqp}_pytest.log
export LCQP_CUDA_VERBOSE=1
LOG=gpurun_out/${TAG:-osqp}_osqp.log
: > $LOG
C4_BATCH=1024 timeout ${TMO:-120} python tools/gpu_osqp_check.py c4 >> $LOG 2>&1
C4_BATCH=4096 timeout ${TMO:-200} python tools/gpu_osqp_check.py c4 >> $LOG 2>&1
C2_BATCH=16384 timeout ${TMO:-120} python tools/gpu_osqp_check.py c2 >> $LOG 2>&1
C5_BATCH=16384 timeout ${TMO:-150} python tools/gpu_osqp_check.py c5 >> $LOG 2>&1
cat $LOG
