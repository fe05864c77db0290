# development aid: GPU tests of the OSQP flavour + throughput of its kernel at large batches (scratch in shared memory vs global)
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_osqp_flavour.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG:-osqp}_pytest.log; cat gpurun_out/${TAG:-osqp}_pytest.log
export LCQP_CUDA_VERBOSE=1
LOG=gpurun_out/${TAG:-osqp}_osqp.log
: > $LOG
C5_BATCH=28416 timeout ${TMO:-120} python tools/gpu_osqp_check.py c5 >> $LOG 2>&1
C4_BATCH=1024 timeout ${TMO:-150} python tools/gpu_osqp_check.py c4 >> $LOG 2>&1
echo "== scratch in global memory" >> $LOG
export LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/liblcqp_cuda_tune.so LCQP_CUDA_OSQP_NOSMEM=1
C5_BATCH=75776 timeout ${TMO:-120} python tools/gpu_osqp_check.py c5 >> $LOG 2>&1
C2_BATCH=75776 timeout ${TMO:-150} python tools/gpu_osqp_check.py c2 >> $LOG 2>&1
cat $LOG
