# usage: TAG=r2q bash tools/gpu_r2q.sh -- OSQP flavour after a kernel change: GPU parity tests of the flavour + section profile of the warp kernel on one wave of C4
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2q}
( timeout 300 python -m pytest tests/test_osqp_flavour.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${TAG}_pytest.log 2>&1
cat gpurun_out/${TAG}_pytest.log
LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/liblcqp_cuda_prof.so LCQP_CUDA_VERBOSE=1 C4_BATCH=${C4_BATCH:-1036} timeout 300 python tools/gpu_osqp_check.py c4 2>&1 | tail -14 | tee gpurun_out/${TAG}_prof.log
