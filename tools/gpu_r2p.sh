# usage: TAG=r2p bash tools/gpu_r2p.sh -- the OSQP flavour with streamed triangular solves (cp.async.bulk ring): GPU parity tests of the flavour, timing at C4 / C2 shapes
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2p}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 400 python -m pytest tests/test_osqp_flavour.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log 2>&1
cat gpurun_out/${TAG}_pytest.log
LCQP_CUDA_VERBOSE=1 C4_BATCH=${C4_BATCH:-4096} C2_BATCH=16384 timeout 500 python tools/gpu_osqp_check.py c2 c4 2>&1 | tail -8 | tee gpurun_out/${TAG}_osqp.log
