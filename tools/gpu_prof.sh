# usage: bash tools/gpu_prof.sh -- section timing with the -DLCQP_PROFILE build (lib/liblcqp_cuda_prof.so, built by hand)
cd $GRAFT_REPO_ROOT
LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/liblcqp_cuda_prof.so LCQP_CUDA_VERBOSE=1 python bench.py --batch ${BATCH:-16384} --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/prof.err | cut -c1-200
grep -A 17 "lcqp_cuda profile" gpurun_out/prof.err | tail -18
