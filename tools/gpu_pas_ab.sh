# usage: bash tools/gpu_pas_ab.sh -- throughput of the PAS kernel for several builds / launch shapes (development aid)
cd $GRAFT_REPO_ROOT
for cfg in "LIB=liblcqp_cuda.so" "LIB=liblcqp_cuda_t768.so" "LIB=liblcqp_cuda_t1024.so" "LIB=liblcqp_cuda_t1024.so,LCQP_CUDA_THREADS=64" "LIB=liblcqp_cuda.so,LCQP_CUDA_THREADS=64" "LIB=liblcqp_cuda.so,LCQP_CUDA_THREADS=256"; do
  lib=$(echo $cfg | tr ',' '\n' | grep LIB= | cut -d= -f2)
  envs=$(echo $cfg | tr ',' '\n' | grep -v LIB= | tr '\n' ' ')
  echo "== $cfg"
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib python tools/gpu_pas_prof.py circle ${BATCH:-16384} 2>&1 | grep -E "^circle|grid" | tail -2
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib python tools/gpu_pas_prof.py dense 8192 2>&1 | grep -E "^dense|grid" | tail -2
done
