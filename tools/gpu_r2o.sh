# usage: TAG=r2o bash tools/gpu_r2o.sh -- launch-shape A/B of the PAS kernel (variant builds) + a short c2 bench line with the work counters
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2o}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for cfg in "LIB=liblcqp_cuda.so" "LIB=liblcqp_cuda_t640.so" "LIB=liblcqp_cuda_t576.so,LCQP_CUDA_THREADS=96" "LIB=liblcqp_cuda_tune.so,LCQP_CUDA_THREADS=96" "LIB=liblcqp_cuda_tune.so,LCQP_CUDA_GROUPS=3"; do
  lib=$(echo $cfg | tr ',' '\n' | grep LIB= | cut -d= -f2)
  envs=$(echo $cfg | tr ',' '\n' | grep -v LIB= | tr '\n' ' ')
  echo "== $cfg"
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib timeout 300 python tools/gpu_pas_prof.py circle 32768 2>&1 | grep -E "^circle|grid|Error|error" | tail -2
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib timeout 300 python tools/gpu_pas_prof.py dense 16384 2>&1 | grep -E "^dense|grid|Error|error" | tail -2
done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 300 python bench.py --config c2 --batch 65536 --steps 2 --warmup 3 --parity 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
cat gpurun_out/${TAG}_bench_c2.json; tail -3 gpurun_out/${TAG}_bench_c2.err
