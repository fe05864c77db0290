# usage: TAG=r2m bash tools/gpu_r2m.sh -- launch-shape A/B of the PAS kernel, bench lines of the OSQP flavour (c4, c2)
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2m}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for cfg in "LIB=liblcqp_cuda.so" "LIB=liblcqp_cuda_t640.so" "LIB=liblcqp_cuda_t576.so,LCQP_CUDA_THREADS=96" "LIB=liblcqp_cuda_tune.so,LCQP_CUDA_THREADS=96"; do
  lib=$(echo $cfg | tr ',' '\n' | grep LIB= | cut -d= -f2)
  envs=$(echo $cfg | tr ',' '\n' | grep -v LIB= | tr '\n' ' ')
  echo "== $cfg"
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib timeout 300 python tools/gpu_pas_prof.py circle 32768 2>&1 | grep -E "^circle|grid" | tail -2
  env $envs LCQP_CUDA_LIB=$GRAFT_REPO_ROOT/lcqpow_b200/lib/$lib timeout 300 python tools/gpu_pas_prof.py dense 16384 2>&1 | grep -E "^dense|grid" | tail -2
done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 900 python bench.py --config c4 --batch ${C4_BATCH:-1628} --steps 1 --warmup 1 --parity 16 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 600 python bench.py --config c2 --flavour osqp --batch 16384 --steps 2 --warmup 3 --parity 64 > gpurun_out/${TAG}_bench_c2_osqp.json 2> gpurun_out/${TAG}_bench_c2_osqp.err
cat gpurun_out/${TAG}_bench_c2_osqp.json; tail -3 gpurun_out/${TAG}_bench_c2_osqp.err
