# usage: TAG=r2w bash tools/gpu_r2w.sh -- after the OSQP-kernel work: bench line of c4 at its full batch, a short c2 line (the active-set kernel is untouched), c2 under the OSQP flavour
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2w}
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --parity 16 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 300 python bench.py --config c2 --batch 131072 --steps 2 --warmup 3 --parity 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_short.json 2> gpurun_out/${TAG}_bench_c2_short.err
cat gpurun_out/${TAG}_bench_c2_short.json; tail -3 gpurun_out/${TAG}_bench_c2_short.err
timeout 300 python bench.py --config c2 --flavour osqp --batch 16384 --steps 2 --warmup 3 --parity 64 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_osqp.json 2> gpurun_out/${TAG}_bench_c2_osqp.err
cat gpurun_out/${TAG}_bench_c2_osqp.json; tail -3 gpurun_out/${TAG}_bench_c2_osqp.err
