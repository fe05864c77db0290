# usage: TAG=r3d bash tools/gpu_r3d.sh -- bench line of c4 at its full batch with the final OSQP warp kernel, then smoke()
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r3d}
timeout 400 python bench.py --config c4 --steps 1 --warmup 1 --parity 16 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
python __graft_entry__.py --smoke 2>&1 | tail -5
