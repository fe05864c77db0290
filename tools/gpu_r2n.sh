# usage: TAG=r2n bash tools/gpu_r2n.sh -- GPU parity tests (whole -m gpu suite), bench lines of c3 (bounds family), c4 (one wave) and c2 under the OSQP flavour
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2n}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log 2>&1
cat gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
cat gpurun_out/${TAG}_bench_c3.json; tail -3 gpurun_out/${TAG}_bench_c3.err
timeout 400 python bench.py --config c4 --batch ${C4_BATCH:-1024} --steps 1 --warmup 1 --parity 16 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
timeout 300 python bench.py --config c2 --flavour osqp --batch 16384 --steps 2 --warmup 3 --parity 64 > gpurun_out/${TAG}_bench_c2_osqp.json 2> gpurun_out/${TAG}_bench_c2_osqp.err
cat gpurun_out/${TAG}_bench_c2_osqp.json; tail -3 gpurun_out/${TAG}_bench_c2_osqp.err
