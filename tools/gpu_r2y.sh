# usage: TAG=r2y bash tools/gpu_r2y.sh -- the driver's own command lines at N = 1 (wall time of each), after the whole -m gpu suite
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2y}
( time python __graft_entry__.py --smoke 2>&1 | tail -6 ) > gpurun_out/${TAG}_smoke.log 2>&1
cat gpurun_out/${TAG}_smoke.log
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_pytest.log 2>&1
cat gpurun_out/${TAG}_pytest.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err ) 2>&1 | tail -4
cut -c1-300 gpurun_out/${TAG}_bench_ref.json
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | tail -4
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
