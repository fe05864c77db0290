set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_pytest.log
cat gpurun_out/r1b_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err
cat gpurun_out/r1b_bench.json; tail -3 gpurun_out/r1b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_ref.json 2>> gpurun_out/r1b_bench.err
cat gpurun_out/r1b_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --batch 4096 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_ncu_bench.log 2>&1
tail -5 gpurun_out/r1b_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k lcqp_solve_kernel -s 1 -c 1 -f -o gpurun_out/r1b_solve python bench.py --batch 2048 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
tail -3 gpurun_out/r1b_ncu_full.log
ls -la gpurun_out
