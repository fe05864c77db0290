# usage: TAG=r2a bash tools/gpu_round2.sh -- GPU parity tests, bench (c2 at the full 2^20 batch, c5, c3, reference arm),
# ncu launch list at the bench batch, one ncu --set full capture of the solve kernel
set -x
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2x}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
cat gpurun_out/${TAG}_bench_c5.json; tail -3 gpurun_out/${TAG}_bench_c5.err
python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
cat gpurun_out/${TAG}_bench_c3.json; tail -3 gpurun_out/${TAG}_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -8 gpurun_out/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lcqp_pas_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_solve python bench.py --batch ${NCU_BATCH:-4736} --steps 1 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -20
