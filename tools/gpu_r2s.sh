# usage: TAG=r2s bash tools/gpu_r2s.sh -- the round's final GPU pass: whole -m gpu suite, bench lines (c2 at 2^20, c5, reference arm, c4 at 4096),
# ncu launch list of the c2 bench at its full batch, ncu --set full of the active-set kernel and of the streamed OSQP warp kernel
set -x
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2s}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${TAG}_pytest.log 2>&1
cat gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
cat gpurun_out/${TAG}_bench_c5.json; tail -3 gpurun_out/${TAG}_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -12 gpurun_out/${TAG}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lcqp_pas_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_pas python bench.py --batch ${NCU_BATCH:-4736} --steps 1 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
MAXIT=2 C4_BATCH=1036 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lcqp_osqpw_kernel -c 1 -f -o gpurun_out/${TAG}_osqpw python tools/gpu_osqp_check.py c4 > gpurun_out/${TAG}_ncu_osqpw.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_osqpw.log
timeout 700 python bench.py --config c4 --steps 1 --warmup 1 --parity 16 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
cat gpurun_out/${TAG}_bench_c4.json; tail -3 gpurun_out/${TAG}_bench_c4.err
ls -la gpurun_out | tail -20
