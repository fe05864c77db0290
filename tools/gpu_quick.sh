# usage: TAG=x bash tools/gpu_quick.sh -- GPU parity tests + a short bench at a fixed batch (no CPU baseline, no ncu)
cd $GRAFT_REPO_ROOT
TAG=${TAG:-quick}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
for cfg in ${CONFIGS:-"LCQP_CUDA_THREADS=128"}; do
  echo "== $cfg"
  env $(echo $cfg | tr ',' ' ') LCQP_CUDA_VERBOSE=1 python bench.py --batch ${BATCH:-32768} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_bench.err | tee -a gpurun_out/${TAG}_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   value %.0f LCQP/s  e2e %.0f  kernel_ms %.1f  solved %.4f  units/lcqp %.1f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['solved_frac'], d['kkt_solves_per_lcqp']))
"
  grep -m1 "lcqp_cuda:" gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
done
