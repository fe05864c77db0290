# usage: CONFIGS="A=1,B=2 C=3" bash tools/gpu_ab.sh -- bench at a fixed batch under several env configurations (no tests)
cd $GRAFT_REPO_ROOT
for cfg in ${CONFIGS}; do
  echo "== $cfg"
  env $(echo $cfg | tr ',' ' ') LCQP_CUDA_VERBOSE=1 python bench.py --batch ${BATCH:-16384} --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   value %.0f LCQP/s  kernel_ms %.1f  solved %.4f  units/lcqp %.1f' % (d['value'], d['roofline']['kernel_ms'], d['solved_frac'], d['kkt_solves_per_lcqp']))
"
  grep -m1 "lcqp_cuda:" gpurun_out/ab.err; grep -i "error\|Traceback" gpurun_out/ab.err | head -3
done
