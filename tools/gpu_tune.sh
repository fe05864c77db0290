# usage: bash tools/gpu_tune.sh  -- parity tests, then the bench at a fixed batch under several plan knobs
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B=${BATCH:-16384}
run() { echo "== $*"; env "$@" LCQP_CUDA_VERBOSE=1 python bench.py --batch $B --steps 1 --warmup 1 --no-cpu-baseline 2> gpurun_out/tune.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   value %.0f LCQP/s  kernel_ms %.1f  solved %.4f  units/lcqp %.1f' % (d['value'], d['roofline']['kernel_ms'], d['solved_frac'], d['kkt_solves_per_lcqp']))
"; grep -m1 "lcqp_cuda:" gpurun_out/tune.err; }
for cfg in ${CONFIGS:-"LCQP_CUDA_THREADS=128"}; do run $(echo $cfg | tr ',' ' '); done
