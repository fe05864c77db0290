# usage: LIBV=/root/repo/lcqpow_b200/lib/v_x.so bash tools/gpu_variant.sh -- GPU parity tests + short bench with an alternative build of the library
cd $GRAFT_REPO_ROOT
export LCQP_CUDA_LIB=$LIBV
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --batch ${BATCH:-32768} --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   value %.0f LCQP/s  e2e %.0f  kernel_ms %.1f  solved %.4f  units/lcqp %.1f  inst0 %s' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['solved_frac'], d['kkt_solves_per_lcqp'], d['instance0_matches_shipped_solution']))
"
