# usage: CONFIGS="A=1 B=2" bash tools/gpu_stalls.sh -- warp-stall breakdown of lcqp_solve_kernel (ncu, a few metrics) per configuration
cd $GRAFT_REPO_ROOT
M=smsp__issue_active.avg.per_cycle_active,smsp__inst_executed.sum,sm__cycles_elapsed.max,smsp__average_warp_latency_per_inst_issued.ratio
for s in barrier branch_resolving dispatch_stall drain lg_throttle long_scoreboard math_pipe_throttle membar mio_throttle misc no_instruction not_selected selected short_scoreboard sleeping tex_throttle wait; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum
for cfg in ${CONFIGS}; do
  echo "== $cfg"
  env $(echo $cfg | tr ',' ' ') ncu --metrics $M --clock-control none -k regex:lcqp_solve -s 1 -c 1 --csv python bench.py --batch ${BATCH:-4736} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/stalls_$(echo $cfg | tr -c 'A-Za-z0-9' '_').csv 2>&1; tail -40 gpurun_out/stalls_$(echo $cfg | tr -c 'A-Za-z0-9' '_').csv | cut -d, -f 13-15
done
