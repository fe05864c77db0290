# usage: TAG=r1c bash tools/gpu_ncu.sh -- one ncu --set full capture of lcqp_solve_kernel (second launch) at a small batch
cd $GRAFT_REPO_ROOT
TAG=${TAG:-ncu}
timeout 900 ncu --set full --clock-control none --import-source on -k lcqp_solve_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_solve python bench.py --batch ${BATCH:-2368} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
