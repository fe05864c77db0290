# usage (gpurun --gpus 2): TAG=r2x bash tools/gpu_r2x.sh -- both bench arms under torchrun at N = 2, as the driver launches them
cd $GRAFT_REPO_ROOT
TAG=${TAG:-r2x}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_2gpu.json 2> gpurun_out/${TAG}_ref_2gpu.err
cat gpurun_out/${TAG}_ref_2gpu.json | cut -c1-400; tail -2 gpurun_out/${TAG}_ref_2gpu.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 3 --batch ${BATCH:-262144} > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
cat gpurun_out/${TAG}_bench_2gpu.json; tail -3 gpurun_out/${TAG}_bench_2gpu.err
