cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg_gpus.csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py > gpurun_out/mg_check.log 2>&1; tail -5 gpurun_out/mg_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/mg_bench2.json 2> gpurun_out/mg_bench2.err; tail -c 1500 gpurun_out/mg_bench2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --shard-edges > gpurun_out/mg_bench2_shard.json 2> gpurun_out/mg_bench2_shard.err; tail -c 1200 gpurun_out/mg_bench2_shard.json
