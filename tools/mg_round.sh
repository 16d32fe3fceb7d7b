#!/bin/bash
# N-GPU round (default 2): row-partitioned solve check, bench (both arms). Run: gpurun --gpus 2 -- 'bash tools/mg_round.sh 2'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/amg_check.py 2>&1 | grep -E "world=|AMG_CHECK" | cut -c1-600 | tee gpurun_out/mg_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/mg_bench2.err | tail -1 > gpurun_out/mg_bench2.json; cut -c1-400 gpurun_out/mg_bench2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>gpurun_out/mg_ref2.err | tail -1 > gpurun_out/mg_ref2.json; cut -c1-300 gpurun_out/mg_ref2.json
