cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/kitti_timeline.py kitti 2 > gpurun_out/timeline.log 2>&1; grep -m2 "timeline" gpurun_out/timeline.log | cut -c1-700; grep -E "create|pgo host" gpurun_out/timeline.log | tail -3
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-large 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
PGO_CHOL_SHAPE=cluster timeout 300 python bench.py --steps 20 --warmup 3 --no-large 2>>gpurun_out/bench.err | tee gpurun_out/bench_cluster.json | cut -c1-200
