cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for occ in 2 3; do for kind in info identity; do PGO_LIN_OCC=$occ timeout 200 python tools/run_large_kernels.py 1000 $kind 2>&1 | grep linearize | tail -2; done; done | tee gpurun_out/lin_variants.log
