#!/bin/bash
# One GPU round: tests, smoke, bench (both arms), launch list, full captures of the hot kernels.
# Run under gpurun:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh profile'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | tee gpurun_out/bench_reference.json
timeout 900 python tools/config_runs.py --large 2>&1 | tee gpurun_out/config_runs.log
if [ "$1" == "profile" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-large > gpurun_out/bench_under_ncu.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 2 -c 1 -f -o gpurun_out/prof_linearize \
      python tools/run_large_kernels.py 1000 > gpurun_out/prof_linearize.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 2 -c 1 -f -o gpurun_out/prof_spmv \
      python tools/run_large_kernels.py 1000 > gpurun_out/prof_spmv.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:level_chol_pcg_kernel -s 3 -c 1 -f -o gpurun_out/prof_chol \
      python tools/kitti_step.py 2 > gpurun_out/prof_chol.log 2>&1
  # the multilevel PCG on the 1M-pose grid: launch list of its iterations and full captures of its two heaviest kernels
  PGO_AMG_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 300 --csv --log-file gpurun_out/launches_amg_grid.csv \
      python tools/amg_check.py --cases grid1000 --no-oracle --reps 1 > gpurun_out/amg_under_ncu.log 2>&1
  PGO_AMG_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:amg_smooth_kernel -s 20 -c 1 -f -o gpurun_out/prof_amg_smooth \
      python tools/amg_check.py --cases grid1000 --no-oracle --reps 1 > gpurun_out/prof_amg_smooth.log 2>&1
  PGO_AMG_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:amg_dense_gj_kernel -s 1 -c 1 -f -o gpurun_out/prof_amg_gj \
      python tools/amg_check.py --cases grid1000 --no-oracle --reps 1 > gpurun_out/prof_amg_gj.log 2>&1
  for k in linearize spmv chol amg_smooth amg_gj; do python tools/ncu_summary.py kernel gpurun_out/prof_$k.ncu-rep > gpurun_out/summary_$k.txt 2>&1; done
  python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/summary_launches_kitti.txt 2>&1
  python tools/ncu_summary.py launches gpurun_out/launches_amg_grid.csv > gpurun_out/summary_launches_amg_grid.txt 2>&1
fi
ls -la gpurun_out
