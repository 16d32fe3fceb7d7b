#!/bin/bash
# One GPU round: tests, bench, launch list, full captures of the hot kernels. Run under gpurun.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 | tee gpurun_out/bench_reference.json
if [ "$1" == "profile" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-large > gpurun_out/bench_under_ncu.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:linearize_kernel -s 2 -c 1 -f -o gpurun_out/prof_linearize \
      python tools/run_large_kernels.py 1000 > gpurun_out/prof_linearize.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 2 -c 1 -f -o gpurun_out/prof_spmv \
      python tools/run_large_kernels.py 1000 > gpurun_out/prof_spmv.log 2>&1
fi
