#!/bin/bash
# Short GPU loop for development: parity tests, then the KITTI bench line (no 1M-grid section).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-large 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'solver ms/launch', d['roofline']['ms_per_launch'], 'split', d['time_split_ms_per_step'], 'pcg', d['config']['pcg_iterations_per_solve'])
PY
