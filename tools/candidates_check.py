"""Loop-edge candidate search on the GPU vs the CPU oracle: exactness on the golden file and timings.
Run under gpurun:  python tools/candidates_check.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import posegraph_ceres_b200.api as pgo   # noqa: E402
import oracle_py as oracle               # noqa: E402

f = np.load(os.path.join(ROOT, "tests", "golden", "kitti00_fixture.npz"))
pos = f["poses_before"][:, :3]
pgo.edge_candidates(pos)                 # warm-up (context, pools)
t = time.perf_counter(); ptr, idx = pgo.edge_candidates(pos); t_gpu = time.perf_counter() - t
t = time.perf_counter(); optr, oidx = oracle.edge_candidates(pos); t_cpu = time.perf_counter() - t
print(f"KITTI-00 4541 frames: golden exact {np.array_equal(idx, f['cand_idx']) and np.array_equal(ptr[1:], f['cand_ptr'])}, "
      f"GPU {t_gpu * 1e3:.2f} ms (host call, two passes + copies), CPU oracle {t_cpu * 1e3:.2f} ms, {idx.size} candidates")
rng = np.random.default_rng(0)
for n in (50_000, 400_000):
    p = np.cumsum(rng.normal(0, 0.8, (n, 3)), axis=0) % 300.0
    t = time.perf_counter(); ptr, idx = pgo.edge_candidates(p, 6.0, 100); t_gpu = time.perf_counter() - t
    pairs = n * (n - 201) / 2
    line = f"random walk {n} frames: GPU {t_gpu * 1e3:.1f} ms, {2 * pairs / t_gpu / 1e9:.1f} G pair tests/s (two passes), {idx.size} candidates"
    if n <= 50_000:
        t = time.perf_counter(); optr, oidx = oracle.edge_candidates(p, 6.0, 100); t_cpu = time.perf_counter() - t
        line += f"; CPU oracle {t_cpu * 1e3:.0f} ms, exact {np.array_equal(ptr, optr) and np.array_equal(idx, oidx)}"
    print(line)
