// pgo_candidates.cuh -- loop-edge candidate search, the caller side of the optimisation path:
// REF/test/generate_edges_from_trajectory_origion.cpp:58-110 (getCandidatesIndex / isInSearchRange) produces
// config/Edge_Candidates_index.txt, which pose_graph_ceres_plus_finial.cpp:72,300-323 reads back to decide which
// frame pairs become edges.  For frame c the candidates are  c-1  followed by every  i < c - min_gap  (ascending)
// whose camera centre lies within the search radius, decided in FLOAT arithmetic exactly as the reference does:
//     dist = dx*dx + dy*dy + dz*dz   (three products, two sums, each rounded to fp32; no FMA contraction)
//     candidate  <=>  !(dist > radius*radius)
// One warp per frame; ballot + popc keep the ascending order; a counting pass sizes the CSR output.
#pragma once

#include "pgo_common.cuh"

namespace pgo {

__device__ __forceinline__ bool in_search_range(float cx, float cy, float cz, float x, float y, float z, float r2) {
  const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
  const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return !(dist > r2);
}

// kFill = false: counts[c] = number of candidates of frame c (c = 1 .. n-1, counts[0] = 0)
// kFill = true : idx[row_ptr[c] ...] = the candidates of frame c
template <bool kFill>
__global__ void __launch_bounds__(256) edge_candidates_kernel(int n, const float* __restrict__ px, const float* __restrict__ py,
                                                              const float* __restrict__ pz, float r2, int min_gap,
                                                              int* __restrict__ counts, const long long* __restrict__ row_ptr,
                                                              int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += warps) {
    if (c == 0) { if (!kFill && lane == 0) counts[0] = 0; continue; }
    const float cx = __ldg(px + c), cy = __ldg(py + c), cz = __ldg(pz + c);
    long long base = kFill ? row_ptr[c] : 0;
    int count = 1;                               // c - 1 is always a candidate (REF :61)
    if (kFill && lane == 0) idx[base] = c - 1;
    const int limit = c - min_gap;               // i < c - 100 (REF :65) and c - i > 100 (REF :70) are the same condition
    for (int i0 = 0; i0 < limit; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < limit && in_search_range(cx, cy, cz, __ldg(px + i), __ldg(py + i), __ldg(pz + i), r2);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (kFill && hit) idx[base + count + __popc(m & ((1u << lane) - 1u))] = i;
      count += __popc(m);
    }
    if (!kFill && lane == 0) counts[c] = count;
  }
}

}  // namespace pgo
