// pgo_candidates.cuh -- loop-edge candidate search, the caller side of the optimisation path:
// REF/test/generate_edges_from_trajectory_origion.cpp:58-110 (getCandidatesIndex / isInSearchRange) produces
// config/Edge_Candidates_index.txt, which pose_graph_ceres_plus_finial.cpp:72,300-323 reads back to decide which
// frame pairs become edges.  For frame c the candidates are  c-1  followed by every  i < c - min_gap  (ascending)
// whose camera centre lies within the search radius, decided in FLOAT arithmetic exactly as the reference does:
//     dist = dx*dx + dy*dy + dz*dz   (three products, two sums, each rounded to fp32; no FMA contraction)
//     candidate  <=>  !(dist > radius*radius)
// Warps own frames; ballot + popc keep the ascending order; a counting pass sizes the CSR output.
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
// A warp owns kFrames consecutive frames and tests each loaded position against all of them (the position array of a long
// trajectory lives in L2, so one frame per warp is L2-bandwidth bound: 12 B per pair test); groups are dealt longest
// first because frame c scans c - min_gap positions.
template <bool kFill, int kFrames>
__global__ void __launch_bounds__(256) edge_candidates_kernel(int n, const float* __restrict__ px, const float* __restrict__ py,
                                                              const float* __restrict__ pz, float r2, int min_gap,
                                                              int* __restrict__ counts, const long long* __restrict__ row_ptr,
                                                              int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int groups = (n + kFrames - 1) / kFrames;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
    const int c0 = (groups - 1 - g) * kFrames;
    float cx[kFrames], cy[kFrames], cz[kFrames];
    int limit[kFrames], count[kFrames];
    long long base[kFrames];
    int max_limit = 0;
#pragma unroll
    for (int f = 0; f < kFrames; ++f) {
      const int c = c0 + f;
      const bool valid = c >= 1 && c < n;
      const int cc = valid ? c : 0;
      cx[f] = __ldg(px + cc); cy[f] = __ldg(py + cc); cz[f] = __ldg(pz + cc);
      limit[f] = valid ? c - min_gap : 0;        // i < c - 100 (REF :65) and c - i > 100 (REF :70) are the same condition
      count[f] = valid ? 1 : 0;                  // c - 1 is always a candidate (REF :61)
      base[f] = (kFill && valid) ? row_ptr[c] : 0;
      if (kFill && valid && lane == 0) idx[base[f]] = c - 1;
      max_limit = max(max_limit, limit[f]);
    }
    for (int i0 = 0; i0 < max_limit; i0 += 32) {
      const int i = i0 + lane;
      const int ii = min(i, n - 1);
      const float x = __ldg(px + ii), y = __ldg(py + ii), z = __ldg(pz + ii);
#pragma unroll
      for (int f = 0; f < kFrames; ++f) {
        const bool hit = i < limit[f] && in_search_range(cx[f], cy[f], cz[f], x, y, z, r2);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (kFill && hit) idx[base[f] + count[f] + __popc(m & lt_mask)] = i;
        count[f] += __popc(m);
      }
    }
    if (!kFill && lane == 0) {
#pragma unroll
      for (int f = 0; f < kFrames; ++f)
        if (c0 + f < n) counts[c0 + f] = count[f];
    }
  }
}

}  // namespace pgo
