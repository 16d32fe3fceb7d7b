// pgo_common.cuh -- shared layouts and small device helpers of the B200 pose-graph solver.
//
// HBM layout (all fp64 unless noted), N poses, E edges, tiles of 32 edges (one warp per tile):
//   poses      [N][8]     x y z qx qy qz qw pad         64 B / pose, two 32 B halves
//   scale      [N][6]     Jacobi column scaling (0 for constant / unused poses)
//   edge_core  [T] tiles  { int a[32]; int b[32]; int slot_ab[32]; int slot_ba[32]; double meas[7][32]; }
//                         2304 B / tile, field-major so that lane l reads consecutive words
//   edge_info  [T] tiles  { double S[36][32]; }  9216 B / tile, absent when every sqrt_information is I
//   Hdiag      [N][36]    diagonal 6x6 blocks,   "panel" layout: (r,c) at (c/2)*12 + r*2 + (c&1)
//   Hoff       [nnz][36]  off-diagonal blocks of the block-CSR Hessian, same panel layout
//   row_ptr    [N+1], col_idx [nnz]   (int32) block-CSR of the off-diagonal part
//   vectors    [N][6]     gradient, PCG vectors, steps
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pgo {

constexpr int kTile = 32;                         // edges per tile == warp size
constexpr int kCoreTileBytes = 4 * 32 * 4 + 7 * 32 * 8;   // 2304
constexpr int kInfoTileBytes = 36 * 32 * 8;               // 9216
constexpr int kLinWarps = 4;                      // warps per CTA in the linearize kernel
constexpr int kRowsPerWarp = 5;                   // 6 lanes per block row, 30 active lanes
constexpr int kPcgThreads = 256;

__host__ __device__ __forceinline__ constexpr int pidx(int r, int c) { return (c >> 1) * 12 + r * 2 + (c & 1); }

struct EdgeCoreTile {
  int a[32];
  int b[32];
  int slot_ab[32];
  int slot_ba[32];
  double meas[7][32];
};
static_assert(sizeof(EdgeCoreTile) == kCoreTileBytes, "core tile layout");

struct EdgeInfoTile {
  double S[36][32];
};
static_assert(sizeof(EdgeInfoTile) == kInfoTileBytes, "info tile layout");

// Scalars reduced on the device and read back by the host once per LM iteration.
struct DeviceScalars {
  double cost;          // sum 0.5 rho(|r|^2)
  double step_norm2;    // |x - x_cand|^2 over active poses (ambient)
  double x_norm2;       // |x_cand|^2 over active poses (ambient)
  unsigned long long gmax_bits;  // max |x - Plus(x,-g)| as raw double bits (non-negative)
  double gnorm2;
  // PCG results
  double pcg_gamma0;
  double pcg_gamma;
  double xtb;
  double xtAx;
  double xtDx;
  int pcg_iterations;
  int pcg_flag;
};

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA, SASS: UBLKCP) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ceres::HuberLoss / CauchyLoss / NULL: returns rho(s), writes rho'(s).
__device__ __forceinline__ double loss_eval(int type, double a, double s, double& rho1) {
  if (type == 1) {
    const double b = a * a;
    if (s > b) {
      const double r = sqrt(s);
      rho1 = fmax(a / r, 2.2250738585072014e-308);
      return 2.0 * a * r - b;
    }
    rho1 = 1.0;
    return s;
  } else if (type == 2) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    rho1 = fmax(1.0 / sum, 2.2250738585072014e-308);
    return b * log(sum);
  }
  rho1 = 1.0;
  return s;
}

// EigenQuaternionParameterization::Plus + identity on p. x: 7 values, d: 6 values.
__device__ __forceinline__ void pose_plus(const double* x, const double* d, double* out) {
  out[0] = x[0] + d[0]; out[1] = x[1] + d[1]; out[2] = x[2] + d[2];
  const double n = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
  if (n > 0.0) {
    double sn, cs;
    sincos(n, &sn, &cs);
    const double k = sn / n;
    const double ax = k * d[3], ay = k * d[4], az = k * d[5], aw = cs;
    const double bx = x[3], by = x[4], bz = x[5], bw = x[6];
    out[3] = aw * bx + ax * bw + ay * bz - az * by;
    out[4] = aw * by + ay * bw + az * bx - ax * bz;
    out[5] = aw * bz + az * bw + ax * by - ay * bx;
    out[6] = aw * bw - ax * bx - ay * by - az * bz;
  } else {
    out[3] = x[3]; out[4] = x[4]; out[5] = x[5]; out[6] = x[6];
  }
}

}  // namespace pgo
