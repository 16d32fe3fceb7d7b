// pgo_pcg_multi.cuh -- stream-ordered block-Jacobi PCG (one GPU): separate launches at full occupancy for graphs of
// >= 200 k poses, where the grid barriers of the persistent kernel no longer pay.  (Multi-GPU solves use the
// row-partitioned multilevel PCG of pgo_amg.cuh.)  Included by pgo_b200.cu.
#pragma once

namespace pgo {

struct PcgMultiState {
  double gamma, gamma_old, delta, alpha, beta, gamma0, acc0, acc1;
  int iter, done, flag, pad;
};

// Reductions are two-stage and fixed-order (per-CTA partial -> one CTA sums the partials): every rank computes
// bit-identical scalars from bit-identical inputs, so all ranks take the same CG / LM decisions and issue the same
// sequence of collectives.
constexpr int kPcgmThreads = 256;
__device__ __forceinline__ void pcgm_block_partial(double v, double* out) {
  __shared__ double red[kPcgmThreads / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kPcgmThreads / 32; ++k) t += red[k];
    out[blockIdx.x] = t;
  }
  __syncthreads();
}
__device__ __forceinline__ double pcgm_sum_partials(const double* part, int n) {   // one CTA, result on every thread
  __shared__ double red[kPcgmThreads / 32];
  __shared__ double total;
  double t = 0.0;
  for (int k = threadIdx.x; k < n; k += kPcgmThreads) t += part[k];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPcgmThreads / 32; ++k) s += red[k];
    total = s;
  }
  __syncthreads();
  const double r = total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kPcgmThreads) pcgm_init_kernel(int n, const double* __restrict__ b, const double* __restrict__ Minv,
                                                                 double* x, double* r, double* u, double* p, double* s,
                                                                 double* part0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    double rv[6], uv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) rv[k] = b[6 * (size_t)i + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) t = fma(Minv[36 * (size_t)i + 6 * k + c], rv[c], t);
      uv[k] = t;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const size_t q = 6 * (size_t)i + k;
      x[q] = 0.0; r[q] = rv[k]; u[q] = uv[k]; p[q] = 0.0; s[q] = 0.0;
      acc += rv[k] * uv[k];
    }
  }
  pcgm_block_partial(acc, part0);
}

// delta = w.u (after the all-reduce of w): per-CTA partials
__global__ void __launch_bounds__(kPcgmThreads) pcgm_dot_kernel(int n6, const double* __restrict__ w, const double* __restrict__ u,
                                                                const PcgMultiState* st, double* part1) {
  double acc = 0.0;
  if (!st->done)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) acc = fma(w[k], u[k], acc);
  pcgm_block_partial(acc, part1);
}

// scalar step between the reductions (one CTA): sums the partials in a fixed order, thread 0 updates the state
__global__ void __launch_bounds__(kPcgmThreads) pcgm_scalar_kernel(PcgMultiState* st, int phase, int max_iterations, double tol,
                                                                   const double* part, int nparts) {
  const double sum = pcgm_sum_partials(part, nparts);
  if (threadIdx.x != 0) return;
  if (phase == 0) {           // after init: gamma0
    st->gamma = st->gamma0 = sum; st->iter = 0; st->flag = 0;
    st->done = (st->gamma0 > 0.0) ? 0 : 1;
  } else if (phase == 1) {    // after delta
    if (st->done) return;
    st->delta = sum;
    if (st->iter == 0) { st->beta = 0.0; st->alpha = st->gamma / st->delta; }
    else { st->beta = st->gamma / st->gamma_old; st->alpha = st->gamma / (st->delta - st->beta * st->gamma / st->alpha); }
    if (!(st->alpha > 0.0) || !isfinite(st->alpha)) { st->flag = 2; st->done = 1; }
    else st->iter++;
  } else {                    // after the update: new gamma
    if (st->done) return;
    st->gamma_old = st->gamma; st->gamma = sum;
    if (st->gamma <= tol * tol * st->gamma0) { st->flag = 0; st->done = 1; }
    else if (st->iter >= max_iterations) { st->flag = 1; st->done = 1; }
  }
}

// p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s, u = Minv r, partial of r.u.  Six lanes per pose (lane =
// component), five poses per warp: the vector traffic is contiguous and a Minv row is three 128-bit loads.
__global__ void __launch_bounds__(kPcgmThreads) pcgm_update_kernel(int n, const double* __restrict__ Minv, const double* __restrict__ w,
                                                                   double* x, double* r, double* u, double* p, double* s,
                                                                   const PcgMultiState* st, double* part0) {
  const int lane = threadIdx.x & 31;
  const int grp = lane / 6, c6 = lane - grp * 6;
  const int warps_per_cta = kPcgmThreads / 32;
  const int i = (blockIdx.x * warps_per_cta + (threadIdx.x >> 5)) * kRowsPerWarp + grp;
  const bool on = !st->done && grp < kRowsPerWarp && i < n;
  double acc = 0.0, rv = 0.0;
  const size_t q = 6 * (size_t)(on ? i : 0) + c6;
  if (on) {
    const double alpha = st->alpha, beta = st->beta;
    const double pv = u[q] + beta * p[q];
    const double sv = w[q] + beta * s[q];
    p[q] = pv; s[q] = sv;
    x[q] += alpha * pv;
    rv = r[q] - alpha * sv;
    r[q] = rv;
  }
  const int g0 = grp * 6;
  const double r0 = __shfl_sync(0xffffffffu, rv, g0), r1 = __shfl_sync(0xffffffffu, rv, g0 + 1), r2 = __shfl_sync(0xffffffffu, rv, g0 + 2);
  const double r3 = __shfl_sync(0xffffffffu, rv, g0 + 3), r4 = __shfl_sync(0xffffffffu, rv, g0 + 4), r5 = __shfl_sync(0xffffffffu, rv, g0 + 5);
  if (on) {
    const double2* mi = reinterpret_cast<const double2*>(Minv + 36 * (size_t)i + 6 * c6);
    const double2 m0 = __ldg(mi), m1 = __ldg(mi + 1), m2 = __ldg(mi + 2);
    const double t = m0.x * r0 + m0.y * r1 + m1.x * r2 + m1.y * r3 + m2.x * r4 + m2.y * r5;
    u[q] = t;
    acc = rv * t;
  }
  pcgm_block_partial(acc, part0);
}

// epilogue: x^T b, x^T A x (w = A x all-reduced), x^T D x -> per-CTA partials, then one CTA -> DeviceScalars
__global__ void __launch_bounds__(kPcgmThreads) pcgm_final_kernel(int n6, const double* __restrict__ x, const double* __restrict__ b,
                                                                  const double* __restrict__ w, const double* __restrict__ d,
                                                                  double* part /* [3][gridDim.x] */) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) {
    const double xv = x[k];
    a0 = fma(xv, b[k], a0); a1 = fma(xv, w[k], a1); a2 = fma(xv * xv, d[k], a2);
  }
  pcgm_block_partial(a0, part);
  pcgm_block_partial(a1, part + gridDim.x);
  pcgm_block_partial(a2, part + 2 * gridDim.x);
}
__global__ void __launch_bounds__(kPcgmThreads) pcgm_final_reduce_kernel(const double* part, int nparts, const PcgMultiState* st,
                                                                         DeviceScalars* sc) {
  const double e0 = pcgm_sum_partials(part, nparts);
  const double e1 = pcgm_sum_partials(part + nparts, nparts);
  const double e2 = pcgm_sum_partials(part + 2 * nparts, nparts);
  if (threadIdx.x == 0) {
    sc->xtb = e0; sc->xtAx = e1; sc->xtDx = e2;
    sc->pcg_gamma0 = st->gamma0; sc->pcg_gamma = st->gamma; sc->pcg_iterations = st->iter; sc->pcg_flag = st->flag;
  }
}

}  // namespace pgo

// (H + diag(dlm)) x = b. Stream-ordered; the host polls `done` every
// kCheckEvery iterations (kernels after convergence are no-ops).
static int pcg_multi(pgo_graph* g, const pgo_solver_options* o, const double* b) {
  using namespace pgo;
  const int N = g->N, n6 = 6 * N;
  const int nb = (N + kPcgmThreads - 1) / kPcgmThreads;
  const int nbu = (N + (kPcgmThreads / 32) * kRowsPerWarp - 1) / ((kPcgmThreads / 32) * kRowsPerWarp);   // update kernel: 5 poses per warp
  const int warps = (N + kRowsPerWarp - 1) / kRowsPerWarp;
  const int sp_ctas = std::max(1, std::min((warps + 7) / 8, 8 * g->num_sms));
  const int dot_ctas = std::max(1, std::min((n6 + kPcgmThreads - 1) / kPcgmThreads, 4 * g->num_sms));
  if (!g->pcgm_state) {
    PGO_TRY(dev_alloc(g, &g->pcgm_state, 1));
    PGO_TRY(dev_alloc(g, &g->pcgm_part0, (size_t)std::max(nb, nbu)));
    PGO_TRY(dev_alloc(g, &g->pcgm_part1, (size_t)std::max(3 * dot_ctas, sp_ctas)));
    CUDA_TRY(pool_pinned(g->device, reinterpret_cast<void**>(&g->pcgm_state_h)));
  }
  PcgMultiState* st = g->pcgm_state;
  PcgMultiState* st_h = g->pcgm_state_h;
  double *part0 = g->pcgm_part0, *part1 = g->pcgm_part1;
  const bool with_diag = true;
  CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(PcgMultiState), g->stream));
  pcgm_init_kernel<<<nb, kPcgmThreads, 0, g->stream>>>(N, b, g->Minv, g->vx, g->vr, g->vu, g->vp, g->vs, part0);
  pcgm_scalar_kernel<<<1, kPcgmThreads, 0, g->stream>>>(st, 0, o->pcg_max_iterations, o->pcg_tolerance, part0, nb);
  g->launches += 2;
  const int kCheckEvery = 32;
  int launched = 0;
  for (;;) {
    for (int k = 0; k < kCheckEvery; ++k) {
      // w . u rides on the SpMV (per-CTA partials in a fixed order)
      spmv_kernel<true><<<sp_ctas, 256, 0, g->stream>>>(bsr_view(g), g->vu, g->dlm, g->vw, with_diag, part1, &st->done);
      pcgm_scalar_kernel<<<1, kPcgmThreads, 0, g->stream>>>(st, 1, o->pcg_max_iterations, o->pcg_tolerance, part1, sp_ctas);
      pcgm_update_kernel<<<nbu, kPcgmThreads, 0, g->stream>>>(N, g->Minv, g->vw, g->vx, g->vr, g->vu, g->vp, g->vs, st, part0);
      pcgm_scalar_kernel<<<1, kPcgmThreads, 0, g->stream>>>(st, 2, o->pcg_max_iterations, o->pcg_tolerance, part0, nbu);
      g->launches += 4;
      ++launched;
    }
    CUDA_TRY(cudaMemcpyAsync(st_h, st, sizeof(PcgMultiState), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (st_h->done || launched >= o->pcg_max_iterations + kCheckEvery) break;
  }
  // epilogue: w = A x for the model cost change
  spmv_kernel<false><<<sp_ctas, 256, 0, g->stream>>>(bsr_view(g), g->vx, g->dlm, g->vw, with_diag);
  pcgm_final_kernel<<<dot_ctas, kPcgmThreads, 0, g->stream>>>(n6, g->vx, b, g->vw, g->dlm, part1);
  pcgm_final_reduce_kernel<<<1, kPcgmThreads, 0, g->stream>>>(part1, dot_ctas, st, g->scalars);
  g->launches += 3;
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}
