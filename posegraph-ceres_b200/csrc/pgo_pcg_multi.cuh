// pgo_pcg_multi.cuh -- stream-ordered block-Jacobi PCG for the multi-GPU path: every rank holds an
// edge shard (a partial off-diagonal Hessian) and replicas of all vectors; the one exchange per
// iteration is an NCCL all-reduce of the SpMV product over NVLink.  Included by pgo_b200.cu.
#pragma once

namespace pgo {

struct PcgMultiState {
  double gamma, gamma_old, delta, alpha, beta, gamma0, acc0, acc1;
  int iter, done, flag, pad;
};

__global__ void __launch_bounds__(256) pcgm_init_kernel(int n, const double* __restrict__ b, const double* __restrict__ Minv,
                                                        double* x, double* r, double* u, double* p, double* s,
                                                        PcgMultiState* st) {
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
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(&st->acc0, acc);
}

// delta = w.u (after the all-reduce of w)
__global__ void __launch_bounds__(256) pcgm_dot_kernel(int n6, const double* __restrict__ w, const double* __restrict__ u,
                                                       PcgMultiState* st) {
  if (st->done) return;
  double acc = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) acc = fma(w[k], u[k], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(&st->acc1, acc);
}

// scalar step between the reductions (1 thread)
__global__ void pcgm_scalar_kernel(PcgMultiState* st, int phase, int max_iterations, double tol) {
  if (phase == 0) {           // after init: gamma0
    st->gamma = st->gamma0 = st->acc0; st->acc0 = 0.0; st->iter = 0; st->flag = 0;
    st->done = (st->gamma0 > 0.0) ? 0 : 1;
  } else if (phase == 1) {    // after delta
    if (st->done) return;
    st->delta = st->acc1; st->acc1 = 0.0;
    if (st->iter == 0) { st->beta = 0.0; st->alpha = st->gamma / st->delta; }
    else { st->beta = st->gamma / st->gamma_old; st->alpha = st->gamma / (st->delta - st->beta * st->gamma / st->alpha); }
    if (!(st->alpha > 0.0) || !isfinite(st->alpha)) { st->flag = 2; st->done = 1; }
    else st->iter++;
  } else {                    // after the update: new gamma
    if (st->done) return;
    st->gamma_old = st->gamma; st->gamma = st->acc0; st->acc0 = 0.0;
    if (st->gamma <= tol * tol * st->gamma0) { st->flag = 0; st->done = 1; }
    else if (st->iter >= max_iterations) { st->flag = 1; st->done = 1; }
  }
}

__global__ void __launch_bounds__(256) pcgm_update_kernel(int n, const double* __restrict__ Minv, const double* __restrict__ w,
                                                          double* x, double* r, double* u, double* p, double* s,
                                                          PcgMultiState* st) {
  if (st->done) return;
  const double alpha = st->alpha, beta = st->beta;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    double rv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const size_t q = 6 * (size_t)i + k;
      const double pv = u[q] + beta * p[q];
      const double sv = w[q] + beta * s[q];
      p[q] = pv; s[q] = sv;
      x[q] += alpha * pv;
      rv[k] = r[q] - alpha * sv;
      r[q] = rv[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) t = fma(Minv[36 * (size_t)i + 6 * k + c], rv[c], t);
      u[6 * (size_t)i + k] = t;
      acc += rv[k] * t;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(&st->acc0, acc);
}

// epilogue: x^T b, x^T A x (w = A x all-reduced), x^T D x -> DeviceScalars
__global__ void __launch_bounds__(256) pcgm_final_kernel(int n6, const double* __restrict__ x, const double* __restrict__ b,
                                                         const double* __restrict__ w, const double* __restrict__ d,
                                                         const PcgMultiState* st, DeviceScalars* sc) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) {
    const double xv = x[k];
    a0 = fma(xv, b[k], a0); a1 = fma(xv, w[k], a1); a2 = fma(xv * xv, d[k], a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sc->xtb, a0); atomicAdd(&sc->xtAx, a1); atomicAdd(&sc->xtDx, a2); }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->pcg_gamma0 = st->gamma0; sc->pcg_gamma = st->gamma; sc->pcg_iterations = st->iter; sc->pcg_flag = st->flag;
  }
}

}  // namespace pgo

// (H + diag(dlm)) x = b across g->world ranks. Stream-ordered; the host polls `done` every
// kCheckEvery iterations (kernels after convergence are no-ops).
static int pcg_multi(pgo_graph* g, const pgo_solver_options* o, const double* b) {
  using namespace pgo;
  static thread_local PcgMultiState* st = nullptr;
  static thread_local PcgMultiState* st_h = nullptr;
  if (!st) {
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&st), sizeof(PcgMultiState)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&st_h), sizeof(PcgMultiState)));
  }
  const int N = g->N, n6 = 6 * N;
  const int nb = (N + 255) / 256;
  const int warps = (N + kRowsPerWarp - 1) / kRowsPerWarp;
  const int sp_ctas = std::max(1, std::min((warps + 7) / 8, 8 * g->num_sms));
  const int dot_ctas = std::max(1, std::min((n6 + 255) / 256, 4 * g->num_sms));
  const bool with_diag = (g->rank == 0);
  CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(PcgMultiState), g->stream));
  pcgm_init_kernel<<<nb, 256, 0, g->stream>>>(N, b, g->Minv, g->vx, g->vr, g->vu, g->vp, g->vs, st);
  pcgm_scalar_kernel<<<1, 1, 0, g->stream>>>(st, 0, o->pcg_max_iterations, o->pcg_tolerance);
  g->launches += 2;
  const int kCheckEvery = 32;
  int launched = 0;
  for (;;) {
    for (int k = 0; k < kCheckEvery; ++k) {
      spmv_kernel<<<sp_ctas, 256, 0, g->stream>>>(bsr_view(g), g->vu, g->dlm, g->vw, with_diag);
      PGO_TRY(allreduce_sum(g, g->vw, (size_t)n6));
      pcgm_dot_kernel<<<dot_ctas, 256, 0, g->stream>>>(n6, g->vw, g->vu, st);
      pcgm_scalar_kernel<<<1, 1, 0, g->stream>>>(st, 1, o->pcg_max_iterations, o->pcg_tolerance);
      pcgm_update_kernel<<<nb, 256, 0, g->stream>>>(N, g->Minv, g->vw, g->vx, g->vr, g->vu, g->vp, g->vs, st);
      pcgm_scalar_kernel<<<1, 1, 0, g->stream>>>(st, 2, o->pcg_max_iterations, o->pcg_tolerance);
      g->launches += 5;
      ++launched;
    }
    CUDA_TRY(cudaMemcpyAsync(st_h, st, sizeof(PcgMultiState), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (st_h->done || launched >= o->pcg_max_iterations + kCheckEvery) break;
  }
  // epilogue: w = A x (all-reduced) for the model cost change
  spmv_kernel<<<sp_ctas, 256, 0, g->stream>>>(bsr_view(g), g->vx, g->dlm, g->vw, with_diag);
  PGO_TRY(allreduce_sum(g, g->vw, (size_t)n6));
  pcgm_final_kernel<<<dot_ctas, 256, 0, g->stream>>>(n6, g->vx, b, g->vw, g->dlm, st, g->scalars);
  g->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}
