// pgo_level_chol.cuh -- level-scheduled block Cholesky preconditioner (placeholder: analysis
// reports "not usable", so PGO_LINEAR_AUTO resolves to block-Jacobi PCG).
#pragma once
#include "pgo_kernels.cuh"

namespace pgo {
struct LevelChol {
  bool usable = false;
  long long factor_blocks = 0;
  int num_levels = 0;
};
static int level_chol_analyze(LevelChol** out, int, const unsigned char*, const int*, const int*, double, cudaStream_t) {
  *out = new LevelChol();
  return 0;
}
static void level_chol_destroy(LevelChol* c) { delete c; }
static int level_chol_factor(LevelChol*, BsrView, const double*, cudaStream_t, long long*) { return -5; }
static int level_chol_pcg(LevelChol*, BsrView, const double*, const double*, double*, double*, double*, double*, double*,
                          int, double, DeviceScalars*, cudaStream_t, long long*) { return -5; }
}  // namespace pgo
