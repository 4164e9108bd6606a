// fg_bq.cu - translation unit of the DMMA block-Householder fused kernel (fg_blockqr.cuh) and its
// launcher.  Kept apart from fg_api.cu so that either side recompiles on its own.
#include <algorithm>

#include "fg_blockqr.cuh"
#include "fg_bq.h"

namespace fg {

bool bq_supported(const Plan& P) { return P.ms >= 1 && P.ms <= 6 && P.p1 <= 80 && P.nj <= kMaxJoints; }

static int bq_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

BqLaunch bq_plan_launch(const Plan& P, int64_t N, int sm_count, int max_smem) {
  BqLaunch bl;
  bl.ncb = 10; bl.nrb = 3;
  bl.nwarps = bq_env_int("FG_BQ_WARPS", 8);
  bl.wpr = bq_env_int("FG_BQ_WPR", 4);
  if (bl.nwarps < 1 || bl.nwarps > 12) bl.nwarps = 8;
  if (bl.wpr < 1 || bl.nwarps % bl.wpr) bl.wpr = bl.nwarps;
  bl.nR = bl.nwarps / bl.wpr;
  const BqLayout L = bq_layout(P.nj, P.ms, P.nv, bl.ncb, bl.nrb, bl.nwarps, bl.nR);
  bl.smem = L.total;
  bl.ok = L.total <= max_smem;
  const int64_t ntiles = (N + BQ_S - 1) / BQ_S;
  int64_t grid = std::min<int64_t>((ntiles + bl.nwarps - 1) / bl.nwarps, sm_count);
  bl.grid = (int)std::max<int64_t>(grid, 1);
  bl.nparts = bl.grid * bl.nR;
  return bl;
}

cudaError_t bq_launch(const BqLaunch& bl, const DevPlanPtrs& pp, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* parts, cudaStream_t st) {
  auto kern = bq_accumulate_samples_kernel<10, 3, 256>;
  if (bl.nwarps > 10) kern = bq_accumulate_samples_kernel<10, 3, 384>;
  else if (bl.nwarps > 8) kern = bq_accumulate_samples_kernel<10, 3, 320>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bl.smem);
  if (e != cudaSuccess) return e;
  kern<<<bl.grid, bl.nwarps * 32, bl.smem, st>>>(pp, q, v, a, tau, ld_tau, N, parts, bl.wpr);
  return cudaGetLastError();
}

}  // namespace fg
