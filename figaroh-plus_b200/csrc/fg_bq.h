// fg_bq.h - host interface of the DMMA block-Householder fused kernel (fg_bq.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

struct BqLaunch {
  int ncb, nrb, nwarps, wpr, nR, grid, smem, nparts;
  bool ok;
};

bool bq_supported(const Plan& P);
BqLaunch bq_plan_launch(const Plan& P, int64_t N, int sm_count, int max_smem);
cudaError_t bq_launch(const BqLaunch& bl, const DevPlanPtrs& pp, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* parts, cudaStream_t st);

}  // namespace fg
