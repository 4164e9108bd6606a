// fg_wide.h - host interface of the CTA-wide register-tile kernels (fg_wide.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

constexpr int WD_MAX_P1 = 768;  // three columns per thread

// samples per tile the fused kernel uses for (row groups, columns incl. rhs); 0 = shape not supported
int wide_samples_per_tile(int ms, int p1);
// resident CTAs per SM of the fused kernel for that shape (grid = SMs x this)
int wide_ctas_per_sm(int ms, int p1);
// rows per tile of the dense kernel; 0 = too wide
int wide_dense_tile_rows(int p1);

// doubles of forward-kinematics state per sample ([nj][FW_STRIDE])
int64_t wide_fw_doubles_per_sample(const Plan& P);
// CTA b of `grid` folds tiles b, b + grid, ... into parts[b] (packed triangle + squared column norms).  The
// samples are processed in chunks of fw_chunk: a thread-per-sample kernel leaves the forward kinematics of a
// chunk in fw_buf (fw_chunk * wide_fw_doubles_per_sample doubles), the fold kernel streams it tile by tile.
cudaError_t wide_launch_samples(const Plan& P, const DevPlanPtrs& pp, const double* q, const double* v, const double* a,
                                const double* tau, int64_t ld_tau, int64_t N, double* parts, int grid, int max_smem,
                                double* fw_buf, int64_t fw_chunk, int* launches, cudaStream_t st);
// tau (optional) is the last column: entry `row` at tau[row * tau_stride]
cudaError_t wide_launch_dense(const double* W, int64_t ldw, const double* tau, int64_t tau_stride, int64_t rows, int p1,
                              double* parts, int grid, cudaStream_t st);
// one level of the pairwise tree: parts[(2b + 1) * stride] folded into parts[2b * stride]
cudaError_t wide_launch_merge_level(double* parts, int nparts, int stride, int p1, cudaStream_t st);

}  // namespace fg
