// fg_coop.cu - host side of the cooperative blocked TSQR (fg_coop.cuh).
#include <cstring>

#include "fg_coop.cuh"
#include "fg_coop.h"

namespace fg {

static int cq_smem_bytes(int rpl, int nb) {
  return (int)(((size_t)rpl * nb * GQ_BLK + (size_t)nb * 64 + (size_t)CQ_NW * GQ_SCRATCH + 2 * (size_t)nb) * 8);
}

int cq_max_p1(int max_smem) {
  int nb = 0;
  while (cq_smem_bytes(1, nb + 1) <= max_smem) ++nb;
  return nb * 8;
}

int64_t cq_r_doubles(int p1) {
  const int64_t nb = (p1 + 7) / 8;
  return nb * (nb + 1) / 2 * 64;
}

static cudaError_t cq_launch(const CqParams& P, int grid, int max_smem, cudaStream_t st) {
  const int rpl = cq_smem_bytes(2, P.nb) <= max_smem ? 2 : 1;
  const int smem = cq_smem_bytes(rpl, P.nb);
  if (smem > max_smem) return cudaErrorInvalidValue;
  cudaError_t e;
  if (rpl == 2) {
    e = cudaFuncSetAttribute(cq_accumulate_dense_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cq_accumulate_dense_kernel<2><<<grid, CQ_NW * 32, smem, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(cq_accumulate_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cq_accumulate_dense_kernel<1><<<grid, CQ_NW * 32, smem, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t cq_launch_dense(const CqArgs& a, double* r_all, int grid, int max_smem, cudaStream_t st) {
  CqParams P;
  memset(&P, 0, sizeof P);
  P.W = a.W; P.ldw = a.ldw; P.colsel = a.colsel; P.tau = a.tau; P.rows = a.rows; P.p1 = a.p1;
  P.nb = (a.p1 + 7) / 8;
  P.r_all = r_all;
  P.tau_blk = a.tau_blk; P.tau_ld = a.tau_ld;
  P.first = a.first ? 1 : 0;
  P.rhs_in_w = a.rhs_in_w ? 1 : 0;
  return cq_launch(P, grid, max_smem, st);
}

cudaError_t cq_reduce(double* r_all, int grid, int p1, int max_smem, int* launches, cudaStream_t st) {
  CqParams P;
  memset(&P, 0, sizeof P);
  P.p1 = p1; P.nb = (p1 + 7) / 8; P.r_all = r_all; P.nfac = grid;
  for (int stride = 1; stride < grid; stride *= 2) {
    P.merge_stride = stride;
    const int pairs = (grid + 2 * stride - 1) / (2 * stride);
    cudaError_t e = cq_launch(P, pairs, max_smem, st);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  return cudaSuccess;
}

cudaError_t cq_export(const double* r_all, int p1, double* part_out, cudaStream_t st) {
  cq_export_kernel<<<32, 256, 0, st>>>(r_all, p1, (p1 + 7) / 8, part_out);
  return cudaGetLastError();
}

cudaError_t cq_scatter_rows(const double* R, int p1g, const int* colmap, int p1, double* D, cudaStream_t st) {
  cq_scatter_rows_kernel<<<32, 256, 0, st>>>(R, p1g, (p1g + 7) / 8, colmap, p1, D);
  return cudaGetLastError();
}

cudaError_t cq_import_dense(const double* packed, int n, int p1, double* r_all, double* rows_acc, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(r_all, 0, sizeof(double) * (size_t)n * cq_r_doubles(p1), st);
  if (e != cudaSuccess) return e;
  cq_import_dense_kernel<<<dim3(16, n), 256, 0, st>>>(packed, n, p1, (p1 + 7) / 8, r_all, rows_acc);
  return cudaGetLastError();
}

cudaError_t cq_reduce_export(double* r_all, int grid, int p1, double* part_out, int max_smem, int* launches, cudaStream_t st) {
  cudaError_t e = cq_reduce(r_all, grid, p1, max_smem, launches, st);
  if (e != cudaSuccess) return e;
  if (launches) ++*launches;
  return cq_export(r_all, p1, part_out, st);
}

}  // namespace fg
