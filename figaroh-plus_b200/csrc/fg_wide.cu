// fg_wide.cu - launchers of the CTA-wide register-tile kernels (fg_wide.cuh): fused samples -> factor for any
// tree / row-group count / up to 767 columns, dense matrix -> factor, and the pairwise merge of wide parts.
//
// Shape classes by the number of state columns p1 (rhs included):
//   A  p1 <= 256 : 256 threads, one column per thread, 48..64-row tiles
//   B  p1 <= 512 : 512 threads, one column per thread, 24..32-row tiles (twice the warps: the column step is
//                  latency bound, profiles/ncu_wide_r02.md)
//   C  p1 <= 768 : 256 threads, three columns per thread, 24-row tiles
#include <algorithm>
#include <cstdlib>

#include "fg_wide.cuh"
#include "fg_wide.h"

namespace fg {

static int wide_class(int p1) { return p1 <= 256 ? 0 : (p1 <= 512 ? 1 : (p1 <= 768 ? 2 : -1)); }

template <int NT, int MSMAX, int S, int NSLOT>
static cudaError_t launch_samples(const Plan& P, const DevPlanPtrs& pp, const double* v, const double* a,
                                  const double* tau, int64_t ld_tau, int64_t n, const double* fw, double* parts,
                                  int grid, int max_smem, int first, cudaStream_t st) {
  constexpr int TR = MSMAX * S;
  const WideLayout L = wide_layout(P.nj, P.ms, S, TR);
  if (L.total > max_smem) return cudaErrorInvalidConfiguration;
  auto kern = wide_accumulate_samples_kernel<NT, MSMAX, S, NSLOT, 1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return e;
  kern<<<grid, NT, L.total, st>>>(pp, v, a, tau, ld_tau, n, fw, parts, first, 0);
  return cudaGetLastError();
}

int wide_ctas_per_sm(int, int) { return 1; }

int wide_samples_per_tile(int ms, int p1) {
  const int cls = wide_class(p1);
  if (cls < 0 || ms > 48) return 0;
  if (cls == 0) return ms <= 6 ? 8 : (ms <= 8 ? 6 : (ms <= 24 ? 2 : 1));
  if (cls == 1) return ms <= 6 ? 5 : (ms <= 8 ? 4 : 1);
  return ms <= 6 ? 4 : (ms <= 8 ? 3 : (ms <= 24 ? 1 : 0));
}

int64_t wide_fw_doubles_per_sample(const Plan& P) { return (int64_t)P.nj * FW_STRIDE; }

cudaError_t wide_launch_samples(const Plan& P, const DevPlanPtrs& pp, const double* q, const double* v, const double* a,
                                const double* tau, int64_t ld_tau, int64_t N, double* parts, int grid, int max_smem,
                                double* fw_buf, int64_t fw_chunk, int* launches, cudaStream_t st) {
  const int cls = wide_class(P.p1), ms = P.ms;
  const int fw_smem = align16((int)sizeof(Plan)) + P.nj * (int)sizeof(DevJoint);
  for (int64_t c0 = 0; c0 < N; c0 += fw_chunk) {
    const int64_t n = std::min(fw_chunk, N - c0);
    wide_forward_kernel<<<(unsigned)((n + 127) / 128), 128, fw_smem, st>>>(pp, q + c0 * P.nq, v + c0 * P.nv, a + c0 * P.nv, n, fw_buf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const double* vc = v + c0 * P.nv;
    const double* ac = a + c0 * P.nv;
    const double* tc = tau ? tau + c0 : nullptr;
    const int first = c0 == 0 ? 1 : 0;
    e = cudaErrorInvalidConfiguration;
#define FG_WD(NT_, MS_, S_, NS_) e = launch_samples<NT_, MS_, S_, NS_>(P, pp, vc, ac, tc, ld_tau, n, fw_buf, parts, grid, max_smem, first, st)
    if (cls == 0) {
      if (ms <= 6) FG_WD(256, 6, 8, 1);
      else if (ms <= 8) FG_WD(256, 8, 6, 1);
      else if (ms <= 24) FG_WD(256, 24, 2, 1);
      else if (ms <= 48) FG_WD(256, 48, 1, 1);
    } else if (cls == 1) {
      if (ms <= 6) FG_WD(512, 6, 5, 1);
      else if (ms <= 8) FG_WD(512, 8, 4, 1);
      else if (ms <= 24) FG_WD(512, 24, 1, 1);
      else if (ms <= 48) FG_WD(512, 48, 1, 1);
    } else if (cls == 2) {
      if (ms <= 6) FG_WD(256, 6, 4, 3);
      else if (ms <= 8) FG_WD(256, 8, 3, 3);
      else if (ms <= 24) FG_WD(256, 24, 1, 3);
    }
#undef FG_WD
    if (e != cudaSuccess) return e;
    if (launches) *launches += 2;
  }
  return cudaSuccess;
}

template <int NT, int TR, int NSLOT>
static cudaError_t launch_dense_t(const double* W, int64_t ldw, const double* tau, int64_t tau_stride, int64_t rows,
                                  int p1, double* parts, int grid, cudaStream_t st) {
  const int smem = 2 * (((TR + 7) / 8) * 8 + 4) * 8;
  wide_accumulate_dense_kernel<NT, TR, NSLOT, 1><<<grid, NT, smem, st>>>(W, ldw, tau, tau_stride, rows, p1, parts);
  return cudaGetLastError();
}

int wide_dense_tile_rows(int p1) {
  const int cls = wide_class(p1);
  return cls == 0 ? 64 : (cls == 1 ? 32 : (cls == 2 ? 24 : 0));
}

cudaError_t wide_launch_dense(const double* W, int64_t ldw, const double* tau, int64_t tau_stride, int64_t rows, int p1,
                              double* parts, int grid, cudaStream_t st) {
  const int cls = wide_class(p1);
  if (cls == 0) return launch_dense_t<256, 64, 1>(W, ldw, tau, tau_stride, rows, p1, parts, grid, st);
  if (cls == 1) return launch_dense_t<512, 32, 1>(W, ldw, tau, tau_stride, rows, p1, parts, grid, st);
  if (cls == 2) return launch_dense_t<256, 24, 3>(W, ldw, tau, tau_stride, rows, p1, parts, grid, st);
  return cudaErrorInvalidConfiguration;
}

template <int NT, int TR, int NSLOT>
static cudaError_t launch_merge_t(double* parts, int nparts, int stride, int p1, cudaStream_t st) {
  const int pairs = (nparts + 2 * stride - 1) / (2 * stride);
  const int smem = 2 * (((TR + 7) / 8) * 8 + 4) * 8;
  wide_merge_pairs_kernel<NT, TR, NSLOT><<<pairs, NT, smem, st>>>(parts, nparts, stride, p1);
  return cudaGetLastError();
}

cudaError_t wide_launch_merge_level(double* parts, int nparts, int stride, int p1, cudaStream_t st) {
  const int cls = wide_class(p1);
  if (cls == 0) return launch_merge_t<256, 64, 1>(parts, nparts, stride, p1, st);
  if (cls == 1) return launch_merge_t<512, 32, 1>(parts, nparts, stride, p1, st);
  if (cls == 2) return launch_merge_t<256, 24, 3>(parts, nparts, stride, p1, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace fg
