// fg_prefilter.cuh - the step immediately ahead of the regressor kernels:
//
//   differentiate_kernel : calculate_first_second_order_differentiation
//                          (src/figaroh/identification/identification_tools.py:334-387 of the reference):
//                          dq = forward difference / dt, ddq = np.gradient(dq) / dt with the LAST column left
//                          at zero (the reference loops over range(model.nq - 1)), trimmed to N - 2 rows;
//   iir_*_kernel         : scipy.signal.filtfilt(b, a, x, axis=0, padtype="odd", padlen=...) of
//                          low_pass_filter_data (identification_tools.py:390-424), parallel in time.
//
// filtfilt is two runs of the direct-form-II-transposed recurrence  y = b0 x + z0,
// z_i = b_{i+1} x - a_{i+1} y + z_{i+1}  (scipy's lfilter) over the odd-extended signal, forwards and
// backwards.  The recurrence is linear in (state, input), so the signal is cut into chunks of IIR_L samples:
//   1. iir_chunk_kernel<false>: every (chunk, column) runs the recurrence from a ZERO state and keeps the
//      final state s_k;
//   2. iir_propagate_kernel   : one thread per column walks the chunks, z_{k+1} = A^L z_k + s_k, from
//      scipy's initial state zi * x[0] (A: the homogeneous part of the recurrence, A^L from the host);
//   3. iir_chunk_kernel<true> : every (chunk, column) reruns its chunk from its true initial state z_k and
//      writes y.
// Everything fp64; the result equals the sequential filter up to summation order (1e-13 relative).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fg {

constexpr int IIR_MAXN = 8;   // filter order (len(a) - 1)
constexpr int IIR_L = 256;    // samples per chunk

struct IirCoef {
  int n;                      // number of states: the order of the (b, a) filter, or 2 * nsec
  int nsec;                   // > 0: cascade of second-order sections (scipy.signal.sosfilt), state z[2 s], z[2 s + 1]
  double sos[IIR_MAXN / 2][6];  // b0 b1 b2 1 a1 a2 of every section
  double b[IIR_MAXN + 1];     // normalised (a[0] == 1)
  double a[IIR_MAXN + 1];
  double zi[IIR_MAXN];        // lfilter_zi(b, a)
  double AL[IIR_MAXN * IIR_MAXN];  // A^IIR_L, row-major n x n
};

// Where the samples of a pass come from / go to.  Forward pass: input = odd extension of x (N rows), output
// y1 (M = N + 2 padlen rows).  Backward pass: input = y1 read backwards, output either y2 (M rows, backward
// order) or, for the final result, the rows [padlen + trim, padlen + N - trim) of the filtered signal.
struct IirIo {
  const double* x;     // forward: original signal; backward: y1
  int64_t ldx;
  int64_t N;           // rows of the original signal
  int64_t M;           // rows of the extended signal
  int padlen;
  int ncols;
  int backward;
  double* y;           // forward: y1 (ld = ncols); backward: out
  int64_t ldy;
  int64_t trim;        // backward only: rows dropped at both ends of the N-row result
  int64_t step;        // backward only: keep every step-th row of the (trimmed) result (scipy.signal.decimate's y[::q])
};

__device__ __forceinline__ double iir_input(const IirIo& io, int64_t e, int c) {
  if (io.backward) return io.x[(io.M - 1 - e) * io.ldx + c];
  const int64_t N = io.N, p = io.padlen;
  if (e < p) return 2.0 * io.x[c] - io.x[(p - e) * io.ldx + c];                       // 2 x[0] - x[p - e]
  if (e < p + N) return io.x[(e - p) * io.ldx + c];
  return 2.0 * io.x[(N - 1) * io.ldx + c] - io.x[(N - 2 - (e - p - N)) * io.ldx + c];  // 2 x[-1] - x[-2 - i]
}

// APPLY = false: zero initial state, final state -> S[(k * ncols + c) * IIR_MAXN + i]
// APPLY = true : initial state from Z0 (same layout), outputs written
template <bool APPLY>
__global__ void __launch_bounds__(128) iir_chunk_kernel(const __grid_constant__ IirCoef F, const __grid_constant__ IirIo io,
                                                        double* __restrict__ S, const double* __restrict__ Z0) {
  const int64_t nchunks = (io.M + IIR_L - 1) / IIR_L;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nchunks * io.ncols) return;
  const int64_t k = tid / io.ncols;
  const int c = (int)(tid % io.ncols);
  double z[IIR_MAXN];
#pragma unroll
  for (int i = 0; i < IIR_MAXN; ++i) z[i] = APPLY ? Z0[tid * IIR_MAXN + i] : 0.0;
  const int64_t e0 = k * IIR_L, e1 = e0 + IIR_L < io.M ? e0 + IIR_L : io.M;
  for (int64_t e = e0; e < e1; ++e) {
    const double x = iir_input(io, e, c);
    double y;
    if (F.nsec > 0) {  // uniform: cascade of direct-form-II-transposed biquads, like scipy's sosfilt
      double u = x;
#pragma unroll
      for (int sct = 0; sct < IIR_MAXN / 2; ++sct) {
        if (sct < F.nsec) {
          const double ys = fma(F.sos[sct][0], u, z[2 * sct]);
          z[2 * sct] = fma(F.sos[sct][1], u, fma(-F.sos[sct][4], ys, z[2 * sct + 1]));
          z[2 * sct + 1] = fma(F.sos[sct][2], u, -F.sos[sct][5] * ys);
          u = ys;
        }
      }
      y = u;
    } else {
      y = fma(F.b[0], x, z[0]);
#pragma unroll
      for (int i = 0; i < IIR_MAXN; ++i) {
        if (i < F.n) {
          const double zn = (i + 1 < IIR_MAXN && i + 1 < F.n) ? z[i + 1] : 0.0;
          z[i] = fma(F.b[i + 1], x, fma(-F.a[i + 1], y, zn));
        }
      }
    }
    if (APPLY) {
      if (!io.backward) {
        io.y[e * io.ldy + c] = y;
      } else {
        // backward index e is row M - 1 - e of the extended signal = row M - 1 - e - padlen of the result
        const int64_t r = io.M - 1 - e - io.padlen - io.trim;
        if (r >= 0 && r < io.N - 2 * io.trim && r % io.step == 0) io.y[(r / io.step) * io.ldy + c] = y;
      }
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int i = 0; i < IIR_MAXN; ++i) S[tid * IIR_MAXN + i] = z[i];
  }
}

// One thread per column: true initial state of every chunk.  S in: zero-state final states; Z0 out.
__global__ void iir_propagate_kernel(const __grid_constant__ IirCoef F, const __grid_constant__ IirIo io,
                                     const double* __restrict__ S, double* __restrict__ Z0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= io.ncols) return;
  const int64_t nchunks = (io.M + IIR_L - 1) / IIR_L;
  const double x0 = iir_input(io, 0, c);  // scipy: zi * (first sample of the pass)
  double z[IIR_MAXN];
#pragma unroll
  for (int i = 0; i < IIR_MAXN; ++i) z[i] = i < F.n ? F.zi[i] * x0 : 0.0;
  for (int64_t k = 0; k < nchunks; ++k) {
    const int64_t o = (k * io.ncols + c) * IIR_MAXN;
    double zn[IIR_MAXN];
#pragma unroll
    for (int i = 0; i < IIR_MAXN; ++i) {
      Z0[o + i] = z[i];
      double acc = S[o + i];
#pragma unroll
      for (int j = 0; j < IIR_MAXN; ++j)
        if (i < F.n && j < F.n) acc = fma(F.AL[i * F.n + j], z[j], acc);
      zn[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < IIR_MAXN; ++i) z[i] = zn[i];
  }
}

// dq[i] = (q[i + 1] - q[i]) / dt_i (pin.difference for vector-space joints), i = 0 .. N - 2;
// ddq[:, j] = np.gradient(dq[:, j], edge_order=1) / dt for j < nq - 1, 0 for the last column;
// the reference then drops the last row of dq / ddq and the last two rows of q: N - 2 rows are written.
__global__ void differentiate_kernel(const double* __restrict__ q, int64_t N, int nq, int64_t ldq, double dt,
                                     const double* __restrict__ dtv, double* __restrict__ dq, double* __restrict__ ddq,
                                     int64_t ldo) {
  const int64_t rows = N - 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * nq;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / nq;
    const int j = (int)(idx % nq);
    auto d = [&](int64_t r) { return (q[(r + 1) * ldq + j] - q[r * ldq + j]) / (dtv ? dtv[r] : dt); };
    const double di = d(i);
    dq[i * ldo + j] = di;
    double g = 0.0;
    if (j < nq - 1) {
      // np.gradient over the N - 1 entries of dq: one-sided at i == 0 (i == N - 2 is dropped)
      g = i == 0 ? d(1) - di : 0.5 * (d(i + 1) - d(i - 1));
      g /= dtv ? dtv[i] : dt;
    }
    ddq[i * ldo + j] = g;
  }
}

// scipy.signal.medfilt(x[:, c], k) for every column c: median of the k-sample window centred on the row, zeros
// beyond both ends (examples/tiago/identification.py:70,80 of the reference: medfilt(., 5) ahead of filtfilt).
constexpr int MEDFILT_MAXK = 15;
__global__ void medfilt_kernel(const double* __restrict__ x, int64_t N, int ncols, int64_t ldx, int k,
                               double* __restrict__ y, int64_t ldy) {
  const int h = k / 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < N * ncols;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    double w[MEDFILT_MAXK];
    for (int t = 0; t < k; ++t) {
      const int64_t r = i - h + t;
      w[t] = (r >= 0 && r < N) ? x[r * ldx + c] : 0.0;
    }
    for (int a = 1; a < k; ++a) {  // insertion sort of a handful of values
      const double v = w[a];
      int b = a - 1;
      while (b >= 0 && w[b] > v) { w[b + 1] = w[b]; --b; }
      w[b + 1] = v;
    }
    y[i * ldy + c] = w[h];
  }
}

// np.gradient(x[:, c]) (unit spacing, edge_order 1) for every column, optionally divided by np.gradient(t)
// (examples/tiago/identification.py:91-98: acceleration = gradient(dq) / gradient(t)).
__global__ void gradient_kernel(const double* __restrict__ x, int64_t N, int ncols, int64_t ldx,
                                const double* __restrict__ t, double* __restrict__ y, int64_t ldy) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < N * ncols;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / ncols;
    const int c = (int)(idx % ncols);
    const int64_t lo = i == 0 ? 0 : i - 1, hi = i == N - 1 ? N - 1 : i + 1;
    const double den = (i == 0 || i == N - 1) ? 1.0 : 2.0;
    double g = (x[hi * ldx + c] - x[lo * ldx + c]) / den;
    if (t != nullptr) g /= (t[hi] - t[lo]) / den;
    y[i * ldy + c] = g;
  }
}

}  // namespace fg
