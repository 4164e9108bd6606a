// fg_coop.cuh - kernel (2) for WIDE dense problems: blocked Householder TSQR of a tall matrix that is already in
// HBM, one CTA per SM working on ONE 64-row tile at a time, the column blocks shared out among its warps.
//
// Where it is used: Talos with free-flyer in external-wrench mode (6 rows x 330 kept columns per sample,
// regressor.py:89-192 of the reference: every row depends on every body, there is no row-group sparsity to
// exploit), the drop-in double_QR / get_baseParams on a wide W_e (qrdecomposition.py:89-296: two np.linalg.qr of
// the tall matrix), and any model the row-group streams of fg_ringqr.cuh do not take.  W is materialised a chunk
// of samples at a time by kernel (1) (fg_matz.cuh) - 2 x 3.7 KB of HBM traffic per row against 2.6 Mflop of
// Householder work per row, i.e. the detour through HBM is 5 % of the FP64 time.
//
// What the column-at-a-time register-tile kernel of fg_wide.cuh showed (profiles/ncu_wide_talos_r02.md: 6.8 % of the
// FP64 roofline): one CTA-wide barrier per column, 331 of them per tile, and no tensor-core work.  Here (Talos:
// 40 % of the roofline, profiles/ncu_cq_talos_r02_final.md):
//   * the tile (64 rows x nb 8-column blocks, fg_tile.cuh layout) lives in shared memory, the CTA's
//     block-upper-triangular factor R in an L2-resident scratch (nb (nb + 1) / 2 blocks of 8 x 8);
//   * warp w OWNS the column blocks j = w (mod NW): nobody else ever touches block j of the tile or column j of R,
//     so there is no CTA barrier inside a tile - only two mbarriers per panel, "V_k is published" and "T_k is published";
//   * panel k (8 columns, two rows per lane) is factored by its owner with the row-group streams' Householder panel
//     (gq_panel_factor: LAPACK dlarfg reflectors, compact-WY T from a DMMA); the owner leaves V_k in place of the
//     block and publishes it, then T_k in a per-panel slot (the others start Y = R_kj + V^T B_j while T_k is computed);
//   * every warp then applies panel k to its own blocks j > k with DMMA (Y = R_kj + V^T B_j, W = T^T Y,
//     R_kj -= W, B_j -= V W: gq_fold_h's trailing update).  The owner of block k + 1 does that block FIRST and
//     factors panel k + 1 at once (look-ahead), so the serial chain of a tile is nb x (one update + one panel)
//     while the other nb (nb - 1) / 2 updates fill the FP64 pipe around it; the R blocks a warp needs after a wait
//     are requested from L2 before it;
//   * merge mode: the same kernel folds one per-CTA factor into another, 64 of its rows per tile, the panels
//     starting at the tile's first non-zero column (pairwise tree over the factors, multi-GPU exchange, small stage).
#pragma once
#define FG_GQ_DEVICE_ONLY
#include "fg_groupqr.cuh"
#include "fg_async.cuh"

namespace fg {

constexpr int CQ_NW = 8;          // warps of the CTA
// RPL rows per lane: 64-row tiles (RPL = 2) up to 49 column blocks, 32-row tiles (RPL = 1) up to 88

struct CqParams {
  const double* W;        // rows x ldw, row-major
  int64_t ldw;
  const int* colsel;      // p column indices into W, or nullptr (identity)
  const double* tau;      // rhs: one entry per row, or nullptr
  int64_t rows;
  int p1;                 // columns incl. the rhs
  int nb;                 // 8-column blocks
  double* r_all;          // grid x nb (nb + 1) / 2 x 64 doubles
  int64_t tau_blk, tau_ld;  // rhs of row r at tau[(r / tau_blk) * tau_ld + r % tau_blk] (tau_blk = 0: tau[r])
  int first;              // first launch of a sequence zeroes the factors
  // merge mode (merge_stride > 0): CTA b folds factor (2b + 1) * merge_stride into factor 2b * merge_stride; W / tau unused
  int merge_stride, nfac;
  int rhs_in_w;           // all p1 columns come from W (the small stage: the rhs is the last column of the matrix)
};

// "panel k is published": one mbarrier per panel and stage (V_k, T_k), arrival count 32: EVERY lane of the owner warp
// arrives (release of its own shared-memory writes), phase = parity of the CTA's tile counter.  Waiting warps sit in
// mbarrier.try_wait (the hardware suspends them) instead of polling a flag in shared memory (first version: the polls
// were 42 % of all executed instructions, profiles/ncu_cq_talos_r02_v1.md).  The owner also observes the completed
// phase before it goes on (one try_wait that succeeds as soon as its 32 lanes have arrived): every thread that
// arrives also waits, which is the pattern compute-sanitizer's racecheck / synccheck can follow.
__device__ __forceinline__ void cq_wait(uint64_t* bar, uint32_t parity) {
  while (!rq_mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void cq_publish(uint64_t* bar, uint32_t parity) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rq_smem_u32(bar)) : "memory");
  while (!rq_mbar_try_wait(bar, parity)) {}
}

// reflector fragments of panel k (kept in registers for the updates that follow)
template <int RPL>
struct CqFrag {
  double vc0[RPL][4], vc1[RPL][4], vb0[RPL][4], vb1[RPL][4];
  double t0, t1;
};

template <int RPL>
__device__ __forceinline__ void cq_load_frag(CqFrag<RPL>& f, const double* tile, int nb, int k, const GqOffs& o) {
  const double* Xk = tile + k * GQ_BLK;
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      f.vc0[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc0[nn]];
      f.vc1[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc1[nn]];
      f.vb0[sl][nn] = Xk[sl * nb * GQ_BLK + o.ob0[nn]];
      f.vb1[sl][nn] = Xk[sl * nb * GQ_BLK + o.ob1[nn]];
    }
}
template <int RPL>
__device__ __forceinline__ void cq_load_T(CqFrag<RPL>& f, int k, const double* Tb, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
  const double* TS = Tb + k * 64;
  f.t0 = 2 * rq <= cq ? TS[(2 * rq) * 8 + cq] : 0.0;
  f.t1 = 2 * rq + 1 <= cq ? TS[(2 * rq + 1) * 8 + cq] : 0.0;
}

// panel k applied to block j of the tile and to R_kj (the trailing update of gq_fold_h)
// WT: T_k is published after V_k; the first update with a panel waits for it only once Y is under way
template <int RPL, bool WT>
__device__ __forceinline__ void cq_update(CqFrag<RPL>& f, double* tile, int nb, int j, double* Rb, double r0, double r1,
                                          const GqOffs& o, int lane, uint64_t* barT = nullptr, uint32_t parity = 0,
                                          int k = 0, const double* Tb = nullptr) {
  const int cq = lane >> 2, rq = lane & 3;
  double* Bj = tile + j * GQ_BLK;
  double c0[RPL][4], c1[RPL][4];
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      c0[sl][nn] = Bj[sl * nb * GQ_BLK + o.oc0[nn]];
      c1[sl][nn] = Bj[sl * nb * GQ_BLK + o.oc1[nn]];
    }
  double y0 = 0.0, y1 = 0.0, z0 = 0.0, z1 = 0.0;
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      tq_dmma(y0, y1, c0[sl][nn], f.vc0[sl][nn]);
      tq_dmma(z0, z1, c1[sl][nn], f.vc1[sl][nn]);
    }
  y0 += z0 + r0;
  y1 += z1 + r1;
  if (WT) {
    cq_wait(barT, parity);
    cq_load_T(f, k, Tb, lane);
  }
  double w0 = 0.0, w1 = 0.0;
  tq_dmma(w0, w1, y0, f.t0);
  tq_dmma(w0, w1, y1, f.t1);
  Rb[(2 * rq) * 8 + cq] = r0 - w0;
  Rb[(2 * rq + 1) * 8 + cq] = r1 - w1;
  const double nw0 = -w0, nw1 = -w1;
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      tq_dmma(c0[sl][nn], c1[sl][nn], nw0, f.vb0[sl][nn]);
      tq_dmma(c0[sl][nn], c1[sl][nn], nw1, f.vb1[sl][nn]);
      Bj[sl * nb * GQ_BLK + o.oc0[nn]] = c0[sl][nn];
      Bj[sl * nb * GQ_BLK + o.oc1[nn]] = c1[sl][nn];
    }
}

// panel k: Householder factorisation of [R_kk; block k of the tile]; V_k stays in the block, T_k goes to Tb[k]
template <int RPL>
// d0, d1: R_kk as loaded by the caller (entries lane, lane + 32), requested before it waited for panel k - 1
__device__ __forceinline__ void cq_panel(double* tile, int nb, int k, double* R, double d0, double d1, double* Tb, double* scr,
                                         const GqOffs& o, int lane, uint64_t* barV, uint64_t* barT, uint32_t parity) {
  double* Rd = R + (size_t)tq_rblk(nb, k, k) * 64;
  double* Xk = tile + k * GQ_BLK;
  double* Rin = scr;
  double* sc = scr + 64;
  Rin[lane] = d0;
  Rin[lane + 32] = d1;
  double x[RPL][8];
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int c = 0; c < 8; ++c) x[sl][c] = Xk[sl * nb * GQ_BLK + tq_idx(c, lane)];
  __syncwarp();
  gq_panel_factor<RPL>(x, Rin, Rin, sc, Xk, lane);
  if ((lane >> 3) <= (lane & 7)) Rd[lane] = Rin[lane];
  if ((lane >> 3) + 4 <= (lane & 7)) Rd[lane + 32] = Rin[lane + 32];
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int c = 0; c < 8; ++c) Xk[sl * nb * GQ_BLK + tq_idx(c, lane)] = x[sl][c];
  cq_publish(barV, parity);  // the reflectors are out: the others start on Y = R + V^T B (all lanes pass the wait: warp-wide order)
  double vc0[RPL][4], vc1[RPL][4];
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      vc0[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc0[nn]];
      vc1[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc1[nn]];
    }
  gq_panel_T<RPL>(vc0, vc1, sc, Tb + k * 64, lane);
  cq_publish(barT, parity);
}

template <int RPL>
__global__ void __launch_bounds__(CQ_NW * 32, 1)
cq_accumulate_dense_kernel(const __grid_constant__ CqParams P) {
  extern __shared__ __align__(16) double cq_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = P.nb, p1 = P.p1, p = P.rhs_in_w ? p1 : p1 - 1;  // the last column is the rhs (zero without tau)
  double* tile = cq_smem;                               // RPL x nb blocks
  double* Tb = tile + (size_t)RPL * nb * GQ_BLK;     // nb x 64
  double* scr = Tb + (size_t)nb * 64 + (size_t)warp * GQ_SCRATCH;
  uint64_t* barsV = reinterpret_cast<uint64_t*>(Tb + (size_t)nb * 64 + (size_t)CQ_NW * GQ_SCRATCH);  // nb: V_k published
  uint64_t* barsT = barsV + nb;                                                                        // nb: T_k published
  const size_t rsize = (size_t)nb * (nb + 1) / 2 * 64;
  double* R = P.r_all + (size_t)blockIdx.x * rsize;
  if (P.first && P.merge_stride == 0)
    for (size_t i = threadIdx.x; i < rsize; i += blockDim.x) R[i] = 0.0;
  for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) rq_mbar_init(barsV + i, 32);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  GqOffs offs;
  gq_make_offs(offs, lane);
  constexpr int CQ_TR = 32 * RPL;
  const bool merge = P.merge_stride > 0;
  const double* Rs = nullptr;  // merge mode: the factor that is folded into R, 64 of its rows per tile
  if (merge) {
    const int dst = 2 * (int)blockIdx.x * P.merge_stride, src = dst + P.merge_stride;
    if (src >= P.nfac) return;
    R = P.r_all + (size_t)dst * rsize;
    Rs = P.r_all + (size_t)src * rsize;
  }
  const int64_t ntiles = merge ? (p1 + CQ_TR - 1) / CQ_TR : (P.rows + CQ_TR - 1) / CQ_TR;
  uint32_t parity = 0;
  for (int64_t tl = merge ? 0 : blockIdx.x; tl < ntiles; tl += merge ? 1 : gridDim.x, parity ^= 1u) {
    // rows 64 t .. of an upper-triangular factor are zero left of column 64 t: the panels start there
    const int k0 = merge ? (int)(tl * CQ_TR / 8) : 0;
    __syncthreads();  // the previous tile is finished with by every warp (and R is zeroed before the first one)
    // ---- the tile: rows row0 .. row0 + 32 RPL - 1 of [W(:, colsel) | tau], zero padded.  One warp instruction moves one
    // column block x 4 rows (lane = (column, row mod 4): 64-byte row segments from HBM, conflict-free stores); eight
    // of them are in flight per warp.
    const int64_t row0 = tl * CQ_TR;
    {
      const int cb = lane >> 2, rr = lane & 3;
      for (int b = warp; b < nb; b += CQ_NW) {
        if (merge) {
          if (b < k0) continue;
          double* dstb = tile + (size_t)b * GQ_BLK;
#pragma unroll 4
          for (int u = 0; u < CQ_TR / 4; ++u) {
            const int r = 4 * u + rr;
            const int sr = (int)row0 + r, br = sr >> 3;
            double val = 0.0;
            if (sr < p1 && b >= br) val = __ldcg(Rs + (size_t)tq_rblk(nb, br, b) * 64 + (sr & 7) * 8 + cb);
            dstb[(size_t)(r >> 5) * nb * GQ_BLK + tq_idx(cb, r & 31)] = val;
          }
          continue;
        }
        const int c = b * 8 + cb;
        const double* src = nullptr;
        int64_t stride = 0;
        if (c < p) { src = P.W + (P.colsel ? P.colsel[c] : c); stride = P.ldw; }
        double* dstb = tile + (size_t)b * GQ_BLK;
#pragma unroll
        for (int r0 = 0; r0 < CQ_TR; r0 += 32) {
          double val[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int64_t r = row0 + r0 + 4 * u + rr;
            val[u] = 0.0;
            if (r < P.rows) {
              if (src != nullptr) val[u] = __ldg(src + r * stride);
              else if (c == p && P.tau != nullptr) val[u] = __ldg(P.tau + (P.tau_blk ? (r / P.tau_blk) * P.tau_ld + r % P.tau_blk : r));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) dstb[(size_t)(r0 >> 5) * nb * GQ_BLK + tq_idx(cb, 4 * u + rr)] = val[u];
        }
      }
      // next tile of this CTA towards L2 while this one is folded
      const int64_t nrow0 = (tl + gridDim.x) * CQ_TR;
      if (!merge && nrow0 < P.rows) {
        const int64_t row_bytes = (int64_t)P.ldw * 8;
        const int lines = (int)((row_bytes + 127) / 128) + 1;
        for (int e = threadIdx.x; e < CQ_TR * lines; e += blockDim.x) {
          const int r = e / lines, l = e - r * lines;
          if (nrow0 + r < P.rows) {
            const char* ptr = reinterpret_cast<const char*>(P.W + (nrow0 + r) * P.ldw) + (int64_t)l * 128;
            if (l * 128 < row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
          }
        }
      }
    }
    __syncwarp();  // every warp loaded its OWN blocks: nobody waits for anybody here
    const int cq = lane >> 2, rq = lane & 3;
    if (warp == k0 % CQ_NW) {
      const double* Rd = R + (size_t)tq_rblk(nb, k0, k0) * 64;
      cq_panel<RPL>(tile, nb, k0, R, __ldcg(Rd + lane), __ldcg(Rd + lane + 32), Tb, scr, offs, lane, barsV + k0, barsT + k0, parity);
    }
    for (int k = k0; k + 1 < nb; ++k) {
      int j = k + 1 + ((warp - (k + 1)) % CQ_NW + CQ_NW) % CQ_NW;  // my first block to the right of k
      if (j >= nb) continue;
      // what the first update needs from L2 is requested BEFORE the wait (R_kj; for the owner of k + 1 also R_jj)
      double* Rb = R + (size_t)tq_rblk(nb, k, j) * 64;
      double r0 = __ldcg(Rb + (2 * rq) * 8 + cq), r1 = __ldcg(Rb + (2 * rq + 1) * 8 + cq);
      double d0 = 0.0, d1 = 0.0;
      if (j == k + 1) {
        const double* Rd = R + (size_t)tq_rblk(nb, j, j) * 64;
        d0 = __ldcg(Rd + lane);
        d1 = __ldcg(Rd + lane + 32);
      }
      cq_wait(barsV + k, parity);
      CqFrag<RPL> f;
      cq_load_frag(f, tile, nb, k, offs);
      if (j == k + 1) {  // look-ahead: this block first, then its panel, then the rest of panel k's updates
        cq_update<RPL, true>(f, tile, nb, j, Rb, r0, r1, offs, lane, barsT + k, parity, k, Tb);
        __syncwarp();
        cq_panel<RPL>(tile, nb, j, R, d0, d1, Tb, scr, offs, lane, barsV + j, barsT + j, parity);
        j += CQ_NW;
        if (j >= nb) continue;
        cq_load_frag(f, tile, nb, k, offs);
        cq_load_T(f, k, Tb, lane);
        Rb = R + (size_t)tq_rblk(nb, k, j) * 64;
        r0 = __ldcg(Rb + (2 * rq) * 8 + cq);
        r1 = __ldcg(Rb + (2 * rq + 1) * 8 + cq);
      } else {
        double* Rn = Rb + (size_t)CQ_NW * 64;
        double nr0 = 0.0, nr1 = 0.0;
        if (j + CQ_NW < nb) { nr0 = __ldcg(Rn + (2 * rq) * 8 + cq); nr1 = __ldcg(Rn + (2 * rq + 1) * 8 + cq); }
        cq_update<RPL, true>(f, tile, nb, j, Rb, r0, r1, offs, lane, barsT + k, parity, k, Tb);
        Rb = Rn; r0 = nr0; r1 = nr1;
        j += CQ_NW;
      }
      for (; j < nb; j += CQ_NW) {
        double nr0 = 0.0, nr1 = 0.0;
        double* Rn = Rb + (size_t)CQ_NW * 64;  // blocks of one block row are consecutive
        if (j + CQ_NW < nb) { nr0 = __ldcg(Rn + (2 * rq) * 8 + cq); nr1 = __ldcg(Rn + (2 * rq + 1) * 8 + cq); }
        cq_update<RPL, false>(f, tile, nb, j, Rb, r0, r1, offs, lane);
        Rb = Rn; r0 = nr0; r1 = nr1;
      }
    }
  }
}

// factor 0 of r_all as a packed part: upper triangle + squared column norms (the format of merge_pairs / fg_state)
__global__ void cq_export_kernel(const double* __restrict__ R, int p1, int nb, double* __restrict__ out) {
  const int ntri = p1 * (p1 + 1) / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p1 * p1; idx += gridDim.x * blockDim.x) {
    const int r = idx / p1, c = idx - r * p1;
    if (c >= r) out[tri_off(r, p1) + c - r] = R[(size_t)tq_rblk(nb, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7)];
  }
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < p1; c += gridDim.x * blockDim.x) {
    double sacc = 0.0;
    for (int r = 0; r <= c; ++r) {
      const double x = R[(size_t)tq_rblk(nb, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7)];
      sacc = fma(x, x, sacc);
    }
    out[ntri + c] = sacc;
  }
}


// n dense upper-triangular factors (fg_state_export layout: p1 x p1 row-major R | p1 norms | rows | p1 - 1 perm entries)
// into block layout (r_all zeroed by the caller); the row counts are summed into rows_acc
__global__ void cq_import_dense_kernel(const double* __restrict__ packed, int n, int p1, int nb, double* __restrict__ r_all,
                                       double* __restrict__ rows_acc) {
  const int64_t pk = (int64_t)p1 * p1 + 2 * (int64_t)p1;
  const size_t rsize = (size_t)nb * (nb + 1) / 2 * 64;
  for (int k = blockIdx.y; k < n; k += gridDim.y) {
    const double* src = packed + (size_t)k * pk;
    double* R = r_all + (size_t)k * rsize;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p1 * p1; idx += gridDim.x * blockDim.x) {
      const int r = idx / p1, c = idx - r * p1;
      if (c >= r) R[(size_t)tq_rblk(nb, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7)] = src[idx];
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += packed[(size_t)k * pk + (size_t)p1 * p1 + p1];
    *rows_acc += acc;
  }
}

// rows of a (narrower) block-triangular factor as dense rows of width p1: D[r][colmap[c]] = R(r, c); D is zeroed by
// the caller.  Used to fold the factor of a row-block group with fewer columns into the full-width factor.
__global__ void cq_scatter_rows_kernel(const double* __restrict__ R, int p1g, int nbg, const int* __restrict__ colmap, int p1,
                                       double* __restrict__ D) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p1g * p1g; idx += gridDim.x * blockDim.x) {
    const int r = idx / p1g, c = idx - r * p1g;
    if (c >= r) D[(size_t)r * p1 + colmap[c]] = R[(size_t)tq_rblk(nbg, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7)];
  }
}

}  // namespace fg
