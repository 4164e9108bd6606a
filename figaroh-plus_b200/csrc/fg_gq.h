// fg_gq.h - tables of the row-group TSQR streams (fg_groupqr.cuh: one independent Householder TSQR stream per
// (warp, row group)) and the host interface of the merge of the per-stream factors.
//
// The regressor of row group g (the rows S_g^T B of joint g, regressor.py:49-70 of the reference) is
// non-zero only in the columns of the bodies that joint supports plus the joint's own Ia/fv/fs/off
// columns.  Rows of different row groups are different rows of W, so
//     W^T W = sum_g W_g^T W_g      and      R = qr([R_0; R_1; ...; R_{ms-1}])
// where R_g is the factor of the w_g non-zero columns of row group g alone.  Every R_g is an
// ordinary dense TSQR with no fill-in from the other row groups: 2 sum_g w_g^2 flop per sample
// (UR10: 19.6 k) instead of 2 m_s p^2 (65.7 k), and m_s independent streams instead of one.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

constexpr int GQ_S = 32;         // samples (= rows) of one tile: one row per lane
constexpr int GQ_MAXJ = 32;      // joints incl. universe
constexpr int GQ_MAXG = 8;       // row groups per sample
constexpr int GQ_MAXB = 16;      // 8-column blocks of one row group's local problem
constexpr int GQ_MAXBF = 24;     // 8-column blocks of the full factor (state columns incl. the rhs <= 192)
constexpr int GQ_MAXW = 8 * GQ_MAXB;
constexpr int GQ_MAXSIMPLE = 16; // input-fed columns (friction / inertia / offset / coupling / rhs) per row group
constexpr int GQ_BLK = 8 * GQ_S; // doubles of one tile block (8 columns x 32 rows)
constexpr int GQ_SCRATCH = 64 + 16;  // T / diagonal block (shared), taus + reduction results
constexpr int GQ_MAXPROC = 16;   // joints walked by the forward pass
constexpr int GQ_MAXNW = 16;     // warps (independent streams) of one CTA; one CTA per SM

struct GqSimple {
  short kind;   // COL_EXTRA / COL_COUPLING / COL_RHS
  short src;    // ExtraSrc, or the coupling column 0..2
  short vidx;   // velocity index of the source
  short slot;   // (local block << 3) | column inside the block
};

struct GqGroup {
  int nb;        // local 8-column blocks (the rhs is the last real column)
  int w;         // real local columns incl. the rhs
  int r_off;     // offset (doubles) of this group's factors in the scratch: instance i at r_off + i * r_size
  int r_size;    // doubles of one block-triangular factor, nb (nb + 1) / 2 * 64
  int ninst;     // warps (instances) that work on this row group in a full launch
  int joint;     // joint whose motion axis the rows project on
  int tjoint;    // its position in the forward walk
  int rowblock;  // reference row-block index (add_coupling_TX40 addresses rows by it)
  int nsimple;
  double S[6];   // motion axis in the joint frame [lin; ang]
};

// What the regressor generator reads per joint / per column: copied to shared memory once per CTA (indexed
// constant-bank loads miss the small constant L1 when 14 warps walk different parts of it).
struct GqGen {
  DevJoint pj[GQ_MAXPROC];             // joints in walk order
  int emit[GQ_MAXG][GQ_MAXPROC];       // (first local slot << 10) | mask of generated pinocchio parameters; 0 = none
  GqGroup grp[GQ_MAXG];
  GqSimple simple[GQ_MAXG][GQ_MAXSIMPLE];
};

constexpr int GQ_GEN_DOUBLES = (int)((sizeof(GqGen) + 15) / 16) * 2;  // shared-memory doubles reserved for it

// Passed by value as a __grid_constant__ kernel parameter: read through the constant cache with
// warp-uniform addresses.
struct GqParams {
  int nj, nq, nv, ms;
  int nproc;            // joints walked by the forward pass
  int p1;               // state columns incl. the rhs
  int nbf;              // 8-column blocks of the full factor
  int maxnb;            // largest local block count
  int maxinst;          // largest instance count of a row group
  double gravity[3];
  GqGen gen;
  unsigned char gcol[GQ_MAXG][GQ_MAXW]; // local column -> state column
};

// tree-merge the per-instance factors of every row group (instances that received no tile are skipped when
// ntiles_hint > 0), fold the roots into one full factor and write it as a packed part (upper triangle +
// squared column norms) to part_out.  counters: GQ_MAXG * maxinst + 2 unsigneds; smem_merge: dynamic shared memory
// of a merge warp ((2 max(maxnb, nbf) GQ_BLK + GQ_SCRATCH) doubles).
cudaError_t gq_merge(int smem_merge, const GqParams& prm, double* r_all, int64_t ntiles_hint, unsigned* counters,
                     double* rfull, double* part_out, cudaStream_t st);

}  // namespace fg
