// fg_tq.h - host interface and launch parameters of the third-generation fused kernel
// (fg_tileqr.cuh): sample-per-thread regressor generation into a shared-memory tile whose rows
// are grouped by joint, folded into R by a structure-driven blocked Householder QR with fp64
// tensor-core (DMMA) trailing updates.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

constexpr int TQ_S = 32;         // samples (= rows per row block) of one tile
constexpr int TQ_WARPS = 4;      // warps per CTA
constexpr int TQ_MAXJ = 32;      // joints incl. universe
constexpr int TQ_MAXG = 8;       // row groups per sample (2 per warp)
constexpr int TQ_MAXCB = 32;     // 8-column blocks
constexpr int TQ_MAXP = 8 * TQ_MAXCB;
constexpr int TQ_MAXSIMPLE = 4 * TQ_MAXJ + 8;

// a column whose entries are read off the inputs (friction / inertia / offset / coupling / rhs)
struct TqSimple {
  int kind;   // COL_EXTRA / COL_COUPLING / COL_RHS
  int src;    // ExtraSrc, or the coupling column 0..2
  int vidx;   // velocity index of the source
  int grp;    // row group, or -1 = every row group
  int slot;   // (column block << 3) | column inside the block
};

// Everything the kernel needs to know about (model, configuration, column selection); passed by
// value as a __grid_constant__ kernel parameter, so all of it is read through the constant cache
// with warp-uniform addresses.
struct TqParams {
  int nj, nq, nv, ms, mode, flags;
  int ncb;             // column blocks: the first `npriv` are private to one row group
  int npriv;
  int p1;              // state columns incl. the rhs
  int nsimple;
  int nproc;           // joints walked by the forward pass
  int tile_doubles;    // shared-memory doubles of one tile
  int scratch_doubles;
  int sm_count;
  double gravity[3];
  DevJoint joints[TQ_MAXJ];
  double rowS[TQ_MAXG][6];          // motion axis of row group g in its joint's frame [lin; ang]
  int rowjoint[TQ_MAXG];
  int rowblock[TQ_MAXG];            // reference row-block index (add_coupling_TX40 addresses rows by it)
  int proc[TQ_MAXJ];
  int blkoff[TQ_MAXG][TQ_MAXCB];    // offset (doubles) of tile block (row group, column block), -1 = not stored
  unsigned gmask[TQ_MAXCB];         // row groups that take part in panel j
  unsigned trmask[TQ_MAXCB];        // column blocks updated by panel j
  short islot[TQ_MAXJ][10];         // (body, pinocchio parameter) -> slot, -1 = column not generated
  unsigned char anc[TQ_MAXJ];       // row groups whose joint supports body b
  unsigned char swz[8];             // tq_sw(c)
  unsigned char ncols[TQ_MAXCB];    // real columns of block j
  unsigned char sc_blk[TQ_MAXP];    // state column -> column block
  unsigned char sc_c[TQ_MAXP];      // state column -> column inside the block
  TqSimple simple[TQ_MAXSIMPLE];
};

struct TqLaunch {
  bool ok = false;
  int grid = 0, smem = 0, maxg = 0;
  int64_t r_doubles = 0;            // per-CTA R scratch (global memory, L2 resident)
  std::vector<int> perm;            // state column -> index into the plan's column list (size p1-1)
};

// Build the launch parameters from a host-side plan.  Returns ok = false (with `why` filled) when
// the model / configuration is outside what this kernel handles; the caller then uses an older path.
TqLaunch tq_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const ColDesc* cols, int sm_count,
                  int max_smem, TqParams* out, const char** why);

cudaError_t tq_launch(const TqLaunch& tl, const TqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_scratch, double* parts, int grid,
                      int phase, cudaStream_t st);

enum { TQ_PHASE_FIRST = 1, TQ_PHASE_LAST = 2 };

}  // namespace fg
