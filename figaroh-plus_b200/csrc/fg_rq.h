// fg_rq.h - host interface and launch parameters of the fifth-generation fused kernel (fg_ringqr.cuh):
// regressor PRODUCER warps and Householder-TSQR CONSUMER warps in one persistent kernel, connected by an
// L2-resident ring of regressor tiles.
//
// What the generation-4 profile showed (profiles/ncu_gq_r01_v4i.md): 39 % of the kernel is regressor
// generation, because every row-group stream re-runs the forward pass of pinocchio's
// computeJointTorqueRegressor (reference call sites regressor.py:49,93) for its own tile - 6 walks per
// 32 samples on UR10, 8 on TIAGo - and fetches its inputs with per-lane loads at a 48-byte stride.
//
// Generation 5 keeps the row-group streams (one independent TSQR per (warp, row group), fg_gq.h) but takes
// the generator out of them:
//   * a few PRODUCER warps per CTA walk the kinematic tree ONCE per 32-sample tile (one sample per lane),
//     carry the motion axes of all row groups down the tree, and write the entries of every row group
//     straight into one slot of a global ring in the consumers' shared-memory tile layout (every store is
//     one coalesced 256-byte row of a column);
//   * a CONSUMER warp waits for "its" tile, pulls its row group's blocks from the ring into its private
//     shared-memory tile with cp.async.bulk (TMA bulk copy, completion on an mbarrier) and folds them
//     into its private block-triangular factor exactly as generation 4 did.  The copy of the NEXT tile is
//     issued block by block as the panels of the current one retire, so it hides behind the fold.
// The ring is a few tens of MB (ring_slots x sum_g nb_g x 2 KB) and lives in L2: W still never goes to HBM.
// Slot hand-over uses two monotonic counters per slot in global memory (ready = tile index + 1 published by the
// producer with release semantics; done = number of consumer copies, bumped after the mbarrier reports the
// bytes have landed); all CTAs are co-resident (cooperative launch), every wait is bounded and raises an
// error flag instead of hanging.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "fg_gq.h"

namespace fg {

constexpr int RQ_MAXOPS = 2 * GQ_MAXPROC;  // descend / ascend steps of the tree walk
constexpr int RQ_MAXSLOTS = 4;             // saved (v, a) states / joint transforms of a side trip

// One step of the producer's depth-first walk.  DESCEND: move from the current body to child `t` (joint
// transform from q, velocity / acceleration propagation, axes of the live row groups transported into the
// child frame, entries of every live row group emitted).  ASCEND: go back up through the joint transform saved
// in slot `mslot` (axes transported back with the inverse action), optionally restoring the saved (v, a).
struct RqOp {
  unsigned char kind;      // 0 descend, 1 ascend
  unsigned char t;         // walk index of the joint (GqGen::pj)
  unsigned char vsave;     // descend: save the current (v, a) to this slot first (255 = no)
  unsigned char msave;     // descend: save the joint transform to this slot (255 = no)
  unsigned char vload;     // ascend: restore (v, a) from this slot (255 = no)
  unsigned char mslot;     // ascend: joint transform to go back through
  unsigned char from_universe;  // descend: the parent is the universe (v = 0, a = -gravity)
  unsigned char pad;
  unsigned live;           // row groups whose axis is live BEFORE this step and must be transported
  unsigned own;            // descend: row groups carried by this joint (their axis starts here)
};

// Input-fed columns that depend on ONE single-dof joint's (qd, qdd) are written when the walk is at that joint (the
// values are in registers there); entry = (column position inside the ring slot << 2) | ExtraSrc.
constexpr int RQ_STEPCOLS = 32;
struct RqStepTab {
  unsigned short col[GQ_MAXPROC][RQ_STEPCOLS];
  unsigned char n[GQ_MAXPROC];
};
constexpr int RQ_STEP_DOUBLES = (int)((sizeof(RqStepTab) + 15) / 16) * 2;

struct RqParams {
  GqParams gq;             // tables shared with the generation-4 merge kernel (groups, emit masks, simple columns)
  int nops;
  RqOp ops[RQ_MAXOPS];
  int goff[GQ_MAXG];       // offset (doubles) of row group g inside a ring slot
  int slot_doubles;        // doubles of one ring slot (sum_g nb_g * GQ_BLK)
  int ring;                // slots
  int nprod;               // producer warps of the whole launch
  int npc;                 // producer CTAs (a CTA is all producers or all consumers)
  int prod_smem_doubles;   // shared memory of one producer warp
  int nvslots, nmslots;
  RqStepTab step;
  int sleep_ns;            // back-off of a waiting warp between two polls of a ring flag
};

// Measured cost of one 32-sample tile in warp cycles (rq_accumulate_kernel<.., true> fills the counters): fold of row
// group g by a consumer warp, walk + emits by a producer warp.  rq_build apportions the warps by these instead of its
// instruction-count model when they are given.
constexpr int RQ_PROF_WORDS = 32;  // u64: [g] fold cycles, [8+g] tiles, [16+g] wait cycles, [24] producer cycles, [25] tiles, [26] wait
struct RqCost {
  bool have = false;
  double grp[GQ_MAXG] = {0};
  double prod = 0.0;
  int force_npcta = 0;   // > 0: this many producer SMs (the calibration's timing trials)
};

struct RqLaunch {
  bool ok = false;
  int nw = 0, sm_count = 0, smem = 0;
  int64_t r_doubles = 0;        // per-instance factors
  int64_t rfull_doubles = 0;
  int64_t ring_doubles = 0;     // ring slots
  int smem_merge = 0;
  std::vector<unsigned char> tab_g;   // [sm_count * nw] row group of warp gw, 255 = producer
  std::vector<unsigned short> tab_i;  // instance index / producer index
  const unsigned char* d_tab_g = nullptr;
  const unsigned short* d_tab_i = nullptr;
};

// walk program over the joints marked in `need` (see fg_rq.cu); order[t] = joint index of walk index t
struct RqWalk {
  bool ok = false;
  std::vector<RqOp> ops;
  std::vector<int> order;
  int nvslots = 0, nmslots = 0;
};
RqWalk rq_build_walk(int nj, const DevJoint* joints, const std::vector<char>& need, const std::vector<unsigned>& own,
                     const std::vector<unsigned>& live, int max_proc, int max_ops, int max_slots = RQ_MAXSLOTS);

RqLaunch rq_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const ColDesc* cols, int sm_count,
                  int max_smem, RqParams* out, const char** why, const RqCost* measured = nullptr);

// ring_ctl: 2 * ring + 4 unsigneds (ready[ring], done[ring], error flag, ...), zeroed by the launcher.
cudaError_t rq_launch(const RqLaunch& rl, const RqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_all, double* ring, unsigned* ring_ctl,
                      int first, cudaStream_t st, unsigned long long* prof = nullptr);

}  // namespace fg
