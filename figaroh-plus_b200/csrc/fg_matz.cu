// fg_matz.cu - host side of the materialising kernel (fg_matz.cuh): tables of the walk over the whole tree for all row
// blocks, launch.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fg_matz.cuh"
#include "fg_matzg.cuh"
#include "fg_matz.h"

namespace fg {

size_t matz_tables_bytes() { return sizeof(MatzTables); }

MatzPlan matz_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const double* mass, int ncols_full,
                    int sm_count, int max_smem) {
  MatzPlan mp;
  const int nj = P.nj, ms = P.ms, nb = nj - 1;
  if (ms < 1 || ms > 32 || nj > kMaxJoints) return mp;
  std::vector<char> need(nj, 1);
  need[0] = 0;
  std::vector<unsigned> own(nj, 0u), live(nj, 0u);
  for (int g = 0; g < ms; ++g) own[rows[g].joint] |= 1u << g;
  for (int i = 1; i < nj; ++i) live[i] = live[joints[i].parent] | own[i];
  const RqWalk walk = rq_build_walk(nj, joints, need, own, live, kMaxJoints, 2 * kMaxJoints, 24);
  if (!walk.ok || (int)walk.order.size() != nb) return mp;
  mp.tab.reset(new MatzTables);
  MatzTables& T = *mp.tab;
  memset(&T, 0, sizeof T);
  MatzParams& M = mp.prm;
  memset(&M, 0, sizeof M);
  M.nj = nj; M.nq = P.nq; M.nv = P.nv; M.ms = ms; M.nops = (int)walk.ops.size(); M.nproc = nb;
  M.mode = P.mode; M.flags = P.flags; M.nvslots = walk.nvslots; M.nmslots = walk.nmslots;
  M.coupling = ncols_full > 14 * nb ? 1 : 0;
  for (int k = 0; k < 3; ++k) M.gravity[k] = P.gravity[k];
  for (int t = 0; t < nb; ++t) { T.pj[t] = joints[walk.order[t]]; T.body_of_t[t] = walk.order[t]; }
  for (int o = 0; o < M.nops; ++o) T.ops[o] = walk.ops[o];
  for (int g = 0; g < ms; ++g) {
    for (int k = 0; k < 6; ++k) T.S[g][k] = rows[g].S[k];
    T.rowblock[g] = rows[g].rowblock;
    T.rowjoint[g] = rows[g].joint;
  }
  for (int b = 1; b < nj; ++b) {
    T.inertial[b] = !(P.mode == 1 /* external wrench */ && mass[b] == 0.0);
    int eg = -2;
    if (b - 1 < P.nv) {  // the reference indexes v[i, b] (regressor.py:145-158)
      if (P.mode == 1) eg = -1;
      else
        for (int g = 0; g < ms; ++g)
          if (rows[g].rowblock == b - 1) eg = g;
    }
    T.extra_grp[b] = eg;
  }
  // walk | patch stage | staged inputs of one tile | mbarrier.  Wide models (Talos: 39 + 2 x 38 inputs and eight saved
  // transforms per sample) keep their inputs in global memory: staging them would leave two warps per SM.
  const int tab_bytes = (int)((sizeof(MatzTables) + 15) / 16) * 16;
  int nw = 0;
  for (int stage_in = 1; stage_in >= 0; --stage_in) {
    M.stage_inputs = stage_in;
    M.smem_per_warp = (ms * 6 + 12 * (M.nvslots + M.nmslots)) * 32 + MATZ_STAGE + (MATZ_STAGE & 1) +
                      (stage_in ? 32 * (P.nq + 2 * P.nv) : 0) + 2;
    M.smem_per_warp += M.smem_per_warp & 1;
    nw = MATZ_WARPS;
    while (nw > 0 && tab_bytes + nw * M.smem_per_warp * 8 > max_smem) --nw;
    if (nw >= (stage_in ? 6 : 1)) break;
  }
  if (nw < 1) { mp.tab.reset(); return mp; }
  {  // serial chain of single-dof joints, one row block per joint, no coupling columns: the row-block-major kernel
    bool chain = P.mode == 0 && ms == nb && !M.coupling && !(getenv("FG_USE_MATZG") && getenv("FG_USE_MATZG")[0] == '0');
    for (int t = 0; t < nb && chain; ++t) {
      const DevJoint& jt = joints[t + 1];
      chain = walk.order[t] == t + 1 && jt.parent == t && jt.jtype >= 1 && jt.jtype <= 3 && rows[t].joint == t + 1 &&
              rows[t].rowblock == t && T.extra_grp[t + 1] == t;
    }
    if (chain) {
      MatzgParams& Gp = mp.gprm;
      memset(&Gp, 0, sizeof Gp);
      Gp.nb = nb; Gp.ms = ms; Gp.nq = P.nq; Gp.nv = P.nv; Gp.flags = P.flags; Gp.ncols = 14 * nb; Gp.pitch = 14 * nb;  // dense rows: the block leaves as ONE bulk copy
      for (int k = 0; k < 3; ++k) Gp.gravity[k] = P.gravity[k];
      Gp.smem_per_warp = nb * MZG_KIN * 32 + 32 * Gp.pitch;
      Gp.smem_per_warp += Gp.smem_per_warp & 1;
      int gw = MATZG_WARPS;
      while (gw > 0 && tab_bytes + gw * Gp.smem_per_warp * 8 > max_smem) --gw;
      if (gw >= 4) { mp.g_ok = true; mp.g_nw = gw; mp.g_smem = tab_bytes + gw * Gp.smem_per_warp * 8; }
    }
  }
  mp.nw = nw;
  mp.smem = tab_bytes + nw * M.smem_per_warp * 8;
  mp.sm_count = sm_count;
  mp.ok = true;
  return mp;
}

cudaError_t matz_launch(const MatzPlan& mp, const MatzTables* d_tab, const double* q, const double* v, const double* a,
                        int64_t N, double* W, int64_t ldw, cudaStream_t st) {
  if (mp.g_ok && ldw == mp.gprm.ncols && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    cudaError_t e = cudaFuncSetAttribute(matzg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mp.g_smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (N + 31) / 32;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ntiles + mp.g_nw - 1) / mp.g_nw, mp.sm_count));
    matzg_kernel<<<grid, mp.g_nw * 32, mp.g_smem, st>>>(mp.gprm, d_tab, q, v, a, N, W, ldw);
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(matz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mp.smem);
  if (e != cudaSuccess) return e;
  const int64_t ntiles = (N + 31) / 32;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ntiles + mp.nw - 1) / mp.nw, mp.sm_count));
  matz_kernel<<<grid, mp.nw * 32, mp.smem, st>>>(mp.prm, d_tab, q, v, a, N, W, ldw);
  return cudaGetLastError();
}

}  // namespace fg
