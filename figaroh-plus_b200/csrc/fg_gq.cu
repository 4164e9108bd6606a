// fg_gq.cu - host side of the fourth-generation fused kernel (fg_groupqr.cuh): the symbolic pass that
// splits (kinematic tree, row groups, column selection) into one local dense problem per row group,
// the assignment of row groups to warp roles, and the launchers.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fg_groupqr.cuh"

namespace fg {

static void gq_fail(const char** why, const char* msg) {
  if (why) *why = msg;
}

GqLaunch gq_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const ColDesc* cols, int sm_count,
                  int max_smem, GqParams* out, const char** why) {
  GqLaunch gl;
  if (why) *why = "";
  const int ms = P.ms, nj = P.nj, ncols = P.ncols;
  if (ms < 1 || ms > GQ_MAXG) { gq_fail(why, "more than 8 row groups"); return gl; }
  if (nj > GQ_MAXJ) { gq_fail(why, "more than 32 joints"); return gl; }
  if (ncols + 1 > 8 * GQ_MAXBF) { gq_fail(why, "too many columns"); return gl; }
  GqParams& T = *out;
  memset(&T, 0, sizeof T);
  T.nj = nj; T.nq = P.nq; T.nv = P.nv; T.ms = ms; T.p1 = ncols + 1;
  T.nbf = (ncols + 1 + 7) / 8;
  for (int k = 0; k < 3; ++k) T.gravity[k] = P.gravity[k];
  const unsigned allg = (1u << ms) - 1u;
  // ---- row-group mask of every selected column
  std::vector<unsigned> cmask(ncols, 0u);
  for (int c = 0; c < ncols; ++c) {
    const ColDesc& cd = cols[c];
    if (cd.kind == COL_INERTIAL) {
      cmask[c] = (unsigned)(col_ancmask(cd) & allg);
    } else if (cd.kind == COL_EXTRA) {
      if (cd.grp >= 0) cmask[c] = (1u << cd.grp) & allg;
      else if (cd.grp == -1) cmask[c] = allg;
    } else if (cd.kind == COL_COUPLING) {
      for (int g = 0; g < ms; ++g)
        if (rows[g].rowblock == 4 || rows[g].rowblock == 5) cmask[c] |= 1u << g;
    }
  }
  // ---- forward pass: the joints that support a selected inertial column or carry a row group; they must
  //      form chains hanging off the universe when walked in index order
  std::vector<char> need(nj, 0);
  for (int c = 0; c < ncols; ++c)
    if (cols[c].kind == COL_INERTIAL && cmask[c]) {
      if (cols[c].body < 1 || cols[c].body >= nj || cols[c].k < 0 || cols[c].k > 9) { gq_fail(why, "internal: bad inertial column"); return gl; }
      for (int i = cols[c].body; i > 0; i = joints[i].parent) need[i] = 1;
    }
  for (int g = 0; g < ms; ++g)
    for (int i = rows[g].joint; i > 0; i = joints[i].parent) need[i] = 1;
  std::vector<int> tof(nj, -1);
  int nproc = 0, prev = 0;
  for (int i = 1; i < nj; ++i) {
    if (!need[i]) continue;
    if (joints[i].parent != 0 && joints[i].parent != prev) { gq_fail(why, "kinematic tree branches below the base"); return gl; }
    if (nproc >= GQ_MAXPROC) { gq_fail(why, "more than 12 joints to walk"); return gl; }
    T.gen.pj[nproc] = joints[i];
    tof[i] = nproc++;
    prev = i;
  }
  if (nproc < 1) { gq_fail(why, "no joint to walk"); return gl; }
  T.nproc = nproc;
  // ---- the local problem of every row group: per walked body its selected inertial columns in pinocchio
  //      parameter order, then the input-fed columns in the caller's order, then the rhs
  int maxnb = 0;
  for (int g = 0; g < ms; ++g) {
    GqGroup& G = T.gen.grp[g];
    for (int k = 0; k < 6; ++k) G.S[k] = rows[g].S[k];
    G.joint = rows[g].joint;
    G.tjoint = tof[rows[g].joint];
    G.rowblock = rows[g].rowblock;
    int l = 0, ns = 0;
    for (int t = 0; t < nproc; ++t) {
      int colof[10];
      for (int k = 0; k < 10; ++k) colof[k] = -1;
      for (int c = 0; c < ncols; ++c)
        if (cols[c].kind == COL_INERTIAL && tof[cols[c].body] == t && ((cmask[c] >> g) & 1u)) colof[cols[c].k] = c;
      int mask = 0;
      const int base = l;
      for (int k = 0; k < 10; ++k)
        if (colof[k] >= 0) {
          if (l >= GQ_MAXW) { gq_fail(why, "a row group has too many columns"); return gl; }
          mask |= 1 << k;
          T.gcol[g][l++] = (unsigned char)colof[k];
        }
      if (mask) {
        // the row group's joint must support the body
        int i = 0;
        for (int b = 1; b < nj; ++b) if (tof[b] == t) i = b;
        while (i > rows[g].joint) i = joints[i].parent;
        if (i != rows[g].joint) { gq_fail(why, "internal: column outside its row group's subtree"); return gl; }
      }
      T.gen.emit[g][t] = mask ? ((base << 10) | mask) : 0;
    }
    for (int c = 0; c <= ncols; ++c) {
      if (c < ncols && (cols[c].kind == COL_INERTIAL || !((cmask[c] >> g) & 1u))) continue;
      if (l >= GQ_MAXW || ns >= GQ_MAXSIMPLE) { gq_fail(why, "a row group has too many columns"); return gl; }
      GqSimple& s = T.gen.simple[g][ns++];
      if (c == ncols) { s.kind = COL_RHS; s.src = 0; s.vidx = 0; }
      else { s.kind = (short)cols[c].kind; s.src = (short)cols[c].k; s.vidx = (short)cols[c].vidx; }
      s.slot = (short)l;
      T.gcol[g][l++] = (unsigned char)c;
    }
    G.w = l;
    G.nb = (l + 7) / 8;
    G.nsimple = ns;
    G.r_size = G.nb * (G.nb + 1) / 2 * 64;
    maxnb = std::max(maxnb, G.nb);
  }
  if (maxnb > GQ_MAXB || T.nbf > GQ_MAXBF) { gq_fail(why, "too many column blocks"); return gl; }
  T.maxnb = maxnb;
  gl.smem_merge = (2 * std::max(maxnb, T.nbf) * GQ_BLK + GQ_SCRATCH) * 8;  // 64-row slabs
  if (gl.smem_merge > max_smem) { gq_fail(why, "tile does not fit shared memory"); return gl; }
  // ---- warps per row group, proportional to the instruction cost of one tile of that group (forward pass,
  //      one emit per supported body, nb panels + T factors, nb (nb - 1) / 2 trailing blocks; calibrated on
  //      the per-line instruction counts of profiles/ncu_gq_r01_v4c.md)
  std::vector<double> cost(ms);
  double csum = 0.0;
  for (int g = 0; g < ms; ++g) {
    const GqGroup& G = T.gen.grp[g];
    int emits = 0;
    for (int t = 0; t < nproc; ++t) emits += T.gen.emit[g][t] != 0;
#ifndef GQ_COST_JOINT
#define GQ_COST_JOINT 300.0  // A/B on the GPU: 250-320 are 1 % faster than 400, 500 is 1 % slower
#endif
    cost[g] = 700.0 + GQ_COST_JOINT * nproc + 233.0 * emits + 20.0 * G.nsimple + 1070.0 * G.nb + 100.0 * G.nb * (G.nb - 1) / 2;
    csum += cost[g];
  }
  gl.sm_count = std::max(1, sm_count);
  for (int nw = GQ_MAXNW; nw >= 1; --nw) {
    const int W = gl.sm_count * nw;
    if (W < ms) continue;
    // largest-remainder apportionment, at least one warp per row group
    std::vector<int> cnt(ms);
    int left = W;
    std::vector<std::pair<double, int>> rem;
    for (int g = 0; g < ms; ++g) {
      const double t = W * cost[g] / csum;
      cnt[g] = std::max(1, (int)t);
      left -= cnt[g];
      rem.push_back({t - (int)t, g});
    }
    std::sort(rem.begin(), rem.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
      return a.first != b.first ? a.first > b.first : a.second < b.second;
    });
    for (int k = 0; left > 0; k = (k + 1) % ms, --left) ++cnt[rem[k].second];
    for (int k = 0; left < 0; k = (k + 1) % ms)
      if (cnt[rem[ms - 1 - k].second] > 1) { --cnt[rem[ms - 1 - k].second]; ++left; }
    // interleave: position gw goes to the row group that is furthest behind its proportional schedule
    gl.tab_g.assign(W, 0);
    gl.tab_i.assign(W, 0);
    std::vector<int> given(ms, 0);
    int worst = 0;
    for (int b = 0; b < gl.sm_count; ++b) {
      int blocks = 0;
      for (int w = 0; w < nw; ++w) {
        const int gw = b * nw + w;
        int best = -1;
        double lag = -1e300;
        for (int g = 0; g < ms; ++g) {
          if (given[g] >= cnt[g]) continue;
          const double l = (double)cnt[g] * (gw + 1) / W - given[g];
          if (l > lag) { lag = l; best = g; }
        }
        gl.tab_g[gw] = (unsigned char)best;
        gl.tab_i[gw] = (unsigned short)given[best]++;
        blocks += T.gen.grp[best].nb;
      }
      worst = std::max(worst, blocks);
    }
    const int smem = (GQ_GEN_DOUBLES + worst * GQ_BLK + nw * GQ_SCRATCH) * 8;
    if (smem > max_smem) continue;
    gl.nw = nw;
    gl.smem = smem;
    int64_t off = 0;
    int maxinst = 0;
    for (int g = 0; g < ms; ++g) {
      GqGroup& G = T.gen.grp[g];
      G.ninst = cnt[g];
      G.r_off = (int)off;
      off += (int64_t)cnt[g] * G.r_size;
      maxinst = std::max(maxinst, cnt[g]);
    }
    if (off > 0x7fffffff || maxinst > 65535) { gq_fail(why, "factor scratch too large"); return gl; }
    T.maxinst = maxinst;
    gl.r_doubles = off;
    break;
  }
  if (gl.nw < 1) { gq_fail(why, "tile does not fit shared memory"); return gl; }
  gl.rfull_doubles = (int64_t)T.nbf * (T.nbf + 1) / 2 * 64;
  gl.ok = true;
  return gl;
}

cudaError_t gq_launch(const GqLaunch& gl, const GqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_all, int first, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(gq_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gl.smem);
  if (e != cudaSuccess) return e;
  gq_accumulate_kernel<<<gl.sm_count, gl.nw * 32, gl.smem, st>>>(prm, gl.d_tab_g, gl.d_tab_i, q, v, a, tau, ld_tau, N,
                                                                r_all, first);
  return cudaGetLastError();
}

cudaError_t gq_merge(const GqLaunch& gl, const GqParams& prm, double* r_all, int64_t ntiles_hint, unsigned* counters,
                     double* rfull, double* part_out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(gq_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gl.smem_merge);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(counters, 0, sizeof(unsigned) * ((size_t)GQ_MAXG * prm.maxinst + 2), st);
  if (e != cudaSuccess) return e;
  GqMergeArgs M;
  int blocks = 0;
  for (int g = 0; g < GQ_MAXG; ++g) {
    M.blk0[g] = blocks;
    M.leaves[g] = 0;
    if (g >= prm.ms) continue;
    int leaves = prm.gen.grp[g].ninst;
    if (ntiles_hint > 0) leaves = (int)std::max<int64_t>(1, std::min<int64_t>(leaves, ntiles_hint));
    M.leaves[g] = leaves;
    blocks += (leaves + 1) / 2;
  }
  M.blk0[GQ_MAXG] = blocks;
  gq_merge_kernel<<<blocks, 32, gl.smem_merge, st>>>(prm, M, r_all, counters, rfull, part_out);
  return cudaGetLastError();
}

}  // namespace fg
