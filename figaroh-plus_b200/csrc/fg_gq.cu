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
  if (ncols + 1 > GQ_MAXW) { gq_fail(why, "too many columns"); return gl; }
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
  int r_off = 0, maxnb = 0;
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
    G.r_off = r_off;
    r_off += G.nb * (G.nb + 1) / 2 * 64;
    maxnb = std::max(maxnb, G.nb);
  }
  if (maxnb > GQ_MAXB || T.nbf > GQ_MAXB) { gq_fail(why, "more than 12 column blocks"); return gl; }
  T.maxnb = maxnb;
  T.r_leaf_doubles = r_off;
  // ---- roles: longest-processing-time assignment of the row groups to GQ_ROLES warps
  {
    std::vector<std::pair<double, int>> cost;
    for (int g = 0; g < ms; ++g) {
      const int nb = T.gen.grp[g].nb;
      cost.push_back({3000.0 + 1600.0 * nb + 300.0 * nb * (nb - 1) / 2, g});
    }
    std::sort(cost.begin(), cost.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
      return a.first != b.first ? a.first > b.first : a.second < b.second;
    });
    double load[GQ_ROLES] = {0, 0, 0, 0};
    for (auto& cg : cost) {
      int best = 0;
      for (int r = 1; r < GQ_ROLES; ++r)
        if (load[r] < load[best]) best = r;
      T.role_g[best][T.role_n[best]++] = cg.second;
      load[best] += cg.first;
    }
  }
  gl.smem_warp = (maxnb * GQ_BLK + GQ_SCRATCH) * 8;
  gl.smem_merge = (std::max(maxnb, T.nbf) * GQ_BLK + GQ_SCRATCH) * 8;
  if (gl.smem_merge > max_smem) { gq_fail(why, "tile does not fit shared memory"); return gl; }
  gl.nw = std::min(GQ_MAXNW, (max_smem - (int)sizeof(GqGen)) / gl.smem_warp);
  if (gl.nw < 1) { gq_fail(why, "tile does not fit shared memory"); return gl; }
  gl.sm_count = std::max(1, sm_count);
  gl.leaves = std::max(1, gl.sm_count * gl.nw / GQ_ROLES);
  gl.r_leaf_doubles = r_off;
  gl.rfull_doubles = (int64_t)T.nbf * (T.nbf + 1) / 2 * 64;
  gl.ok = true;
  return gl;
}

static int gq_grid_for(const GqLaunch& gl, int64_t N, bool whole) {
  if (!whole) return gl.sm_count;
  const int64_t ntiles = std::max<int64_t>(1, (N + GQ_S - 1) / GQ_S);
  const int64_t warps = std::min<int64_t>(ntiles, gl.leaves) * GQ_ROLES;
  return (int)std::min<int64_t>(gl.sm_count, (warps + gl.nw - 1) / gl.nw);
}

int gq_leaves_for(const GqLaunch& gl, int64_t N, bool whole) {
  return std::max(1, gq_grid_for(gl, N, whole) * gl.nw / GQ_ROLES);
}

cudaError_t gq_launch(const GqLaunch& gl, const GqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_all, int leaves, int first,
                      cudaStream_t st) {
  const int smem = gl.nw * gl.smem_warp + (int)sizeof(GqGen);
  cudaError_t e = cudaFuncSetAttribute(gq_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int grid = (leaves * GQ_ROLES + gl.nw - 1) / gl.nw;
  gq_accumulate_kernel<<<grid, gl.nw * 32, smem, st>>>(prm, q, v, a, tau, ld_tau, N, r_all, leaves, first);
  return cudaGetLastError();
}

cudaError_t gq_merge(const GqLaunch& gl, const GqParams& prm, double* r_all, int leaves, unsigned* counters,
                     double* rfull, double* part_out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(gq_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gl.smem_merge);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(counters, 0, sizeof(unsigned) * ((size_t)GQ_MAXG * leaves + 2), st);
  if (e != cudaSuccess) return e;
  const int npairs = (leaves + 1) / 2;
  gq_merge_kernel<<<prm.ms * npairs, 32, gl.smem_merge, st>>>(prm, r_all, leaves, npairs, counters, rfull, part_out);
  return cudaGetLastError();
}

}  // namespace fg
