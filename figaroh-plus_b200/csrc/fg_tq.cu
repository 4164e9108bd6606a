// fg_tq.cu - host side of the third-generation fused kernel (fg_tileqr.cuh): the symbolic pass
// that turns (kinematic tree, row groups, column selection) into the block lists the kernel
// walks, and the launcher.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fg_tileqr.cuh"

namespace fg {

static const char* fail(const char** why, const char* msg) {
  if (why) *why = msg;
  return msg;
}

TqLaunch tq_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const ColDesc* cols, int sm_count,
                  int max_smem, TqParams* out, const char** why) {
  TqLaunch tl;
  if (why) *why = "";
  const int ms = P.ms, nj = P.nj, ncols = P.ncols;
  if (ms < 1 || ms > TQ_MAXG) { fail(why, "more than 8 row groups"); return tl; }
  if (nj > TQ_MAXJ) { fail(why, "more than 32 joints"); return tl; }
  if (ncols + 1 > TQ_MAXP) { fail(why, "too many columns"); return tl; }
  TqParams& T = *out;
  memset(&T, 0, sizeof T);
  T.nj = nj; T.nq = P.nq; T.nv = P.nv; T.ms = ms; T.mode = P.mode; T.flags = P.flags;
  for (int k = 0; k < 3; ++k) T.gravity[k] = P.gravity[k];
  for (int i = 0; i < nj; ++i) T.joints[i] = joints[i];
  for (int g = 0; g < ms; ++g) {
    for (int k = 0; k < 6; ++k) T.rowS[g][k] = rows[g].S[k];
    T.rowjoint[g] = rows[g].joint;
    T.rowblock[g] = rows[g].rowblock;
  }
  const unsigned allg = (1u << ms) - 1u;
  // ---- row-group mask of every selected column
  std::vector<unsigned> cmask(ncols, 0u);
  std::vector<char> priv(ncols, 0);
  for (int c = 0; c < ncols; ++c) {
    const ColDesc& cd = cols[c];
    if (cd.kind == COL_INERTIAL) {
      cmask[c] = (unsigned)(col_ancmask(cd) & allg);
    } else if (cd.kind == COL_EXTRA) {
      if (cd.grp >= 0) { cmask[c] = 1u << cd.grp; priv[c] = 1; }
      else if (cd.grp == -1) cmask[c] = allg;
    } else if (cd.kind == COL_COUPLING) {
      for (int g = 0; g < ms; ++g)
        if (rows[g].rowblock == 4 || rows[g].rowblock == 5) cmask[c] |= 1u << g;
    }
  }
  // ---- column order: private blocks (one per row group), general columns in the given order,
  //      columns that are zero in every row group, rhs
  struct Blk { std::vector<int> cols; int priv_g; };
  std::vector<Blk> blocks;
  for (int g = 0; g < ms; ++g) {
    Blk b; b.priv_g = g;
    for (int c = 0; c < ncols; ++c)
      if (priv[c] && cols[c].grp == g) b.cols.push_back(c);
    if (b.cols.size() > 4) { fail(why, "more than 4 private columns in a row group"); return tl; }
    if (!b.cols.empty()) blocks.push_back(b);
  }
  const int npriv = (int)blocks.size();
  std::vector<int> gen;
  for (int c = 0; c < ncols; ++c) if (!priv[c] && cmask[c]) gen.push_back(c);
  for (int c = 0; c < ncols; ++c) if (!priv[c] && !cmask[c]) gen.push_back(c);
  gen.push_back(ncols);  // rhs
  for (size_t i = 0; i < gen.size(); i += 8) {
    Blk b; b.priv_g = -1;
    for (size_t k = i; k < std::min(gen.size(), i + 8); ++k) b.cols.push_back(gen[k]);
    blocks.push_back(b);
  }
  const int ncb = (int)blocks.size();
  if (ncb > TQ_MAXCB) { fail(why, "more than 32 column blocks"); return tl; }
  T.ncb = ncb; T.npriv = npriv; T.p1 = ncols + 1;
  T.sm_count = std::max(1, sm_count);
  // ---- activity of (row group, column block), with the fill-in of the general panels
  std::vector<unsigned> act(ncb, 0u);  // bit g
  for (int j = 0; j < ncb; ++j)
    for (int c : blocks[j].cols) act[j] |= c == ncols ? allg : cmask[c];
  for (int e = 0; e < npriv; ++e) {
    T.gmask[e] = act[e];  // == 1 << priv_g
    unsigned tm = 0;
    for (int j = npriv; j < ncb; ++j) if (act[j] & act[e]) tm |= 1u << j;
    T.trmask[e] = tm;
  }
  for (int k = npriv; k < ncb; ++k) {
    const unsigned gm = act[k];
    T.gmask[k] = gm;
    unsigned tm = 0;
    if (gm)
      for (int j = k + 1; j < ncb; ++j)
        if (act[j] & gm) { act[j] |= gm; tm |= 1u << j; }
    T.trmask[k] = tm;
  }
  // ---- tile storage
  int cursor = 0;
  for (int g = 0; g < TQ_MAXG; ++g)
    for (int j = 0; j < TQ_MAXCB; ++j) T.blkoff[g][j] = -1;
  for (int j = 0; j < ncb; ++j)
    for (int g = 0; g < ms; ++g)
      if ((act[j] >> g) & 1u) { T.blkoff[g][j] = cursor; cursor += (j < npriv ? 4 : 8) * TQ_S; }
  T.tile_doubles = cursor;
  // ---- state columns, slots, simple columns
  for (int b = 0; b < TQ_MAXJ; ++b)
    for (int k = 0; k < 10; ++k) T.islot[b][k] = -1;
  tl.perm.assign(ncols, -1);
  int sc = 0, nsimple = 0;
  for (int j = 0; j < ncb; ++j) {
    T.ncols[j] = (unsigned char)blocks[j].cols.size();
    for (size_t ci = 0; ci < blocks[j].cols.size(); ++ci, ++sc) {
      const int c = blocks[j].cols[ci];
      const int slot = (j << 3) | (int)ci;
      T.sc_blk[sc] = (unsigned char)j;
      T.sc_c[sc] = (unsigned char)ci;
      if (c == ncols) {
        if (sc != ncols) { fail(why, "internal: rhs is not the last state column"); return tl; }
        TqSimple& s = T.simple[nsimple++];
        s.kind = COL_RHS; s.src = 0; s.vidx = 0; s.grp = -1; s.slot = slot;
        continue;
      }
      tl.perm[sc] = c;
      const ColDesc& cd = cols[c];
      if (cd.kind == COL_INERTIAL) {
        if (cd.body < 1 || cd.body >= nj || cd.k < 0 || cd.k > 9) { fail(why, "internal: bad inertial column"); return tl; }
        T.islot[cd.body][cd.k] = (short)slot;
      } else if (cd.kind == COL_EXTRA || cd.kind == COL_COUPLING) {
        if (nsimple >= TQ_MAXSIMPLE - 1) { fail(why, "too many friction / offset columns"); return tl; }
        TqSimple& s = T.simple[nsimple++];
        s.kind = cd.kind; s.src = cd.k; s.vidx = cd.vidx; s.grp = cd.kind == COL_EXTRA ? cd.grp : -1; s.slot = slot;
      }
    }
  }
  T.nsimple = nsimple;
  for (int c = 0; c < 8; ++c) T.swz[c] = (unsigned char)tq_sw(c);
  // ---- forward pass: the joints that support a generated body, which must form chains hanging
  //      off the universe when walked in index order (parent = previously walked joint, or 0)
  std::vector<char> need(nj, 0);
  for (int c = 0; c < ncols; ++c)
    if (cols[c].kind == COL_INERTIAL && cmask[c])
      for (int i = cols[c].body; i > 0; i = joints[i].parent) need[i] = 1;
  for (int g = 0; g < ms; ++g)
    for (int i = rows[g].joint; i > 0; i = joints[i].parent) need[i] = 1;
  int nproc = 0, prev = 0;
  for (int i = 1; i < nj; ++i) {
    if (!need[i]) continue;
    if (joints[i].parent != 0 && joints[i].parent != prev) { fail(why, "kinematic tree branches below the base"); return tl; }
    T.proc[nproc++] = i;
    prev = i;
  }
  T.nproc = nproc;
  for (int b = 1; b < nj; ++b) {
    unsigned m = 0;
    for (int g = 0; g < ms; ++g) {
      int i = b;
      while (i > rows[g].joint) i = joints[i].parent;
      if (i == rows[g].joint) m |= 1u << g;
    }
    bool has = false;
    for (int k = 0; k < 10; ++k) has = has || T.islot[b][k] >= 0;
    T.anc[b] = has ? (unsigned char)m : 0;
  }
  T.scratch_doubles = std::max(nproc * 2 * TQ_S + nj * (int)(sizeof(DevJoint) / 8), 4 * 112);
  tl.smem = (T.tile_doubles + T.scratch_doubles) * 8;
  if (tl.smem > max_smem) { fail(why, "tile does not fit shared memory"); return tl; }
  tl.maxg = 0;
  for (int k = npriv; k < ncb; ++k) tl.maxg = std::max(tl.maxg, __builtin_popcount(T.gmask[k]));
  tl.r_doubles = (int64_t)(ncb * (ncb + 1) / 2) * 64;
  const int per_sm = std::max(1, std::min(3, (max_smem + 1024) / (tl.smem + 1024)));
  tl.grid = sm_count * per_sm;
  tl.ok = true;
  return tl;
}

cudaError_t tq_launch(const TqLaunch& tl, const TqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_scratch, double* parts, int grid,
                      int phase, cudaStream_t st) {
  auto kern = tl.maxg <= 6 ? tq_accumulate_kernel<6> : tq_accumulate_kernel<8>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tl.smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, TQ_WARPS * 32, tl.smem, st>>>(prm, q, v, a, tau, ld_tau, N, r_scratch, tl.r_doubles, parts, phase);
  return cudaGetLastError();
}

}  // namespace fg
