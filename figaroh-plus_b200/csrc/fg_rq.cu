// fg_rq.cu - host side of the fifth-generation fused kernel (fg_ringqr.cuh): the symbolic pass that turns
// (kinematic tree, row groups, column selection) into the producers' depth-first walk program and one local
// dense problem per row group, the split of the warps into producers and consumers, and the launcher.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "fg_ringqr.cuh"

namespace fg {

static void rq_fail(const char** why, const char* msg) {
  if (why) *why = msg;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  return atoi(e);
}

// Depth-first walk program over the joints marked in `need` (fg_rq.h: RqOp).  Children are visited with the deepest
// needed subtree last, so the main chain is never walked back up; every other child subtree is a side trip: the
// parent's (v, a) is saved before it, the joint transforms inside it are saved as it descends, and it is walked
// back up (axes transported with the inverse action) when it is done.
RqWalk rq_build_walk(int nj, const DevJoint* joints, const std::vector<char>& need, const std::vector<unsigned>& own,
                     const std::vector<unsigned>& live, int max_proc, int max_ops, int max_slots) {
  RqWalk W;
  std::vector<std::vector<int>> kids(nj);
  std::vector<int> height(nj, 0);
  for (int i = nj - 1; i >= 1; --i)
    if (need[i]) {
      const int pa = joints[i].parent;
      height[pa] = std::max(height[pa], height[i] + 1);
    }
  for (int i = 1; i < nj; ++i)
    if (need[i]) kids[joints[i].parent].push_back(i);
  for (int b = 0; b < nj; ++b)
    std::stable_sort(kids[b].begin(), kids[b].end(), [&](int x, int y) { return height[x] < height[y]; });
  int vdepth = 0;
  bool fail = false;
  std::vector<int> path;   // nodes of the current side trip (every one of them has its transform saved)
  std::vector<int> mslot_of(nj, 255), tof(nj, -1);
  auto push_op = [&](const RqOp& op) {
    if ((int)W.ops.size() >= max_ops) { fail = true; return; }
    W.ops.push_back(op);
  };
  std::function<void(int, bool)> visit = [&](int b, bool in_trip) {
    for (size_t ci = 0; ci < kids[b].size() && !fail; ++ci) {
      const int c = kids[b][ci];
      const bool last = ci + 1 == kids[b].size();
      const bool side = !last && b != 0;          // we come back to b afterwards
      const bool trip = in_trip || side;
      if ((int)W.order.size() >= max_proc) { fail = true; return; }
      tof[c] = (int)W.order.size();
      W.order.push_back(c);
      RqOp op;
      memset(&op, 0, sizeof op);
      op.kind = 0; op.t = (unsigned char)tof[c];
      op.vsave = 255; op.msave = 255; op.vload = 255; op.mslot = 255; op.pad = 255;
      op.from_universe = b == 0;
      int my_vslot = 255;
      if (side) {
        if (vdepth >= max_slots) { fail = true; return; }
        my_vslot = vdepth++;
        W.nvslots = std::max(W.nvslots, vdepth);
        op.vsave = (unsigned char)my_vslot;
      }
      const size_t path_mark = path.size();
      if (trip) {
        if ((int)path.size() >= max_slots) { fail = true; return; }
        mslot_of[c] = (int)path.size();
        op.msave = (unsigned char)mslot_of[c];
        path.push_back(c);
        W.nmslots = std::max(W.nmslots, (int)path.size());
      }
      op.live = b == 0 ? 0u : live[b];
      op.own = own[c];
      push_op(op);
      visit(c, trip);
      if (side) {  // walk back up to b
        while (path.size() > path_mark && !fail) {
          const int x = path.back();
          path.pop_back();
          RqOp up;
          memset(&up, 0, sizeof up);
          up.kind = 1; up.t = (unsigned char)tof[x];
          up.vsave = 255; up.msave = 255; up.vload = 255; up.pad = 255;
          up.mslot = (unsigned char)mslot_of[x];
          up.live = live[joints[x].parent];
          if (path.size() == path_mark) up.vload = (unsigned char)my_vslot;
          push_op(up);
        }
        --vdepth;
      } else if (b == 0 && !last) {
        path.resize(path_mark);  // a chain hanging off the universe is simply left behind
      }
    }
  };
  visit(0, false);
  if (fail) return W;
  {  // input prefetch: walk index of the next descend
    int nxt = 255;
    for (int o = (int)W.ops.size() - 1; o >= 0; --o)
      if (W.ops[o].kind == 0) { W.ops[o].pad = (unsigned char)nxt; nxt = W.ops[o].t; }
  }
  W.ok = true;
  return W;
}

RqLaunch rq_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const ColDesc* cols, int sm_count,
                  int max_smem, RqParams* out, const char** why, const RqCost* measured) {
  RqLaunch rl;
  if (why) *why = "";
  const int ms = P.ms, nj = P.nj, ncols = P.ncols;
  if (ms < 1 || ms > GQ_MAXG) { rq_fail(why, "more than 8 row groups"); return rl; }
  if (nj > kMaxJoints) { rq_fail(why, "too many joints"); return rl; }
  if (ncols + 1 > 8 * GQ_MAXBF) { rq_fail(why, "too many columns"); return rl; }
  RqParams& R = *out;
  memset(&R, 0, sizeof R);
  GqParams& T = R.gq;
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
      // a row group whose motion axis has no angular part (prismatic joint, force rows of a free-flyer) never sees
      // the six inertia-tensor parameters: s^T B has s_ang = 0 in every frame (the transport keeps it zero)
      if (cd.k >= 4)
        for (int g = 0; g < ms; ++g)
          if (rows[g].S[3] == 0.0 && rows[g].S[4] == 0.0 && rows[g].S[5] == 0.0) cmask[c] &= ~(1u << g);
    } else if (cd.kind == COL_EXTRA) {
      if (cd.grp >= 0) cmask[c] = (1u << cd.grp) & allg;
      else if (cd.grp == -1) cmask[c] = allg;
    } else if (cd.kind == COL_COUPLING) {
      for (int g = 0; g < ms; ++g)
        if (rows[g].rowblock == 4 || rows[g].rowblock == 5) cmask[c] |= 1u << g;
    }
  }
  // ---- joints to walk: those that support a selected inertial column of some row group or carry a row group
  std::vector<char> need(nj, 0);
  for (int c = 0; c < ncols; ++c)
    if (cols[c].kind == COL_INERTIAL && cmask[c]) {
      if (cols[c].body < 1 || cols[c].body >= nj || cols[c].k < 0 || cols[c].k > 9) { rq_fail(why, "internal: bad inertial column"); return rl; }
      for (int i = cols[c].body; i > 0; i = joints[i].parent) need[i] = 1;
    }
  for (int g = 0; g < ms; ++g)
    for (int i = rows[g].joint; i > 0; i = joints[i].parent) need[i] = 1;
  // row groups carried by a joint / live at a body
  std::vector<unsigned> own(nj, 0u), live(nj, 0u);
  for (int g = 0; g < ms; ++g) own[rows[g].joint] |= 1u << g;
  for (int i = 1; i < nj; ++i) live[i] = live[joints[i].parent] | own[i];
  const RqWalk walk = rq_build_walk(nj, joints, need, own, live, GQ_MAXPROC, RQ_MAXOPS);
  if (!walk.ok) { rq_fail(why, "kinematic tree too large for the walk program"); return rl; }
  const int nproc = (int)walk.order.size(), nops = (int)walk.ops.size();
  if (nproc < 1) { rq_fail(why, "no joint to walk"); return rl; }
  std::vector<int> tof(nj, -1);
  for (int t = 0; t < nproc; ++t) { tof[walk.order[t]] = t; T.gen.pj[t] = joints[walk.order[t]]; }
  for (int o = 0; o < nops; ++o) R.ops[o] = walk.ops[o];
  const int nv_used = walk.nvslots, nm_used = walk.nmslots;
  T.nproc = nproc;
  R.nops = nops;
  R.nvslots = nv_used;
  R.nmslots = nm_used;
  // ---- the local problem of every row group: per walked body its selected inertial columns in pinocchio
  //      parameter order, then the input-fed columns in the caller's order, then the rhs
  int maxnb = 0, n_emits = 0, n_simple = 0, n_transports = 0;
  for (int o = 0; o < nops; ++o) n_transports += __builtin_popcount(R.ops[o].live);
  std::vector<int> joint_of_t(nproc, 0);
  for (int b = 1; b < nj; ++b) if (tof[b] >= 0) joint_of_t[tof[b]] = b;
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
          if (l >= GQ_MAXW) { rq_fail(why, "a row group has too many columns"); return rl; }
          mask |= 1 << k;
          T.gcol[g][l++] = (unsigned char)colof[k];
        }
      if (mask) {
        if (!((live[joint_of_t[t]] >> g) & 1u)) { rq_fail(why, "internal: column outside its row group's subtree"); return rl; }
        ++n_emits;
      }
      T.gen.emit[g][t] = mask ? ((base << 10) | mask) : 0;
    }
    for (int c = 0; c <= ncols; ++c) {
      if (c < ncols && (cols[c].kind == COL_INERTIAL || !((cmask[c] >> g) & 1u))) continue;
      if (l >= GQ_MAXW || ns >= GQ_MAXSIMPLE) { rq_fail(why, "a row group has too many columns"); return rl; }
      GqSimple& s = T.gen.simple[g][ns++];
      if (c == ncols) { s.kind = COL_RHS; s.src = 0; s.vidx = 0; }
      else { s.kind = (short)cols[c].kind; s.src = (short)cols[c].k; s.vidx = (short)cols[c].vidx; }
      s.slot = (short)l;
      T.gcol[g][l++] = (unsigned char)c;
    }
    G.w = l;
    G.nb = (l + 7) / 8;
    G.nsimple = ns;
    n_simple += ns;
    G.r_size = G.nb * (G.nb + 1) / 2 * 64;
    maxnb = std::max(maxnb, G.nb);
  }
  if (maxnb > GQ_MAXB || T.nbf > GQ_MAXBF) { rq_fail(why, "too many column blocks"); return rl; }
  static_assert(sizeof(RqStepTab) % 4 == 0, "RqStepTab is copied word by word");
  {  // input-fed columns of single-dof joints move from the end of the tile to the joint's walk step
    int goff_cols = 0;
    for (int g = 0; g < ms; ++g) {
      GqGroup& G = T.gen.grp[g];
      int keep = 0;
      for (int t = 0; t < G.nsimple; ++t) {
        const GqSimple s1 = T.gen.simple[g][t];
        int at = -1;
        if (s1.kind == COL_EXTRA)
          for (int w = 0; w < nproc; ++w) {
            const DevJoint& jw = T.gen.pj[w];
            if (jw.jtype >= 1 && jw.jtype <= 3 && jw.idx_v == s1.vidx) at = w;
          }
        const int colpos = goff_cols + s1.slot;
        if (at >= 0 && R.step.n[at] < RQ_STEPCOLS && colpos < (1 << 14)) {
          R.step.col[at][R.step.n[at]++] = (unsigned short)((colpos << 2) | (s1.src & 3));
          --n_simple;
        } else {
          T.gen.simple[g][keep++] = s1;
        }
      }
      G.nsimple = keep;
      goff_cols += G.nb * 8;
    }
  }
  T.maxnb = maxnb;
  rl.smem_merge = (2 * std::max(maxnb, T.nbf) * GQ_BLK + GQ_SCRATCH) * 8;  // 64-row slabs
  if (rl.smem_merge > max_smem) { rq_fail(why, "tile does not fit shared memory"); return rl; }
  {
    int off = 0;
    for (int g = 0; g < ms; ++g) { R.goff[g] = off; off += T.gen.grp[g].nb * GQ_BLK; }
    R.slot_doubles = off;
  }
  R.prod_smem_doubles = (ms * 6 + 12 * (R.nvslots + R.nmslots)) * 32;
  // ---- cost model (warp instructions per 32-sample tile; calibrated on the generation-4 per-line counts)
  std::vector<double> cost(ms);
  double csum = 0.0;
  double c_base = 250.0, c_panel = 1070.0, c_trail = 100.0;
  if (const char* e = getenv("FG_RQ_COST")) sscanf(e, "%lf,%lf,%lf", &c_base, &c_panel, &c_trail);  // A/B only
  for (int g = 0; g < ms; ++g) {
    const GqGroup& G = T.gen.grp[g];
    cost[g] = c_base + c_panel * G.nb + c_trail * G.nb * (G.nb - 1) / 2;
    csum += cost[g];
  }
  double pcost = 0.97 * (400.0 + 420.0 * nproc + 45.0 * n_transports + 190.0 * n_emits + 25.0 * n_simple + 60.0 * (nops - nproc));  // 0.97: the timed calibration's pick on UR10 (22 producer SMs of 148) and TIAGo (21..22)
  if (measured && measured->have) {  // warp cycles per tile, measured on this device with this plan
    csum = 0.0;
    for (int g = 0; g < ms; ++g) { cost[g] = std::max(1.0, measured->grp[g]); csum += cost[g]; }
    pcost = std::max(1.0, measured->prod);
  }
  rl.sm_count = std::max(1, sm_count);
  const int gen_doubles = GQ_GEN_DOUBLES + (int)(sizeof(RqOp) * RQ_MAXOPS / 8) + RQ_STEP_DOUBLES;
  // Roles are assigned per CTA (= per SM): a producer SM runs only the generator, a consumer SM only the fold, so
  // each SM's instruction cache holds one of the two code paths and the 16 producer warps of an SM hide each
  // other's load latencies.  (First version: producers and consumers mixed inside every CTA - the generator's code
  // was evicted between two uses, `no_instruction` became the top stall, profiles/ncu_rq_r02_v5a.md.)
  int npcta = (int)(rl.sm_count * pcost / (pcost + csum) + 0.5);
  if (measured && measured->force_npcta > 0) npcta = measured->force_npcta;
  npcta = env_int("FG_RQ_NPCTA", npcta);
  npcta = std::max(1, std::min(npcta, rl.sm_count - 1));
  if (rl.sm_count < 2) { rq_fail(why, "needs at least two SMs"); return rl; }
  const int ncta = rl.sm_count - npcta;
  const int nw_cap = std::max(1, std::min(env_int("FG_RQ_NW", 16), 20));
  for (int nw = nw_cap; nw >= 1; --nw) {
    if (R.prod_smem_doubles * nw * 8 + gen_doubles * 8 > max_smem) continue;
    const int W = ncta * nw;
    if (W < ms) continue;
    // largest-remainder apportionment of the consumer warps, at least one per row group
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
    // interleave: consumer position cw goes to the row group that is furthest behind its proportional schedule;
    // producer CTAs are spread evenly over the grid
    rl.tab_g.assign((size_t)rl.sm_count * nw, 0);
    rl.tab_i.assign((size_t)rl.sm_count * nw, 0);
    std::vector<int> given(ms, 0);
    int worst = 0, cw = 0, pw = 0;
    for (int b = 0; b < rl.sm_count; ++b) {
      const bool prod = (int64_t)(b + 1) * npcta / rl.sm_count > (int64_t)b * npcta / rl.sm_count;
      int doubles = 0;
      for (int w = 0; w < nw; ++w) {
        const int gw = b * nw + w;
        if (prod) {
          rl.tab_g[gw] = 255;
          rl.tab_i[gw] = (unsigned short)pw++;
          doubles += R.prod_smem_doubles;
          continue;
        }
        int best = -1;
        double lag = -1e300;
        for (int g = 0; g < ms; ++g) {
          if (given[g] >= cnt[g]) continue;
          const double l = (double)cnt[g] * (cw + 1) / W - given[g];
          if (l > lag) { lag = l; best = g; }
        }
        ++cw;
        rl.tab_g[gw] = (unsigned char)best;
        rl.tab_i[gw] = (unsigned short)given[best]++;
        doubles += T.gen.grp[best].nb * GQ_BLK + GQ_SCRATCH + 2;
      }
      worst = std::max(worst, doubles);
    }
    const int smem = (gen_doubles + worst) * 8;
    if (smem > max_smem) continue;
    rl.nw = nw;
    rl.smem = smem;
    R.npc = npcta;
    R.nprod = pw;
    int64_t off = 0;
    int maxinst = 0;
    for (int g = 0; g < ms; ++g) {
      GqGroup& G = T.gen.grp[g];
      G.ninst = cnt[g];
      G.r_off = (int)off;
      off += (int64_t)cnt[g] * G.r_size;
      maxinst = std::max(maxinst, cnt[g]);
    }
    if (off > 0x7fffffff || maxinst > 65535) { rq_fail(why, "factor scratch too large"); return rl; }
    T.maxinst = maxinst;
    rl.r_doubles = off;
    break;
  }
  if (rl.nw < 1) { rq_fail(why, "tile does not fit shared memory"); return rl; }
  // Ring size: every producer warp holds one slot while it fills it, so the ring needs the producer warps plus a little
  // slack; beyond that, slots only cost L2 (ring + per-warp factors must stay inside the 126 MB).  A/B on B200 - UR10
  // (352 producer warps, slots of 62 KB): 320 slots 3.90 ms, 352 3.59, 384 3.39, 416 3.42, 448 3.45, 512 3.46; TIAGo
  // (220 producer warps, slots of 132 KB): 192 11.22 ms, 256..320 10.70, 384 10.86, 512 10.95 (8.4 GB of DRAM traffic
  // per 1 M samples at 512 slots, profiles/ncu_rq_tiago_r02_final.md).
  const int ring_auto = std::min(512, std::max(64, (R.nprod + 32 + 31) / 32 * 32));
  R.ring = std::max(8, env_int("FG_RQ_RING", ring_auto));
  R.sleep_ns = std::max(0, env_int("FG_RQ_SLEEP", 400));
  rl.ring_doubles = (int64_t)R.ring * R.slot_doubles;
  rl.rfull_doubles = (int64_t)T.nbf * (T.nbf + 1) / 2 * 64;
  rl.ok = true;
  return rl;
}

cudaError_t rq_launch(const RqLaunch& rl, const RqParams& prm, const double* q, const double* v, const double* a,
                      const double* tau, int64_t ld_tau, int64_t N, double* r_all, double* ring, unsigned* ring_ctl,
                      int first, cudaStream_t st, unsigned long long* prof) {
  const void* kern = prof ? (rl.nw > 16 ? (const void*)rq_accumulate_kernel<20, true> : (const void*)rq_accumulate_kernel<16, true>)
                          : (rl.nw > 16 ? (const void*)rq_accumulate_kernel<20, false> : (const void*)rq_accumulate_kernel<16, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, rl.smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ring_ctl, 0, sizeof(unsigned) * (2 * (size_t)prm.ring), st);  // ready / done; the error word stays until it is reported
  if (e != cudaSuccess) return e;
  const unsigned char* tg = rl.d_tab_g;
  const unsigned short* ti = rl.d_tab_i;
  void* args[] = {(void*)&prm, (void*)&tg, (void*)&ti, (void*)&q, (void*)&v, (void*)&a, (void*)&tau,
                  (void*)&ld_tau, (void*)&N, (void*)&r_all, (void*)&ring, (void*)&ring_ctl, (void*)&first, (void*)&prof};
  // every CTA must be resident (producers and consumers wait for each other): cooperative launch
  return cudaLaunchCooperativeKernel(kern, dim3(rl.sm_count), dim3(rl.nw * 32), args, (size_t)rl.smem, st);
}

}  // namespace fg
