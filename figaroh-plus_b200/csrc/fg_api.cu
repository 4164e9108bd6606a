// fg_api.cu - host side of libfigaroh_b200.so: model upload, regressor plans, kernel
// dispatch, TSQR state management and the extern "C" ABI declared in include/figaroh_b200.h.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/figaroh_b200.h"
#include "fg_kernels.cuh"
#include "fg_tile.cuh"
#include "fg_gq.h"
#include "fg_rq.h"
#include "fg_wide.h"
#include "fg_matz.h"
#include "fg_coop.h"
#include "fg_prefilter.cuh"

using namespace fg;

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define FG_CUDA(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess)                                                                           \
      return set_err(FG_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// number of kernels this library has launched (diagnostic: bench.py reports it as gpu_launches)
static std::atomic<long long> g_launches{0};
#define FG_LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// ------------------------------------------------------------------------------------ objects
// Row blocks with the same set of structurally non-zero columns (cooperative path): blocks = positions in the
// materialised W, d_colsel = the group's columns as indices into the full W, d_colmap = local column -> column of group 0
// (the group that sees every column), both with the rhs as last entry (d_colsel[p1-1] unused).
struct CqGroup {
  std::vector<int> blocks;
  int p1 = 0;
  int* d_colsel = nullptr;
  int* d_colmap = nullptr;
};

struct DevPlan {
  Plan P;
  std::vector<ColDesc> cols;
  std::vector<RowDesc> rows;
  std::vector<int> nact, fullcol;
  void* d_blob = nullptr;  // one allocation holding everything
  DevPlanPtrs ptrs{};
  // row-block groups of the cooperative path (build_cq_groups)
  bool cqg_built = false;
  std::vector<CqGroup> cqg;
  // materialising kernel, second version (fg_matz.cuh)
  bool matz_tried = false;
  MatzPlan matz;
  void* d_matz_tab = nullptr;
  // fifth-generation kernel (fg_ringqr.cuh)
  bool rq_tried = false;
  bool rq_calibrated = false;   // warps re-apportioned by measured per-tile cycles (rq_calibrate)
  RqLaunch rql;
  std::unique_ptr<RqParams> rq;
  void* d_rq_tab = nullptr;
  ~DevPlan() { if (d_blob) cudaFree(d_blob); if (d_rq_tab) cudaFree(d_rq_tab); if (d_matz_tab) cudaFree(d_matz_tab);
    for (auto& g : cqg) { if (g.d_colsel) cudaFree(g.d_colsel); if (g.d_colmap) cudaFree(g.d_colmap); } }
};

struct fg_model {
  int njoints = 0, nq = 0, nv = 0;
  std::vector<DevJoint> joints;
  std::vector<double> mass;
  double gravity[3] = {0, 0, -9.81};
  DevJoint* d_joints = nullptr;
  std::map<std::string, std::unique_ptr<DevPlan>> plans;
  int sm_count = 148;
  int max_smem = 227 * 1024;
};

struct fg_state {
  int p = 0, p1 = 0;
  int64_t rows = 0;
  double* d_rows_pending = nullptr;  // rows that arrived inside packed factors, summed on the device
  bool rows_pending = false;
  bool empty = true;
  double* d_part = nullptr;     // merged factor: packed triangle + norms   [part_doubles(p1)]
  double* d_parts = nullptr;    // scratch partial factors
  int64_t parts_cap = 0;        // in parts
  double* d_gR = nullptr;       // per-warp factors in HBM (only when they do not fit shared memory)
  int64_t gR_cap = 0;           // in doubles
  double* d_ring = nullptr;     // generation 5: ring of regressor tiles + its control words (L2 resident)
  int64_t ring_cap = 0;         // in doubles
  const void* ring_plan = nullptr;  // the plan whose tile layout the ring currently holds (padding columns stay zero)
  int ring_slots = 0;               // ... and its slot count (the control words sit behind the last slot)
  int64_t hint_total = 0;       // host-buffer path: samples of the whole call (the chunks are shorter)
  int cq_grid = 0;              // CTAs (= per-CTA factors in d_gR) of the running cooperative-TSQR sequence
  unsigned* h_rq_err = nullptr; // pinned: error flag of the last generation-5 launch
  double* d_A = nullptr;        // (p1 x p1) dense scratch of the small stage
  double* d_Re = nullptr;       // reduced factor (dense, (nk+1)^2)
  double* d_X = nullptr;        // solve output
  int* d_idx = nullptr;         // 2*p1 ints: column map + extents
  double* d_solve = nullptr;    // output blob of fg_state_solve (4 + 3 p1 + 2 p1^2 doubles)
  int nk = -1;                  // columns kept by the last fg_state_reduce
  std::vector<int> keep;
  // column order of the factor: state column i holds caller column perm[i] (empty = identity).
  // The tile-QR kernel eliminates the friction / inertia / offset columns first (fg_tq.cu).
  std::vector<int> perm;
  // staging for the host-buffer path
  double* d_stage = nullptr;
  int64_t stage_cap = 0;
  cudaStream_t copy_streams[4] = {nullptr, nullptr, nullptr, nullptr};   // host-buffer path: one per input array
  cudaEvent_t ev_copy[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // optional timing of the dominant (fold) kernel
  bool timing = false;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  double main_ms_sum = 0.0;
  long long main_launches = 0;
  bool t_pending = false;
};

static int check_rq_error(fg_state* s);

static int timing_begin(fg_state* s, cudaStream_t st) {
  if (!s->timing) return FG_OK;
  if (s->t_pending) {  // collect the previous launch first
    FG_CUDA(cudaEventSynchronize(s->t1));
    float ms = 0;
    FG_CUDA(cudaEventElapsedTime(&ms, s->t0, s->t1));
    s->main_ms_sum += ms; s->main_launches += 1; s->t_pending = false;
  }
  FG_CUDA(cudaEventRecord(s->t0, st));
  return FG_OK;
}
static int timing_end(fg_state* s, cudaStream_t st) {
  if (!s->timing) return FG_OK;
  FG_CUDA(cudaEventRecord(s->t1, st));
  s->t_pending = true;
  return FG_OK;
}

static const int FIG2PIN[10] = {4, 5, 7, 6, 8, 9, 1, 2, 3, 0};  // figaroh slot -> pinocchio parameter

// ------------------------------------------------------------------------------------ model
extern "C" const char* fg_last_error(void) { return g_err; }
extern "C" int fg_abi_version(void) { return FG_ABI_VERSION; }

// host part of fg_model_create: validate and flatten the description (no device work)
static int model_from_desc(const fg_model_desc* d, fg_model* m) {
  if (d->njoints < 2 || d->njoints > kMaxJoints)
    return set_err(FG_ERR_UNSUPPORTED, "fg_model_create: njoints=%d outside [2,%d]", d->njoints, kMaxJoints);
  m->njoints = d->njoints; m->nq = d->nq; m->nv = d->nv;
  m->joints.resize(d->njoints);
  m->mass.assign(d->njoints, 1.0);
  int nq = 0, nv = 0;
  for (int i = 0; i < d->njoints; ++i) {
    DevJoint& j = m->joints[i];
    memset(&j, 0, sizeof j);
    for (int k = 0; k < 9; ++k) j.plcR[k] = i == 0 ? (k % 4 == 0 ? 1.0 : 0.0) : d->plc_R[9 * i + k];
    for (int k = 0; k < 3; ++k) { j.plcp[k] = i == 0 ? 0.0 : d->plc_p[3 * i + k]; j.axis[k] = d->axis[3 * i + k]; }
    j.parent = d->parent[i]; j.jtype = d->jtype[i]; j.idx_q = d->idx_q[i]; j.idx_v = d->idx_v[i];
    if (d->mass) m->mass[i] = d->mass[i];
    if (i == 0) continue;
    if (j.parent < 0 || j.parent >= i) return set_err(FG_ERR_INVALID, "joint %d: parent %d not before it", i, j.parent);
    int jq, jv;
    switch (j.jtype) {
      case FG_JT_REVOLUTE: jq = 1; jv = 1; break;
      case FG_JT_REVOLUTE_UNB: jq = 2; jv = 1; break;
      case FG_JT_PRISMATIC: jq = 1; jv = 1; break;
      case FG_JT_FREEFLYER: jq = 7; jv = 6; break;
      default: return set_err(FG_ERR_UNSUPPORTED, "joint %d: unsupported joint type %d", i, j.jtype);
    }
    if (j.idx_q != nq || j.idx_v != nv) return set_err(FG_ERR_INVALID, "joint %d: idx_q/idx_v not contiguous", i);
    nq += jq; nv += jv;
  }
  if (nq != d->nq || nv != d->nv) return set_err(FG_ERR_INVALID, "nq/nv (%d,%d) do not match joints (%d,%d)", d->nq, d->nv, nq, nv);
  if (d->gravity) for (int k = 0; k < 3; ++k) m->gravity[k] = d->gravity[k];
  return FG_OK;
}

extern "C" int fg_model_create(const fg_model_desc* d, fg_model** out) {
  if (!d || !out) return set_err(FG_ERR_INVALID, "fg_model_create: null argument");
  std::unique_ptr<fg_model> m(new fg_model);
  int rc0 = model_from_desc(d, m.get());
  if (rc0) return rc0;
  FG_CUDA(cudaMalloc(&m->d_joints, sizeof(DevJoint) * d->njoints));
  FG_CUDA(cudaMemcpy(m->d_joints, m->joints.data(), sizeof(DevJoint) * d->njoints, cudaMemcpyHostToDevice));
  int dev = 0;
  FG_CUDA(cudaGetDevice(&dev));
  FG_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev));
  FG_CUDA(cudaDeviceGetAttribute(&m->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  *out = m.release();
  return FG_OK;
}

extern "C" void fg_model_destroy(fg_model* m) {
  if (!m) return;
  if (m->d_joints) cudaFree(m->d_joints);
  delete m;
}

static int joint_of_vidx(const fg_model* m, int vi, int* local) {
  for (int i = 1; i < m->njoints; ++i) {
    const int nvj = m->joints[i].jtype == FG_JT_FREEFLYER ? 6 : 1;
    if (vi >= m->joints[i].idx_v && vi < m->joints[i].idx_v + nvj) { *local = vi - m->joints[i].idx_v; return i; }
  }
  return -1;
}

static int check_cfg(const fg_model* m, const fg_regressor_cfg* cfg) {
  if (!m || !cfg) return set_err(FG_ERR_INVALID, "null model or cfg");
  const int nb = m->njoints - 1;
  if (cfg->mode == FG_MODE_JOINT_TORQUES) {
    if (m->nv != nb)
      return set_err(FG_ERR_UNSUPPORTED,
                     "joint-torque mode needs nv == njoints-1 (got nv=%d, bodies=%d): the reference cannot "
                     "broadcast a %d-wide regressor row into %d columns (regressor.py:53)", m->nv, nb, 10 * nb, 10 * m->nv);
  } else if (cfg->mode == FG_MODE_EXTERNAL_WRENCH) {
    if (m->nv < 6) return set_err(FG_ERR_UNSUPPORTED, "external-wrench mode needs nv >= 6");
    if (cfg->flags & FG_FLAG_TX40_COUPLING) return set_err(FG_ERR_UNSUPPORTED, "TX40 coupling is a joint-torque feature");
  } else {
    return set_err(FG_ERR_INVALID, "unknown mode %d", cfg->mode);
  }
  if ((cfg->flags & FG_FLAG_TX40_COUPLING) && m->nv < 6) return set_err(FG_ERR_UNSUPPORTED, "TX40 coupling needs 6 joints");
  return FG_OK;
}

extern "C" int fg_regressor_shape(const fg_model* m, const fg_regressor_cfg* cfg, int32_t* rps, int32_t* cols) {
  int rc = check_cfg(m, cfg);
  if (rc) return rc;
  if (rps) *rps = cfg->mode == FG_MODE_JOINT_TORQUES ? (cfg->n_row_sel > 0 ? cfg->n_row_sel : m->nv) : 6;
  if (cols) *cols = 14 * (m->njoints - 1) + ((cfg->flags & FG_FLAG_TX40_COUPLING) ? 3 : 0);
  return FG_OK;
}

// Host part of a plan: row groups, column descriptors, first-active columns (no device work).
static int build_plan_host(const fg_model* m, const fg_regressor_cfg* cfg, const int32_t* col_sel, int n_col_sel,
                           DevPlan* pl) {
  const int nb = m->njoints - 1;
  const int ncols_full = 14 * nb + ((cfg->flags & FG_FLAG_TX40_COUPLING) ? 3 : 0);
  // ---- rows
  std::vector<int> vsel;
  if (cfg->mode == FG_MODE_JOINT_TORQUES) {
    if (cfg->n_row_sel > 0) {
      for (int i = 0; i < cfg->n_row_sel; ++i) {
        const int v = cfg->row_sel[i];
        if (v < 0 || v >= m->nv || (i > 0 && v <= cfg->row_sel[i - 1]))
          return set_err(FG_ERR_INVALID, "row_sel must be ascending velocity indices in [0,%d)", m->nv);
        vsel.push_back(v);
      }
    } else {
      for (int v = 0; v < m->nv; ++v) vsel.push_back(v);
    }
  } else {
    for (int v = 0; v < 6; ++v) vsel.push_back(v);
  }
  if ((int)vsel.size() > kMaxRowGroups) return set_err(FG_ERR_UNSUPPORTED, "more than %d row blocks", kMaxRowGroups);
  for (size_t g = 0; g < vsel.size(); ++g) {
    int local = 0;
    const int j = joint_of_vidx(m, vsel[g], &local);
    if (j < 0) return set_err(FG_ERR_INVALID, "no joint owns velocity index %d", vsel[g]);
    RowDesc rd;
    memset(&rd, 0, sizeof rd);
    const DevJoint& J = m->joints[j];
    const bool enabled = cfg->mode == FG_MODE_JOINT_TORQUES || ((cfg->ft_mask >> g) & 1);
    if (enabled) {
      if (J.jtype == FG_JT_REVOLUTE || J.jtype == FG_JT_REVOLUTE_UNB) for (int k = 0; k < 3; ++k) rd.S[3 + k] = J.axis[k];
      else if (J.jtype == FG_JT_PRISMATIC) for (int k = 0; k < 3; ++k) rd.S[k] = J.axis[k];
      else rd.S[local] = 1.0;
    }
    rd.joint = j;
    rd.rowblock = cfg->mode == FG_MODE_JOINT_TORQUES ? vsel[g] : (int)g;
    rd.start_col = 14 * (j - 1);
    pl->rows.push_back(rd);
  }
  // ---- columns (full reference numbering first)
  std::vector<ColDesc> full(ncols_full);
  for (auto& c : full) { memset(&c, 0, sizeof c); c.kind = COL_ZERO; c.grp = -2; }
  auto grp_of_rowblock = [&](int rb) {
    for (size_t g = 0; g < pl->rows.size(); ++g) if (pl->rows[g].rowblock == rb) return (int)g;
    return -2;
  };
  for (int b = 0; b < nb; ++b) {
    const int body = b + 1;
    uint64_t anc = 0;
    for (size_t g = 0; g < pl->rows.size(); ++g) {
      int j = body;
      while (j > pl->rows[g].joint) j = m->joints[j].parent;
      if (j == pl->rows[g].joint) anc |= 1ull << g;
    }
    const bool massless = cfg->mode == FG_MODE_EXTERNAL_WRENCH && m->mass[body] == 0.0;
    for (int f = 0; f < 10; ++f) {
      ColDesc& c = full[14 * b + f];
      if (massless) continue;
      c.kind = COL_INERTIAL; c.body = body; c.k = FIG2PIN[f];
      c.anc_lo = (unsigned)(anc & 0xffffffffu); c.anc_hi = (unsigned)(anc >> 32);
    }
    const int grp = cfg->mode == FG_MODE_JOINT_TORQUES ? grp_of_rowblock(b) : -1;
    const bool src_ok = b < m->nv;  // the reference indexes v[i, b] (regressor.py:145-158)
    auto extra = [&](int slot, int src) {
      ColDesc& c = full[14 * b + slot];
      if (!src_ok || grp == -2) return;
      c.kind = COL_EXTRA; c.k = src; c.vidx = b; c.grp = grp;
    };
    if (cfg->flags & FG_FLAG_ACTUATOR_INERTIA) extra(10, SRC_QDD);
    if (cfg->flags & FG_FLAG_FRICTION) { extra(11, SRC_QD); extra(12, SRC_SIGN_QD); }
    if (cfg->flags & FG_FLAG_JOINT_OFFSET) extra(13, SRC_ONE);
  }
  if (cfg->flags & FG_FLAG_TX40_COUPLING)
    for (int k = 0; k < 3; ++k) { ColDesc& c = full[14 * nb + k]; c.kind = COL_COUPLING; c.k = k; }
  // ---- column selection
  std::vector<int> sel;
  if (n_col_sel > 0) {
    for (int i = 0; i < n_col_sel; ++i) {
      if (col_sel[i] < 0 || col_sel[i] >= ncols_full || (i > 0 && col_sel[i] <= col_sel[i - 1]))
        return set_err(FG_ERR_INVALID, "col_sel must be ascending column indices in [0,%d)", ncols_full);
      sel.push_back(col_sel[i]);
    }
  } else {
    for (int c = 0; c < ncols_full; ++c) sel.push_back(c);
  }
  const int ncols = (int)sel.size(), p1 = ncols + 1;
  pl->cols.resize(p1);
  pl->fullcol.resize(p1);
  pl->nact.assign(kMaxRowGroups + 1, 0);
  for (int c = 0; c < ncols; ++c) { pl->cols[c] = full[sel[c]]; pl->fullcol[c] = sel[c]; }
  memset(&pl->cols[ncols], 0, sizeof(ColDesc));
  pl->cols[ncols].kind = COL_RHS; pl->cols[ncols].grp = -1;
  pl->fullcol[ncols] = -1;
  // gstart[g]: first (compact) column at which row group g can be non-zero; the rhs column
  // (index ncols) is non-zero in every row.  Row groups are ordered by joint, so this is ascending.
  for (size_t g = 0; g < pl->rows.size(); ++g) {
    int first = ncols;
    for (int c = 0; c < ncols; ++c) if (sel[c] >= pl->rows[g].start_col) { first = c; break; }
    pl->nact[g] = first;
  }
  pl->nact[pl->rows.size()] = p1;
  Plan& P = pl->P;
  memset(&P, 0, sizeof P);
  P.nj = m->njoints; P.nq = m->nq; P.nv = m->nv; P.ms = (int)pl->rows.size();
  P.ncols = ncols; P.p1 = p1; P.mode = cfg->mode; P.flags = cfg->flags;
  for (int k = 0; k < 3; ++k) P.gravity[k] = m->gravity[k];
  return FG_OK;
}

// Build (or fetch) the device plan of (model, cfg, column selection).
static int get_plan(fg_model* m, const fg_regressor_cfg* cfg, const int32_t* col_sel, int n_col_sel, DevPlan** out) {
  int rc = check_cfg(m, cfg);
  if (rc) return rc;
  std::string key = std::to_string(cfg->mode) + ":" + std::to_string(cfg->flags) + ":" + std::to_string(cfg->ft_mask) + ":";
  for (int i = 0; i < cfg->n_row_sel; ++i) key += std::to_string(cfg->row_sel[i]) + ",";
  key += ":";
  for (int i = 0; i < n_col_sel; ++i) key += std::to_string(col_sel[i]) + ",";
  auto it = m->plans.find(key);
  if (it != m->plans.end()) { *out = it->second.get(); return FG_OK; }
  std::unique_ptr<DevPlan> pl(new DevPlan);
  rc = build_plan_host(m, cfg, col_sel, n_col_sel, pl.get());
  if (rc) return rc;
  const Plan& P = pl->P;
  const int p1 = P.p1;
  // ---- upload as one blob
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_plan = 0, o_rows = al(sizeof(Plan)), o_cols = o_rows + al(sizeof(RowDesc) * pl->rows.size());
  const size_t o_nact = o_cols + al(sizeof(ColDesc) * p1), o_full = o_nact + al(sizeof(int) * (kMaxRowGroups + 1));
  const size_t total = o_full + al(sizeof(int) * p1);
  std::vector<unsigned char> blob(total, 0);
  memcpy(&blob[o_plan], &P, sizeof P);
  memcpy(&blob[o_rows], pl->rows.data(), sizeof(RowDesc) * pl->rows.size());
  memcpy(&blob[o_cols], pl->cols.data(), sizeof(ColDesc) * p1);
  memcpy(&blob[o_nact], pl->nact.data(), sizeof(int) * (kMaxRowGroups + 1));
  memcpy(&blob[o_full], pl->fullcol.data(), sizeof(int) * p1);
  FG_CUDA(cudaMalloc(&pl->d_blob, total));
  FG_CUDA(cudaMemcpy(pl->d_blob, blob.data(), total, cudaMemcpyHostToDevice));
  unsigned char* db = (unsigned char*)pl->d_blob;
  pl->ptrs.plan = (const Plan*)(db + o_plan);
  pl->ptrs.joints = m->d_joints;
  pl->ptrs.rows = (const RowDesc*)(db + o_rows);
  pl->ptrs.cols = (const ColDesc*)(db + o_cols);
  pl->ptrs.nact = (const int*)(db + o_nact);
  pl->ptrs.fullcol = (const int*)(db + o_full);
  *out = pl.get();
  m->plans[key] = std::move(pl);
  return FG_OK;
}

// Host-only description of how the fused path would run (model, cfg): the producers' walk program, the local
// problem of every row group and the split of the warps.  Needs no device; used by the CPU tests of the
// symbolic pass and by DESIGN.md's tables.  JSON into buf.
extern "C" int fg_plan_describe(const fg_model_desc* d, const fg_regressor_cfg* cfg, int32_t sm_count, int32_t max_smem,
                                char* buf, int32_t buflen) {
  if (!d || !cfg || !buf || buflen < 64) return set_err(FG_ERR_INVALID, "fg_plan_describe: bad argument");
  fg_model m;
  int rc = model_from_desc(d, &m);
  if (rc) return rc;
  rc = check_cfg(&m, cfg);
  if (rc) return rc;
  DevPlan pl;
  rc = build_plan_host(&m, cfg, cfg->col_sel, cfg->n_col_sel, &pl);
  if (rc) return rc;
  std::unique_ptr<RqParams> rq(new RqParams);
  const char* why = "";
  const RqLaunch rl = rq_build(pl.P, m.joints.data(), pl.rows.data(), pl.cols.data(), sm_count, max_smem, rq.get(), &why);
  std::string o = "{";
  char t[256];
  snprintf(t, sizeof t, "\"ms\": %d, \"ncols\": %d, \"ring_kernel\": %s, \"why\": \"%s\"", pl.P.ms, pl.P.ncols,
           rl.ok ? "true" : "false", why);
  o += t;
  if (rl.ok) {
    const GqParams& G = rq->gq;
    snprintf(t, sizeof t, ", \"nw\": %d, \"producer_ctas\": %d, \"smem\": %d, \"nproc\": %d, \"ring\": %d, \"slot_doubles\": %d, "
             "\"nvslots\": %d, \"nmslots\": %d, \"r_doubles\": %lld, \"nbf\": %d", rl.nw, rq->npc, rl.smem, G.nproc,
             rq->ring, rq->slot_doubles, rq->nvslots, rq->nmslots, (long long)rl.r_doubles, G.nbf);
    o += t;
    o += ", \"walk\": [";
    for (int i = 0; i < G.nproc; ++i) {
      int joint = -1;
      for (int b = 1; b < m.njoints; ++b)
        if (m.joints[b].idx_v == G.gen.pj[i].idx_v && m.joints[b].idx_q == G.gen.pj[i].idx_q) joint = b;
      snprintf(t, sizeof t, "%s%d", i ? ", " : "", joint);
      o += t;
    }
    o += "], \"ops\": [";
    for (int i = 0; i < rq->nops; ++i) {
      const RqOp& op = rq->ops[i];
      snprintf(t, sizeof t, "%s[%d, %d, %d, %d, %d, %d, %d, %u, %u, %d]", i ? ", " : "", op.kind, op.t, op.vsave, op.msave,
               op.vload, op.mslot, op.from_universe, op.live, op.own, op.pad);
      o += t;
    }
    o += "], \"groups\": [";
    for (int g = 0; g < G.ms; ++g) {
      const GqGroup& H = G.gen.grp[g];
      snprintf(t, sizeof t, "%s{\"joint\": %d, \"w\": %d, \"nb\": %d, \"ninst\": %d, \"nsimple\": %d, \"cols\": [", g ? ", " : "",
               H.joint, H.w, H.nb, H.ninst, H.nsimple);
      o += t;
      for (int l = 0; l < H.w; ++l) { snprintf(t, sizeof t, "%s%d", l ? ", " : "", (int)G.gcol[g][l]); o += t; }
      snprintf(t, sizeof t, "], \"rowblock\": %d, \"S\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"emit\": [", H.rowblock,
               H.S[0], H.S[1], H.S[2], H.S[3], H.S[4], H.S[5]);
      o += t;
      bool firste = true;
      for (int i = 0; i < G.nproc; ++i)
        if (G.gen.emit[g][i]) {
          snprintf(t, sizeof t, "%s[%d, %d, %d]", firste ? "" : ", ", i, G.gen.emit[g][i] >> 10, G.gen.emit[g][i] & 1023);
          o += t;
          firste = false;
        }
      o += "], \"simple\": [";
      for (int i = 0; i < H.nsimple; ++i) {
        const GqSimple& sp = G.gen.simple[g][i];
        snprintf(t, sizeof t, "%s[%d, %d, %d, %d]", i ? ", " : "", sp.kind, sp.src, sp.vidx, sp.slot);
        o += t;
      }
      o += "]}";
    }
    o += "], \"stepcols\": [";  // [walk step, row group, local column, ExtraSrc]: written when the walk is at that joint
    {
      bool firsts = true;
      for (int i = 0; i < G.nproc; ++i)
        for (int k = 0; k < rq->step.n[i]; ++k) {
          const int e = rq->step.col[i][k];
          int colpos = e >> 2, g = 0;
          while (g + 1 < G.ms && colpos >= G.gen.grp[g].nb * 8) { colpos -= G.gen.grp[g].nb * 8; ++g; }
          snprintf(t, sizeof t, "%s[%d, %d, %d, %d]", firsts ? "" : ", ", i, g, colpos, e & 3);
          o += t;
          firsts = false;
        }
    }
    o += "]";
  }
  o += "}";
  if ((int)o.size() + 1 > buflen) return set_err(FG_ERR_INVALID, "fg_plan_describe: buffer too small (%d needed)", (int)o.size() + 1);
  memcpy(buf, o.c_str(), o.size() + 1);
  return FG_OK;
}

// ------------------------------------------------------------------------------------ materialise
template <int MSMAX, int S>
static int launch_materialize(fg_model* m, DevPlan* pl, const double* q, const double* v, const double* a, int64_t N,
                              double* W, int64_t ldw, cudaStream_t st) {
  int nwarps = 8;
  while (nwarps > 1 &&
         make_layout(pl->P.nj, pl->P.ms, pl->P.p1, pl->P.nv, S, MSMAX * S, nwarps, false, true).total > m->max_smem / 2)
    nwarps /= 2;  // aim for two CTAs per SM
  const SmemLayout L = make_layout(pl->P.nj, pl->P.ms, pl->P.p1, pl->P.nv, S, MSMAX * S, nwarps, false, true);
  if (L.total > m->max_smem) return set_err(FG_ERR_UNSUPPORTED, "materialise: %d bytes of shared memory needed", L.total);
  auto kern = regressor_materialize_kernel<MSMAX, S>;
  FG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const int64_t ntiles = (N + S - 1) / S;
  int64_t grid = std::min<int64_t>((ntiles + nwarps - 1) / nwarps, (int64_t)m->sm_count * 16);
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, nwarps * 32, L.total, st>>>(pl->ptrs, q, v, a, N, W, ldw);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

static int materialize_dev(fg_model* m, const fg_regressor_cfg* cfg, const double* q, const double* v, const double* a,
                           int64_t N, double* W, int64_t ldw, cudaStream_t st, int* ncols_out = nullptr) {
  DevPlan* pl = nullptr;
  fg_regressor_cfg c2 = *cfg;
  c2.n_row_sel = 0; c2.row_sel = nullptr;  // the materialised matrix always has every row block
  c2.n_col_sel = 0; c2.col_sel = nullptr;  // ... and every column
  int rc = get_plan(m, &c2, nullptr, 0, &pl);
  if (rc) return rc;
  if (ncols_out) *ncols_out = pl->P.ncols;
  if (N <= 0) return FG_OK;
  if (ldw < pl->P.ncols) return set_err(FG_ERR_INVALID, "ldw=%lld smaller than %d columns", (long long)ldw, pl->P.ncols);
  const int ms = pl->P.ms;
  {  // second version: the producers' walk, one sample per lane, 14-column patches (fg_matz.cuh)
    const char* e = getenv("FG_USE_MATZ");
    const bool use = !(e && e[0] == '0');
    if (use && !pl->matz_tried) {
      pl->matz_tried = true;
      pl->matz = matz_build(pl->P, m->joints.data(), pl->rows.data(), m->mass.data(), pl->P.ncols, m->sm_count, m->max_smem);
      if (pl->matz.ok) {
        FG_CUDA(cudaMalloc(&pl->d_matz_tab, matz_tables_bytes()));
        FG_CUDA(cudaMemcpy(pl->d_matz_tab, pl->matz.tab.get(), matz_tables_bytes(), cudaMemcpyHostToDevice));
      }
    }
    if (use && pl->matz.ok) {
      cudaError_t ce = matz_launch(pl->matz, static_cast<const MatzTables*>(pl->d_matz_tab), q, v, a, N, W, ldw, st);
      if (ce != cudaSuccess) return set_err(FG_ERR_CUDA, "materialise kernel failed: %s", cudaGetErrorString(ce));
      FG_LAUNCHED();
      return FG_OK;
    }
  }
  if (ms <= 8) return launch_materialize<8, 4>(m, pl, q, v, a, N, W, ldw, st);
  if (ms <= 24) return launch_materialize<24, 1>(m, pl, q, v, a, N, W, ldw, st);
  if (ms <= 48) return launch_materialize<48, 1>(m, pl, q, v, a, N, W, ldw, st);
  return set_err(FG_ERR_UNSUPPORTED, "materialise: %d row blocks per sample not supported", ms);
}

extern "C" int fg_regressor_materialize(const fg_model* mc, const fg_regressor_cfg* cfg, const double* q,
                                        const double* v, const double* a, int64_t N, double* W, int64_t ldw,
                                        void* stream) {
  if (!mc || !cfg) return set_err(FG_ERR_INVALID, "fg_regressor_materialize: null argument");
  return materialize_dev(const_cast<fg_model*>(mc), cfg, q, v, a, N, W, ldw, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ state
extern "C" int64_t fg_state_packed_doubles(int32_t p) { return (int64_t)(p + 1) * (p + 1) + 2 * (p + 1); }
extern "C" int fg_state_p(const fg_state* s) { return s ? s->p : -1; }
extern "C" int64_t fg_state_rows(const fg_state* sc) {
  if (!sc) return -1;
  fg_state* s = const_cast<fg_state*>(sc);
  if (s->rows_pending) {  // row counts of merged packed factors: one small synchronous copy, on demand
    double r = 0.0;
    if (cudaMemcpy(&r, s->d_rows_pending, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    cudaMemset(s->d_rows_pending, 0, sizeof(double));
    s->rows += (int64_t)r;
    s->rows_pending = false;
  }
  return s->rows;
}

extern "C" int fg_state_create(int32_t p, fg_state** out) {
  if (p < 1 || p + 1 > WD_MAX_P1 || !out) return set_err(FG_ERR_INVALID, "fg_state_create: p=%d outside [1,%d]", p, WD_MAX_P1 - 1);
  std::unique_ptr<fg_state> s(new fg_state);
  s->p = p; s->p1 = p + 1;
  const int p1 = s->p1;
  FG_CUDA(cudaMalloc(&s->d_part, sizeof(double) * part_doubles(p1)));
  FG_CUDA(cudaMemset(s->d_part, 0, sizeof(double) * part_doubles(p1)));
  FG_CUDA(cudaMalloc(&s->d_A, sizeof(double) * p1 * p1));
  FG_CUDA(cudaMalloc(&s->d_Re, sizeof(double) * p1 * p1));
  FG_CUDA(cudaMalloc(&s->d_X, sizeof(double) * p1 * p1));
  FG_CUDA(cudaMalloc(&s->d_idx, sizeof(int) * 2 * p1));
  *out = s.release();
  return FG_OK;
}

extern "C" void fg_state_destroy(fg_state* s) {
  if (!s) return;
  cudaFree(s->d_rows_pending);
  cudaFree(s->d_part); cudaFree(s->d_parts); cudaFree(s->d_gR); cudaFree(s->d_A); cudaFree(s->d_Re);
  cudaFree(s->d_X); cudaFree(s->d_idx); cudaFree(s->d_stage); cudaFree(s->d_solve); cudaFree(s->d_ring);
  if (s->h_rq_err) cudaFreeHost(s->h_rq_err);
  for (auto& cs : s->copy_streams) if (cs) cudaStreamDestroy(cs);
  for (auto& e : s->ev_copy) if (e) cudaEventDestroy(e);
  for (auto& e : s->ev) if (e) cudaEventDestroy(e);
  if (s->t0) cudaEventDestroy(s->t0);
  if (s->t1) cudaEventDestroy(s->t1);
  delete s;
}

extern "C" long long fg_launch_count(void) { return g_launches.load(); }

extern "C" int fg_state_set_timing(fg_state* s, int on) {
  if (!s) return set_err(FG_ERR_INVALID, "null state");
  if (on && !s->t0) {
    FG_CUDA(cudaEventCreate(&s->t0));
    FG_CUDA(cudaEventCreate(&s->t1));
  }
  s->timing = on != 0;
  s->main_ms_sum = 0.0; s->main_launches = 0; s->t_pending = false;
  return FG_OK;
}

extern "C" int fg_state_get_timing(fg_state* s, double* total_ms, long long* launches) {
  if (!s) return set_err(FG_ERR_INVALID, "null state");
  if (s->t_pending) {
    FG_CUDA(cudaEventSynchronize(s->t1));
    float ms = 0;
    FG_CUDA(cudaEventElapsedTime(&ms, s->t0, s->t1));
    s->main_ms_sum += ms; s->main_launches += 1; s->t_pending = false;
  }
  if (total_ms) *total_ms = s->main_ms_sum;
  if (launches) *launches = s->main_launches;
  return FG_OK;
}

extern "C" int fg_state_reset(fg_state* s, void* stream) {
  if (!s) return set_err(FG_ERR_INVALID, "null state");
  FG_CUDA(cudaMemsetAsync(s->d_part, 0, sizeof(double) * part_doubles(s->p1), (cudaStream_t)stream));
  if (s->rows_pending) {
    FG_CUDA(cudaMemsetAsync(s->d_rows_pending, 0, sizeof(double), (cudaStream_t)stream));
    s->rows_pending = false;
  }
  s->rows = 0; s->empty = true; s->nk = -1;
  s->perm.clear();
  return FG_OK;
}

// inverse of the state's column order: caller column -> state column
static std::vector<int> inverse_perm(const fg_state* s) {
  std::vector<int> inv(s->p);
  for (int i = 0; i < s->p; ++i) inv[s->perm.empty() ? i : s->perm[i]] = i;
  return inv;
}

extern "C" int fg_state_get_perm(const fg_state* s, int32_t* perm_out) {
  if (!s || !perm_out) return set_err(FG_ERR_INVALID, "fg_state_get_perm: null argument");
  for (int i = 0; i < s->p; ++i) perm_out[i] = s->perm.empty() ? i : s->perm[i];
  return FG_OK;
}

extern "C" int fg_state_set_perm(fg_state* s, const int32_t* perm) {
  if (!s || !perm) return set_err(FG_ERR_INVALID, "fg_state_set_perm: null argument");
  if (!s->empty) return set_err(FG_ERR_INVALID, "fg_state_set_perm: the state already holds rows");
  std::vector<char> seen(s->p, 0);
  bool ident = true;
  for (int i = 0; i < s->p; ++i) {
    if (perm[i] < 0 || perm[i] >= s->p || seen[perm[i]]) return set_err(FG_ERR_INVALID, "fg_state_set_perm: not a permutation");
    seen[perm[i]] = 1;
    ident = ident && perm[i] == i;
  }
  s->perm.clear();
  if (!ident) s->perm.assign(perm, perm + s->p);
  return FG_OK;
}

static int ensure_parts(fg_state* s, int64_t nparts) {
  if (nparts <= s->parts_cap) return FG_OK;
  if (s->d_parts) cudaFree(s->d_parts);
  s->d_parts = nullptr; s->parts_cap = 0;
  FG_CUDA(cudaMalloc(&s->d_parts, sizeof(double) * part_doubles(s->p1) * nparts));
  s->parts_cap = nparts;
  return FG_OK;
}
static int ensure_gR(fg_state* s, int64_t doubles) {
  if (doubles <= s->gR_cap) return FG_OK;
  if (s->d_gR) cudaFree(s->d_gR);
  s->d_gR = nullptr; s->gR_cap = 0;
  FG_CUDA(cudaMalloc(&s->d_gR, sizeof(double) * doubles));
  s->gR_cap = doubles;
  return FG_OK;
}


template <int NSLOT, int TR>
static int launch_merge_level(double* parts, int nparts, int stride, int p1, int max_smem, cudaStream_t st) {
  const int pairs = (nparts + 2 * stride - 1) / (2 * stride);
  const int ntri = p1 * (p1 + 1) / 2;
  const int xs_bytes = align16(2 * (TR + 2) * 8);
  int smem = xs_bytes + ntri * 8;
  int use_smem = 1;
  if (smem > max_smem) { smem = xs_bytes; use_smem = 0; }
  auto kern = merge_pairs_kernel<NSLOT, TR>;
  FG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<pairs, 32, smem, st>>>(parts, nparts, stride, p1, use_smem);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

static int merge_level(double* parts, int nparts, int stride, int p1, int max_smem, cudaStream_t st) {
  if (p1 > 128) {  // wide factors: the whole CTA folds one pair (fg_wide.cuh)
    cudaError_t e = wide_launch_merge_level(parts, nparts, stride, p1, st);
    if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "wide merge failed: %s", cudaGetErrorString(e));
    FG_LAUNCHED();
    return FG_OK;
  }
  const int nslot = (p1 + 31) / 32;
  if (nslot <= 1) return launch_merge_level<1, 24>(parts, nparts, stride, p1, max_smem, st);
  if (nslot <= 2) return launch_merge_level<2, 24>(parts, nparts, stride, p1, max_smem, st);
  if (nslot <= 3) return launch_merge_level<3, 24>(parts, nparts, stride, p1, max_smem, st);
  if (nslot <= 5) return launch_merge_level<5, 16>(parts, nparts, stride, p1, max_smem, st);
  if (nslot <= 8) return launch_merge_level<8, 12>(parts, nparts, stride, p1, max_smem, st);
  if (nslot <= 11) return launch_merge_level<11, 8>(parts, nparts, stride, p1, max_smem, st);
  return launch_merge_level<15, 8>(parts, nparts, stride, p1, max_smem, st);
}

// Tree-merge parts[0..nparts) (slot `nparts` may hold the state's current factor) into s->d_part.
static int merge_parts_into_state(fg_state* s, int nparts, int max_smem, cudaStream_t st) {
  const int64_t pd = part_doubles(s->p1);
  int total = nparts;
  if (!s->empty) {  // the existing factor is one more part (placed last: deterministic order)
    FG_CUDA(cudaMemcpyAsync(s->d_parts + pd * nparts, s->d_part, sizeof(double) * pd, cudaMemcpyDeviceToDevice, st));
    total = nparts + 1;
  }
  for (int stride = 1; stride < total; stride *= 2) {
    int rc = merge_level(s->d_parts, total, stride, s->p1, max_smem, st);
    if (rc) return rc;
  }
  FG_CUDA(cudaMemcpyAsync(s->d_part, s->d_parts, sizeof(double) * pd, cudaMemcpyDeviceToDevice, st));
  s->empty = false;
  s->nk = -1;
  return FG_OK;
}

static int device_limits(int* sm_count, int* max_smem) {
  int dev = 0;
  FG_CUDA(cudaGetDevice(&dev));
  FG_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  FG_CUDA(cudaDeviceGetAttribute(max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return FG_OK;
}

// ---- dense accumulate
template <int NSLOT, int TR>
static int launch_dense(fg_state* s, const double* W, int64_t ldw, const double* tau, int64_t rows, cudaStream_t st) {
  int sm_count, max_smem;
  int rc = device_limits(&sm_count, &max_smem);
  if (rc) return rc;
  const int p1 = s->p1, ntri = p1 * (p1 + 1) / 2;
  const int xs_bytes = align16(2 * (TR + 2) * 8);
  const int static_smem = 8 * 32 * NSLOT * 8 + 64;
  int nwarps = 8;
  bool in_smem = true;
  while (nwarps > 1 && nwarps * (xs_bytes + align16(ntri * 8)) + static_smem > max_smem) --nwarps;
  if (nwarps * (xs_bytes + align16(ntri * 8)) + static_smem > max_smem || nwarps < 4) { in_smem = false; nwarps = 8; }
  const int64_t ntiles = (rows + TR - 1) / TR;
  int grid = (int)std::min<int64_t>((ntiles + nwarps - 1) / nwarps, sm_count);
  if (grid < 1) grid = 1;
  if (!in_smem) {  // factors live in HBM/L2: keep the footprint well inside L2
    const int64_t cap = std::max<int64_t>(1, (64ll << 20) / ((int64_t)ntri * 8 * nwarps));
    grid = (int)std::min<int64_t>(grid, cap);
    rc = ensure_gR(s, (int64_t)grid * nwarps * ntri);
    if (rc) return rc;
  }
  rc = ensure_parts(s, grid + 1);
  if (rc) return rc;
  const int smem = in_smem ? nwarps * (xs_bytes + align16(ntri * 8)) : nwarps * xs_bytes;
  auto kern = accumulate_dense_kernel<NSLOT, TR>;
  FG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rc = timing_begin(s, st);
  if (rc) return rc;
  kern<<<grid, nwarps * 32, smem, st>>>(W, ldw, tau, rows, p1, s->d_parts, in_smem ? nullptr : s->d_gR);
  rc = timing_end(s, st);
  if (rc) return rc;
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  rc = merge_parts_into_state(s, grid, max_smem, st);
  if (rc) return rc;
  s->rows += rows;
  return FG_OK;
}

extern "C" int fg_state_accumulate_dense(fg_state* s, const double* W, int64_t ldw, const double* tau, int64_t rows,
                                         void* stream) {
  if (!s || (!W && rows > 0)) return set_err(FG_ERR_INVALID, "fg_state_accumulate_dense: null argument");
  if (rows <= 0) return FG_OK;
  if (ldw < s->p) return set_err(FG_ERR_INVALID, "ldw=%lld smaller than p=%d", (long long)ldw, s->p);
  if (!s->perm.empty())
    return set_err(FG_ERR_UNSUPPORTED, "fg_state_accumulate_dense: the state holds a permuted factor (fused samples); use a fresh state");
  cudaStream_t st = (cudaStream_t)stream;
  const int nslot = (s->p1 + 31) / 32;
  if (s->p1 > 96) {  // wide: cooperative blocked TSQR (fg_coop.cuh); beyond its width the register tile (fg_wide.cuh)
    int sm_count, max_smem;
    int rc = device_limits(&sm_count, &max_smem);
    if (rc) return rc;
    const char* ecq = getenv("FG_USE_CQ");
    if (!(ecq && ecq[0] == '0') && s->p1 <= cq_max_p1(max_smem)) {
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + 63) / 64, sm_count));
      rc = ensure_parts(s, grid + 1);
      if (rc) return rc;
      rc = ensure_gR(s, (int64_t)grid * cq_r_doubles(s->p1));
      if (rc) return rc;
      rc = timing_begin(s, st);
      if (rc) return rc;
      CqArgs ca;
      ca.W = W; ca.ldw = ldw; ca.tau = tau; ca.rows = rows; ca.p1 = s->p1;
      cudaError_t e = cq_launch_dense(ca, s->d_gR, grid, max_smem, st);
      rc = timing_end(s, st);
      if (rc) return rc;
      if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "cooperative dense kernel failed: %s", cudaGetErrorString(e));
      FG_LAUNCHED();
      int nl = 0;
      e = cq_reduce_export(s->d_gR, grid, s->p1, s->d_parts, max_smem, &nl, st);
      if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "cooperative merge failed: %s", cudaGetErrorString(e));
      for (int i = 0; i < nl; ++i) FG_LAUNCHED();
      rc = merge_parts_into_state(s, 1, max_smem, st);
      if (rc) return rc;
      s->rows += rows;
      return FG_OK;
    }
    const int tr = wide_dense_tile_rows(s->p1);
    if (tr <= 0) return set_err(FG_ERR_UNSUPPORTED, "fg_state_accumulate_dense: %d columns are too many", s->p);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + tr - 1) / tr, sm_count));
    rc = ensure_parts(s, grid + 1);
    if (rc) return rc;
    rc = timing_begin(s, st);
    if (rc) return rc;
    cudaError_t e = wide_launch_dense(W, ldw, tau, 1, rows, s->p1, s->d_parts, grid, st);
    rc = timing_end(s, st);
    if (rc) return rc;
    if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "wide dense kernel failed: %s", cudaGetErrorString(e));
    FG_LAUNCHED();
    rc = merge_parts_into_state(s, grid, max_smem, st);
    if (rc) return rc;
    s->rows += rows;
    return FG_OK;
  }
  if (nslot <= 1) return launch_dense<1, 24>(s, W, ldw, tau, rows, st);
  if (nslot <= 2) return launch_dense<2, 24>(s, W, ldw, tau, rows, st);
  if (nslot <= 3) return launch_dense<3, 24>(s, W, ldw, tau, rows, st);
  if (nslot <= 5) return launch_dense<5, 16>(s, W, ldw, tau, rows, st);
  if (nslot <= 8) return launch_dense<8, 12>(s, W, ldw, tau, rows, st);
  if (nslot <= 11) return launch_dense<11, 8>(s, W, ldw, tau, rows, st);
  return launch_dense<15, 8>(s, W, ldw, tau, rows, st);
}

// fifth-generation fused kernel (fg_ringqr.cuh): producer warps write every row group's tile into an L2-resident
// ring, consumer warps (one TSQR stream per (warp, row group)) pull their blocks with bulk copies.  Factors and
// merge exactly as generation 4.
static int check_rq_error(fg_state* s) {
  if (s->h_rq_err && *s->h_rq_err != 0u) {
    const unsigned code = *s->h_rq_err;
    *s->h_rq_err = 0u;
    if (s->d_ring) cudaMemset(s->d_ring, 0, sizeof(double) * s->ring_cap);  // also clears the device flag
    return set_err(FG_ERR_CUDA, "fused kernel: ring hand-over timed out (code %u: 1 consumer, 2 producer, 3 mbarrier)", code);
  }
  return FG_OK;
}

static int rq_upload_tables(DevPlan* pl) {
  RqLaunch& rl = pl->rql;
  if (pl->d_rq_tab) { cudaFree(pl->d_rq_tab); pl->d_rq_tab = nullptr; }
  const size_t nb_i = rl.tab_i.size() * sizeof(unsigned short), nb_g = (rl.tab_g.size() + 15) / 16 * 16;
  FG_CUDA(cudaMalloc(&pl->d_rq_tab, nb_g + nb_i));
  FG_CUDA(cudaMemcpy(pl->d_rq_tab, rl.tab_g.data(), rl.tab_g.size(), cudaMemcpyHostToDevice));
  FG_CUDA(cudaMemcpy(static_cast<char*>(pl->d_rq_tab) + nb_g, rl.tab_i.data(), nb_i, cudaMemcpyHostToDevice));
  rl.d_tab_g = static_cast<const unsigned char*>(pl->d_rq_tab);
  rl.d_tab_i = reinterpret_cast<const unsigned short*>(static_cast<char*>(pl->d_rq_tab) + nb_g);
  return FG_OK;
}

static int launch_rq(fg_state* s, fg_model* m, DevPlan* pl, const double* q, const double* v, const double* a,
                     const double* tau, int64_t ld_tau, int64_t N, int phase, cudaStream_t st);

// The row groups advance in lockstep (a ring of a few hundred tiles couples them) and the producers feed all of them,
// so the launch runs at the pace of whichever role has too small a share of the SMs: a cost model that is 10 % off
// for one row group costs 10 %.  The first large launch of a plan therefore calibrates the split on the caller's
// first samples: timed trial launches walk the number of producer SMs downhill from the model's guess (TIAGo: 22
// instead of 26, -2 % kernel time; UR10: the model's 26 confirmed).  A launch of the profiling instantiation of the
// kernel also measures the warp cycles of a fold per row group and of a produced tile (diagnosis).  One-off per plan,
// ~10 ms; FG_RQ_CALIBRATE=0 keeps the instruction-count model, FG_RQ_CALIBRATE=2 prints what was measured.
// pseudo-random doubles in (-3, 3) (calibration inputs only)
__global__ void synth_inputs_kernel(double* x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    x[i] = ((double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 6.0;
  }
}

static int rq_calibrate(fg_state* s, fg_model* m, DevPlan* pl, const double* q, const double* v, const double* a,
                        const double* tau, int64_t ld_tau, int64_t N, cudaStream_t st) {
  const char* e = getenv("FG_RQ_CALIBRATE");
  const int mode = e && *e ? atoi(e) : 1;
  pl->rq_calibrated = true;
  if (mode == 0 || getenv("FG_RQ_NPCTA") || getenv("FG_RQ_COST")) return FG_OK;
  int64_t ncal = std::min<int64_t>(N, 1 << 18);
  // a short first chunk (host-buffer path) cannot carry timed trials: synthetic inputs then (what a tile costs does not
  // depend on the values)
  double* synth = nullptr;
  if (N < (1 << 17)) {
    ncal = 1 << 18;
    const int64_t per = m->nq + 2 * m->nv + pl->P.ms;
    FG_CUDA(cudaMalloc(&synth, sizeof(double) * ncal * per));
    synth_inputs_kernel<<<1024, 256, 0, st>>>(synth, ncal * per);
    q = synth; v = q + ncal * m->nq; a = v + ncal * m->nv; tau = a + ncal * m->nv; ld_tau = ncal;
  }
  unsigned long long* d_prof = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  FG_CUDA(cudaMalloc(&d_prof, sizeof(unsigned long long) * RQ_PROF_WORDS));
  FG_CUDA(cudaEventCreate(&ev0));
  FG_CUDA(cudaEventCreate(&ev1));
  int rc = FG_OK;
  // one launch of the current pl->rql / pl->rq over the first ncal samples; ms < 0 on failure
  auto trial = [&](unsigned long long* prof, float* ms) -> int {
    const RqLaunch& rl = pl->rql;
    const RqParams& prm = *pl->rq;
    const int64_t cnt_doubles = ((int64_t)GQ_MAXG * prm.gq.maxinst + 3) / 2 + 1;
    int r = ensure_gR(s, rl.r_doubles + rl.rfull_doubles + cnt_doubles);
    if (r) return r;
    const int64_t ring_need = rl.ring_doubles + (2 * (int64_t)prm.ring + 4 + 1) / 2;
    if (ring_need > s->ring_cap) {
      if (s->d_ring) cudaFree(s->d_ring);
      s->d_ring = nullptr; s->ring_cap = 0; s->ring_plan = nullptr;
      if (cudaMalloc(&s->d_ring, sizeof(double) * ring_need) != cudaSuccess) return set_err(FG_ERR_CUDA, "ring allocation failed");
      s->ring_cap = ring_need;
    }
    if (s->ring_plan != (const void*)pl || s->ring_slots != prm.ring) {
      cudaMemsetAsync(s->d_ring, 0, sizeof(double) * s->ring_cap, st);
      s->ring_plan = (const void*)pl;
      s->ring_slots = prm.ring;
    }
    unsigned* ctl = reinterpret_cast<unsigned*>(s->d_ring + rl.ring_doubles);
    if (prof) cudaMemsetAsync(prof, 0, sizeof(unsigned long long) * RQ_PROF_WORDS, st);
    cudaEventRecord(ev0, st);
    cudaError_t ce = rq_launch(rl, prm, q, v, a, tau, ld_tau, ncal, s->d_gR, s->d_ring, ctl, 1, st, prof);
    cudaEventRecord(ev1, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return set_err(FG_ERR_CUDA, "calibration launch failed: %s", cudaGetErrorString(ce));
    if (ms) cudaEventElapsedTime(ms, ev0, ev1);
    return FG_OK;
  };
  auto rebuild = [&](const RqCost& mc) -> bool {
    std::unique_ptr<RqParams> np(new RqParams);
    const char* why = "";
    RqLaunch nl = rq_build(pl->P, m->joints.data(), pl->rows.data(), pl->cols.data(), m->sm_count, m->max_smem, np.get(),
                           &why, &mc);
    if (!nl.ok) return false;
    pl->rql = std::move(nl);
    pl->rq = std::move(np);
    return rq_upload_tables(pl) == FG_OK;
  };
  do {
    // (1) fold cycles per row group
    rc = trial(d_prof, nullptr);
    if (rc) break;
    unsigned long long h[RQ_PROF_WORDS];
    if (cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = set_err(FG_ERR_CUDA, "calibration read-back failed"); break; }
    const int ms_ = pl->rq->gq.ms;
    RqCost mc;
    mc.have = h[25] > 0;
    for (int g = 0; g < ms_; ++g) {
      if (h[8 + g] == 0) mc.have = false;
      else mc.grp[g] = (double)h[g] / (double)h[8 + g];
    }
    if (!mc.have) break;
    mc.prod = (double)h[24] / (double)h[25];
    if (mode >= 2) {
      fprintf(stderr, "[fg] rq calibration (%lld samples, %d producer SMs, ring %d slots of %d KB):", (long long)ncal, pl->rq->npc,
              pl->rq->ring, pl->rq->slot_doubles / 128);
      for (int g = 0; g < ms_; ++g)
        fprintf(stderr, " g%d[%d warps] fold %.0f wait %.0f |", g, pl->rq->gq.gen.grp[g].ninst, mc.grp[g],
                (double)h[16 + g] / (double)h[8 + g]);
      fprintf(stderr, " producer tile %.0f wait %.0f\n", mc.prod, (double)h[26] / (double)h[25]);
    }
    // The measured fold cycles are not reliably the better apportionment of the consumer warps (a fold's cycles depend
    // on what its neighbours on the SM are doing: UR10 3.53 ms against 3.45 ms with the instruction-count model, TIAGo
    // 10.73 ms against 11.0 ms), so both are TIMED below and the faster one is kept (mode 3 forces the measured one).
    const bool have_meas = mc.have;
    if (mode < 3) mc.have = false;
    // (2) producer SMs: timed trials, downhill from the model's count
    const int start = pl->rq->npc;
    int best_n = start;
    float best_ms = 0.f;
    auto eval = [&](int n, float* ms) -> bool {
      if (n < 1 || n > m->sm_count - 1) return false;
      mc.force_npcta = n;
      if (!rebuild(mc)) return false;
      float t1 = 0.f, t2 = 0.f, t3 = 0.f;   // the fastest of three: a trial is ~1 ms and one outlier must not decide
      if (trial(nullptr, &t1) || trial(nullptr, &t2) || trial(nullptr, &t3)) return false;
      *ms = std::min(t1, std::min(t2, t3));
      if (mode >= 2) fprintf(stderr, "[fg]   %d producer SMs: %.4f ms\n", n, *ms);
      return true;
    };
    if (!eval(start, &best_ms)) { rc = check_rq_error(s); break; }
    for (int dir = -1; dir <= 1; dir += 2) {
      bool moved = false;
      for (int n = best_n + 2 * dir;; n += 2 * dir) {
        float t = 0.f;
        if (!eval(n, &t) || t >= 0.998f * best_ms) break;   // a little hysteresis: the model's count is tuned
        best_ms = t; best_n = n; moved = true;
      }
      if (moved) break;
    }
    for (int dir = -1; dir <= 1; dir += 2) {  // odd neighbours
      float t = 0.f;
      if (eval(best_n + dir, &t) && t < 0.998f * best_ms) { best_ms = t; best_n += dir; break; }
    }
    bool use_meas = mc.have;
    if (have_meas && !mc.have) {  // (3) row groups by measured fold cycles at the chosen producer count: kept if >= 0.5 % faster
      mc.have = true;
      float t = 0.f;
      if (eval(best_n, &t) && t < 0.995f * best_ms) {
        best_ms = t; use_meas = true;
        for (int dir = -1; dir <= 1; dir += 2) {  // the producer count once more around the new apportionment
          float t2 = 0.f;
          if (eval(best_n + dir, &t2) && t2 < best_ms) { best_ms = t2; best_n += dir; break; }
        }
      }
      mc.have = use_meas;
    }
    mc.force_npcta = best_n;
    if (!rebuild(mc)) rc = set_err(FG_ERR_INVALID, "internal: calibrated split could not be rebuilt");
    if (mode >= 2)
      fprintf(stderr, "[fg] rq calibration: %d producer SMs (model %d), row groups by %s\n", best_n, start,
              use_meas ? "measured fold cycles" : "the instruction-count model");
  } while (false);
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaFree(d_prof);
  if (synth) { cudaStreamSynchronize(st); cudaFree(synth); }
  if (rc == FG_OK) rc = check_rq_error(s);
  return rc;
}

static int launch_rq(fg_state* s, fg_model* m, DevPlan* pl, const double* q, const double* v, const double* a,
                     const double* tau, int64_t ld_tau, int64_t N, int phase, cudaStream_t st) {
  if ((phase & FG_PHASE_FIRST) && !pl->rq_calibrated && std::max<int64_t>(N, s->hint_total) >= (1 << 17)) {
    if (!pl->d_rq_tab) { int rc0 = rq_upload_tables(pl); if (rc0) return rc0; }
    int rc0 = rq_calibrate(s, m, pl, q, v, a, tau, ld_tau, N, st);
    if (rc0) return rc0;
  }
  RqLaunch& rl = pl->rql;
  const RqParams& prm = *pl->rq;
  if (!pl->d_rq_tab) {
    const size_t nb_i = rl.tab_i.size() * sizeof(unsigned short), nb_g = (rl.tab_g.size() + 15) / 16 * 16;
    FG_CUDA(cudaMalloc(&pl->d_rq_tab, nb_g + nb_i));
    FG_CUDA(cudaMemcpy(pl->d_rq_tab, rl.tab_g.data(), rl.tab_g.size(), cudaMemcpyHostToDevice));
    FG_CUDA(cudaMemcpy(static_cast<char*>(pl->d_rq_tab) + nb_g, rl.tab_i.data(), nb_i, cudaMemcpyHostToDevice));
    rl.d_tab_g = static_cast<const unsigned char*>(pl->d_rq_tab);
    rl.d_tab_i = reinterpret_cast<const unsigned short*>(static_cast<char*>(pl->d_rq_tab) + nb_g);
  }
  const bool whole = phase == (FG_PHASE_FIRST | FG_PHASE_LAST);
  const int64_t cnt_doubles = ((int64_t)GQ_MAXG * prm.gq.maxinst + 3) / 2 + 1;
  const int64_t need = rl.r_doubles + rl.rfull_doubles + cnt_doubles;
  int rc;
  if (phase & FG_PHASE_FIRST) {
    rc = ensure_gR(s, need);
    if (rc) return rc;
  } else if (s->gR_cap < need) {
    return set_err(FG_ERR_INVALID, "internal: continued launch without a first one");
  }
  const int64_t ctl_doubles = (2 * (int64_t)prm.ring + 4 + 1) / 2;
  const int64_t ring_need = rl.ring_doubles + ctl_doubles;
  if (ring_need > s->ring_cap) {
    if (s->d_ring) cudaFree(s->d_ring);
    s->d_ring = nullptr; s->ring_cap = 0; s->ring_plan = nullptr;
    FG_CUDA(cudaMalloc(&s->d_ring, sizeof(double) * ring_need));
    s->ring_cap = ring_need;
  }
  // the padding columns of a slot are never written: they must read as zero; the control words (error flag) sit behind the
  // last slot, so a different slot count (the calibration changes the producer count) also means a fresh ring
  if (s->ring_plan != (const void*)pl || s->ring_slots != prm.ring) {
    FG_CUDA(cudaMemsetAsync(s->d_ring, 0, sizeof(double) * s->ring_cap, st));
    s->ring_plan = (const void*)pl;
    s->ring_slots = prm.ring;
  }
  if (!s->h_rq_err) {
    FG_CUDA(cudaMallocHost(&s->h_rq_err, sizeof(unsigned)));
    *s->h_rq_err = 0u;
  }
  rc = ensure_parts(s, 2);
  if (rc) return rc;
  unsigned* ctl = reinterpret_cast<unsigned*>(s->d_ring + rl.ring_doubles);
  rc = timing_begin(s, st);
  if (rc) return rc;
  cudaError_t e = rq_launch(rl, prm, q, v, a, tau, ld_tau, N, s->d_gR, s->d_ring, ctl, (phase & FG_PHASE_FIRST) ? 1 : 0, st);
  rc = timing_end(s, st);
  if (rc) return rc;
  if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "rq_launch failed: %s", cudaGetErrorString(e));
  FG_LAUNCHED();
  s->rows += N * pl->P.ms;
  // the flag of every launch is OR-ed into the pinned word by the copy of the last one; a timeout aborts all warps
  FG_CUDA(cudaMemcpyAsync(s->h_rq_err, ctl + 2 * (size_t)prm.ring, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  if (!(phase & FG_PHASE_LAST)) return FG_OK;
  double* rfull = s->d_gR + rl.r_doubles;
  unsigned* counters = reinterpret_cast<unsigned*>(rfull + rl.rfull_doubles);
  e = gq_merge(rl.smem_merge, prm.gq, s->d_gR, whole ? (N + GQ_S - 1) / GQ_S : 0, counters, rfull, s->d_parts, st);
  if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "gq_merge failed: %s", cudaGetErrorString(e));
  FG_LAUNCHED();
  return merge_parts_into_state(s, 1, m->max_smem, st);
}

static int build_cq_groups(DevPlan* pl) {
  const int ms = pl->P.ms, ncols = pl->P.ncols;
  std::vector<std::vector<char>> mask(ms, std::vector<char>(ncols, 1));
  bool split_ok = ms <= 32 && getenv("FG_CQ_NOSPLIT") == nullptr;
  for (int g = 0; g < ms && split_ok; ++g) {
    if (pl->rows[g].rowblock != g) split_ok = false;  // tau rows and W row blocks must line up
    const double* S = pl->rows[g].S;
    const bool no_ang = S[3] == 0.0 && S[4] == 0.0 && S[5] == 0.0;
    for (int c = 0; c < ncols; ++c) {
      const ColDesc& cd = pl->cols[c];
      bool nz = true;
      if (cd.kind == COL_INERTIAL) nz = ((col_ancmask(cd) >> g) & 1ull) && !(cd.k >= 4 && no_ang);
      else if (cd.kind == COL_EXTRA) nz = cd.grp == -1 || cd.grp == g;
      else if (cd.kind == COL_COUPLING) nz = pl->rows[g].rowblock == 4 || pl->rows[g].rowblock == 5;
      mask[g][c] = nz ? 1 : 0;
    }
  }
  std::vector<int> grp_of(ms, 0);
  std::vector<int> rep;  // representative row block of every group
  if (split_ok) {
    for (int g = 0; g < ms; ++g) {
      int k = 0;
      for (; k < (int)rep.size(); ++k)
        if (mask[rep[k]] == mask[g]) break;
      if (k == (int)rep.size()) rep.push_back(g);
      grp_of[g] = k;
    }
    // group 0 must see every column; a narrower group must save at least a quarter of the columns to pay for itself
    int full = -1;
    for (int k = 0; k < (int)rep.size(); ++k)
      if (std::count(mask[rep[k]].begin(), mask[rep[k]].end(), 1) == ncols) full = k;
    if (full < 0) split_ok = false;
    else {
      for (int k = 0; k < (int)rep.size(); ++k)
        if (k != full && std::count(mask[rep[k]].begin(), mask[rep[k]].end(), 1) * 4 > ncols * 3)
          for (int g = 0; g < ms; ++g)
            if (grp_of[g] == k) grp_of[g] = full;
      std::vector<int> order{full};
      for (int k = 0; k < (int)rep.size(); ++k)
        if (k != full && std::count(grp_of.begin(), grp_of.end(), k) > 0) order.push_back(k);
      std::vector<int> renum(rep.size(), -1);
      for (size_t i = 0; i < order.size(); ++i) renum[order[i]] = (int)i;
      std::vector<int> nrep(order.size());
      for (size_t i = 0; i < order.size(); ++i) nrep[i] = rep[order[i]];
      for (int g = 0; g < ms; ++g) grp_of[g] = renum[grp_of[g]];
      rep = nrep;
    }
  }
  if (!split_ok) { rep.assign(1, 0); std::fill(grp_of.begin(), grp_of.end(), 0); for (auto& v : mask[0]) v = 1; }
  pl->cqg.assign(rep.size(), CqGroup());
  for (size_t k = 0; k < rep.size(); ++k) {
    CqGroup& G = pl->cqg[k];
    for (int g = 0; g < ms; ++g)
      if (grp_of[g] == (int)k) G.blocks.push_back(pl->rows[g].rowblock);
    std::sort(G.blocks.begin(), G.blocks.end());
    std::vector<int> colsel, colmap;
    for (int c = 0; c < ncols; ++c)
      if (k == 0 || mask[rep[k]][c]) { colsel.push_back(pl->fullcol[c]); colmap.push_back(c); }
    colsel.push_back(0);
    colmap.push_back(ncols);  // rhs
    G.p1 = (int)colsel.size();
    FG_CUDA(cudaMalloc(&G.d_colsel, sizeof(int) * colsel.size()));
    FG_CUDA(cudaMalloc(&G.d_colmap, sizeof(int) * colmap.size()));
    FG_CUDA(cudaMemcpy(G.d_colsel, colsel.data(), sizeof(int) * colsel.size(), cudaMemcpyHostToDevice));
    FG_CUDA(cudaMemcpy(G.d_colmap, colmap.data(), sizeof(int) * colmap.size(), cudaMemcpyHostToDevice));
  }
  return FG_OK;
}

static int accumulate_samples_dev(fg_state* s, fg_model* m, const fg_regressor_cfg* cfg, const double* q,
                                  const double* v, const double* a, const double* tau, int64_t ld_tau, int64_t N,
                                  cudaStream_t st, int phase = FG_PHASE_FIRST | FG_PHASE_LAST, bool* chunkable = nullptr) {
  DevPlan* pl = nullptr;
  int rc = get_plan(m, cfg, cfg->col_sel, cfg->n_col_sel, &pl);
  if (rc) return rc;
  if (pl->P.ncols != s->p)
    return set_err(FG_ERR_INVALID, "state has p=%d but the regressor has %d columns", s->p, pl->P.ncols);
  if (N <= 0) return FG_OK;
  if (tau && ld_tau < N) return set_err(FG_ERR_INVALID, "ld_tau=%lld smaller than N=%lld", (long long)ld_tau, (long long)N);
  const int ms = pl->P.ms;
  {
    const char* e = getenv("FG_USE_RQ");
    const bool use_rq = !(e && e[0] == '0');
    if (use_rq && !pl->rq_tried) {
      pl->rq_tried = true;
      pl->rq.reset(new RqParams);
      const char* why = "";
      pl->rql = rq_build(pl->P, m->joints.data(), pl->rows.data(), pl->cols.data(), m->sm_count, m->max_smem,
                         pl->rq.get(), &why);
      if (!pl->rql.ok) pl->rq.reset();
    }
    if (use_rq && pl->rql.ok && s->perm.empty()) {
      if (chunkable) *chunkable = true;
      return launch_rq(s, m, pl, q, v, a, tau, ld_tau, N, phase, st);
    }
  }
  {  // models the row-group streams do not take (Talos with free-flyer: 6 dense rows x 330 columns per sample): W is
     // materialised a chunk of samples at a time and folded by the cooperative blocked TSQR (fg_coop.cuh)
    const char* e = getenv("FG_USE_CQ");
    const bool use_cq = !(e && e[0] == '0');
    if (use_cq && s->perm.empty() && pl->P.p1 > 96 && pl->P.p1 <= cq_max_p1(m->max_smem)) {
      fg_regressor_cfg cfull = *cfg;
      cfull.n_col_sel = 0; cfull.col_sel = nullptr;
      int ncols_full = 0;
      rc = materialize_dev(m, &cfull, nullptr, nullptr, nullptr, 0, nullptr, 0, st, &ncols_full);
      if (rc) return rc;
      // the materialised matrix has every row block of the mode; the plan may select some of them
      int ms_full = 0, cols_chk = 0;
      rc = fg_regressor_shape(m, &cfull, &ms_full, &cols_chk);
      if (rc) return rc;
      if (ms_full == ms) {
        // Row blocks that see the same set of columns form a group with its own (narrower) problem: the three force
        // rows of a free-flyer never see the inertia-tensor columns (133 of Talos' 331 columns).  The widest group
        // must see every column; the others are folded into its factor at the end.
        if (!pl->cqg_built) {
          pl->cqg_built = true;
          rc = build_cq_groups(pl);
          if (rc) return rc;
        }
        const std::vector<CqGroup>& G = pl->cqg;
        const int ng = (int)G.size();
        const int64_t chunk = std::max<int64_t>(32, std::min<int64_t>(N, (int64_t)(768ll << 20) / ((int64_t)ms * ncols_full * 8)));
        const int64_t w_doubles = chunk * ms * ncols_full;
        if (w_doubles > s->ring_cap) {
          if (s->d_ring) cudaFree(s->d_ring);
          s->d_ring = nullptr; s->ring_cap = 0;
          FG_CUDA(cudaMalloc(&s->d_ring, sizeof(double) * w_doubles));
          s->ring_cap = w_doubles;
        }
        s->ring_plan = nullptr;
        // a sequence of calls (FIRST ... LAST: the chunks of the host-buffer path) shares the per-CTA factors
        std::vector<int64_t> roff(ng + 1, 0);
        if (phase & FG_PHASE_FIRST)
          s->cq_grid = (int)std::max<int64_t>(1, std::min<int64_t>((std::min(chunk, N) + 63) / 64, m->sm_count));
        if (s->cq_grid <= 0) return set_err(FG_ERR_INVALID, "internal: continued launch without a first one");
        const int grid = s->cq_grid;
        for (int k = 0; k < ng; ++k) roff[k + 1] = roff[k] + (int64_t)grid * cq_r_doubles(G[k].p1);
        if (phase & FG_PHASE_FIRST) {
          rc = ensure_gR(s, roff[ng]);
          if (rc) return rc;
        } else if (s->gR_cap < roff[ng]) {
          return set_err(FG_ERR_INVALID, "internal: continued launch without a first one");
        }
        if (chunkable) *chunkable = true;
        rc = ensure_parts(s, 2);
        if (rc) return rc;
        rc = timing_begin(s, st);
        if (rc) return rc;
        for (int64_t i0 = 0; i0 < N; i0 += chunk) {
          const int64_t n = std::min(chunk, N - i0);
          rc = materialize_dev(m, &cfull, q + i0 * m->nq, v + i0 * m->nv, a + i0 * m->nv, n, s->d_ring, ncols_full, st);
          if (rc) return rc;
          for (int k = 0; k < ng; ++k) {
            bool first_run = true;
            for (size_t r0 = 0; r0 < G[k].blocks.size();) {  // runs of consecutive row blocks
              size_t r1 = r0 + 1;
              while (r1 < G[k].blocks.size() && G[k].blocks[r1] == G[k].blocks[r1 - 1] + 1) ++r1;
              const int b0 = G[k].blocks[r0], nblk = (int)(r1 - r0);
              CqArgs ca;
              ca.W = s->d_ring + (int64_t)b0 * n * ncols_full; ca.ldw = ncols_full; ca.colsel = G[k].d_colsel;
              ca.tau = tau ? tau + (int64_t)b0 * ld_tau + i0 : nullptr; ca.tau_blk = n; ca.tau_ld = ld_tau;
              ca.rows = n * nblk; ca.p1 = G[k].p1;
              ca.first = (phase & FG_PHASE_FIRST) && i0 == 0 && first_run;
              cudaError_t ce = cq_launch_dense(ca, s->d_gR + roff[k], grid, m->max_smem, st);
              if (ce != cudaSuccess) return set_err(FG_ERR_CUDA, "cooperative dense kernel failed: %s", cudaGetErrorString(ce));
              FG_LAUNCHED();
              first_run = false;
              r0 = r1;
            }
          }
        }
        rc = timing_end(s, st);
        if (rc) return rc;
        s->rows += N * ms;
        if (!(phase & FG_PHASE_LAST)) return FG_OK;
        int nl = 0;
        cudaError_t ce = cq_reduce(s->d_gR + roff[0], grid, G[0].p1, m->max_smem, &nl, st);
        for (int k = 1; k < ng && ce == cudaSuccess; ++k) {  // narrower groups: tree, rows scattered to full width, folded
          ce = cq_reduce(s->d_gR + roff[k], grid, G[k].p1, m->max_smem, &nl, st);
          if (ce != cudaSuccess) break;
          const int64_t dd = (int64_t)G[k].p1 * G[0].p1;
          if (dd > s->ring_cap) { ce = cudaErrorMemoryAllocation; break; }
          ce = cudaMemsetAsync(s->d_ring, 0, sizeof(double) * dd, st);
          if (ce != cudaSuccess) break;
          ce = cq_scatter_rows(s->d_gR + roff[k], G[k].p1, G[k].d_colmap, G[0].p1, s->d_ring, st);
          if (ce != cudaSuccess) break;
          CqArgs ca;
          ca.W = s->d_ring; ca.ldw = G[0].p1; ca.rhs_in_w = true; ca.rows = G[k].p1; ca.p1 = G[0].p1; ca.first = false;
          ce = cq_launch_dense(ca, s->d_gR + roff[0], 1, m->max_smem, st);
          nl += 2;
        }
        if (ce == cudaSuccess) { ce = cq_export(s->d_gR + roff[0], G[0].p1, s->d_parts, st); ++nl; }
        if (ce != cudaSuccess) return set_err(FG_ERR_CUDA, "cooperative merge failed: %s", cudaGetErrorString(ce));
        for (int i = 0; i < nl; ++i) FG_LAUNCHED();
        return merge_parts_into_state(s, 1, m->max_smem, st);
      }
    }
  }
  {
    const char* e = getenv("FG_USE_WIDE");
    const bool use_wide = !(e && e[0] == '0');
    const int spt = wide_samples_per_tile(ms, pl->P.p1);
    if (use_wide && spt > 0 && s->perm.empty()) {
      const int64_t ntiles = (N + spt - 1) / spt;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)m->sm_count * wide_ctas_per_sm(ms, pl->P.p1)));
      rc = ensure_parts(s, grid + 1);
      if (rc) return rc;
      rc = timing_begin(s, st);
      if (rc) return rc;
      // forward kinematics of a chunk of samples live in a scratch buffer (shared with the ring: never both in use)
      const int64_t fw_chunk = std::min<int64_t>(N, 1 << 16);
      const int64_t fw_need = fw_chunk * wide_fw_doubles_per_sample(pl->P);
      if (fw_need > s->ring_cap) {
        if (s->d_ring) cudaFree(s->d_ring);
        s->d_ring = nullptr; s->ring_cap = 0;
        FG_CUDA(cudaMalloc(&s->d_ring, sizeof(double) * fw_need));
        s->ring_cap = fw_need;
      }
      s->ring_plan = nullptr;  // whatever tile layout the ring held is gone
      int nl = 0;
      cudaError_t ce = wide_launch_samples(pl->P, pl->ptrs, q, v, a, tau, ld_tau, N, s->d_parts, grid, m->max_smem,
                                           s->d_ring, fw_chunk, &nl, st);
      rc = timing_end(s, st);
      if (rc) return rc;
      if (ce != cudaSuccess) return set_err(FG_ERR_CUDA, "wide fused kernel failed: %s", cudaGetErrorString(ce));
      for (int i = 0; i < nl; ++i) FG_LAUNCHED();
      rc = merge_parts_into_state(s, grid, m->max_smem, st);
      if (rc) return rc;
      s->rows += N * ms;
      return FG_OK;
    }
  }
  return set_err(FG_ERR_UNSUPPORTED,
                 "no fused kernel for %d row blocks x %d columns; materialise W and use fg_state_accumulate_dense",
                 ms, pl->P.ncols);
}

extern "C" int fg_state_accumulate_samples(fg_state* s, const fg_model* m, const fg_regressor_cfg* cfg,
                                           const double* q, const double* v, const double* a, const double* tau,
                                           int64_t ld_tau, int64_t N, void* stream) {
  if (!s || !m || !cfg) return set_err(FG_ERR_INVALID, "fg_state_accumulate_samples: null argument");
  return accumulate_samples_dev(s, const_cast<fg_model*>(m), cfg, q, v, a, tau, ld_tau, N, (cudaStream_t)stream);
}

// ---- host-buffer path: chunked H2D staging overlapped with the fused kernel
extern "C" int fg_state_accumulate_samples_host(fg_state* s, const fg_model* mc, const fg_regressor_cfg* cfg,
                                                const double* q, const double* v, const double* a, const double* tau,
                                                int64_t ld_tau, int64_t N, void* stream) {
  if (!s || !mc || !cfg) return set_err(FG_ERR_INVALID, "fg_state_accumulate_samples_host: null argument");
  fg_model* m = const_cast<fg_model*>(mc);
  int32_t ms = 0, cols = 0;
  int rc = fg_regressor_shape(m, cfg, &ms, &cols);
  if (rc) return rc;
  if (N <= 0) return FG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // The copies are the bottleneck (PCIe: 192 B per UR10 sample against 12 B/ns of kernel appetite; tools/probe_h2d.py on
  // this pool, pinned memory: 1 MB copies 30 GB/s, 6 MB 40, 25 MB 42, 192 MB 55).  Chunk sizes ramp up at the start (the
  // first copy, which nothing can overlap, is short) and down at the end (so is the last kernel, which overlaps nothing).
  // Measured and rejected (UR10, 1 M samples, end to end 5.70 ms = 33.7 GB/s): chunks of 256 Ki / 512 Ki samples (6.2 /
  // 6.9 ms), the four arrays on two / four copy streams (6.2 / 6.1 ms); FG_H2D_CAP / FG_H2D_STREAMS keep the A/B.
  const char* ecap = getenv("FG_H2D_CAP");      // log2 of the largest chunk (A/B)
  const char* estr = getenv("FG_H2D_STREAMS");  // 1 .. 4 copy streams (A/B)
  const int64_t cap = 1ll << (ecap && *ecap ? std::max(15, std::min(22, atoi(ecap))) : 17);
  const int nstr = estr && *estr ? std::max(1, std::min(4, atoi(estr))) : 1;
  const int64_t chunk = std::min<int64_t>(N, cap);
  const int64_t per_sample = m->nq + 2 * m->nv + ms;
  const int64_t need = 2 * chunk * per_sample;
  if (need > s->stage_cap) {
    if (s->d_stage) cudaFree(s->d_stage);
    s->d_stage = nullptr; s->stage_cap = 0;
    FG_CUDA(cudaMalloc(&s->d_stage, sizeof(double) * need));
    s->stage_cap = need;
  }
  for (auto& cs : s->copy_streams) if (!cs) FG_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  for (auto& e : s->ev) if (!e) FG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : s->ev_copy) if (!e) FG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // ev_copy[4 b + k]: copy k into buffer b done; ev[2 + b]: compute on buffer b done
  FG_CUDA(cudaEventRecord(s->ev[2], st));
  FG_CUDA(cudaEventRecord(s->ev[3], st));
  std::vector<int64_t> sizes;
  {
    int64_t left = N;
    std::vector<int64_t> down;
    for (int64_t d = cap / 2; d >= (1 << 15) && left > 2 * d; d /= 2) { down.push_back(d); left -= d; }
    for (int64_t u = 1 << 15; u < cap && left > u; u *= 2) { sizes.push_back(u); left -= u; }
    if (left % cap) { sizes.push_back(left % cap); left -= left % cap; }
    for (; left > 0; left -= cap) sizes.push_back(cap);
    sizes.insert(sizes.end(), down.begin(), down.end());
  }
  s->hint_total = N;
  int b = 0;
  bool chunkable = false;  // set by the first chunk when a chunk-continuing kernel took it: later chunks continue its factors
  int64_t i0 = 0;
  for (size_t ci = 0; ci < sizes.size(); ++ci, b ^= 1) {
    const int64_t n = sizes[ci];
    int phase = FG_PHASE_FIRST | FG_PHASE_LAST;
    if (i0 == 0) phase = N > n ? FG_PHASE_FIRST : phase;
    else if (chunkable) phase = i0 + n >= N ? FG_PHASE_LAST : 0;
    double* base = s->d_stage + (int64_t)b * chunk * per_sample;
    double* dq = base;
    double* dv = dq + chunk * m->nq;
    double* da = dv + chunk * m->nv;
    double* dt = da + chunk * m->nv;
    for (int k = 0; k < 4; ++k) FG_CUDA(cudaStreamWaitEvent(s->copy_streams[k], s->ev[2 + b], 0));
    FG_CUDA(cudaMemcpyAsync(dq, q + i0 * m->nq, sizeof(double) * n * m->nq, cudaMemcpyHostToDevice, s->copy_streams[0]));
    FG_CUDA(cudaMemcpyAsync(dv, v + i0 * m->nv, sizeof(double) * n * m->nv, cudaMemcpyHostToDevice, s->copy_streams[1 % nstr]));
    FG_CUDA(cudaMemcpyAsync(da, a + i0 * m->nv, sizeof(double) * n * m->nv, cudaMemcpyHostToDevice, s->copy_streams[2 % nstr]));
    if (tau)
      FG_CUDA(cudaMemcpy2DAsync(dt, sizeof(double) * chunk, tau + i0, sizeof(double) * ld_tau, sizeof(double) * n, ms,
                                cudaMemcpyHostToDevice, s->copy_streams[3 % nstr]));
    for (int k = 0; k < 4; ++k) {
      FG_CUDA(cudaEventRecord(s->ev_copy[4 * b + k], s->copy_streams[k]));
      FG_CUDA(cudaStreamWaitEvent(st, s->ev_copy[4 * b + k], 0));
    }
    rc = accumulate_samples_dev(s, m, cfg, dq, dv, da, tau ? dt : nullptr, chunk, n, st, phase, i0 == 0 ? &chunkable : nullptr);
    if (rc) return rc;
    FG_CUDA(cudaEventRecord(s->ev[2 + b], st));
    i0 += n;
  }
  s->hint_total = 0;
  return FG_OK;
}

// ---- preprocessing ahead of the regressor
// scipy.signal.filtfilt(b, a, x, axis=0, padtype="odd", padlen) followed by dropping `trim` rows at both ends
// (low_pass_filter_data, identification_tools.py:390-424).  b, a, zi are HOST arrays (a[0] != 0; zi =
// scipy.signal.lfilter_zi(b, a), order entries); x and y are device matrices.
// Filtering core shared by the (b, a) and the second-order-section entries.  F: coefficients with F.n states and the
// scipy initial-state vector zi; A^L (the homogeneous part of the recurrence over one chunk) is obtained here by
// running the recurrence itself from every unit state, so it is exactly what the kernels iterate.
static int filtfilt_run(IirCoef& F, const double* x, int64_t N, int32_t ncols, int64_t ldx, int32_t padlen, int64_t trim,
                        int64_t step, double* y, int64_t ldy, void* stream) {
  const int n = F.n;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n; ++i0) {
    double z[IIR_MAXN] = {0};
    z[i0] = 1.0;
    for (int l = 0; l < IIR_L; ++l) {
      if (F.nsec > 0) {
        double u = 0.0;
        for (int s = 0; s < F.nsec; ++s) {
          const double ys = F.sos[s][0] * u + z[2 * s];
          z[2 * s] = F.sos[s][1] * u - F.sos[s][4] * ys + z[2 * s + 1];
          z[2 * s + 1] = F.sos[s][2] * u - F.sos[s][5] * ys;
          u = ys;
        }
      } else {
        const double yv = z[0];
        for (int i = 0; i < n; ++i) z[i] = -F.a[i + 1] * yv + (i + 1 < n ? z[i + 1] : 0.0);
      }
    }
    for (int i = 0; i < n; ++i) F.AL[i * n + i0] = z[i];
  }
  const int64_t M = N + 2 * (int64_t)padlen;
  const int64_t nchunks = (M + IIR_L - 1) / IIR_L;
  double *y1 = nullptr, *S = nullptr, *Z0 = nullptr;
  FG_CUDA(cudaMallocAsync(&y1, sizeof(double) * M * ncols, st));
  FG_CUDA(cudaMallocAsync(&S, sizeof(double) * nchunks * ncols * IIR_MAXN, st));
  FG_CUDA(cudaMallocAsync(&Z0, sizeof(double) * nchunks * ncols * IIR_MAXN, st));
  const int threads = 128;
  const int blocks = (int)((nchunks * ncols + threads - 1) / threads);
  for (int pass = 0; pass < 2; ++pass) {
    IirIo io;
    io.N = N; io.M = M; io.padlen = padlen; io.ncols = ncols; io.backward = pass; io.trim = trim; io.step = step;
    if (pass == 0) { io.x = x; io.ldx = ldx; io.y = y1; io.ldy = ncols; }
    else { io.x = y1; io.ldx = ncols; io.y = y; io.ldy = ldy; }
    iir_chunk_kernel<false><<<blocks, threads, 0, st>>>(F, io, S, nullptr);
    iir_propagate_kernel<<<(ncols + 63) / 64, 64, 0, st>>>(F, io, S, Z0);
    iir_chunk_kernel<true><<<blocks, threads, 0, st>>>(F, io, nullptr, Z0);
    FG_CUDA(cudaGetLastError());
    FG_LAUNCHED(); FG_LAUNCHED(); FG_LAUNCHED();
  }
  FG_CUDA(cudaFreeAsync(y1, st));
  FG_CUDA(cudaFreeAsync(S, st));
  FG_CUDA(cudaFreeAsync(Z0, st));
  return FG_OK;
}

static int filtfilt_impl(const double* x, int64_t N, int32_t ncols, int64_t ldx, const double* b, int32_t nb,
                         const double* a, int32_t na, const double* zi, int32_t padlen, int64_t trim, int64_t step,
                         double* y, int64_t ldy, void* stream) {
  if (!x || !y || !b || !a || !zi || N < 1 || ncols < 1 || ldx < ncols || ldy < ncols || nb < 1 || na < 1 || a[0] == 0.0 ||
      step < 1)
    return set_err(FG_ERR_INVALID, "fg_filtfilt: bad argument");
  const int n = std::max(nb, na) - 1;
  if (n < 1 || n > IIR_MAXN) return set_err(FG_ERR_UNSUPPORTED, "fg_filtfilt: filter order %d outside [1,%d]", n, IIR_MAXN);
  if (padlen < 0 || padlen >= N) return set_err(FG_ERR_INVALID, "fg_filtfilt: padlen=%d needs more than %d samples", padlen, padlen);
  if (trim < 0 || 2 * trim >= N) return set_err(FG_ERR_INVALID, "fg_filtfilt: trim=%lld leaves no row", (long long)trim);
  IirCoef F;
  memset(&F, 0, sizeof F);
  F.n = n;
  for (int i = 0; i <= n; ++i) { F.b[i] = (i < nb ? b[i] : 0.0) / a[0]; F.a[i] = (i < na ? a[i] : 0.0) / a[0]; }
  for (int i = 0; i < n; ++i) F.zi[i] = zi[i];
  return filtfilt_run(F, x, N, ncols, ldx, padlen, trim, step, y, ldy, stream);
}

// scipy.signal.sosfiltfilt(sos, x, axis=0, padtype="odd", padlen) [::step] for every column.  sos: nsec x 6 host
// array (a0 == 1), zi = scipy.signal.sosfilt_zi(sos) flattened (nsec x 2).  This is what scipy.signal.decimate(x, q,
// zero_phase=True) runs (cheby1(8, 0.05, 0.8 / q, output="sos")) - examples/tiago/identification.py:142-187.
extern "C" int fg_sosfiltfilt(const double* x, int64_t N, int32_t ncols, int64_t ldx, const double* sos, int32_t nsec,
                              const double* zi, int32_t padlen, int64_t trim, int32_t step, double* y, int64_t ldy,
                              void* stream) {
  if (!x || !y || !sos || !zi || N < 1 || ncols < 1 || ldx < ncols || ldy < ncols || step < 1)
    return set_err(FG_ERR_INVALID, "fg_sosfiltfilt: bad argument");
  if (nsec < 1 || 2 * nsec > IIR_MAXN) return set_err(FG_ERR_UNSUPPORTED, "fg_sosfiltfilt: %d sections outside [1,%d]", nsec, IIR_MAXN / 2);
  if (padlen < 0 || padlen >= N) return set_err(FG_ERR_INVALID, "fg_sosfiltfilt: padlen=%d needs more than %d samples", padlen, padlen);
  if (trim < 0 || 2 * trim >= N) return set_err(FG_ERR_INVALID, "fg_sosfiltfilt: trim=%lld leaves no row", (long long)trim);
  IirCoef F;
  memset(&F, 0, sizeof F);
  F.n = 2 * nsec;
  F.nsec = nsec;
  for (int s = 0; s < nsec; ++s) {
    if (sos[6 * s + 3] != 1.0) return set_err(FG_ERR_INVALID, "fg_sosfiltfilt: sos[%d][3] must be 1", s);
    for (int k = 0; k < 6; ++k) F.sos[s][k] = sos[6 * s + k];
  }
  for (int i = 0; i < 2 * nsec; ++i) F.zi[i] = zi[i];
  return filtfilt_run(F, x, N, ncols, ldx, padlen, trim, step, y, ldy, stream);
}

extern "C" int fg_filtfilt(const double* x, int64_t N, int32_t ncols, int64_t ldx, const double* b, int32_t nb,
                           const double* a, int32_t na, const double* zi, int32_t padlen, int64_t trim, double* y,
                           int64_t ldy, void* stream) {
  return filtfilt_impl(x, N, ncols, ldx, b, nb, a, na, zi, padlen, trim, 1, y, ldy, stream);
}

// scipy.signal.decimate(x, q, zero_phase=True) for every column: filtfilt with the caller's (Chebyshev) design,
// then every q-th row - only those rows are written (ceil(N / q) of them).
extern "C" int fg_filtfilt_decimate(const double* x, int64_t N, int32_t ncols, int64_t ldx, const double* b, int32_t nb,
                                    const double* a, int32_t na, const double* zi, int32_t padlen, int32_t q, double* y,
                                    int64_t ldy, void* stream) {
  if (q < 1) return set_err(FG_ERR_INVALID, "fg_filtfilt_decimate: q must be >= 1");
  return filtfilt_impl(x, N, ncols, ldx, b, nb, a, na, zi, padlen, 0, q, y, ldy, stream);
}

extern "C" int fg_medfilt(const double* x, int64_t N, int32_t ncols, int64_t ldx, int32_t k, double* y, int64_t ldy,
                          void* stream) {
  if (!x || !y || N < 1 || ncols < 1 || ldx < ncols || ldy < ncols || k < 1 || k > MEDFILT_MAXK || (k & 1) == 0)
    return set_err(FG_ERR_INVALID, "fg_medfilt: bad argument (odd kernel size up to %d)", MEDFILT_MAXK);
  const int blocks = (int)std::min<int64_t>((N * ncols + 255) / 256, 148 * 16);
  medfilt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, N, ncols, ldx, k, y, ldy);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

extern "C" int fg_gradient(const double* x, int64_t N, int32_t ncols, int64_t ldx, const double* t_dev, double* y,
                           int64_t ldy, void* stream) {
  if (!x || !y || N < 2 || ncols < 1 || ldx < ncols || ldy < ncols) return set_err(FG_ERR_INVALID, "fg_gradient: bad argument (needs N >= 2)");
  const int blocks = (int)std::min<int64_t>((N * ncols + 255) / 256, 148 * 16);
  gradient_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, N, ncols, ldx, t_dev, y, ldy);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

// calculate_first_second_order_differentiation (identification_tools.py:334-387) for vector-space joints
// (nq == nv): q (N x nq, device) -> dq, ddq ((N - 2) x nq, device); dt_dev == NULL: constant step dt.
extern "C" int fg_differentiate(const double* q, int64_t N, int32_t nq, int64_t ldq, double dt, const double* dt_dev,
                                double* dq, double* ddq, int64_t ldo, void* stream) {
  if (!q || !dq || !ddq || N < 3 || nq < 1 || ldq < nq || ldo < nq || (!dt_dev && dt == 0.0))
    return set_err(FG_ERR_INVALID, "fg_differentiate: bad argument (needs N >= 3)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t work = (N - 2) * nq;
  const int blocks = (int)std::min<int64_t>((work + 255) / 256, 148 * 16);
  differentiate_kernel<<<blocks, 256, 0, st>>>(q, N, nq, ldq, dt, dt_dev, dq, ddq, ldo);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

// ---- export / merge (multi-GPU exchange)
extern "C" int fg_state_export(const fg_state* s, double* packed, void* stream) {
  if (!s || !packed) return set_err(FG_ERR_INVALID, "fg_state_export: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  // [R dense p1 x p1 | squared column norms p1 | rows | column order p] - everything written by the kernel
  tri_to_dense_kernel<<<64, 256, 0, st>>>(s->d_part, s->p1, packed, (double)s->rows);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

// Merge n packed factors (fg_state_export layout, e.g. the output of an all-gather) into the state.  No host
// round trip: the row counts are summed on the device and fetched only when somebody asks (fg_state_rows).
extern "C" int fg_state_merge_packed(fg_state* s, const double* packed, int32_t n, void* stream) {
  if (!s || !packed || n < 1) return set_err(FG_ERR_INVALID, "fg_state_merge_packed: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int sm_count, max_smem;
  int rc = device_limits(&sm_count, &max_smem);
  if (rc) return rc;
  rc = ensure_parts(s, n + 1);
  if (rc) return rc;
  if (!s->d_rows_pending) {
    FG_CUDA(cudaMalloc(&s->d_rows_pending, sizeof(double)));
    FG_CUDA(cudaMemsetAsync(s->d_rows_pending, 0, sizeof(double), st));
  }
  {  // pairwise tree with the cooperative blocked TSQR (fg_coop.cuh): 40 us per level against 95 us (UR10, 74 columns)
     // and 0.7 ms against 2.4 ms (Talos, 331 columns) with the column-owner merge kernels
    const char* ecq = getenv("FG_USE_CQ");
    if (n >= 2 && !(ecq && ecq[0] == '0') && s->p1 <= cq_max_p1(max_smem)) {
      rc = ensure_gR(s, (int64_t)n * cq_r_doubles(s->p1));
      if (rc) return rc;
      cudaError_t e = cq_import_dense(packed, n, s->p1, s->d_gR, s->d_rows_pending, st);
      int nl = 1;
      if (e == cudaSuccess) e = cq_reduce_export(s->d_gR, n, s->p1, s->d_parts, max_smem, &nl, st);
      if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "merge of packed factors failed: %s", cudaGetErrorString(e));
      for (int i = 0; i < nl; ++i) FG_LAUNCHED();
      s->rows_pending = true;
      return merge_parts_into_state(s, 1, max_smem, st);
    }
  }
  dense_to_tri_kernel<<<dim3(16, n), 256, 0, st>>>(packed, n, s->p1, s->d_parts, s->d_rows_pending);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  s->rows_pending = true;
  return merge_parts_into_state(s, n, max_smem, st);
}

// diag(W^T W) of a dense device matrix of any width (get_index_eliminate, regressor.py:258-279): no QR needed
extern "C" int fg_colsq_dense(const double* W, int64_t ldw, int64_t rows, int32_t p, double* out_host, void* stream) {
  if (!W || !out_host || rows < 0 || p < 1 || ldw < p) return set_err(FG_ERR_INVALID, "fg_colsq_dense: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nslab = (int)std::max<int64_t>(1, std::min<int64_t>(rows / 64 + 1, 592));
  double* d = nullptr;
  FG_CUDA(cudaMallocAsync(&d, sizeof(double) * ((size_t)nslab + 1) * p, st));
  colsq_partial_kernel<<<dim3((p + 255) / 256, nslab), 256, 0, st>>>(W, ldw, rows, p, d);
  colsq_final_kernel<<<(p + 255) / 256, 256, 0, st>>>(d, nslab, p, d + (size_t)nslab * p);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED(); FG_LAUNCHED();
  FG_CUDA(cudaMemcpyAsync(out_host, d + (size_t)nslab * p, sizeof(double) * p, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaFreeAsync(d, st));
  FG_CUDA(cudaStreamSynchronize(st));
  return FG_OK;
}

// ------------------------------------------------------------------------------------ small stage
extern "C" int fg_state_colsq(fg_state* s, double* out, void* stream) {
  if (!s || !out) return set_err(FG_ERR_INVALID, "fg_state_colsq: null argument");
  const int ntri = s->p1 * (s->p1 + 1) / 2;
  std::vector<double> tmp(s->p1);
  FG_CUDA(cudaMemcpyAsync(tmp.data(), s->d_part + ntri, sizeof(double) * s->p1, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (int rc = check_rq_error(s)) return rc;
  const std::vector<int> inv = inverse_perm(s);
  for (int i = 0; i < s->p; ++i) out[i] = tmp[inv[i]];
  out[s->p] = tmp[s->p];
  return FG_OK;
}

// packed upper triangle (n columns) -> dense n x n (ld = n), zeros below the diagonal
__global__ void tri_to_square_kernel(const double* __restrict__ part, int n, double* __restrict__ A) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
    const int r = idx / n, c = idx % n;
    A[idx] = c >= r ? part[tri_off(r, n) + c - r] : 0.0;
  }
}

// In-place QR of the (mrows x n, ld = n) matrix A on the device: the upper triangle of its first n rows becomes R.
// Small matrices: one CTA in shared memory (small_qr_kernel).  Wide ones (n > 128): the rows are folded by the
// CTA-wide register-tile kernel of fg_wide.cuh (one CTA; the last column rides as the rhs) into `scratch`
// (part_doubles(n) doubles) and unpacked.  Only the first min(mrows, n) rows of the result are meaningful.
static int run_small_qr(double* A, int mrows, int n, const int* d_ext, cudaStream_t st, double* scratch = nullptr) {
  int sm_count, max_smem;
  int rc = device_limits(&sm_count, &max_smem);
  if (rc) return rc;
  if (n > 128 && scratch != nullptr && n <= cq_max_p1(max_smem)) {
    // cooperative blocked TSQR with one CTA (fg_coop.cuh); scratch: one part for the result + two for the block factor
    CqArgs ca;
    ca.W = A; ca.ldw = n; ca.rhs_in_w = true; ca.rows = mrows; ca.p1 = n;
    double* Rblk = scratch + part_doubles(n);
    cudaError_t e = cq_launch_dense(ca, Rblk, 1, max_smem, st);
    int nl = 1;
    if (e == cudaSuccess) e = cq_reduce_export(Rblk, 1, n, scratch, max_smem, &nl, st);
    if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "small stage (cooperative) failed: %s", cudaGetErrorString(e));
    for (int i = 0; i < nl; ++i) FG_LAUNCHED();
    tri_to_square_kernel<<<64, 256, 0, st>>>(scratch, n, A);
    FG_CUDA(cudaGetLastError());
    FG_LAUNCHED();
    return FG_OK;
  }
  if (n > 128 && scratch != nullptr && wide_dense_tile_rows(n) > 0) {
    cudaError_t e = wide_launch_dense(A, n, A + (n - 1), n, mrows, n, scratch, 1, st);
    if (e != cudaSuccess) return set_err(FG_ERR_CUDA, "small stage (wide) failed: %s", cudaGetErrorString(e));
    FG_LAUNCHED();
    tri_to_square_kernel<<<64, 256, 0, st>>>(scratch, n, A);
    FG_CUDA(cudaGetLastError());
    FG_LAUNCHED();
    return FG_OK;
  }
  int smem = (mrows * n + n) * 8 + n * 4, use_smem = 1;
  if (smem > max_smem) { smem = n * 8 + n * 4; use_smem = 0; }
  FG_CUDA(cudaFuncSetAttribute(small_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int threads = n <= 128 ? 128 : (n <= 256 ? 256 : 512);
  small_qr_kernel<<<1, threads, smem, st>>>(A, mrows, n, d_ext, use_smem);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  return FG_OK;
}

extern "C" int fg_state_reduce(fg_state* s, const int32_t* keep, int32_t nk, double* diag, void* stream) {
  if (!s || !keep || nk < 1 || nk > s->p || !diag) return set_err(FG_ERR_INVALID, "fg_state_reduce: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int p1 = s->p1, n = nk + 1;
  std::vector<int> idx(2 * n);
  const std::vector<int> inv = inverse_perm(s);
  for (int j = 0; j < nk; ++j) {
    if (keep[j] < 0 || keep[j] >= s->p || (j > 0 && keep[j] <= keep[j - 1]))
      return set_err(FG_ERR_INVALID, "fg_state_reduce: keep must be ascending indices in [0,%d)", s->p);
    idx[j] = inv[keep[j]];       // where the factor keeps caller column keep[j]
    idx[n + j] = idx[j] + 1;     // rows 0..idx[j] of that column can be non-zero
  }
  idx[nk] = s->p;       // rhs column
  idx[n + nk] = p1;
  FG_CUDA(cudaMemcpyAsync(s->d_idx, idx.data(), sizeof(int) * 2 * n, cudaMemcpyHostToDevice, st));
  gather_cols_kernel<<<64, 256, 0, st>>>(s->d_part, p1, 1, s->d_idx, p1, n, s->d_A);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  int rc = ensure_parts(s, 3);
  if (rc) return rc;
  rc = run_small_qr(s->d_A, p1, n, s->d_idx + n, st, s->d_parts);
  if (rc) return rc;
  // keep the leading n x n block as the reduced factor (dense, ld = n)
  FG_CUDA(cudaMemcpyAsync(s->d_Re, s->d_A, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
  std::vector<double> Rh((size_t)n * n);
  FG_CUDA(cudaMemcpyAsync(Rh.data(), s->d_Re, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  if (int rc2 = check_rq_error(s)) return rc2;
  for (int j = 0; j < n; ++j) diag[j] = std::abs(Rh[(size_t)j * n + j]);
  s->nk = nk;
  s->keep.assign(keep, keep + nk);
  return FG_OK;
}

// Column-pivoted QR of the kept columns (QR_pivoting, qrdecomposition.py:24-86) on the factor.
extern "C" int fg_state_pivot_qr(fg_state* s, const int32_t* keep, int32_t nk, int32_t* perm_host, double* R_host,
                                 void* stream) {
  if (!s || !keep || nk < 1 || nk > s->p || !perm_host || !R_host)
    return set_err(FG_ERR_INVALID, "fg_state_pivot_qr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int p1 = s->p1, n = nk + 1;
  int sm_count, max_smem;
  int rc = device_limits(&sm_count, &max_smem);
  if (rc) return rc;
  const int smem = (p1 * n + n) * 8 + n * 4 + 16;
  if (smem > max_smem) return set_err(FG_ERR_UNSUPPORTED, "fg_state_pivot_qr: factor does not fit shared memory");
  std::vector<int> idx(2 * n);
  const std::vector<int> inv = inverse_perm(s);
  for (int j = 0; j < nk; ++j) {
    if (keep[j] < 0 || keep[j] >= s->p || (j > 0 && keep[j] <= keep[j - 1]))
      return set_err(FG_ERR_INVALID, "fg_state_pivot_qr: keep must be ascending indices in [0,%d)", s->p);
    idx[j] = inv[keep[j]];
  }
  idx[nk] = s->p;
  FG_CUDA(cudaMemcpyAsync(s->d_idx, idx.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  gather_cols_kernel<<<64, 256, 0, st>>>(s->d_part, p1, 1, s->d_idx, p1, n, s->d_A);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  FG_CUDA(cudaFuncSetAttribute(pivot_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pivot_qr_kernel<<<1, 256, smem, st>>>(s->d_A, p1, n, nk, s->d_idx + n);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  std::vector<int> perm(n);
  FG_CUDA(cudaMemcpyAsync(R_host, s->d_A, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaMemcpyAsync(perm.data(), s->d_idx + n, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  if (int rc2 = check_rq_error(s)) return rc2;
  for (int j = 0; j < nk; ++j) perm_host[j] = perm[j];
  return FG_OK;
}

// get_index_eliminate + get_baseIndex + double_QR on the factor in ONE launch and ONE device-to-host copy
// (fused_solve_kernel); the three-call sequence colsq / reduce / regroup_solve stays for callers that decide
// the column sets themselves.
extern "C" int fg_state_solve(fg_state* s, double tol_e, const int32_t* drop, int32_t n_drop, double tol_qr,
                              int32_t* n_keep, int32_t* keep, double* colsq, double* diag, int32_t* n_base,
                              int32_t* order, double* beta, double* phi, double* R1, double* resid2, void* stream) {
  if (!s || !n_keep || !keep || !n_base || !order) return set_err(FG_ERR_INVALID, "fg_state_solve: null argument");
  if (!s->perm.empty()) return set_err(FG_ERR_UNSUPPORTED, "fg_state_solve: the state holds a permuted factor");
  const int p = s->p, p1 = s->p1;
  if (p > 32 * US_M) return set_err(FG_ERR_UNSUPPORTED, "fg_state_solve: p=%d above %d", p, 32 * US_M);
  int sm_count, max_smem;
  int rc = device_limits(&sm_count, &max_smem);
  if (rc) return rc;
  const int smem = (2 * p1 * p1 + p1) * 8 + 3 * p1 * 4 + 16;
  if (smem > max_smem) return set_err(FG_ERR_UNSUPPORTED, "fg_state_solve: factor does not fit shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t blob = 4 + 3 * (size_t)p1 + 2 * (size_t)p1 * p1;
  if (!s->d_solve) FG_CUDA(cudaMalloc(&s->d_solve, sizeof(double) * blob));
  std::vector<int> flags(p, 0);
  for (int i = 0; i < n_drop; ++i) {
    if (!drop || drop[i] < 0 || drop[i] >= p) return set_err(FG_ERR_INVALID, "fg_state_solve: bad column index in drop");
    flags[drop[i]] = 1;
  }
  if (n_drop > 0) FG_CUDA(cudaMemcpyAsync(s->d_idx, flags.data(), sizeof(int) * p, cudaMemcpyHostToDevice, st));
  FG_CUDA(cudaFuncSetAttribute(fused_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fused_solve_kernel<<<1, 256, smem, st>>>(s->d_part, p1, n_drop > 0 ? s->d_idx : nullptr, tol_e, tol_qr, s->d_Re,
                                           s->d_solve);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  std::vector<double> h(blob);
  FG_CUDA(cudaMemcpyAsync(h.data(), s->d_solve, sizeof(double) * blob, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  if (int rc2 = check_rq_error(s)) return rc2;
  const int nk = (int)h[0], r = (int)h[1], n = nk + 1, ng = nk - r, nrhs = ng + 1;
  const double* h_colsq = h.data() + 4;
  const double* h_diag = h_colsq + p1;
  const double* h_order = h_diag + p1;
  const double* h_Rg = h_order + p1;
  const double* h_X = h_Rg + (size_t)n * n;
  *n_keep = nk;
  *n_base = r;
  s->keep.clear();
  {
    int k = 0;
    for (int c = 0; c < p; ++c)
      if (!(flags[c] || (tol_e >= 0.0 && h_colsq[c] < tol_e))) { keep[k++] = c; s->keep.push_back(c); }
    if (k != nk) return set_err(FG_ERR_CUDA, "fg_state_solve: internal: keep count mismatch (%d vs %d)", k, nk);
  }
  s->nk = nk;
  if (colsq) for (int c = 0; c < p1; ++c) colsq[c] = h_colsq[c];
  if (diag) for (int j = 0; j < n; ++j) diag[j] = h_diag[j];
  for (int j = 0; j < nk; ++j) order[j] = (int)h_order[j];
  if (r == 0) return FG_OK;
  for (int i = 0; i < r; ++i) {
    if (beta) for (int j = 0; j < ng; ++j) beta[(size_t)i * ng + j] = h_X[(size_t)i * nrhs + j];
    if (phi) phi[i] = h_X[(size_t)i * nrhs + ng];
    if (R1) for (int j = 0; j < r; ++j) R1[(size_t)i * r + j] = h_Rg[(size_t)i * n + j];
  }
  if (resid2) {
    double r2 = 0.0;
    for (int i = r; i < n; ++i) r2 += h_Rg[(size_t)i * n + (n - 1)] * h_Rg[(size_t)i * n + (n - 1)];
    *resid2 = r2;
  }
  return FG_OK;
}

extern "C" int fg_state_regroup_solve(fg_state* s, const int32_t* base, int32_t nb, const int32_t* regroup, int32_t ng,
                                      double* beta, double* phi, double* R1, double* resid2, void* stream) {
  if (!s || s->nk < 0) return set_err(FG_ERR_INVALID, "fg_state_regroup_solve: call fg_state_reduce first");
  if (nb < 1 || nb + ng != s->nk || !base || (ng > 0 && !regroup))
    return set_err(FG_ERR_INVALID, "fg_state_regroup_solve: base (%d) + regroup (%d) must cover the %d kept columns", nb, ng, s->nk);
  cudaStream_t st = (cudaStream_t)stream;
  const int n = s->nk + 1;
  std::vector<int> idx(2 * n);
  std::vector<char> seen(s->nk, 0);
  for (int j = 0; j < nb + ng; ++j) {
    const int c = j < nb ? base[j] : regroup[j - nb];
    if (c < 0 || c >= s->nk || seen[c]) return set_err(FG_ERR_INVALID, "fg_state_regroup_solve: bad column index %d", c);
    seen[c] = 1;
    idx[j] = c;
    idx[n + j] = c + 1;
  }
  idx[n - 1] = s->nk;
  idx[2 * n - 1] = n;
  FG_CUDA(cudaMemcpyAsync(s->d_idx, idx.data(), sizeof(int) * 2 * n, cudaMemcpyHostToDevice, st));
  gather_cols_kernel<<<64, 256, 0, st>>>(s->d_Re, n, 0, s->d_idx, n, n, s->d_A);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  int rc = ensure_parts(s, 3);
  if (rc) return rc;
  rc = run_small_qr(s->d_A, n, n, s->d_idx + n, st, s->d_parts);
  if (rc) return rc;
  const int nrhs = ng + 1;
  if (nb <= 32 * US_M) upper_solve_kernel<<<(nrhs + 3) / 4, 128, 0, st>>>(s->d_A, n, nb, nrhs, s->d_X);
  else upper_solve_serial_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(s->d_A, n, nb, nrhs, s->d_X);
  FG_CUDA(cudaGetLastError());
  FG_LAUNCHED();
  std::vector<double> X((size_t)nb * nrhs), Rg((size_t)n * n);
  FG_CUDA(cudaMemcpyAsync(X.data(), s->d_X, sizeof(double) * nb * nrhs, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaMemcpyAsync(Rg.data(), s->d_A, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
  FG_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < nb; ++i) {
    if (beta) for (int j = 0; j < ng; ++j) beta[(size_t)i * ng + j] = X[(size_t)i * nrhs + j];
    if (phi) phi[i] = X[(size_t)i * nrhs + ng];
    if (R1) for (int j = 0; j < nb; ++j) R1[(size_t)i * nb + j] = Rg[(size_t)i * n + j];
  }
  if (resid2) {
    double r2 = 0.0;
    for (int i = nb; i < n; ++i) r2 += Rg[(size_t)i * n + (n - 1)] * Rg[(size_t)i * n + (n - 1)];
    *resid2 = r2;
  }
  return FG_OK;
}
