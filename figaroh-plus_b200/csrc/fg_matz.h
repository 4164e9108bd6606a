// fg_matz.h - host interface of the materialising kernel (fg_matz.cuh).
#pragma once
#include <cstdint>
#include <memory>
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

struct MatzTables;

}  // namespace fg

#include "fg_matz_params.h"
#include "fg_matzg_params.h"

namespace fg {

struct MatzPlan {
  bool ok = false;
  int nw = 0, smem = 0, sm_count = 0;
  MatzParams prm;
  std::shared_ptr<MatzTables> tab;   // host copy; the caller uploads it (matz_tables_bytes() bytes)
  // row-block-major kernel for serial chains (fg_matzg.cuh); it needs contiguous 16-byte aligned rows (checked per launch)
  bool g_ok = false;
  int g_nw = 0, g_smem = 0;
  MatzgParams gprm;
};

size_t matz_tables_bytes();
// rows: ALL row blocks of the mode (materialised matrices have every row block); mass: per joint (massless bodies have no
// inertial columns in external-wrench mode); ncols_full: 14 bodies (+3 TX40 coupling columns)
MatzPlan matz_build(const Plan& P, const DevJoint* joints, const RowDesc* rows, const double* mass, int ncols_full,
                    int sm_count, int max_smem);
cudaError_t matz_launch(const MatzPlan& mp, const MatzTables* d_tab, const double* q, const double* v, const double* a,
                        int64_t N, double* W, int64_t ldw, cudaStream_t st);

}  // namespace fg
