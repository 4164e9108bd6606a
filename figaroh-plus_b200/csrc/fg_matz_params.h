// fg_matz_params.h - launch parameters of the materialising kernel (shared by fg_matz.cuh and its host interface).
#pragma once

namespace fg {

struct MatzParams {
  int nj, nq, nv, ms, nops, nproc;
  int mode, flags;
  int nvslots, nmslots;
  int coupling;             // 3 extra TX40 columns behind the bodies' blocks
  int smem_per_warp;        // doubles
  int stage_inputs;         // the tile's (q, qd, qdd) are staged in shared memory by bulk copy (small models)
  double gravity[3];
};

}  // namespace fg
