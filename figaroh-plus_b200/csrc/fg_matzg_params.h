// fg_matzg_params.h - launch parameters of the row-block-major materialising kernel (fg_matzg.cuh)
#pragma once

namespace fg {

struct MatzgParams {
  int nb, ms, nq, nv, flags, ncols, pitch;   // pitch: doubles per staged row (= ncols: dense, one bulk copy per block)
  int smem_per_warp;                         // doubles
  double gravity[3];
};

}  // namespace fg
