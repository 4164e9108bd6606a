// fg_gq.cu - launcher of the merge of the per-stream factors (fg_groupqr.cuh::gq_merge_kernel).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fg_groupqr.cuh"

namespace fg {

cudaError_t gq_merge(int smem_merge, const GqParams& prm, double* r_all, int64_t ntiles_hint, unsigned* counters,
                     double* rfull, double* part_out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(gq_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_merge);
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
  gq_merge_kernel<<<blocks, 32, smem_merge, st>>>(prm, M, r_all, counters, rfull, part_out);
  return cudaGetLastError();
}

}  // namespace fg
