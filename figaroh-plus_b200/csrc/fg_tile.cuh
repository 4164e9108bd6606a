// fg_tile.cuh - layout of the shared-memory regressor tiles and of the block-triangular factors that the
// row-group streams (fg_groupqr.cuh, fg_ringqr.cuh) work on, and the fp64 tensor-core primitive.
#pragma once
#include "fg_device.cuh"

namespace fg {

constexpr int TQ_S = 32;  // samples (= rows) of one tile: one row per lane

// phases of a chunked accumulation: FIRST starts the per-stream factors from zero, LAST merges them into the state
enum { FG_PHASE_FIRST = 1, FG_PHASE_LAST = 2 };

// mma.sync.m8n8k4.f64 (DMMA): the fp64 tensor path of sm_100a (tcgen05 has no fp64 mode)
__device__ __forceinline__ void tq_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Tile block = 8 columns x 32 rows, column-major with the row index XOR-swizzled per column so that
// every access pattern of the kernels (lane = row; lane = (column, row mod 4); lane = (row mod 8,
// column pair)) hits 16 distinct 8-byte banks per half warp.
__host__ __device__ __forceinline__ int tq_sw(int c) {
  const int f = (c + (c >> 2)) & 3;
  return (f << 2) | (f & 1);
}
__host__ __device__ __forceinline__ int tq_idx(int c, int s) { return c * TQ_S + (s ^ tq_sw(c)); }

// index of R block (k, j), j >= k, in a block-upper-triangular factor (64 doubles per block, row-major)
__host__ __device__ __forceinline__ int tq_rblk(int ncb, int k, int j) { return k * ncb - (k * (k - 1)) / 2 + (j - k); }

}  // namespace fg
