// fg_coop.h - host interface of the cooperative blocked TSQR of a dense matrix in HBM (fg_coop.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fg {

// widest problem (columns incl. the rhs) the kernel takes with max_smem bytes of shared memory per CTA; 0 = none
int cq_max_p1(int max_smem);
// doubles of one CTA's block-triangular factor
int64_t cq_r_doubles(int p1);

struct CqArgs {
  const double* W = nullptr;   // rows x ldw, row-major (device)
  int64_t ldw = 0;
  const int* colsel = nullptr; // p1 - 1 column indices into W (device), or nullptr: columns 0 .. p1 - 2
  const double* tau = nullptr; // rhs (device) or nullptr
  int64_t tau_blk = 0, tau_ld = 0;  // rhs of row r at tau[(r / tau_blk) * tau_ld + r % tau_blk]; tau_blk = 0: tau[r]
  int64_t rows = 0;
  int p1 = 0;
  bool rhs_in_w = false;       // all p1 columns (incl. the rhs) are columns 0 .. p1 - 1 of W
  bool first = true;           // a sequence of launches over row chunks shares the per-CTA factors: the first one zeroes them
};
// CTA b of `grid` folds tiles b, b + grid, ... into its factor r_all + b * cq_r_doubles(p1).
cudaError_t cq_launch_dense(const CqArgs& a, double* r_all, int grid, int max_smem, cudaStream_t st);
// after the last launch of a sequence: pairwise tree over the `grid` factors (log2 launches of the same kernel: 64
// rows of the right-hand factor per tile, panels starting at the tile's first non-zero column), then factor 0 as one
// packed part (upper triangle + squared column norms) in part_out.  *launches += kernels launched.
// the two halves of cq_reduce_export, and the scatter of a narrower factor's rows into dense rows of width p1 (D must
// be zeroed; colmap: p1g device ints, local column -> column of the wide problem)
cudaError_t cq_reduce(double* r_all, int grid, int p1, int max_smem, int* launches, cudaStream_t st);
cudaError_t cq_export(const double* r_all, int p1, double* part_out, cudaStream_t st);
cudaError_t cq_scatter_rows(const double* R, int p1g, const int* colmap, int p1, double* D, cudaStream_t st);
// n gathered factors in the fg_state_export layout -> block layout in r_all (zeroed here), row counts added to *rows_acc
cudaError_t cq_import_dense(const double* packed, int n, int p1, double* r_all, double* rows_acc, cudaStream_t st);
cudaError_t cq_reduce_export(double* r_all, int grid, int p1, double* part_out, int max_smem, int* launches, cudaStream_t st);

}  // namespace fg
