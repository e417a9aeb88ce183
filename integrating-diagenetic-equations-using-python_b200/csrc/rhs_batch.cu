// rhs_batch.cu — one evaluation of the five-field right-hand side for a batch of columns.
//
// Replaces one call of LMAHeureuxPorosityDiff.fun_numba -> pde_rhs per column
// (marlpde/LHeureux_model.py:290-359, :361-522): y[column][field][cell] -> dy/dt, same layout.
// One thread per depth cell; a CTA covers kCellsPerCta consecutive cells of one column, stages
// its 5 x (cells + 2 halo) inputs in shared memory with coalesced loads, and every thread then
// reads its (i-1, i, i+1) triple from the tile.  Memory traffic is 40 B in + 40 B out per cell
// against ~350 fp64 instructions, so the kernel is fp64-pipe bound, not HBM bound.
#include <cuda_runtime.h>

#include "lheureux_device.cuh"
#include "rk45_persistent.cuh"

namespace marlpde {

constexpr int kCellsPerCta = 256;

__global__ void __launch_bounds__(kCellsPerCta)
rhs_batch_kernel(const double* __restrict__ g_y, const marlpde_column_params* __restrict__ g_params,
                 int n_cells, int tiles_per_col, double* __restrict__ g_out) {
  __shared__ ColumnConsts kc;
  __shared__ double tile[5][kCellsPerCta + 2];
  __shared__ __align__(16) unsigned char tab_raw[fm::kTableBytes];
  const fm::Tables tb = fm::stage_tables(tab_raw, threadIdx.x, kCellsPerCta);
  const int col = blockIdx.x / tiles_per_col;   // 1-D grid: 65 536+ columns exceed gridDim.y
  const int cell0 = (blockIdx.x - col * tiles_per_col) * kCellsPerCta;
  const int cell = cell0 + threadIdx.x;
  const double* ycol = g_y + (size_t)col * 5 * n_cells;
  if (threadIdx.x == 0) make_consts(g_params[col], n_cells, kc);
  // tile index j <-> cell (cell0 - 1 + j), clamped into the column; ghosts are rebuilt below
  for (int j = threadIdx.x; j < kCellsPerCta + 2; j += kCellsPerCta) {
    int i = cell0 - 1 + j;
    i = i < 0 ? 0 : (i > n_cells - 1 ? n_cells - 1 : i);
#pragma unroll
    for (int f = 0; f < 5; ++f) tile[f][j] = ycol[(size_t)f * n_cells + i];
  }
  __syncthreads();
  if (cell >= n_cells) return;
  double c[5], m[5], p[5];
  load_triple(kc, cell, [&](int f, int i) { return tile[f][i - cell0 + 1]; }, c, m, p);
  CellRates r;
  cell_rhs(kc, tb, c, m, p, cell >= kc.mask_lo && cell < kc.mask_hi, r);
  double* ocol = g_out + (size_t)col * 5 * n_cells;
#pragma unroll
  for (int f = 0; f < 5; ++f) ocol[(size_t)f * n_cells + cell] = r.r[f];
}

cudaError_t launch_rhs_batch(const double* d_y, const marlpde_column_params* d_params, int n_columns,
                             int n_cells, double* d_out, cudaStream_t stream) {
  if (n_columns == 0) return cudaSuccess;
  const int tiles = (n_cells + kCellsPerCta - 1) / kCellsPerCta;
  const long long blocks = (long long)tiles * n_columns;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  rhs_batch_kernel<<<(unsigned)blocks, kCellsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out);
  return cudaGetLastError();
}

}  // namespace marlpde
