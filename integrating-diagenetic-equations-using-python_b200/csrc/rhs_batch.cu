// rhs_batch.cu — one evaluation of the five-field right-hand side for a batch of columns.
//
// Replaces one call of LMAHeureuxPorosityDiff.fun_numba -> pde_rhs per column
// (marlpde/LHeureux_model.py:290-359, :361-522): y[column][field][cell] -> dy/dt, same layout.
// One thread per PAIR of adjacent depth cells (the same rhs_pair code the persistent RK45 kernel
// runs, so the single-call parity tests exercise the integrators' arithmetic — every instruction schedule of
// rhs_pair_own that an integrator instantiates can be selected with MARLPDE_RHS_SCHEDULE=0..4); a CTA covers
// 2 * kPairsPerCta consecutive cells of one column.  Each thread reads its own two cells and one
// neighbour on either side straight from global memory (the neighbours are L1 hits: they are the
// adjacent threads' own cells).  40 B in + 40 B out per cell against ~200 fp64 instructions:
// fp64-pipe bound, not HBM bound.
#include <cuda_runtime.h>

#include <cstdlib>

#include "lheureux_device.cuh"
#include "rk45_persistent.cuh"

namespace marlpde {

constexpr int kPairsPerCta = 128;

template <int kSched>
__global__ void __launch_bounds__(kPairsPerCta)
rhs_batch_kernel(const double* __restrict__ g_y, const marlpde_column_params* __restrict__ g_params,
                 int n_cells, int tiles_per_col, double* __restrict__ g_out) {
  __shared__ ColumnConsts kc;
  __shared__ __align__(16) unsigned char tab_raw[fm::kTableBytes];
  const fm::Tables tb = fm::stage_tables(tab_raw, threadIdx.x, kPairsPerCta);
  const int col = blockIdx.x / tiles_per_col;   // 1-D grid: 65 536+ columns exceed gridDim.y
  const int pair = (blockIdx.x - col * tiles_per_col) * kPairsPerCta + threadIdx.x;
  const int cell0 = 2 * pair;
  const double* ycol = g_y + (size_t)col * 5 * n_cells;
  if (threadIdx.x == 0) make_consts(g_params[col], n_cells, kc);
  __syncthreads();

  const bool valid0 = cell0 < n_cells;          // lanes past the column end run on benign values:
  const bool valid1 = cell0 + 1 < n_cells;      // rhs_pair votes need all 32 lanes of the warp
  double c[5][2], mlo[5], phi[5];
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    const double* yf = ycol + (size_t)f * n_cells;
    c[f][0] = valid0 ? yf[cell0] : 0.5;
    c[f][1] = valid1 ? yf[cell0 + 1] : 0.5;
    mlo[f] = (valid0 && cell0 > 0) ? yf[cell0 - 1] : top_ghost(kc, f, c[f][0]);
    if (cell0 + 2 < n_cells) {
      phi[f] = yf[cell0 + 2];
    } else if (valid1) {                         // cell0 + 1 is the last cell of the column
      phi[f] = bottom_ghost(f, c[f][1], c[f][0]);
    } else {                                     // cell0 is the last cell: its ghost sits in slot 1
      c[f][1] = bottom_ghost(f, c[f][0], mlo[f]);
      phi[f] = c[f][1];
    }
  }
  const bool in_mask[2] = {cell0 >= kc.mask_lo && cell0 < kc.mask_hi,
                           cell0 + 1 >= kc.mask_lo && cell0 + 1 < kc.mask_hi};
  double r[5][2], U[2], W[2];
  // a CTA lies inside one column, so the model variant is a CTA-uniform choice between the two instantiations the
  // integrators use (MARLPDE_MODEL_VAR_DPHI: per-cell porosity diffusion coefficient, LHeureux_model.py:430)
  PairFlags fl;
  if (kc.var_dphi) fl = rhs_pair<kSched, true>(kc, tb, c, mlo, phi, in_mask, r, U, W);
  else fl = rhs_pair<kSched, false>(kc, tb, c, mlo, phi, in_mask, r, U, W);
  fl.bad[0] = fl.bad[0] && valid0;
  fl.bad[1] = fl.bad[1] && valid1;
  if (fl.bad[0] || fl.bad[1]) rhs_pair_fixup(kc, tb, fl, c, mlo, phi, in_mask, r, U, W);
  double* ocol = g_out + (size_t)col * 5 * n_cells;
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    if (valid0) ocol[(size_t)f * n_cells + cell0] = r[f][0];
    if (valid1) ocol[(size_t)f * n_cells + cell0 + 1] = r[f][1];
  }
}

cudaError_t launch_rhs_batch(const double* d_y, const marlpde_column_params* d_params, int n_columns,
                             int n_cells, double* d_out, cudaStream_t stream) {
  if (n_columns == 0) return cudaSuccess;
  const int pairs = (n_cells + 1) / 2;
  const int tiles = (pairs + kPairsPerCta - 1) / kPairsPerCta;
  const long long blocks = (long long)tiles * n_columns;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  // default: the schedule of the on-chip RK45 kernel; MARLPDE_RHS_SCHEDULE (read per call) selects another one
  const char* env = std::getenv("MARLPDE_RHS_SCHEDULE");
  const int sched = rhs_schedule(env && env[0] >= '0' && env[0] <= '4' && !env[1] ? env[0] - '0' : kSchedLean);
  const unsigned g = (unsigned)blocks;
  switch (sched) {
    case kSchedSplit: rhs_batch_kernel<kSchedSplit><<<g, kPairsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out); break;
    case kSchedMerged: rhs_batch_kernel<kSchedMerged><<<g, kPairsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out); break;
    default: rhs_batch_kernel<kSchedLean><<<g, kPairsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out); break;
    case kSchedAll: rhs_batch_kernel<kSchedAll><<<g, kPairsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out); break;
    case kSchedTwoArm: rhs_batch_kernel<kSchedTwoArm><<<g, kPairsPerCta, 0, stream>>>(d_y, d_params, n_cells, tiles, d_out); break;
  }
  return cudaGetLastError();
}

}  // namespace marlpde
