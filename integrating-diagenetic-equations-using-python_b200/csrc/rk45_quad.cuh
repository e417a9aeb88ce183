// Host-side declarations of the experimental 4-cells-per-thread RK45 kernel (rk45_quad.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/marlpde_b200.h"

namespace marlpde {

// columns per CTA for this grid, 0 if the build does not take it (N % 4 != 0, N < 32, N > 1024, shared memory)
int rk45_quad_columns_per_cta(int n_cells, int smem_budget);

cudaError_t launch_rk45_quad(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                             int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                             double* d_snap, int32_t* d_ev_counts, double* d_ev_times, int32_t* d_queue, int sm_count,
                             int smem_budget, cudaStream_t stream);

}  // namespace marlpde
