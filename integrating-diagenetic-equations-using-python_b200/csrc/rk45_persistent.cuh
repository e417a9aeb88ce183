// Host-side declarations of the kernels' launchers (internal to the shared library).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/marlpde_b200.h"

namespace marlpde {

// The on-chip RK45 kernel gives every thread two adjacent depth cells; see rk45_persistent.cu for
// the builds (threads per CTA, registers) and the largest grid each accepts.
int rk45_max_cells();
int rk45_columns_per_cta(int n_cells, int smem_budget);

cudaError_t launch_rk45(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                        int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                        double* d_snap, int32_t* d_ev_counts, double* d_ev_times, int32_t* d_queue, int sm_count,
                        int smem_budget, cudaStream_t stream);

cudaError_t launch_rhs_batch(const double* d_y, const marlpde_column_params* d_params, int n_columns,
                             int n_cells, double* d_out, cudaStream_t stream);

}  // namespace marlpde
