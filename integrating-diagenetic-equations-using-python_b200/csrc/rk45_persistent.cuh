// Host-side declarations of the kernels' launchers (internal to the shared library).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/marlpde_b200.h"

namespace marlpde {

// Upper bound on threads per CTA of the on-chip RK45 kernel: 608 threads (3 columns of 200
// cells, 19 warps) x 96 registers (registers are granted per warp in units of 512, so 104 would
// not fit) fill the 64K-register file of one SM with one resident CTA.
constexpr int kRk45MaxThreads = 608;

int rk45_columns_per_cta(int n_cells, int smem_budget);
size_t rk45_smem_bytes(int columns_per_cta, int n_cells);

cudaError_t launch_rk45(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                        int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                        double* d_snap, int32_t* d_queue, int sm_count, int smem_budget, cudaStream_t stream);

cudaError_t launch_rhs_batch(const double* d_y, const marlpde_column_params* d_params, int n_columns,
                             int n_cells, double* d_out, cudaStream_t stream);

}  // namespace marlpde
