// Host-side declarations of the streaming (large depth grid) RK45 launcher (internal to the library).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/marlpde_b200.h"

namespace marlpde {

size_t rk45_stream_workspace_bytes(int n_columns, int n_cells);

// Enqueues `attempts` adaptive step attempts for every column (plus the K1 evaluation and the
// bookkeeping kernels) on `stream`; never synchronises.
cudaError_t launch_rk45_stream(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                               int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                               double* d_snap, int32_t* d_ev_counts, double* d_ev_times, void* d_work,
                               long long attempts, cudaStream_t stream);

}  // namespace marlpde
