// jac_probe.cu — test helper: the block-tridiagonal Jacobian the implicit kernels use (implicit_common.cuh jacobian():
// analytic 5x5 blocks), for ONE column on HOST arrays, so that the parity suite can compare it directly with central
// differences of the oracle RHS.  No reference counterpart (SciPy forms its Jacobian by num_jac on the 27-diagonal
// pattern of marlpde/parameters.py:150-199).
#include "implicit_common.cuh"

namespace marlpde {
namespace {

template <bool VD>
__global__ void __launch_bounds__(32) jac_probe_kernel(const double* y_field_major, const marlpde_column_params* params,
                                                       int N, double* ycell, double* J) {
  __shared__ __align__(16) unsigned char tab_raw[fm::kTableBytes];
  __shared__ ColumnConsts kc;
  const fm::Tables tb = fm::stage_tables(tab_raw, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) make_consts(params[0], N, kc);
  const int lane = threadIdx.x;
  for (int idx = lane; idx < 5 * N; idx += 32) ycell[idx] = y_field_major[(idx % 5) * N + idx / 5];
  __syncthreads();
  imp::jac_analytic<VD>(&kc, &tb, N, lane, ycell, 1e-3, J);   // atol of the reference's Solver (parameters.py:208)
}

}  // namespace
}  // namespace marlpde

// J_out: [n_cells][3][5][5] — blocks L, D, U of every cell, each stored [column][row] (column-major)
extern "C" int marlpde_probe_jacobian(const double* y, const marlpde_column_params* params, int n_cells, double* J_out,
                                      int device) {
  if (!y || !params || !J_out || n_cells < 3) return MARLPDE_EINVAL;
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) return MARLPDE_ENODEVICE;
  if (device < 0 || device >= nd || cudaSetDevice(device) != cudaSuccess) return MARLPDE_EINVAL;
  const size_t nb_y = sizeof(double) * 5 * (size_t)n_cells, nb_J = sizeof(double) * 75 * (size_t)n_cells;
  double *dy = nullptr, *dc = nullptr, *dJ = nullptr;
  marlpde_column_params* dp = nullptr;
  cudaError_t e = cudaMalloc(&dy, nb_y);
  if (e == cudaSuccess) e = cudaMalloc(&dc, nb_y);
  if (e == cudaSuccess) e = cudaMalloc(&dJ, nb_J);
  if (e == cudaSuccess) e = cudaMalloc(&dp, sizeof(marlpde_column_params));
  if (e == cudaSuccess) e = cudaMemcpy(dy, y, nb_y, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dp, params, sizeof(marlpde_column_params), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    if (params->model_flags & MARLPDE_MODEL_VAR_DPHI) marlpde::jac_probe_kernel<true><<<1, 32>>>(dy, dp, n_cells, dc, dJ);
    else marlpde::jac_probe_kernel<false><<<1, 32>>>(dy, dp, n_cells, dc, dJ);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(J_out, dJ, nb_J, cudaMemcpyDeviceToHost);
  cudaFree(dy);
  cudaFree(dc);
  cudaFree(dJ);
  cudaFree(dp);
  return e == cudaSuccess ? MARLPDE_OK : MARLPDE_ECUDA;
}
