// probe.cu — measures the fp64 FMA peak of the device the integrator runs on.
// The RK45 kernel is fp64-pipe bound; MEASURED_PEAKS.json carries no fp64 entry,
// so bench.py measures the roofline denominator on the box with this kernel: every thread runs
// 8 independent DFMA chains (no memory traffic), 2048 threads per SM.
#include <cuda_runtime.h>

#include "../../include/marlpde_b200.h"
#include "lheureux_device.cuh"

namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;   // keeps the chains alive without a store on the hot path
}


// element-wise evaluation of the table-driven device maths (accuracy tests only)
__global__ void math_probe_kernel(int op, const double* x, int n, double* out) {
  __shared__ __align__(16) unsigned char tab_raw[marlpde::fm::kTableBytes];
  const marlpde::fm::Tables tb = marlpde::fm::stage_tables(tab_raw, threadIdx.x, blockDim.x);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (op) {
    case 0: r = marlpde::fm::log(tb, v); break;
    case 1: r = marlpde::fm::exp(tb, v); break;
    case 2: r = marlpde::fm::expm1(tb, v); break;
    case 3: r = marlpde::fm::rcp(v); break;
    case 4: r = marlpde::fm::div(1.0 + v, v); break;
    default: r = marlpde::fv_sigma(tb, v, v, 1e-2, 1e2); break;
  }
  out[i] = r;
}

}  // namespace

extern "C" int marlpde_probe_fp64_peak(int device, int iters, int repeats, double* tflops) {
  if (!tflops || iters <= 0 || repeats <= 0) return MARLPDE_EINVAL;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return MARLPDE_ENODEVICE;
  if (device < 0 || device >= n) return MARLPDE_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return MARLPDE_ECUDA;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return MARLPDE_ECUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8;
  dfma_peak_kernel<<<blocks, 256>>>(d, iters / 8 + 1, 1.0);  // warm-up
  double best = 0.0;
  for (int r = 0; r < repeats; ++r) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, 256>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? MARLPDE_OK : MARLPDE_ECUDA;
}

extern "C" int marlpde_probe_math(int op, const double* x, int n, double* out, int device) {
  if (!x || !out || n < 0 || op < 0 || op > 5) return MARLPDE_EINVAL;
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) return MARLPDE_ENODEVICE;
  if (device < 0 || device >= nd || cudaSetDevice(device) != cudaSuccess) return MARLPDE_EINVAL;
  if (n == 0) return MARLPDE_OK;
  double *dx = nullptr, *dout = nullptr;
  if (cudaMalloc(&dx, sizeof(double) * n) != cudaSuccess) return MARLPDE_ECUDA;
  if (cudaMalloc(&dout, sizeof(double) * n) != cudaSuccess) { cudaFree(dx); return MARLPDE_ECUDA; }
  cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  math_probe_kernel<<<(n + 255) / 256, 256>>>(op, dx, n, dout);
  cudaError_t e = cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dout);
  return e == cudaSuccess ? MARLPDE_OK : MARLPDE_ECUDA;
}
