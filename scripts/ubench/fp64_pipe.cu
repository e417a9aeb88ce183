// fp64_pipe.cu — micro-benchmark of the sm_100a fp64 pipe as the RK45 kernel uses it: issue rate of DFMA / DMUL / DADD /
// DSETP per SM sub-partition, and how far a given number of resident warps x independent chains gets towards it.
// Measurement tool (not part of the library).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_pipe scripts/ubench/fp64_pipe.cu && /tmp/fp64_pipe
#include <cuda_runtime.h>
#include <cstdio>

enum Op { FMA = 0, MUL = 1, ADD = 2, SETP = 3, MIXFM = 4, FMA_INT = 5, MUFU = 6 };

template <int OP, int CH>
__global__ void pipe_kernel(double* out, long long* cyc, int iters, double seed) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
  const double m = 1.0000001, c = 1e-9;
  int k = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (OP == FMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(m), "d"(c));
        if (OP == MUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(m));
        if (OP == ADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(c));
        if (OP == SETP) asm volatile("{ .reg .pred p; setp.lt.f64 p, %0, %1; selp.f64 %0, %0, %1, p; }" : "+d"(a[j]) : "d"(m));
        if (OP == MIXFM) {
          if (u & 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[j]) : "d"(m));
          else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(m), "d"(c));
        }
        if (OP == FMA_INT) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(m), "d"(c));
          asm volatile("mad.lo.s32 %0, %0, 3, 7;" : "+r"(k));
          asm volatile("xor.b32 %0, %0, 0x55;" : "+r"(k));
        }
        if (OP == MUFU) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a[j]));
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678 || k == 0x7fffffff) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int CH>
void run(const char* name, int warps, double* d, long long* dc) {
  const int iters = 2000;
  pipe_kernel<OP, CH><<<1, warps * 32>>>(d, dc, 10, 1.0);
  pipe_kernel<OP, CH><<<1, warps * 32>>>(d, dc, iters, 1.0);
  long long c = 0;
  cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
  const double n_inst = (double)iters * 8 * CH * warps * (OP == FMA_INT ? 1 : 1);   // fp64 warp instructions on the SM
  // cycles per fp64 warp instruction per sub-partition if the warps were spread evenly (warps/4 per SMSP)
  printf("%-8s warps %2d chains %d : %8.0f cycles, %.3f fp64 warp-inst/cycle/SM (peak 2.0 if 16 lanes/SMSP), %.2f cycles per inst per warp\n",
         name, warps, CH, (double)c, n_inst / (double)c, (double)c / ((double)iters * 8 * CH));
}

int main() {
  double* d;
  long long* dc;
  cudaMalloc(&d, 8);
  cudaMalloc(&dc, 8);
  const int W[] = {1, 4, 8, 10, 12, 16, 32};
  for (int w : W) {
    run<FMA, 1>("DFMA", w, d, dc);
    run<FMA, 2>("DFMA", w, d, dc);
    run<FMA, 4>("DFMA", w, d, dc);
    run<FMA, 8>("DFMA", w, d, dc);
  }
  for (int w : {4, 10, 16, 32}) {
    run<MUL, 8>("DMUL", w, d, dc);
    run<ADD, 8>("DADD", w, d, dc);
    run<SETP, 8>("DSETP+S", w, d, dc);
    run<MIXFM, 8>("FMA/MUL", w, d, dc);
    run<FMA_INT, 8>("FMA+2INT", w, d, dc);
    run<FMA_INT, 2>("FMA+2INT", w, d, dc);
    run<MUFU, 4>("MUFU.RCP", w, d, dc);
  }
  run<MUL, 1>("DMUL", 1, d, dc);
  run<ADD, 1>("DADD", 1, d, dc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
