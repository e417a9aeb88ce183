// cluster_sync.cu — micro-benchmark: what a rendezvous of the two CTAs of a thread-block cluster costs on sm_100a in the
// shape the paired RK45 kernel used it (352 threads per CTA, one cluster per SM pair, a block of dependent fp64 work
// between rendezvous, a few spilled values re-read from local memory after each one).  Variants:
//   0  __syncthreads() only (no cluster coupling: the floor)
//   1  barrier.cluster.arrive.release + wait.acquire   (MEMBAR.ALL.GPU ... UCGABAR_ARV / UCGABAR_WAIT + CCTL.IVALL)
//   2  barrier.cluster.arrive.relaxed + wait           (UCGABAR_ARV / UCGABAR_WAIT + CCTL.IVALL)
//   3  __syncthreads() + one remote mbarrier arrive (relaxed) per CTA + CTA-scope try_wait by every thread
// Measurement tool (not part of the library).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_sync scripts/ubench/cluster_sync.cu && /tmp/cluster_sync
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int V>
__global__ void __launch_bounds__(352, 1) sync_kernel(long long* cyc, double* out, int iters, int work, double seed) {
  __shared__ uint64_t bar[2];
  double spill[24];                       // local memory: indexed dynamically so it cannot live in registers
  for (int i = 0; i < 24; ++i) spill[i] = seed + i + threadIdx.x;
  unsigned rank = 0;
  if (V != 0) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1u) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (V != 0) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  unsigned remote = 0;
  unsigned parity[2] = {0u, 0u};
  double a = seed + threadIdx.x, b = seed * 0.5;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int w = 0; w < work; ++w) {      // two dependent chains, like the two cells of a thread
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a) : "d"(1.0000001), "d"(1e-9));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(b) : "d"(0.9999999), "d"(1e-9));
    }
    a += spill[(it + threadIdx.x) % 24];  // re-read "spilled" values: L1 hits unless the rendezvous emptied the L1
    b += spill[(it * 7 + threadIdx.x) % 24];
    spill[it % 24] = a;
    if (V == 0) {
      __syncthreads();
    } else if (V == 1) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (V == 2) {
      asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    } else {
      const int k = it & 1;
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&bar[k])), "r"(rank ^ 1u));
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
      unsigned done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar[k])), "r"(parity[k]) : "memory");
      } while (!done);
      parity[k] ^= 1u;
    }
  }
  const long long t1 = clock64();
  if (V != 0) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (a + b == 12345.678) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int V>
static double run(long long* dc, double* dout, int iters, int work) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(352);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = V == 0 ? 1 : 2;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, sync_kernel<V>, dc, dout, iters, work, 1.0);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
      printf("variant %d failed: %s\n", V, cudaGetErrorString(e));
      return -1.0;
    }
  }
  long long c = 0;
  cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
  return (double)c / iters;
}

int main() {
  long long* dc;
  double* dout;
  cudaMalloc(&dc, 8);
  cudaMalloc(&dout, 8);
  const int iters = 2000;
  const char* names[4] = {"__syncthreads", "cluster release/acquire", "cluster relaxed arrive", "syncthreads + remote mbarrier"};
  for (int work : {0, 300, 2000}) {      // cycles of fp64 work between two rendezvous: ~0 / ~5 k (a stage trip) / ~33 k (an attempt)
    double base = 0;
    for (int v = 0; v < 4; ++v) {
      double c = v == 0 ? run<0>(dc, dout, iters, work) : v == 1 ? run<1>(dc, dout, iters, work)
                 : v == 2 ? run<2>(dc, dout, iters, work) : run<3>(dc, dout, iters, work);
      if (v == 0) base = c;
      printf("work %4d x 2 DFMA  %-32s %8.0f cycles per iteration  (+%.0f over __syncthreads)\n", work, names[v], c, c - base);
    }
  }
  return 0;
}
