// mbar.cuh — split block barrier (arrive now, wait later) on a shared-memory mbarrier object.
// __syncthreads() makes every warp idle until the slowest one has published its values; with
// arrive / wait a thread signals "my values are published", keeps computing what does not need the
// neighbours, and only then waits — the barrier latency hides behind useful fp64 work.
#pragma once
#include <cstdint>
#include <cstring>

// the kernel's dynamic shared memory (the host emulator of tests/emu/ substitutes its own buffer)
#ifndef MARLPDE_DYN_SMEM
#define MARLPDE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// kernel launch (the host emulator runs the blocks of the grid one after the other)
#ifdef MARLPDE_HOST_EMU
#define MARLPDE_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::simt::run_grid((int)(grid), (int)(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#else
#define MARLPDE_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace marlpde {

#ifdef MARLPDE_HOST_EMU   // tests/emu/: kernel control logic on the host, test infrastructure only
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) { ::simt::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { ::simt::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) { ::simt::mbar_wait(bar, parity); }
// bulk copies complete at once on the host: expect_tx is then just the issuing thread's arrival
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned) { ::simt::mbar_arrive(bar); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t*) { std::memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) { std::memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_commit_wait() {}
__device__ __forceinline__ void fence_proxy_async() {}
// thread-block cluster of two (run_cluster2): the peer's shared memory is another host buffer
typedef unsigned char* raddr_t;
__device__ __forceinline__ unsigned cluster_ctarank() { return ::simt::cluster_rank(); }
__device__ __forceinline__ void cluster_sync() { ::simt::cluster_barrier(); }
__device__ __forceinline__ raddr_t peer_addr(const void* p, unsigned rank) {
  return ::simt::peer_smem(rank) + (reinterpret_cast<const unsigned char*>(p) - ::simt::dyn_smem());
}
template <class T> __device__ __forceinline__ void st_peer(raddr_t a, T v) { std::memcpy(a, &v, sizeof(T)); }
__device__ __forceinline__ void red_or_peer(raddr_t a, unsigned v) { *reinterpret_cast<unsigned*>(a) |= v; }
__device__ __forceinline__ void mbar_arrive_peer(raddr_t bar) { ::simt::mbar_arrive(reinterpret_cast<uint64_t*>(bar)); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) { ::simt::mbar_wait(bar, parity); }
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// release semantics at CTA scope: shared-memory writes made before the arrive are visible after the wait
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// ---- 1-D bulk copies of the TMA unit (cp.async.bulk, SASS UBLKCP): addresses and sizes are multiples of 16 bytes --------
// makes a fresh mbarrier.init visible to the async proxy before a bulk copy signals the barrier
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival + the number of bytes the bulk copies will deliver: the phase completes when both are in
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {      // all bulk stores of this thread have been written
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// generic-proxy shared-memory writes become visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- thread-block cluster of two CTAs: distributed shared memory ---------------------------------------------------------
// An address in the PEER CTA's shared memory (shared::cluster window): remote stores, a remote reduction and a remote
// mbarrier arrival.  Ordering: remote stores are made visible to the peer by a later release at cluster scope (the remote
// arrive below, or barrier.cluster.arrive.release) that the consumer acquires at cluster scope (mbar_wait_cluster,
// barrier.cluster.wait.acquire).
typedef unsigned raddr_t;
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
#ifndef MARLPDE_PAIR_SYNC
#define MARLPDE_PAIR_SYNC 0      // 1: experiment — relaxed cluster barrier / remote arrive, CTA-scope waits (no MEMBAR.GPU, no CCTL.IVALL)
#endif
__device__ __forceinline__ void cluster_sync() {       // every thread of both CTAs
#if MARLPDE_PAIR_SYNC == 1
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ raddr_t peer_addr(const void* p, unsigned rank) {
  raddr_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_peer(raddr_t a, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void st_peer(raddr_t a, int v) { asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st_peer(raddr_t a, long long v) { asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void red_or_peer(raddr_t a, unsigned v) {
  asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_peer(raddr_t bar) {
#if MARLPDE_PAIR_SYNC == 1
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#else
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
#if MARLPDE_PAIR_SYNC == 1
  mbar_wait(bar, parity);
  return;
#endif
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
#endif

}  // namespace marlpde
