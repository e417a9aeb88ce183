// tests/emu/cuda_runtime.h — stand-in for <cuda_runtime.h> when the KERNEL SOURCES are compiled for the host by the
// SIMT emulator (tests/emu/simt_emu.h).  TEST INFRASTRUCTURE ONLY: it exists so that the control logic of a kernel
// (indexing, barriers, halo exchange, slot service) can be exercised without a GPU before the kernel is first
// launched on one.  The product library never includes this file and has no CPU path.
#pragma once
#ifndef MARLPDE_HOST_EMU
#error "tests/emu/cuda_runtime.h is only for -DMARLPDE_HOST_EMU builds"
#endif
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static            // block-scope shared variables: one CTA is emulated at a time
#define MARLPDE_DYN_SMEM(name) unsigned char* const name = ::simt::dyn_smem()

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct double2 { double x, y; };
struct float2 { float x, y; };
struct uint3_ { unsigned x, y, z; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }

#include <functional>
namespace simt {
void run_grid(int grid, int n_threads, size_t dyn_smem_bytes, const std::function<void()>& body);
unsigned char* dyn_smem();
extern uint3_ g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
// collectives (simt_emu.cc): every lane named in `mask` must call the same one
unsigned long long warp_collective(unsigned mask, unsigned long long value, int op, int arg);
int block_barrier(int pred, int op);          // op 0: plain, 1: count, 2: or
void mbar_init(uint64_t* bar, unsigned count);
void mbar_arrive(uint64_t* bar);
void mbar_wait(uint64_t* bar, unsigned parity);
enum { OP_SHFL_IDX, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_ANY, OP_ALL, OP_BALLOT, OP_MATCH_ANY, OP_RED_OR, OP_RED_MAX, OP_SYNC };
}  // namespace simt
#define threadIdx (::simt::g_threadIdx)
#define blockIdx (::simt::g_blockIdx)
#define blockDim (::simt::g_blockDim)
#define gridDim (::simt::g_gridDim)

inline unsigned long long simt_bits(double v) { unsigned long long b; std::memcpy(&b, &v, 8); return b; }
inline double simt_dbl(unsigned long long b) { double v; std::memcpy(&v, &b, 8); return v; }

inline int __double2hiint(double v) { return (int)(simt_bits(v) >> 32); }
inline int __double2loint(double v) { return (int)(simt_bits(v) & 0xffffffffull); }
inline double __hiloint2double(int hi, int lo) { return simt_dbl(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }
inline double __longlong_as_double(long long b) { return simt_dbl((unsigned long long)b); }
inline long long __double_as_longlong(double v) { return (long long)simt_bits(v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
// round-to-nearest arithmetic that the compiler may not contract (brent.cuh)
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
using std::signbit;
using std::fabs; using std::fma; using std::fmax; using std::fmin; using std::isfinite; using std::nextafter; using std::sqrt;

inline double __shfl_sync(unsigned m, double v, int src) { return simt_dbl(simt::warp_collective(m, simt_bits(v), simt::OP_SHFL_IDX, src)); }
inline int __shfl_sync(unsigned m, int v, int src) { return (int)simt::warp_collective(m, (unsigned long long)(unsigned)v, simt::OP_SHFL_IDX, src); }
inline double __shfl_up_sync(unsigned m, double v, int d) { return simt_dbl(simt::warp_collective(m, simt_bits(v), simt::OP_SHFL_UP, d)); }
inline double __shfl_down_sync(unsigned m, double v, int d) { return simt_dbl(simt::warp_collective(m, simt_bits(v), simt::OP_SHFL_DOWN, d)); }
inline double __shfl_xor_sync(unsigned m, double v, int d) { return simt_dbl(simt::warp_collective(m, simt_bits(v), simt::OP_SHFL_XOR, d)); }
inline int __any_sync(unsigned m, int p) { return (int)simt::warp_collective(m, p != 0, simt::OP_ANY, 0); }
inline int __all_sync(unsigned m, int p) { return (int)simt::warp_collective(m, p != 0, simt::OP_ALL, 0); }
inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)simt::warp_collective(m, p != 0, simt::OP_BALLOT, 0); }
inline unsigned __match_any_sync(unsigned m, int v) { return (unsigned)simt::warp_collective(m, (unsigned long long)(unsigned)v, simt::OP_MATCH_ANY, 0); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)simt::warp_collective(m, v, simt::OP_RED_OR, 0); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)simt::warp_collective(m, v, simt::OP_RED_MAX, 0); }
inline void __syncwarp(unsigned m = 0xffffffffu) { simt::warp_collective(m, 0, simt::OP_SYNC, 0); }
inline void __syncthreads() { simt::block_barrier(0, 0); }
inline int __syncthreads_count(int p) { return simt::block_barrier(p != 0, 1); }
inline int __syncthreads_or(int p) { return simt::block_barrier(p != 0, 2); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
inline void __threadfence() {}
template <class T> inline T __ldcg(const T* p) { return *p; }
