// tests/emu/simt_emu.h — a tiny SIMT emulator: runs ONE thread block of a CUDA kernel, compiled for the host, with one
// fiber (ucontext) per CUDA thread.  A fiber runs until it reaches a synchronising primitive (__syncthreads*, a
// *_sync warp collective, an mbarrier wait), deposits its contribution and yields; when every participant has
// arrived the results are formed and the fibers continue.  Warps therefore do NOT run in lock-step between
// primitives — like independent thread scheduling on the GPU — which is exactly what exposes a missing barrier or a
// collective that not all named lanes reach (the emulator aborts with a message instead of hanging).
// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h next to this file).
#pragma once
#include <cstddef>
#include <functional>

namespace simt {
// run `body` once per thread of a block of `n_threads` threads with `dyn_smem_bytes` of dynamic shared memory;
// returns 0, or -1 after a deadlock / mismatched collective (message on stderr)
int run_block(int n_threads, size_t dyn_smem_bytes, const std::function<void()>& body, int block = 0, int grid = 1);
// the blocks of a grid, one after the other (kernels without inter-block synchronisation); aborts the process on failure
void run_grid(int grid, int n_threads, size_t dyn_smem_bytes, const std::function<void()>& body);
}  // namespace simt
