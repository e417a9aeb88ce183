// tests/emu/simt_emu.cc — see simt_emu.h.  TEST INFRASTRUCTURE ONLY.
#include "simt_emu.h"

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "cuda_runtime.h"

namespace simt {

uint3_ g_threadIdx, g_blockIdx, g_blockDim = {1, 1, 1}, g_gridDim = {1, 1, 1};

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  bool done = false;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
bool failed = false;
std::vector<unsigned char> smem;
const std::function<void()>* body_fn = nullptr;
long long progress = 0;          // bumped whenever a rendezvous completes or a fiber finishes

void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
void fail(const char* what) {
  if (!failed) std::fprintf(stderr, "[simt_emu] %s (thread %d)\n", what, cur);
  failed = true;
  yield();                        // never resumed
}

// ---- block barrier ----------------------------------------------------------------------
struct BlockBar {
  int arrived = 0, count = 0, any = 0;
  long long gen = 0;
  int res_count = 0, res_or = 0;
} bb;

// ---- warp collectives: one rendezvous state per (warp, mask) ---------------------------
struct Rdv {
  unsigned mask = 0, arrived = 0;
  int op = -1, arg = 0;
  unsigned long long val[32], res[2][32];
  long long gen = 0;
};
std::map<unsigned long long, Rdv> rdv;   // key: (warp, mask) — collectives on different masks are independent

struct MBar {
  unsigned count = 0, pending = 0, phase = 0;
};
std::map<uint64_t*, MBar> mbars;

void trampoline() {
  (*body_fn)();
  fibers[cur].done = true;
  ++progress;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

unsigned char* dyn_smem() { return smem.data(); }

int block_barrier(int pred, int op) {
  (void)op;
  const long long g = bb.gen;
  bb.arrived += 1;
  bb.count += pred ? 1 : 0;
  bb.any |= pred ? 1 : 0;
  if (bb.arrived == (int)fibers.size()) {
    bb.res_count = bb.count;
    bb.res_or = bb.any;
    bb.arrived = bb.count = bb.any = 0;
    bb.gen += 1;
    ++progress;
  } else {
    while (bb.gen == g) yield();
  }
  return op == 1 ? bb.res_count : (op == 2 ? bb.res_or : 0);
}

unsigned long long warp_collective(unsigned mask, unsigned long long value, int op, int arg) {
  const int tid = cur, lane = tid & 31, warp = tid >> 5;
  const int n = (int)fibers.size();
  // lanes that do not exist in the block (partial last warp) cannot be named
  if (warp * 32 + 32 > n) mask &= (n - warp * 32 >= 32) ? 0xffffffffu : ((1u << (n - warp * 32)) - 1u);
  if (!((mask >> lane) & 1u)) fail("warp collective called by a lane that is not in its mask");
  Rdv& r = rdv[((unsigned long long)warp << 32) | mask];
  if (r.arrived == 0) {
    r.mask = mask;
    r.op = op;
    r.arg = arg;
  } else if (r.op != op || (op <= OP_SHFL_XOR && r.arg != arg)) {
    fail("lanes named by one mask reached different collectives");
  }
  r.val[lane] = value;
  r.arrived |= 1u << lane;
  const long long g = r.gen;
  if (r.arrived == mask) {
    unsigned long long* out = r.res[g & 1];
    unsigned long long acc_or = 0, acc_max = 0, all = 1, any = 0, ballot = 0;
    for (int l = 0; l < 32; ++l)
      if ((mask >> l) & 1u) {
        acc_or |= r.val[l];
        if (r.val[l] > acc_max) acc_max = r.val[l];
        all &= r.val[l] ? 1 : 0;
        any |= r.val[l] ? 1 : 0;
        ballot |= (r.val[l] ? 1ull : 0ull) << l;
      }
    for (int l = 0; l < 32; ++l) {
      if (!((mask >> l) & 1u)) continue;
      int src = l;
      switch (op) {
        case OP_SHFL_IDX: src = arg & 31; break;
        case OP_SHFL_UP: src = l - arg; break;
        case OP_SHFL_DOWN: src = l + arg; break;
        case OP_SHFL_XOR: src = l ^ arg; break;
        default: break;
      }
      switch (op) {
        case OP_SHFL_IDX:
        case OP_SHFL_UP:
        case OP_SHFL_DOWN:
        case OP_SHFL_XOR:
          // out-of-range or un-named source lane: CUDA returns the caller's own value (up/down) / undefined; own value here
          out[l] = (src >= 0 && src < 32 && ((mask >> src) & 1u)) ? r.val[src] : r.val[l];
          break;
        case OP_ANY: out[l] = any; break;
        case OP_ALL: out[l] = all; break;
        case OP_BALLOT: out[l] = ballot; break;
        case OP_MATCH_ANY: {
          unsigned long long m = 0;
          for (int k = 0; k < 32; ++k)
            if (((mask >> k) & 1u) && r.val[k] == r.val[l]) m |= 1ull << k;
          out[l] = m;
          break;
        }
        case OP_RED_OR: out[l] = acc_or; break;
        case OP_RED_MAX: out[l] = acc_max; break;
        default: out[l] = 0; break;
      }
    }
    r.arrived = 0;
    r.gen += 1;
    ++progress;
  } else {
    while (r.gen == g) yield();
  }
  return r.res[g & 1][lane];
}

void mbar_init(uint64_t* bar, unsigned count) { mbars[bar] = MBar{count, count, 0}; }
void mbar_arrive(uint64_t* bar) {
  MBar& m = mbars[bar];
  if (m.count == 0) fail("mbarrier used before init");
  if (--m.pending == 0) {
    m.pending = m.count;
    m.phase ^= 1u;
    ++progress;
  }
}
void mbar_wait(uint64_t* bar, unsigned parity) {
  MBar& m = mbars[bar];
  while (m.phase == (parity & 1u)) yield();     // the phase with this parity has not completed yet
}

int run_block(int n_threads, size_t dyn_smem_bytes, const std::function<void()>& body, int block, int grid) {
  if ((int)fibers.size() != n_threads) fibers.resize(n_threads);     // stacks are kept between blocks
  for (Fiber& f : fibers) f.done = false;
  rdv.clear();
  mbars.clear();
  bb = BlockBar();
  failed = false;
  progress = 0;
  if (smem.size() < dyn_smem_bytes + 64) smem.resize(dyn_smem_bytes + 64);
  std::fill(smem.begin(), smem.end(), 0);
  body_fn = &body;
  g_blockDim = {(unsigned)n_threads, 1, 1};
  g_blockIdx = {(unsigned)block, 0, 0};
  g_gridDim = {(unsigned)grid, 1, 1};
  for (int i = 0; i < n_threads; ++i) {
    Fiber& f = fibers[i];
    if (f.stack.size() != kStack) f.stack.resize(kStack);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &sched_ctx;
    makecontext(&f.ctx, trampoline, 0);
  }
  long long last_progress = -1;
  int idle_rounds = 0;
  for (;;) {
    int alive = 0;
    for (int i = 0; i < n_threads && !failed; ++i) {
      if (fibers[i].done) continue;
      ++alive;
      cur = i;
      g_threadIdx = {(unsigned)i, 0, 0};
      swapcontext(&sched_ctx, &fibers[i].ctx);
    }
    if (failed) return -1;
    if (alive == 0) return 0;
    if (progress == last_progress) {
      if (++idle_rounds > 4) {
        std::fprintf(stderr, "[simt_emu] deadlock: %d threads wait and nothing completes\n", alive);
        return -1;
      }
    } else {
      idle_rounds = 0;
      last_progress = progress;
    }
  }
}

void run_grid(int grid, int n_threads, size_t dyn_smem_bytes, const std::function<void()>& body) {
  for (int b = 0; b < grid; ++b)
    if (run_block(n_threads, dyn_smem_bytes, body, b, grid) != 0) {
      std::fprintf(stderr, "[simt_emu] block %d of %d failed\n", b, grid);
      std::abort();
    }
}

}  // namespace simt
