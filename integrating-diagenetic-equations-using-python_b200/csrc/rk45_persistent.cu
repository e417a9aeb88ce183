// rk45_persistent.cu — batched adaptive Dormand-Prince 5(4) for sediment columns, sm_100a.
//
// Replaces the loop `scipy.integrate.solve_ivp(eq.fun_numba, ..., method="RK45")`
// (reference call site marlpde/Evolve_scenario.py:104-109; algorithm: scipy/integrate/_ivp/rk.py
// `RungeKutta._step_impl`, `rk_step`, `RkDenseOutput`; ivp.py main loop for t_eval sampling).
//
// Design (B200-first):
//   * persistent CTAs, one per SM; each CTA integrates C columns side by side ("slots");
//     a thread owns TWO adjacent depth cells of one column (cells 2k, 2k+1): N=200 gives
//     3 columns x 100 threads = 300 of 320 lanes busy.  Two cells per thread double the
//     instruction-level parallelism (the register file only holds ~3 warps per scheduler) and
//     halve every per-instruction overhead (addresses, constants, control) per cell;
//   * the whole integration of a column happens on-chip.  In REGISTERS: the state y, K1 (which
//     is last step's K7 — FSAL costs nothing), the current stage input and the current stage
//     derivative of the thread's two cells.  In SHARED memory: K2..K5 as double2 per thread
//     (K6 re-uses K2's slot once K2 is dead) and a double-buffered halo-exchange tile through
//     which a thread sees cell 2k-1 and cell 2k+2.  48 kB per column at N=200.  All shared
//     arrays use the compile-time stride TP, so every address is `tid*8|16 + immediate`.
//     HBM is touched only to load y0 and to store snapshots / the final state;
//   * every column runs its own adaptive controller (own t, h, accept/reject); the error
//     norm is a warp-shuffle butterfly over aligned thread groups plus one shared-memory hop,
//     summed in a fixed, slot-independent order: all threads of a column take bit-identical
//     decisions and a column's trajectory does not depend on what else shares the CTA;
//   * one block barrier per RHS evaluation (the halo tile is double buffered) and one for
//     the norm: 7 barriers per step attempt;
//   * finished slots claim the next column from a global atomic queue, so columns with
//     different step counts do not leave SMs idle.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "brent.cuh"
#include "dopri.cuh"
#include "events.cuh"
#include "lheureux_device.cuh"
#include "mbar.cuh"
#include "rk45_persistent.cuh"

namespace marlpde {


// ---- shared memory carve-up -------------------------------------------------------------
// K[4][5][TP] double2 | tileE[2][5][TP] | tileO[2][5][TP] | grp[TP] | log/exp tables |
// consts[C] | ctl[C] | slot_col[C] | svc flag
struct SlotCtl {          // per-slot counters, written by the slot's leader thread only
  long long n_acc, n_rej, nfev;
  long long budget;       // the claim ends at the first accepted step with this many attempts (0: unlimited) ...
  long long used;         // ... counted from here (attempts of earlier claims on the column in this launch);
                          // both read by the whole slot after the claim barrier
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) / 16 * 16; }

template <int TP>
struct Smem {
  static constexpr size_t off_K = 0;
  // halo exchange: two full stage-input vectors [2][5][TP] per parity (even / odd cells)
  static constexpr size_t off_tE = off_K + sizeof(double2) * 4 * 5 * TP;
  static constexpr size_t off_tO = off_tE + sizeof(double) * 2 * 5 * TP;
  static constexpr size_t off_grp = off_tO + sizeof(double) * 2 * 5 * TP;
  static constexpr size_t off_tab = off_grp + sizeof(double) * TP;
  static constexpr size_t off_var = off_tab + fm::kTableBytes;          // per-slot arrays start here
  static constexpr size_t slot_bytes = (sizeof(ColumnConsts) + 15) / 16 * 16 + (sizeof(SlotCtl) + 15) / 16 * 16 + 32;
  __host__ __device__ static size_t total(int C) { return off_var + slot_bytes * (size_t)C + 32; }   // + svc flag, mbarrier
};

static int group_log2(int threads_per_column) {
  int logG = 0;
  while (logG < 5 && (threads_per_column % (2 << logG)) == 0) ++logG;
  return logG;
}

template <int TP>
static int columns_per_cta_t(int n_cells, int smem_budget) {
  const int Hc = (n_cells + 1) / 2;
  if (n_cells < 32 || Hc > TP) return 0;
  int C = TP / Hc;
  while (C > 0 && Smem<TP>::total(C) > (size_t)smem_budget) --C;
  return C;
}

// Shape of the kernel (register file: 16K registers per SM sub-partition, so the register cap follows from the warps
// per sub-partition, not from the thread count alone): 320 threads = 10 warps, <= 168 registers, y and K1 in registers
// (n_cells <= 640; N=200: 3 columns per CTA).  Shapes that were measured and dropped (profiles/r02a_ab_candidates.log,
// DESIGN.md section 9): y and K1 in shared memory (same speed), 13 warps / 128 registers / 4 columns with the halo by
// warp shuffles (-12 %), 128-thread CTAs with one column each, 3 per SM (-20 %), 4 cells per thread / 8 warps / 5
// columns (+1.5 %, not worth a second kernel).
constexpr int kRk45Threads = 320;

int rk45_columns_per_cta(int n_cells, int smem_budget) { return columns_per_cta_t<kRk45Threads>(n_cells, smem_budget); }

int rk45_max_cells() { return 2 * kRk45Threads; }


struct Rk45Args {
  double* g_y;
  const marlpde_column_params* g_params;
  marlpde_column_state* g_state;
  const double* g_t_eval;
  double* g_snap;
  int32_t* g_queue;
  int32_t* g_ev_counts;      // [n_columns][7]
  double* g_ev_times;        // [n_columns][7][event_capacity]
  int n_columns, N, C, logG;
  int n_quanta;                   // work items per column (1: a claim covers the column's whole step budget)
  int n_whole;                    // the first n_whole columns are claimed whole, only the others are cut into quanta
  long long quantum;              // step attempts per work item when n_quanta > 1
  unsigned long long warp_perm;   // logical warp of physical warp w = (warp_perm >> 4 w) & 15 (see rk45_warp_perm)
  marlpde_rk45_options opt;
};

// VD: the instantiation for batches with MARLPDE_MODEL_VAR_DPHI columns (opt.flags & MARLPDE_FLAG_VAR_DPHI), see rhs_pair_own
template <int TP, bool VD>
__global__ void __launch_bounds__((TP + 31) / 32 * 32, 1) rk45_persistent_kernel(const Rk45Args A) {
  MARLPDE_DYN_SMEM(smem_raw);
  using L = Smem<TP>;
  // Logical thread index.  Physical warps can be dealt out to the column ranges in any order; the order decides WHICH
  // scheduler (physical warp & 3) runs the warps whose lanes lie in the dissolution zone and therefore evaluate one
  // more real power per RHS.  Ten warps over four schedulers is 3-3-2-2: the automatic order puts those warps on the
  // schedulers that hold fewer warps first (r01g: +1.5 %).  Results do not depend on the order.
  __shared__ unsigned char sPerm[16];
  __shared__ unsigned sHeavy;
  const int n_warps = (int)(blockDim.x >> 5);
  const bool auto_order = A.warp_perm == ~0ull && n_warps <= 16 && A.n_columns > 0;
  if (threadIdx.x == 0) sHeavy = 0u;
  __syncthreads();
  {
    bool hv = false;
    const int t = (int)threadIdx.x, Hc_ = (A.N + 1) >> 1;
    if (auto_order && t < A.C * Hc_) {
      const int mlo = A.g_params[0].mask_lo, mhi = A.g_params[0].mask_hi;   // the sweep's first column stands for all
      const int c0 = 2 * (t % Hc_);
      hv = (c0 >= mlo && c0 < mhi) || (c0 + 1 >= mlo && c0 + 1 < mhi);
    }
    const unsigned b = __ballot_sync(0xffffffffu, hv);
    if ((t & 31) == 0 && b) atomicOr(&sHeavy, 1u << (t >> 5));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!auto_order) {
      const unsigned long long p = A.warp_perm == ~0ull ? 0xfedcba9876543210ull : A.warp_perm;
      for (int w = 0; w < 16; ++w) sPerm[w] = (unsigned char)((p >> (4 * w)) & 15ull);
    } else {
      const unsigned heavy = sHeavy;
      int per_sched[4] = {0, 0, 0, 0};
      for (int w = 0; w < n_warps; ++w) per_sched[w & 3] += 1;
      // physical warps, those on the emptier schedulers first; logical warps, the heavy ones first
      int np = 0, nl = 0;
      unsigned char phys[16], logical[16];
      for (int cnt = 0; cnt <= 4; ++cnt)
        for (int w = 0; w < n_warps; ++w)
          if (per_sched[w & 3] == cnt) phys[np++] = (unsigned char)w;
      for (int pass = 0; pass < 2; ++pass)
        for (int w = 0; w < n_warps; ++w)
          if (((heavy >> w) & 1u) == (pass == 0 ? 1u : 0u)) logical[nl++] = (unsigned char)w;
      for (int i = 0; i < n_warps; ++i) sPerm[phys[i]] = logical[i];
    }
  }
  __syncthreads();
  const int tid = (int)sPerm[threadIdx.x >> 5] * 32 + (int)(threadIdx.x & 31);
  const int N = A.N, C = A.C;
  const int Hc = (N + 1) >> 1;                     // threads per column
  double2* const sK = reinterpret_cast<double2*>(smem_raw + L::off_K) + tid;         // [4][5][TP]
  double* const sE = reinterpret_cast<double*>(smem_raw + L::off_tE);               // [2][5][TP] even cells
  double* const sO = reinterpret_cast<double*>(smem_raw + L::off_tO);               // [2][5][TP] odd cells
  double* const sGrp = reinterpret_cast<double*>(smem_raw + L::off_grp);            // [TP >> logG]
  unsigned char* const var = smem_raw + L::off_var;
  const size_t consts_sz = align16(sizeof(ColumnConsts)), ctl_sz = align16(sizeof(SlotCtl));
  int* const sSlotCol = reinterpret_cast<int*>(var + (consts_sz + ctl_sz) * C);     // [C]
  unsigned* const sEv = reinterpret_cast<unsigned*>(sSlotCol + C);                  // [2][C] bits of y_new
  unsigned* const sEv0 = sEv + 2 * C;                                               // [C] bits of a fresh y
  int* const sSvc = reinterpret_cast<int*>(sEv0 + C);
  // split barrier of the stage loop (16-byte aligned slot at the very end of the carve-up)
  uint64_t* const sBar = reinterpret_cast<uint64_t*>(var + ((L::slot_bytes * (size_t)C + 15) / 16) * 16 + 16);

  const fm::Tables tb = fm::stage_tables(smem_raw + L::off_tab, tid, blockDim.x);
  const bool active = tid < C * Hc;
  const int slot = active ? tid / Hc : 0;
  const int pr = active ? tid - slot * Hc : 0;     // pair index inside the column
  const int cell0 = 2 * pr;
  const bool has1 = cell0 + 1 < N;                 // false only for the last thread of an odd-N column
  const bool first = pr == 0, last = pr == Hc - 1;
  const ColumnConsts& kc = *reinterpret_cast<const ColumnConsts*>(var + consts_sz * slot);
  SlotCtl& ctl = *reinterpret_cast<SlotCtl*>(var + consts_sz * C + ctl_sz * slot);
  const bool leader = active && first;
  // halo reads: cell 2k-1 is the odd cell of thread tid-1, cell 2k+2 the even cell of thread tid+1
  const double* const haloM = sO + (first ? tid : tid - 1);
  const double* const haloP = sE + (last ? tid : tid + 1);
  double* const myE = sE + tid;
  double* const myO = sO + tid;
  // Error-norm reduction tree, identical for every slot so that a column's trajectory does not
  // depend on where it is scheduled: threads are summed in aligned groups of G = 2^k lanes
  // (G = largest power of two <= 32 dividing Hc, hence dividing every slot base) by an xor
  // butterfly, then the Hc/G group sums are added in order from shared memory.
  const int logG = A.logG;
  const int G = 1 << logG;
  const int nGroups = Hc >> logG;
  const double* const grpRow = sGrp + ((slot * Hc) >> logG);
  const double inv_n = 1.0 / (double)(5 * N);
  // event monitors: lanes of my slot inside my warp (slots are not warp aligned)
  const bool ev_on = (A.opt.flags & MARLPDE_FLAG_EVENTS) != 0;
  const unsigned peers = __match_any_sync(0xffffffffu, active ? slot : -1);
  const bool peer_lead = (tid & 31) == (__ffs(peers) - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
  unsigned ev_prev = 0, ev_new_s = 0, ev_todo = 0;   // predicate bits at y / at the parked y_new; monitors to locate
  bool need_prev = false, parked = false;
  double factor_s = 1.0;
  unsigned it = 0;                                    // loop counter (parity selects the sEv buffer)

  // per-thread column state
  int col = -1;                 // column index being integrated by my slot, -1 = idle
  bool exhausted = !active;     // the work queue ran dry (padding lanes never claim)
  bool rejected = false;
  double t = 0.0, h_abs = 0.0, h = 0.0, t_new = 0.0;
  int next_eval = 0;
  int attempts = 0;             // step attempts made for this column in this launch
  long long budget = 0;         // the current claim ends at the first accepted step with attempts >= budget (0: unlimited)
  int pending_item = -1;        // leader: work item whose column was busy at the last claim
  int release_col = -1;         // leader: column to unlock once the slot's stores have passed a block barrier
  bool in_mask[2] = {false, false};
  double y[5][2], k1[5][2], c[5][2], r[5][2];
#pragma unroll
  for (int f = 0; f < 5; ++f)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      y[f][q] = 0.5;
      k1[f][q] = 0.0;
      c[f][q] = 0.5;
      r[f][q] = 0.0;
    }

  auto tile_store = [&](int b) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      myE[(b * 5 + f) * TP] = c[f][0];
      myO[(b * 5 + f) * TP] = c[f][1];
    }
  };
  // raw halo values of my pair for stage parity b: (odd cell of thread tid-1, even cell of thread tid+1)
  auto halo_load = [&](int b, int f, double& hm, double& hp) {
    hm = haloM[(b * 5 + f) * TP];
    hp = haloP[(b * 5 + f) * TP];
  };
  auto Kst = [&](int s, int f, double v0, double v1) { sK[(s * 5 + f) * TP] = make_double2(v0, v1); };
  auto Kld = [&](int s, int f) -> double2 { return sK[(s * 5 + f) * TP]; };
  // the state y and K1 of the thread's two cells live in registers
  auto Yld = [&](int f) -> double2 { return make_double2(y[f][0], y[f][1]); };
  auto K1ld = [&](int f) -> double2 { return make_double2(k1[f][0], k1[f][1]); };
  auto K1st = [&](int f, double v0, double v1) {
    k1[f][0] = v0;
    k1[f][1] = v1;
  };
  auto Yst = [&](int f, double v0, double v1) {
    y[f][0] = v0;
    y[f][1] = v1;
  };

  // scipy _step_impl: min_step = 10 * |nextafter(t, inf) - t| ; clamp h_abs at the start of a step
  auto min_step_at = [&](double tt) { return 10.0 * fabs(nextafter(tt, (double)INFINITY) - tt); };
  auto begin_step = [&]() {
    const double ms = min_step_at(t);
    if (h_abs > A.opt.max_step) h_abs = A.opt.max_step;
    else if (h_abs < ms) h_abs = ms;
    rejected = false;
  };
  auto begin_attempt = [&]() -> bool {   // false: TOO_SMALL_STEP
    if (h_abs < min_step_at(t)) return false;
    h = h_abs;
    t_new = t + h;
    if (t_new - A.opt.t_bound > 0.0) t_new = A.opt.t_bound;
    h = t_new - t;
    h_abs = fabs(h);
    return true;
  };
  auto retire = [&](int status) {  // store the column's end point and free the slot
    double* gy = A.g_y + (size_t)col * 5 * N + cell0;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const double2 yv = Yld(f);
      gy[(size_t)f * N] = yv.x;
      if (has1) gy[(size_t)f * N + 1] = yv.y;
    }
    if (leader) {
      marlpde_column_state st;
      st.t = t;
      st.h_abs = h_abs;
      st.n_accepted = ctl.n_acc;
      st.n_rejected = ctl.n_rej;
      st.nfev = ctl.nfev;
      st.status = status;
      st.next_eval = next_eval;
      A.g_state[col] = st;
      if (A.n_quanta > 1) {
        A.g_queue[1 + A.n_columns + col] = attempts;
        release_col = col;
      }
    }
    col = -1;
#pragma unroll
    for (int f = 0; f < 5; ++f) c[f][0] = c[f][1] = 0.5;   // idle lanes evaluate the RHS on benign values
  };
  auto stage2_input = [&]() {
    const double ha = h * dp::a21;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const double2 yv = Yld(f), kv = K1ld(f);
      c[f][0] = fma(ha, kv.x, yv.x);
      c[f][1] = fma(ha, kv.y, yv.y);
    }
    tile_store(1);
  };
  // quartic dense output of the step just computed (scipy RkDenseOutput): y(t + x h) for my two cells;
  // needs K1 (k1), K3..K5 and K6 (shared memory) and K7 (= r), i.e. must run before the commit
  auto interp_all = [&](double x, double (&out)[5][2]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const double2 K3 = Kld(1, f), K4 = Kld(2, f), K5 = Kld(3, f), K6 = Kld(0, f), yv = Yld(f), kv = K1ld(f);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double v3 = q ? K3.y : K3.x, v4 = q ? K4.y : K4.x, v5 = q ? K5.y : K5.x, v6 = q ? K6.y : K6.x;
        const double v1 = q ? kv.y : kv.x;
        double qq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double sacc = dp::P[0][j] * v1;
          sacc = fma(dp::P[2][j], v3, sacc);
          sacc = fma(dp::P[3][j], v4, sacc);
          sacc = fma(dp::P[4][j], v5, sacc);
          sacc = fma(dp::P[5][j], v6, sacc);
          qq[j] = fma(dp::P[6][j], r[f][q], sacc);
        }
        const double poly = x * (qq[0] + x * (qq[1] + x * (qq[2] + x * qq[3])));
        out[f][q] = fma(h, poly, q ? yv.y : yv.x);
      }
    }
  };
  // an accepted step becomes the column's state: t_eval samples, y <- y_new, K1 <- K7 (FSAL)
  auto commit = [&](double factor) {
    // dense output for t_eval points in (t, t_new] (ivp.py: searchsorted side='right')
    while (next_eval < A.opt.n_eval) {
      const double te = A.g_t_eval[next_eval];
      if (!(te <= t_new)) break;
      double ys[5][2];
      interp_all((te - t) / h, ys);
      double* gs = A.g_snap + ((size_t)col * A.opt.n_eval + next_eval) * 5 * N + cell0;
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        gs[(size_t)f * N] = ys[f][0];
        if (has1) gs[(size_t)f * N + 1] = ys[f][1];
      }
      ++next_eval;
    }
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      Yst(f, c[f][0], c[f][1]);
      K1st(f, r[f][0], r[f][1]);
    }
    t = t_new;
    h_abs *= factor;
    if (leader) ctl.n_acc += 1;
    if (t >= A.opt.t_bound) {
      retire(MARLPDE_STATUS_FINISHED);
    } else if (budget > 0 && (long long)attempts >= budget) {
      retire(MARLPDE_STATUS_STEP_BUDGET);
    } else {
      begin_step();
      if (!begin_attempt()) retire(MARLPDE_STATUS_STEP_TOO_SMALL);
    }
  };
  // value of monitor k on my two cells at y(t + x h), in "min form" (a max is the min of the negation)
  auto event_partial = [&](int k, double x) -> double {
    double ys[5][2];
    interp_all(x, ys);
    double v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double Phi = ys[4][q];
      double UW = 0.0;
      if (k >= 5) {   // same arithmetic as rhs_pair, so detection and location agree
        const double F = 1.0 - fm::exp(tb, fma(-10.0, fm::rcp3(Phi), 10.0));
        const double Phi2 = Phi * Phi;
        UW = k == 5 ? fma(kc.rhorat * (Phi2 * Phi), F * fm::rcp3(1.0 - Phi), kc.presum)
                    : -fma(-kc.rhorat * Phi2, F, kc.presum);
      }
      switch (k) {
        case 0: v[q] = fmin(fmin(fmin(ys[0][q], ys[1][q]), fmin(ys[2][q], ys[3][q])), ys[4][q]); break;
        case 1: v[q] = ys[0][q]; break;
        case 2: v[q] = ys[1][q]; break;
        case 3: v[q] = -(ys[0][q] + ys[1][q]); break;
        case 4: v[q] = -Phi; break;
        default: v[q] = UW; break;
      }
    }
    return has1 ? fmin(v[0], v[1]) : v[0];
  };

  // per-slot constants, counters and flags start as zeros: slots that never receive a column still
  // run the (ignored) RHS evaluations of their lanes, on benign values
  for (int i = tid; i < (int)((L::slot_bytes * (size_t)C + 16) / 4); i += blockDim.x)
    reinterpret_cast<int*>(var)[i] = 0;
  if (tid == 0) mbar_init(sBar, blockDim.x);
  unsigned bar_parity = 0;
  __syncthreads();

  bool fresh = false;           // column just loaded: K1 = f(y) still to be evaluated (stage i = 0)
  for (;;) {
    // ================= stage 2 input + slot service =======================================
    ++it;
    if (col >= 0 && !parked) stage2_input();
    if (leader && col < 0 && !exhausted) atomicOr(sSvc, 1);
    int nlive = __syncthreads_count(col >= 0);
    auto release = [&]() {      // the barrier above ordered the slot's y / state stores before this thread
      if (release_col >= 0) {
        __threadfence();
        atomicExch(A.g_queue + 1 + release_col, 0);
        release_col = -1;
      }
    };
    release();
    int i0 = 1;
    const int svc = *sSvc;
    if (svc) {
      // -- claim columns for idle slots
      if (leader && col < 0 && !exhausted) {
        // Work items: item = quantum * n_columns + column (quantum-major), each "advance this column by up to
        // `quantum` step attempts from its stored state".  With n_quanta == 1 an item is a whole column (the
        // classic queue).  Items of one column may be claimed by different CTAs at about the same time, so a
        // column is guarded by a lock word (g_queue[1 + column]); which item runs first does not matter, they
        // are interchangeable.  A slot that finds its column busy keeps the item and tries again at the next
        // service (the holder is resident and running, so it will release).
        int item = pending_item >= 0 ? pending_item : atomicAdd(A.g_queue, 1);
        pending_item = -1;
        int cc = -1;
        const int n_cut = A.n_columns - A.n_whole;       // columns that are cut into quanta (the last ones of the batch)
        if (item < A.n_whole + n_cut * A.n_quanta) {
          bool cut = false;
          if (item < A.n_whole) {
            cc = item;
          } else {
            const int j = item - A.n_whole, q = j / n_cut;
            cc = A.n_whole + (j - q * n_cut);
            cut = A.n_quanta > 1;
          }
          long long budget = A.opt.max_steps > 0 ? A.opt.max_steps : 0;
          long long used = 0;
          if (cut) {
            if (atomicCAS(A.g_queue + 1 + cc, 0, 1) != 0) {
              pending_item = item;
              cc = -2;
            } else {
              __threadfence();       // acquire: the previous holder's y / state / counter stores are visible from here on
              // attempts this column has used in this launch so far; the claim runs to the next multiple of the
              // quantum (an attempt budget ends at the first ACCEPTED step at or beyond it, so claims may overrun and
              // later items of the column may find nothing left to do)
              used = __ldcg(A.g_queue + 1 + A.n_columns + cc);
              const long long target = (used / A.quantum + 1) * A.quantum;
              if (A.opt.max_steps > 0) budget = target < A.opt.max_steps ? target : A.opt.max_steps;
              else budget = used / A.quantum + 1 >= A.n_quanta ? 0 : target;   // no budget: the column's last item runs to the end
            }
          }
          ctl.budget = budget;
          ctl.used = used;
        }
        sSlotCol[slot] = cc;
      }
      __syncthreads();
      if (tid == 0) *sSvc = 0;
      if (active && col < 0 && !exhausted) {
        col = sSlotCol[slot];
        if (col == -2) {
          col = -1;                  // column busy elsewhere: stay idle, the leader retries
        } else if (col < 0) {
          exhausted = true;
        } else {
          budget = ctl.budget;
          // (L2 loads: another SM may have written this column's state and y a moment ago)
          marlpde_column_state st;
          {
            const double* sp = reinterpret_cast<const double*>(A.g_state + col);
            double w[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldcg(sp + i);
            memcpy(&st, w, sizeof(st));
          }
          if (leader) {
            ColumnConsts tmp;
            make_consts(A.g_params[col], N, tmp);
            *const_cast<ColumnConsts*>(&kc) = tmp;
            ctl.n_acc = st.n_accepted;
            ctl.n_rej = st.n_rejected;
            ctl.nfev = st.nfev;
            sEv0[slot] = 0u;
            // both y_new buffers as well: a slot that sat idle for an odd number of trips (its column was locked by
            // another slot) would otherwise find the previous column's last bits in the buffer of its first attempt
            sEv[slot] = 0u;
            sEv[C + slot] = 0u;
          }
          need_prev = true;
          const int mlo_ = A.g_params[col].mask_lo, mhi_ = A.g_params[col].mask_hi;
          in_mask[0] = cell0 >= mlo_ && cell0 < mhi_;
          in_mask[1] = cell0 + 1 >= mlo_ && cell0 + 1 < mhi_;
          attempts = (int)ctl.used;
          t = st.t;
          h_abs = st.h_abs;
          next_eval = st.next_eval;
          const double* gy = A.g_y + (size_t)col * 5 * N + cell0;
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            c[f][0] = __ldcg(gy + (size_t)f * N);
            c[f][1] = has1 ? __ldcg(gy + (size_t)f * N + 1) : 0.0;
            Yst(f, c[f][0], c[f][1]);
          }
          tile_store(0);
          if (t >= A.opt.t_bound) {            // nothing to integrate
            retire(MARLPDE_STATUS_FINISHED);
          } else if (budget > 0 && (long long)attempts >= budget) {   // earlier claims used the column's whole budget
            retire(st.status);
          } else {
            begin_step();
            if (begin_attempt()) fresh = true;
            else retire(MARLPDE_STATUS_STEP_TOO_SMALL);
          }
        }
      }
      if (svc & 2) {
        // -- locate the events of parked steps (ivp.py handle_events -> brentq on the dense output),
        //    then commit them.  Every function value is a reduction over the column: one barrier each.
        //    Scratch: rows 0/1 of halo tile 0 (dead here), indexed inside the slot's own thread range.
        BrentState bs;
        int k = 0, buf = 0;
        double xeval = t;
        bool working = parked && ev_todo != 0u;
        if (working) {
          k = __ffs(ev_todo) - 1;
          ev_todo &= ev_todo - 1u;
          bs.init(t, t_new);
        }
        for (;;) {
          double* const scr = sE + buf * TP + slot * Hc;
          if (working) {
            double v = event_partial(k, (xeval - t) / h);
            for (int o = G >> 1; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(gmask, v, o));
            if ((pr & (G - 1)) == 0) scr[pr >> logG] = v;
          }
          if (!__syncthreads_or(working)) break;
          if (working) {
            double m = scr[0];
            for (int gi = 1; gi < nGroups; ++gi) m = fmin(m, scr[gi]);
            const double g = (k == 3 || k == 4) ? (-m) - 1.0 : (k == 6 ? -m : m);
            double root = 0.0;
            if (bs.feed(g, xeval, root)) {
              if (leader) {
                int32_t* cnt = A.g_ev_counts + (size_t)col * MARLPDE_NEVENTS + k;
                const int n = *cnt;
                if (n < A.opt.event_capacity)
                  A.g_ev_times[((size_t)col * MARLPDE_NEVENTS + k) * A.opt.event_capacity + n] = root;
                *cnt = n + 1;
              }
              if (ev_todo != 0u) {
                k = __ffs(ev_todo) - 1;
                ev_todo &= ev_todo - 1u;
                bs.init(t, t_new);
                xeval = t;
              } else {
                working = false;
              }
            }
          }
          buf ^= 1;
        }
        if (parked) {
          parked = false;
          ev_prev = ev_new_s;
          commit(factor_s);
          if (col >= 0) stage2_input();
        }
      }
      nlive = __syncthreads_count(col >= 0);   // also publishes consts and tile 0 of the new columns
      release();
      if (nlive == 0) continue;                // everything claimed retired at once: look again
      i0 = 0;
    }
    if (nlive == 0) break;

    // ================= stages: i = 1..6 evaluates K_{i+1} from halo tile (i & 1) ============
    // (i = 0, only after a slot service: K1 = f(y) of freshly loaded columns from tile 0.)
    // ONE code instance of the RHS serves every stage: the kernel stays inside the instruction
    // cache and K1 of a resumed column is bit-identical to the FSAL K7 it replaces.  Every lane
    // evaluates (votes inside need whole warps); only live slots consume the result.
    const bool live = col >= 0;
    double U[2], W[2];
#pragma unroll 1
    for (int i = i0; i <= 6; ++i) {
      // the part of the RHS that needs no neighbour runs while the barrier that publishes the neighbours'
      // stage inputs (arrived at below, at the end of the previous trip) is still pending
      OwnTerms own;
      PairFlags fl = rhs_pair_own<rhs_schedule(kSchedLean), VD>(kc, tb, c, in_mask, own);
      if (i > i0) {
        mbar_wait(sBar, bar_parity);
        bar_parity ^= 1u;
      }
      double mlo[5], phi[5];
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        double hm, hp;
        halo_load(i & 1, f, hm, hp);
        mlo[f] = first ? top_ghost(kc, f, c[f][0]) : hm;
        if (!last) {
          phi[f] = hp;
        } else if (has1) {
          phi[f] = bottom_ghost(f, c[f][1], c[f][0]);
        } else {                               // odd N: the ghost of cell N-1 sits in the pair's slot 1
          c[f][1] = bottom_ghost(f, c[f][0], mlo[f]);
          phi[f] = c[f][1];
        }
      }
      rhs_pair_finish<VD>(kc, c, mlo, phi, own, r);
      U[0] = own.U[0];
      U[1] = own.U[1];
      W[0] = own.W[0];
      W[1] = own.W[1];
      fl.bad[0] = fl.bad[0] && live;
      fl.bad[1] = fl.bad[1] && live && has1;
      if (fl.bad[0] || fl.bad[1]) rhs_pair_fixup(kc, tb, fl, c, mlo, phi, in_mask, r, U, W);
      if (!has1) {
#pragma unroll
        for (int f = 0; f < 5; ++f) r[f][1] = 0.0;
      }
      if (live) switch (i) {
        case 0:
          if (fresh) {
#pragma unroll
            for (int f = 0; f < 5; ++f) {
              K1st(f, r[f][0], r[f][1]);
            }
            if (leader) ctl.nfev += 1;
            if (ev_on) {     // signs of the monitors at the start point (ivp.py: g = event(t0, y0))
              const unsigned bv = __reduce_or_sync(peers, event_bits(c, U, W, has1));
              if (peer_lead && bv) atomicOr(&sEv0[slot], bv);
            }
            stage2_input();
            fresh = false;
          }
          break;
        case 1:   // r = K2
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(0, f, r[f][0], r[f][1]);
            const double2 yv = Yld(f), kv = K1ld(f);
            c[f][0] = fma(h, fma(dp::a31, kv.x, dp::a32 * r[f][0]), yv.x);
            c[f][1] = fma(h, fma(dp::a31, kv.y, dp::a32 * r[f][1]), yv.y);
          }
          tile_store(0);
          break;
        case 2:   // r = K3
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(1, f, r[f][0], r[f][1]);
            const double2 K2 = Kld(0, f), yv = Yld(f), kv = K1ld(f);
            c[f][0] = fma(h, fma(dp::a41, kv.x, fma(dp::a42, K2.x, dp::a43 * r[f][0])), yv.x);
            c[f][1] = fma(h, fma(dp::a41, kv.y, fma(dp::a42, K2.y, dp::a43 * r[f][1])), yv.y);
          }
          tile_store(1);
          break;
        case 3:   // r = K4
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(2, f, r[f][0], r[f][1]);
            const double2 K2 = Kld(0, f), K3 = Kld(1, f), yv = Yld(f), kv = K1ld(f);
            c[f][0] = fma(h, fma(dp::a51, kv.x, fma(dp::a52, K2.x, fma(dp::a53, K3.x, dp::a54 * r[f][0]))),
                          yv.x);
            c[f][1] = fma(h, fma(dp::a51, kv.y, fma(dp::a52, K2.y, fma(dp::a53, K3.y, dp::a54 * r[f][1]))),
                          yv.y);
          }
          tile_store(0);
          break;
        case 4:   // r = K5
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(3, f, r[f][0], r[f][1]);
            const double2 K2 = Kld(0, f), K3 = Kld(1, f), K4 = Kld(2, f), yv = Yld(f), kv = K1ld(f);
            c[f][0] = fma(h, fma(dp::a61, kv.x,
                                 fma(dp::a62, K2.x, fma(dp::a63, K3.x, fma(dp::a64, K4.x, dp::a65 * r[f][0])))),
                          yv.x);
            c[f][1] = fma(h, fma(dp::a61, kv.y,
                                 fma(dp::a62, K2.y, fma(dp::a63, K3.y, fma(dp::a64, K4.y, dp::a65 * r[f][1])))),
                          yv.y);
          }
          tile_store(1);
          break;
        case 5:   // r = K6 (stored over the dead K2): c becomes y_new
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            const double2 K3 = Kld(1, f), K4 = Kld(2, f), K5 = Kld(3, f), yv = Yld(f), kv = K1ld(f);
            Kst(0, f, r[f][0], r[f][1]);
            c[f][0] = fma(h, fma(dp::b1, kv.x, fma(dp::b3, K3.x, fma(dp::b4, K4.x, fma(dp::b5, K5.x, dp::b6 * r[f][0])))),
                          yv.x);
            c[f][1] = fma(h, fma(dp::b1, kv.y, fma(dp::b3, K3.y, fma(dp::b4, K4.y, fma(dp::b5, K5.y, dp::b6 * r[f][1])))),
                          yv.y);
          }
          tile_store(0);
          break;
        default:  // i == 6: r = K7 = f(y_new); monitor signs at y_new ride on the norm barrier
          if (ev_on) {
            const unsigned bv = __reduce_or_sync(peers, event_bits(c, U, W, has1));
            if (peer_lead && bv) atomicOr(&sEv[(it & 1u) * C + slot], bv);
          }
          break;
      }
      if (i < 6) mbar_arrive(sBar);        // my stage input for trip i+1 is published (every thread arrives, live or not)
    }
    // ---- K7 = f(y_new) is in r, y_new in c; error estimate and its norm
    double part = 0.0;
    if (live) {
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        const double2 K3 = Kld(1, f), K4 = Kld(2, f), K5 = Kld(3, f), K6 = Kld(0, f), yv = Yld(f), kv = K1ld(f);
        const double e0 = fma(dp::e1, kv.x, fma(dp::e3, K3.x, fma(dp::e4, K4.x, fma(dp::e5, K5.x, fma(dp::e6, K6.x, dp::e7 * r[f][0])))));
        const double e1_ = fma(dp::e1, kv.y, fma(dp::e3, K3.y, fma(dp::e4, K4.y, fma(dp::e5, K5.y, fma(dp::e6, K6.y, dp::e7 * r[f][1])))));
        const double s0 = fma(fmax(fabs(yv.x), fabs(c[f][0])), A.opt.rtol, A.opt.atol);
        const double s1 = fma(fmax(fabs(yv.y), fabs(c[f][1])), A.opt.rtol, A.opt.atol);
        const double q0 = (h * e0) * fm::rcp3(s0);
        const double q1 = (h * e1_) * fm::rcp3(s1);
        p0 = fma(q0, q0, p0);
        p1 = fma(q1, q1, p1);
      }
      part = has1 ? p0 + p1 : p0;
    }
    {
      double a = part;
      for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (live && (pr & (G - 1)) == 0) sGrp[tid >> logG] = a;
    }
    __syncthreads();
    if (live) {
      // group sums in order (4 interleaved partial sums, fixed association)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int gi = 0;
      for (; gi + 4 <= nGroups; gi += 4) {
        s0 += grpRow[gi];
        s1 += grpRow[gi + 1];
        s2 += grpRow[gi + 2];
        s3 += grpRow[gi + 3];
      }
      for (; gi < nGroups; ++gi) s0 += grpRow[gi];
      const double sum = (s0 + s1) + (s2 + s3);
      const double err_norm = sqrt(sum * inv_n);
      if (leader) ctl.nfev += 6;
      attempts += 1;
      unsigned ev_bits_new = 0u;
      if (ev_on) {
        if (need_prev) {
          ev_prev = sEv0[slot];
          need_prev = false;
        }
        ev_bits_new = sEv[(it & 1u) * C + slot];
        if (leader) sEv[((it + 1u) & 1u) * C + slot] = 0u;   // next attempt's buffer (last read 7 barriers ago)
      }
      if (err_norm < 1.0) {
        double factor = dp::MAX_FACTOR;
        if (err_norm != 0.0) factor = fmin(dp::MAX_FACTOR, dp::SAFETY * fm::exp(tb, -0.2 * fm::log(tb, err_norm)));
        if (rejected) factor = fmin(1.0, factor);
        // Same predicate bits as at the step's start and no monitor exactly on its threshold: the sign classes
        // are equal and none is "zero", so find_active_events reports nothing — skip the 7-monitor classification.
        const unsigned ev_new = ev_bits_new;
        unsigned act = 0u;
        if (ev_on && (ev_new != ev_prev || (ev_new & kEqBitsMask) != 0u))
          act = active_events(event_classes(ev_prev), event_classes(ev_new));
        if (act) {           // park the step: its events are located in the slot-service phase, then it commits
          parked = true;
          ev_todo = act;
          ev_new_s = ev_new;
          factor_s = factor;
          if (leader) atomicOr(sSvc, 2);
        } else {
          ev_prev = ev_new;
          commit(factor);
        }
      } else {
        // NaN error norms land here too: fmax drops the NaN, like Python's max(0.2, nan)
        h_abs *= fmax(dp::MIN_FACTOR, dp::SAFETY * fm::exp(tb, -0.2 * fm::log(tb, err_norm)));
        rejected = true;
        if (leader) ctl.n_rej += 1;
        if (!begin_attempt()) retire(MARLPDE_STATUS_STEP_TOO_SMALL);
      }
    }
  }
}

// Warp order of the kernel (see sPerm there).  MARLPDE_RK45_WARP_PERM = "identity" | "l0,l1,..." (logical warp of
// physical warp 0, 1, ...) overrides the automatic choice; anything that is not a permutation is ignored.
static unsigned long long rk45_warp_perm(int n_warps) {
  const unsigned long long automatic = ~0ull;
  unsigned long long id = 0;
  for (int w = 0; w < 16; ++w) id |= (unsigned long long)w << (4 * w);
  const char* s = std::getenv("MARLPDE_RK45_WARP_PERM");
  if (!s || !*s) return automatic;
  if (s[0] == 'i') return id;
  if (n_warps > 16) return automatic;
  unsigned long long p = 0;
  unsigned seen = 0;
  int w = 0;
  while (*s && w < n_warps) {
    char* end;
    const long v = std::strtol(s, &end, 10);
    if (end == s || v < 0 || v >= n_warps || (seen >> v & 1u)) return automatic;
    seen |= 1u << v;
    p |= (unsigned long long)v << (4 * w++);
    s = *end == ',' ? end + 1 : end;
  }
  if (w != n_warps) return automatic;
  for (; w < 16; ++w) p |= (unsigned long long)w << (4 * w);
  return p;
}

// Work items per column.  A launch whose columns all get the same step budget (a sweep advanced in steps, bench.py)
// runs ceil(columns / slots) rounds and the last round is only partly filled: 4096 columns on 444 slots are 9.22
// rounds of work in 10 rounds of time.  Cutting the budget into quanta that are claimed quantum-major makes the
// partly filled round 1 / n_quanta as long.  Needs the per-column lock and counter words (MARLPDE_FLAG_QUEUE_LOCKS)
// opt.quantum > 0 fixes the quantum, < 0 switches it off.  The end point of a column does not depend on the quanta: it is
// the first accepted step at which the launch's attempts reach max_steps, as without them.  Without a budget (a sweep
// integrated to t_bound in one launch) the columns take 0.5-1.2 M attempts each and the slots of the last round run dry
// one by one over the duration of a whole column (~5 % of a 4096-column sweep to T*).  Cutting EVERY column into quanta is
// wrong there — the columns would advance in lock-step and the longest ones (up to 2.4x the shortest) would finish alone at
// the speed of a lone column (measured, r02i: 151 s instead of 138 s).  Instead the first columns of the batch are claimed
// whole and only the LAST 2 x slots columns are cut into quanta (32 768 attempts; a column's last item runs without a
// limit): with the batch ordered longest first (callers sort by sweep.predicted_cost) the long columns start first and the
// short ones fill the end of the launch evenly.  MARLPDE_FLAG_QUEUE_TAIL asks for the same with a step budget.
static void choose_quanta(Rk45Args& a, int slots) {
  a.n_quanta = 1;
  a.quantum = 0;
  a.n_whole = 0;
  if (!(a.opt.flags & MARLPDE_FLAG_QUEUE_LOCKS) || a.opt.quantum < 0) return;
  const long long budget = a.opt.max_steps;
  if (budget <= 0 || (a.opt.flags & MARLPDE_FLAG_QUEUE_TAIL)) {
    if (a.n_columns <= slots) return;
    a.quantum = a.opt.quantum > 0 ? a.opt.quantum : 32768;
    a.n_quanta = 32;
    if (budget > 0) {
      if ((budget + a.quantum - 1) / a.quantum > 64) a.quantum = (budget + 63) / 64;
      a.n_quanta = (int)((budget + a.quantum - 1) / a.quantum);
    }
    const int n_cut = a.n_columns < 2 * slots ? a.n_columns : 2 * slots;
    a.n_whole = a.n_columns - n_cut;
  } else if (a.opt.quantum > 0) {
    a.quantum = a.opt.quantum;
    a.n_quanta = (int)((budget + a.quantum - 1) / a.quantum);
  } else {
    if (a.n_columns <= slots) return;          // every column has its own slot anyway
    double best = 1e300;
    for (int nq = 1; nq <= 32; ++nq) {
      const long long q = (budget + nq - 1) / nq;
      if (nq > 1 && q < 64) break;
      const double items = (double)a.n_columns * nq, rounds = items / slots;
      const double cost = std::ceil(rounds) / rounds + 0.5 / (double)q;   // idle tail + ~half an attempt per claim
      if (cost < best - 1e-9) {
        best = cost;
        a.n_quanta = nq;
        a.quantum = q;
      }
    }
  }
  if ((long long)a.n_columns * a.n_quanta > 0x7fffffffLL) {
    a.n_quanta = 1;
    a.quantum = 0;
  }
  if (a.n_quanta <= 1) {
    a.n_quanta = 1;
    a.quantum = 0;
    a.n_whole = 0;
  }
}

#ifndef MARLPDE_HOST_EMU
template <int TP, bool VD>
static cudaError_t launch_t(const Rk45Args& a, int sm_count, int smem_budget, cudaStream_t stream) {
  Rk45Args args = a;
  args.C = columns_per_cta_t<TP>(a.N, smem_budget);
  if (args.C <= 0) return cudaErrorInvalidValue;
  const int Hc = (a.N + 1) / 2;
  args.logG = group_log2(Hc);
  const size_t smem = Smem<TP>::total(args.C);
  cudaError_t e = cudaFuncSetAttribute(rk45_persistent_kernel<TP, VD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = (a.n_columns + args.C - 1) / args.C;
  if (grid > sm_count) grid = sm_count;
  if (grid < 1) grid = 1;
  const int threads = ((args.C * Hc + 31) / 32) * 32;
  choose_quanta(args, grid * args.C);
  args.warp_perm = rk45_warp_perm(threads / 32);
  rk45_persistent_kernel<TP, VD><<<grid, threads, smem, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launch_rk45(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                        int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                        double* d_snap, int32_t* d_ev_counts, double* d_ev_times, int32_t* d_queue, int sm_count,
                        int smem_budget, cudaStream_t stream) {
  Rk45Args a;
  a.g_y = d_y;
  a.g_params = d_params;
  a.g_state = d_state;
  a.g_t_eval = d_t_eval;
  a.g_snap = d_snap;
  a.g_queue = d_queue;
  a.g_ev_counts = d_ev_counts;
  a.g_ev_times = d_ev_times;
  a.n_columns = n_columns;
  a.N = n_cells;
  a.C = 0;
  a.logG = 0;
  a.n_quanta = 1;
  a.n_whole = 0;
  a.quantum = 0;
  a.warp_perm = 0;
  a.opt = opt;
  if (opt.flags & MARLPDE_FLAG_VAR_DPHI) return launch_t<kRk45Threads, true>(a, sm_count, smem_budget, stream);
  return launch_t<kRk45Threads, false>(a, sm_count, smem_budget, stream);
}

#endif

}  // namespace marlpde
