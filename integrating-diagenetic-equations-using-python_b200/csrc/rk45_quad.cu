// rk45_quad.cu — EXPERIMENTAL build of the on-chip RK45 kernel: four depth cells per thread, eight warps.
// Selected with MARLPDE_RK45_BUILD=450 (N % 4 == 0, N <= 1024); NOT the default; first GPU contact r01i: counters identical to the default build, +3.6 %
// (DESIGN.md 10.3).  Same algorithm, same controller, same event handling as rk45_persistent.cu — read that
// file first; this one only differs in how the work is laid out:
//
//   * rk45_persistent.cu: thread = 2 cells, 100 threads per column (N = 200), 3 columns = 10 warps, which four
//     schedulers hold as 3-3-2-2.  After r01g that kernel is bound by the instruction throughput of the two
//     schedulers with three warps.
//   * here: thread = 4 cells, 50 threads per column, 5 columns = 250 of 256 lanes = 8 warps = 2-2-2-2.  Per
//     scheduler and stage trip 2 warps x 2 pair evaluations serve 5 columns (0.8 units per column against 1.0).
//   * registers (255 allowed at 8 warps): K1, the stage input and the stage derivative of the four cells; the
//     state y lives in shared memory (only the stage algebra reads it);
//   * shared memory: K2..K5 (160 kB) + y (40 kB) leave no room for a stage-input tile, so a thread gets cell
//     4k-1 and cell 4k+4 from its neighbour lanes by warp shuffle; only the two edge lanes of a warp publish
//     through shared memory (which is what the split stage barrier protects);
//   * the RHS is rhs_pair as everywhere: cells (0, 1) with phi = cell 2, cells (2, 3) with mlo = cell 1.  The
//     own-cell part of the first pair overlaps the stage barrier.
//
// Reference call site and algorithm: marlpde/Evolve_scenario.py:104-109 -> scipy/integrate/_ivp/rk.py
// (RungeKutta._step_impl, rk_step, RkDenseOutput), ivp.py (t_eval sampling, find_active_events).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "brent.cuh"
#include "dopri.cuh"
#include "lheureux_device.cuh"
#include "mbar.cuh"
#include "rk45_quad.cuh"

namespace marlpde {
namespace quad {

constexpr int TP = 256;   // threads per CTA
constexpr int Q = 4;      // cells per thread
constexpr int kWarps = TP / 32;
#ifndef MARLPDE_QUAD_ROLLED
#define MARLPDE_QUAD_ROLLED 0  // 1: one RHS instance in a rolled loop over the two pairs (A/B candidate: half the hot code)
#endif
#ifndef MARLPDE_QUAD_ORDER
#define MARLPDE_QUAD_ORDER 0   // 1: own-cell parts of both pairs before the stage-barrier wait (A/B candidate)
#endif

struct SlotCtl {          // per-slot counters, touched by the slot's leader thread only
  long long n_acc, n_rej, nfev;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) / 16 * 16; }

// ---- shared memory carve-up -------------------------------------------------------------
// K[4 stages][5 fields][2 halves][TP] double2 | Y[5][2][TP] double2 | edge[2][kWarps+1][2][5] | scr[2][TP] |
// grp[TP] | log/exp tables | consts[C] | ctl[C] | slot_col[C] | event words | svc flag | mbarrier
struct Smem {
  static constexpr size_t off_K = 0;
  static constexpr size_t off_Y = off_K + sizeof(double2) * 4 * 5 * 2 * TP;
  static constexpr size_t off_edge = off_Y + sizeof(double2) * 5 * 2 * TP;
  static constexpr size_t edge_stride = (kWarps + 1) * 2 * 5;                       // doubles per parity
  static constexpr size_t off_scr = off_edge + sizeof(double) * 2 * edge_stride;
  static constexpr size_t off_grp = off_scr + sizeof(double) * 2 * TP;
  static constexpr size_t off_tab = off_grp + sizeof(double) * TP;
  static constexpr size_t off_var = (off_tab + fm::kTableBytes + 15) / 16 * 16;
  static constexpr size_t slot_bytes = (sizeof(ColumnConsts) + 15) / 16 * 16 + (sizeof(SlotCtl) + 15) / 16 * 16 + 32;
  __host__ __device__ static size_t total(int C) { return off_var + slot_bytes * (size_t)C + 32 + 16; }
};

static int group_log2(int threads_per_column) {
  int logG = 0;
  while (logG < 5 && (threads_per_column % (2 << logG)) == 0) ++logG;
  return logG;
}

int columns_per_cta(int n_cells, int smem_budget) {
  if (n_cells < 32 || (n_cells % Q) != 0) return 0;
  const int Hc = n_cells / Q;
  if (Hc > TP) return 0;
  int C = TP / Hc;
  while (C > 0 && Smem::total(C) > (size_t)smem_budget) --C;
  return C;
}

// ---- event monitors: same predicate bits as rk45_persistent.cu (3 per monitor: beyond / on / NaN) -----------
constexpr unsigned kMaxTypeMask = (1u << 3) | (1u << 4) | (1u << 6);
constexpr unsigned kEqBitsMask = 0x92492u;

__device__ __forceinline__ unsigned event_bits(const double (&v)[5][Q], const double (&U)[Q], const double (&W)[Q]) {
  unsigned b = 0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    bool lt = false, eq = false, nn = false;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      lt |= v[f][q] < 0.0;
      eq |= v[f][q] == 0.0;
      nn |= v[f][q] != v[f][q];
    }
    const double CA = v[0][q], CC = v[1][q], s = CA + CC, Phi = v[4][q];
    b |= (lt ? 1u : 0u) | (eq ? 2u : 0u) | (nn ? 4u : 0u);
    b |= (CA < 0.0 ? 1u : 0u) << 3 | (CA == 0.0 ? 1u : 0u) << 4 | (CA != CA ? 1u : 0u) << 5;
    b |= (CC < 0.0 ? 1u : 0u) << 6 | (CC == 0.0 ? 1u : 0u) << 7 | (CC != CC ? 1u : 0u) << 8;
    b |= (s > 1.0 ? 1u : 0u) << 9 | (s == 1.0 ? 1u : 0u) << 10 | (s != s ? 1u : 0u) << 11;
    b |= (Phi > 1.0 ? 1u : 0u) << 12 | (Phi == 1.0 ? 1u : 0u) << 13 | (Phi != Phi ? 1u : 0u) << 14;
    b |= (U[q] < 0.0 ? 1u : 0u) << 15 | (U[q] == 0.0 ? 1u : 0u) << 16 | (U[q] != U[q] ? 1u : 0u) << 17;
    b |= (W[q] > 0.0 ? 1u : 0u) << 18 | (W[q] == 0.0 ? 1u : 0u) << 19 | (W[q] != W[q] ? 1u : 0u) << 20;
  }
  return b;
}

__device__ __forceinline__ unsigned event_classes(unsigned bits) {   // 2 bits per monitor: 0 neg, 1 zero, 2 pos, 3 NaN
  unsigned cls = 0;
#pragma unroll
  for (int k = 0; k < MARLPDE_NEVENTS; ++k) {
    const unsigned b3 = (bits >> (3 * k)) & 7u;
    const bool maxtype = (kMaxTypeMask >> k) & 1u;
    unsigned c;
    if (b3 & 4u) c = 3u;
    else if (b3 & 1u) c = maxtype ? 2u : 0u;
    else if (b3 & 2u) c = 1u;
    else c = maxtype ? 0u : 2u;
    cls |= c << (2 * k);
  }
  return cls;
}

__device__ __forceinline__ unsigned active_events(unsigned cls_old, unsigned cls_new) {   // ivp.py find_active_events
  unsigned act = 0;
#pragma unroll
  for (int k = 0; k < MARLPDE_NEVENTS; ++k) {
    const unsigned a = (cls_old >> (2 * k)) & 3u, b = (cls_new >> (2 * k)) & 3u;
    const bool a_le = a <= 1u, a_ge = a == 1u || a == 2u, b_le = b <= 1u, b_ge = b == 1u || b == 2u;
    if ((a_le && b_ge) || (a_ge && b_le)) act |= 1u << k;
  }
  return act;
}

struct Args {
  double* g_y;
  const marlpde_column_params* g_params;
  marlpde_column_state* g_state;
  const double* g_t_eval;
  double* g_snap;
  int32_t* g_queue;
  int32_t* g_ev_counts;      // [n_columns][7]
  double* g_ev_times;        // [n_columns][7][event_capacity]
  int n_columns, N, C, logG;
  marlpde_rk45_options opt;
};

__global__ void __launch_bounds__(TP, 1) rk45_quad_kernel(const Args A) {
  MARLPDE_DYN_SMEM(smem_raw);
  using L = Smem;
  const int tid = threadIdx.x;
  const int N = A.N, C = A.C;
  const int Hc = N / Q;                            // threads per column
  double2* const sK = reinterpret_cast<double2*>(smem_raw + L::off_K) + tid;         // [4][5][2][TP]
  double2* const sY = reinterpret_cast<double2*>(smem_raw + L::off_Y) + tid;         // [5][2][TP]
  double* const sEdge = reinterpret_cast<double*>(smem_raw + L::off_edge);          // [2][kWarps+1][2][5]
  double* const sScr = reinterpret_cast<double*>(smem_raw + L::off_scr);            // [2][TP] event location scratch
  double* const sGrp = reinterpret_cast<double*>(smem_raw + L::off_grp);            // [TP >> logG]
  unsigned char* const var = smem_raw + L::off_var;
  const size_t consts_sz = align16(sizeof(ColumnConsts)), ctl_sz = align16(sizeof(SlotCtl));
  int* const sSlotCol = reinterpret_cast<int*>(var + (consts_sz + ctl_sz) * C);     // [C]
  unsigned* const sEv = reinterpret_cast<unsigned*>(sSlotCol + C);                  // [2][C] bits of y_new
  unsigned* const sEv0 = sEv + 2 * C;                                               // [C] bits of a fresh y
  int* const sSvc = reinterpret_cast<int*>(sEv0 + C);
  uint64_t* const sBar = reinterpret_cast<uint64_t*>(var + ((L::slot_bytes * (size_t)C + 15) / 16) * 16 + 16);

  const fm::Tables tb = fm::stage_tables(smem_raw + L::off_tab, tid, blockDim.x);
  const bool active = tid < C * Hc;
  const int slot = active ? tid / Hc : 0;
  const int pr = active ? tid - slot * Hc : 0;     // index of my group of four cells inside the column
  const int cell0 = Q * pr;
  const bool first = pr == 0, last = pr == Hc - 1;
  const ColumnConsts& kc = *reinterpret_cast<const ColumnConsts*>(var + consts_sz * slot);
  SlotCtl& ctl = *reinterpret_cast<SlotCtl*>(var + consts_sz * C + ctl_sz * slot);
  const bool leader = active && first;
  const int lane_id = tid & 31, warp_id = tid >> 5;
  // error-norm reduction tree, identical for every slot (see rk45_persistent.cu)
  const int logG = A.logG;
  const int G = 1 << logG;
  const int nGroups = Hc >> logG;
  const double* const grpRow = sGrp + ((slot * Hc) >> logG);
  const double inv_n = 1.0 / (double)(5 * N);
  const bool ev_on = (A.opt.flags & MARLPDE_FLAG_EVENTS) != 0;
  const unsigned peers = __match_any_sync(0xffffffffu, active ? slot : -1);
  const bool peer_lead = (tid & 31) == (__ffs(peers) - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
  unsigned ev_prev = 0, ev_new_s = 0, ev_todo = 0;   // predicate bits at y / at the parked y_new; monitors to locate
  bool need_prev = false, parked = false;
  double factor_s = 1.0;
  unsigned it = 0;

  // per-thread column state
  int col = -1;
  bool exhausted = !active;
  bool rejected = false;
  double t = 0.0, h_abs = 0.0, h = 0.0, t_new = 0.0;
  int next_eval = 0;
  int attempts = 0;
  bool in_mask[Q] = {false, false, false, false};
  double k1[5][Q], c[5][Q], r[5][Q];
#pragma unroll
  for (int f = 0; f < 5; ++f)
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      k1[f][q] = 0.0;
      c[f][q] = 0.5;
      r[f][q] = 0.0;
    }

  // only the two edge lanes of a warp publish through shared memory: side 0 = lane 0's first cell, side 1 = lane
  // 31's last cell, for stage parity b
  auto tile_store = [&](int b) {
    if (lane_id == 0 || lane_id == 31) {
      double* e = sEdge + b * L::edge_stride + (warp_id * 2 + (lane_id == 31 ? 1 : 0)) * 5;
#pragma unroll
      for (int f = 0; f < 5; ++f) e[f] = lane_id == 31 ? c[f][Q - 1] : c[f][0];
    }
  };
  // raw halo values for stage parity b: (last cell of thread tid-1, first cell of thread tid+1); all 32 lanes call it
  auto halo_load = [&](int b, int f, double& hm, double& hp) {
    hm = __shfl_up_sync(0xffffffffu, c[f][Q - 1], 1);
    hp = __shfl_down_sync(0xffffffffu, c[f][0], 1);
    const double* e = sEdge + b * L::edge_stride;
    if (lane_id == 0 && warp_id > 0) hm = e[((warp_id - 1) * 2 + 1) * 5 + f];
    if (lane_id == 31) hp = e[((warp_id + 1) * 2) * 5 + f];
  };
  // K slots in shared memory: four doubles per thread, field and stage = two double2
  auto Kst = [&](int s, int f, const double (&v)[5][Q]) {
    sK[((s * 5 + f) * 2 + 0) * TP] = make_double2(v[f][0], v[f][1]);
    sK[((s * 5 + f) * 2 + 1) * TP] = make_double2(v[f][2], v[f][3]);
  };
  auto Kld = [&](int s, int f, double (&v)[Q]) {
    const double2 a = sK[((s * 5 + f) * 2 + 0) * TP], b = sK[((s * 5 + f) * 2 + 1) * TP];
    v[0] = a.x;
    v[1] = a.y;
    v[2] = b.x;
    v[3] = b.y;
  };
  auto Yld = [&](int f, double (&v)[Q]) {
    const double2 a = sY[(f * 2 + 0) * TP], b = sY[(f * 2 + 1) * TP];
    v[0] = a.x;
    v[1] = a.y;
    v[2] = b.x;
    v[3] = b.y;
  };
  auto Yst = [&](int f, const double (&v)[5][Q]) {
    sY[(f * 2 + 0) * TP] = make_double2(v[f][0], v[f][1]);
    sY[(f * 2 + 1) * TP] = make_double2(v[f][2], v[f][3]);
  };

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
      double yv[Q];
      Yld(f, yv);
#pragma unroll
      for (int q = 0; q < Q; ++q) gy[(size_t)f * N + q] = yv[q];
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
    }
    col = -1;
#pragma unroll
    for (int f = 0; f < 5; ++f)
#pragma unroll
      for (int q = 0; q < Q; ++q) c[f][q] = 0.5;   // idle lanes evaluate the RHS on benign values
  };
  auto stage2_input = [&]() {
    const double ha = h * dp::a21;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      double yv[Q];
      Yld(f, yv);
#pragma unroll
      for (int q = 0; q < Q; ++q) c[f][q] = fma(ha, k1[f][q], yv[q]);
    }
    tile_store(1);
  };
  // quartic dense output of the step just computed (scipy RkDenseOutput): y(t + x h) for my four cells; needs K1 (k1),
  // K3..K5 and K6 (shared memory) and K7 (= r), i.e. must run before the commit
  auto interp_all = [&](double x, double (&out)[5][Q]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      double K3[Q], K4[Q], K5[Q], K6[Q], yv[Q];
      Kld(1, f, K3);
      Kld(2, f, K4);
      Kld(3, f, K5);
      Kld(0, f, K6);
      Yld(f, yv);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        double qq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double sacc = dp::P[0][j] * k1[f][q];
          sacc = fma(dp::P[2][j], K3[q], sacc);
          sacc = fma(dp::P[3][j], K4[q], sacc);
          sacc = fma(dp::P[4][j], K5[q], sacc);
          sacc = fma(dp::P[5][j], K6[q], sacc);
          qq[j] = fma(dp::P[6][j], r[f][q], sacc);
        }
        const double poly = x * (qq[0] + x * (qq[1] + x * (qq[2] + x * qq[3])));
        out[f][q] = fma(h, poly, yv[q]);
      }
    }
  };
  // an accepted step becomes the column's state: t_eval samples, y <- y_new, K1 <- K7 (FSAL)
  auto commit = [&](double factor) {
    while (next_eval < A.opt.n_eval) {
      const double te = A.g_t_eval[next_eval];
      if (!(te <= t_new)) break;
      double ys[5][Q];
      interp_all((te - t) / h, ys);
      double* gs = A.g_snap + ((size_t)col * A.opt.n_eval + next_eval) * 5 * N + cell0;
#pragma unroll
      for (int f = 0; f < 5; ++f)
#pragma unroll
        for (int q = 0; q < Q; ++q) gs[(size_t)f * N + q] = ys[f][q];
      ++next_eval;
    }
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      Yst(f, c);
#pragma unroll
      for (int q = 0; q < Q; ++q) k1[f][q] = r[f][q];
    }
    t = t_new;
    h_abs *= factor;
    if (leader) ctl.n_acc += 1;
    if (t >= A.opt.t_bound) {
      retire(MARLPDE_STATUS_FINISHED);
    } else if (A.opt.max_steps > 0 && (long long)attempts >= A.opt.max_steps) {
      retire(MARLPDE_STATUS_STEP_BUDGET);
    } else {
      begin_step();
      if (!begin_attempt()) retire(MARLPDE_STATUS_STEP_TOO_SMALL);
    }
  };
  // value of monitor k on my four cells at y(t + x h), in "min form" (a max is the min of the negation)
  auto event_partial = [&](int k, double x) -> double {
    double ys[5][Q];
    interp_all(x, ys);
    double m = (double)INFINITY;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const double Phi = ys[4][q];
      double UW = 0.0;
      if (k >= 5) {   // same arithmetic as rhs_pair, so detection and location agree
        const double F = 1.0 - fm::exp(tb, fma(-10.0, fm::rcp3(Phi), 10.0));
        const double Phi2 = Phi * Phi;
        UW = k == 5 ? fma(kc.rhorat * (Phi2 * Phi), F * fm::rcp3(1.0 - Phi), kc.presum)
                    : -fma(-kc.rhorat * Phi2, F, kc.presum);
      }
      double v;
      switch (k) {
        case 0: v = fmin(fmin(fmin(ys[0][q], ys[1][q]), fmin(ys[2][q], ys[3][q])), ys[4][q]); break;
        case 1: v = ys[0][q]; break;
        case 2: v = ys[1][q]; break;
        case 3: v = -(ys[0][q] + ys[1][q]); break;
        case 4: v = -Phi; break;
        default: v = UW; break;
      }
      m = fmin(m, v);
    }
    return m;
  };

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
    int i0 = 1;
    const int svc = *sSvc;
    if (svc) {
      if (leader && col < 0 && !exhausted) {
#if MARLPDE_TAIL_SPREAD   // slot s only claims while more than s columns per CTA are left: the tail of a sweep (and a
                          // batch smaller than the machine) is spread one column per SM instead of stacked on a few
        const int left = A.n_columns - *reinterpret_cast<volatile int32_t*>(A.g_queue);
        const int cc = (slot > 0 && left <= slot * (int)gridDim.x) ? A.n_columns : atomicAdd(A.g_queue, 1);
#else
        const int cc = atomicAdd(A.g_queue, 1);
#endif
        sSlotCol[slot] = cc < A.n_columns ? cc : -1;
      }
      __syncthreads();
      if (tid == 0) *sSvc = 0;
      if (active && col < 0 && !exhausted) {
        col = sSlotCol[slot];
        if (col < 0) {
          exhausted = true;
        } else {
          const marlpde_column_state st = A.g_state[col];
          if (leader) {
            ColumnConsts tmp;
            make_consts(A.g_params[col], N, tmp);
            *const_cast<ColumnConsts*>(&kc) = tmp;
            ctl.n_acc = st.n_accepted;
            ctl.n_rej = st.n_rejected;
            ctl.nfev = st.nfev;
            sEv0[slot] = 0u;
          }
          need_prev = true;
          const int mlo_ = A.g_params[col].mask_lo, mhi_ = A.g_params[col].mask_hi;
#pragma unroll
          for (int q = 0; q < Q; ++q) in_mask[q] = cell0 + q >= mlo_ && cell0 + q < mhi_;
          attempts = 0;
          t = st.t;
          h_abs = st.h_abs;
          next_eval = st.next_eval;
          const double* gy = A.g_y + (size_t)col * 5 * N + cell0;
#pragma unroll
          for (int f = 0; f < 5; ++f) {
#pragma unroll
            for (int q = 0; q < Q; ++q) c[f][q] = gy[(size_t)f * N + q];
            Yst(f, c);
          }
          if (t >= A.opt.t_bound) {            // nothing to integrate
            retire(MARLPDE_STATUS_FINISHED);
          } else {
            begin_step();
            if (begin_attempt()) fresh = true;
            else retire(MARLPDE_STATUS_STEP_TOO_SMALL);
          }
        }
      }
      tile_store(0);                           // (every thread: the edge lanes of a warp may belong to any slot)
      if (svc & 2) {
        // -- locate the events of parked steps (ivp.py handle_events -> brentq on the dense output), then commit them
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
          double* const scr = sScr + buf * TP + slot * Hc;
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
      nlive = __syncthreads_count(col >= 0);   // also publishes consts and the edge values of the new columns
      if (nlive == 0) continue;                // everything claimed retired at once: look again
      i0 = 0;
    }
    if (nlive == 0) break;

    // ================= stages: i = 1..6 evaluates K_{i+1} with the edge values of parity (i & 1) ============
    // (i = 0, only after a slot service: K1 = f(y) of freshly loaded columns, parity 0.)
    const bool live = col >= 0;
    double U[Q], W[Q];
#pragma unroll 1
    for (int i = i0; i <= 6; ++i) {
#if MARLPDE_QUAD_ROLLED
      // ONE instance of the RHS in a rolled loop over the thread's two pairs (half the hot code of the two inlined instances;
      // the pair's inputs and outputs change places in registers between the trips)
      double cP[5][2], mloP[5], phiP[5], hpS[5], rP[5][2], UP[2], WP[2];
      bool maskP[2] = {in_mask[0], in_mask[1]};
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        cP[f][0] = c[f][0];
        cP[f][1] = c[f][1];
        mloP[f] = phiP[f] = hpS[f] = 0.0;
      }
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        OwnTerms own;
        PairFlags fl = rhs_pair_own<rhs_schedule(kSchedLean)>(kc, tb, cP, maskP, own);
        if (p == 0) {                              // the own-cell part of the first pair overlapped the pending barrier
          if (i > i0) {
            mbar_wait(sBar, bar_parity);
            bar_parity ^= 1u;
          }
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            double hm, hp;
            halo_load(i & 1, f, hm, hp);
            mloP[f] = first ? top_ghost(kc, f, c[f][0]) : hm;
            phiP[f] = c[f][2];
            hpS[f] = hp;
          }
        }
        rhs_pair_finish(kc, cP, mloP, phiP, own, rP);
        UP[0] = own.U[0];
        UP[1] = own.U[1];
        WP[0] = own.W[0];
        WP[1] = own.W[1];
        fl.bad[0] = fl.bad[0] && live;
        fl.bad[1] = fl.bad[1] && live;
        if (fl.bad[0] || fl.bad[1]) rhs_pair_fixup(kc, tb, fl, cP, mloP, phiP, maskP, rP, UP, WP);
        if (p == 0) {
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            r[f][0] = rP[f][0];
            r[f][1] = rP[f][1];
            cP[f][0] = c[f][2];
            cP[f][1] = c[f][3];
            mloP[f] = c[f][1];
            phiP[f] = last ? bottom_ghost(f, c[f][3], c[f][2]) : hpS[f];
          }
          U[0] = UP[0];
          U[1] = UP[1];
          W[0] = WP[0];
          W[1] = WP[1];
          maskP[0] = in_mask[2];
          maskP[1] = in_mask[3];
        } else {
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            r[f][2] = rP[f][0];
            r[f][3] = rP[f][1];
          }
          U[2] = UP[0];
          U[3] = UP[1];
          W[2] = WP[0];
          W[3] = WP[1];
        }
      }
#else
      // own-cell part of the first pair while the barrier that publishes the warp-edge values is pending
      double cA[5][2], cB[5][2];
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        cA[f][0] = c[f][0];
        cA[f][1] = c[f][1];
        cB[f][0] = c[f][2];
        cB[f][1] = c[f][3];
      }
      const bool maskA[2] = {in_mask[0], in_mask[1]}, maskB[2] = {in_mask[2], in_mask[3]};
      OwnTerms ownA;
      PairFlags flA = rhs_pair_own<rhs_schedule(kSchedLean)>(kc, tb, cA, maskA, ownA);
#if MARLPDE_QUAD_ORDER == 1
      // own-cell parts of BOTH pairs before the wait (more work behind the barrier, 48 more live registers)
      OwnTerms ownB;
      PairFlags flB = rhs_pair_own<rhs_schedule(kSchedLean)>(kc, tb, cB, maskB, ownB);
#endif
      if (i > i0) {
        mbar_wait(sBar, bar_parity);
        bar_parity ^= 1u;
      }
      double mloA[5], phiA[5], mloB[5], phiB[5];
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        double hm, hp;
        halo_load(i & 1, f, hm, hp);
        mloA[f] = first ? top_ghost(kc, f, c[f][0]) : hm;
        phiA[f] = c[f][2];
        mloB[f] = c[f][1];
        phiB[f] = last ? bottom_ghost(f, c[f][3], c[f][2]) : hp;
      }
      double rA[5][2], rB[5][2], UA[2], WA[2], UB[2], WB[2];
      rhs_pair_finish(kc, cA, mloA, phiA, ownA, rA);
      UA[0] = ownA.U[0];
      UA[1] = ownA.U[1];
      WA[0] = ownA.W[0];
      WA[1] = ownA.W[1];
      flA.bad[0] = flA.bad[0] && live;
      flA.bad[1] = flA.bad[1] && live;
      if (flA.bad[0] || flA.bad[1]) rhs_pair_fixup(kc, tb, flA, cA, mloA, phiA, maskA, rA, UA, WA);
#if MARLPDE_QUAD_ORDER == 1
      rhs_pair_finish(kc, cB, mloB, phiB, ownB, rB);
      UB[0] = ownB.U[0];
      UB[1] = ownB.U[1];
      WB[0] = ownB.W[0];
      WB[1] = ownB.W[1];
#else
      PairFlags flB = rhs_pair<rhs_schedule(kSchedLean)>(kc, tb, cB, mloB, phiB, maskB, rB, UB, WB);
#endif
      flB.bad[0] = flB.bad[0] && live;
      flB.bad[1] = flB.bad[1] && live;
      if (flB.bad[0] || flB.bad[1]) rhs_pair_fixup(kc, tb, flB, cB, mloB, phiB, maskB, rB, UB, WB);
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        r[f][0] = rA[f][0];
        r[f][1] = rA[f][1];
        r[f][2] = rB[f][0];
        r[f][3] = rB[f][1];
      }
      U[0] = UA[0];
      U[1] = UA[1];
      U[2] = UB[0];
      U[3] = UB[1];
      W[0] = WA[0];
      W[1] = WA[1];
      W[2] = WB[0];
      W[3] = WB[1];
#endif
      if (live) switch (i) {
        case 0:
          if (fresh) {
#pragma unroll
            for (int f = 0; f < 5; ++f)
#pragma unroll
              for (int q = 0; q < Q; ++q) k1[f][q] = r[f][q];
            if (leader) ctl.nfev += 1;
            if (ev_on) {     // signs of the monitors at the start point (ivp.py: g = event(t0, y0))
              const unsigned bv = __reduce_or_sync(peers, event_bits(c, U, W));
              if (peer_lead && bv) atomicOr(&sEv0[slot], bv);
            }
          }
          break;
        case 1:   // r = K2
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(0, f, r);
            double yv[Q];
            Yld(f, yv);
#pragma unroll
            for (int q = 0; q < Q; ++q) c[f][q] = fma(h, fma(dp::a31, k1[f][q], dp::a32 * r[f][q]), yv[q]);
          }
          break;
        case 2:   // r = K3
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(1, f, r);
            double K2[Q], yv[Q];
            Kld(0, f, K2);
            Yld(f, yv);
#pragma unroll
            for (int q = 0; q < Q; ++q)
              c[f][q] = fma(h, fma(dp::a41, k1[f][q], fma(dp::a42, K2[q], dp::a43 * r[f][q])), yv[q]);
          }
          break;
        case 3:   // r = K4
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(2, f, r);
            double K2[Q], K3[Q], yv[Q];
            Kld(0, f, K2);
            Kld(1, f, K3);
            Yld(f, yv);
#pragma unroll
            for (int q = 0; q < Q; ++q)
              c[f][q] = fma(h, fma(dp::a51, k1[f][q], fma(dp::a52, K2[q], fma(dp::a53, K3[q], dp::a54 * r[f][q]))), yv[q]);
          }
          break;
        case 4:   // r = K5
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            Kst(3, f, r);
            double K2[Q], K3[Q], K4[Q], yv[Q];
            Kld(0, f, K2);
            Kld(1, f, K3);
            Kld(2, f, K4);
            Yld(f, yv);
#pragma unroll
            for (int q = 0; q < Q; ++q)
              c[f][q] = fma(h, fma(dp::a61, k1[f][q],
                                   fma(dp::a62, K2[q], fma(dp::a63, K3[q], fma(dp::a64, K4[q], dp::a65 * r[f][q])))),
                            yv[q]);
          }
          break;
        case 5:   // r = K6 (stored over the dead K2): c becomes y_new
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            double K3[Q], K4[Q], K5[Q], yv[Q];
            Kld(1, f, K3);
            Kld(2, f, K4);
            Kld(3, f, K5);
            Yld(f, yv);
            Kst(0, f, r);
#pragma unroll
            for (int q = 0; q < Q; ++q)
              c[f][q] = fma(h, fma(dp::b1, k1[f][q],
                                   fma(dp::b3, K3[q], fma(dp::b4, K4[q], fma(dp::b5, K5[q], dp::b6 * r[f][q])))),
                            yv[q]);
          }
          break;
        default:  // i == 6: r = K7 = f(y_new); monitor signs at y_new ride on the norm barrier
          if (ev_on) {
            const unsigned bv = __reduce_or_sync(peers, event_bits(c, U, W));
            if (peer_lead && bv) atomicOr(&sEv[(it & 1u) * C + slot], bv);
          }
          break;
      }
      if (i == 0 && live && fresh) {
        stage2_input();                      // (writes c; its tile_store(1) is repeated below for every thread)
        fresh = false;
      }
      // publish the warp-edge values of the next stage input (parity (i + 1) & 1); every thread executes it: an
      // edge lane publishes for whichever slot it belongs to, idle slots publish their benign values
      if (i < 6) {
        tile_store((i + 1) & 1);
        mbar_arrive(sBar);
      }
    }
    // ---- K7 = f(y_new) is in r, y_new in c; error estimate and its norm
    double part = 0.0;
    if (live) {
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        double K3[Q], K4[Q], K5[Q], K6[Q], yv[Q];
        Kld(1, f, K3);
        Kld(2, f, K4);
        Kld(3, f, K5);
        Kld(0, f, K6);
        Yld(f, yv);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double e = fma(dp::e1, k1[f][q],
                               fma(dp::e3, K3[q], fma(dp::e4, K4[q], fma(dp::e5, K5[q], fma(dp::e6, K6[q], dp::e7 * r[f][q])))));
          const double s = fma(fmax(fabs(yv[q]), fabs(c[f][q])), A.opt.rtol, A.opt.atol);
          const double qv = (h * e) * fm::rcp3(s);
          part = fma(qv, qv, part);
        }
      }
    }
    {
      double a = part;
      for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (live && (pr & (G - 1)) == 0) sGrp[tid >> logG] = a;
    }
    __syncthreads();
    if (live) {
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
        if (leader) sEv[((it + 1u) & 1u) * C + slot] = 0u;
      }
      if (err_norm < 1.0) {
        double factor = dp::MAX_FACTOR;
        if (err_norm != 0.0) factor = fmin(dp::MAX_FACTOR, dp::SAFETY * fm::exp(tb, -0.2 * fm::log(tb, err_norm)));
        if (rejected) factor = fmin(1.0, factor);
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
        h_abs *= fmax(dp::MIN_FACTOR, dp::SAFETY * fm::exp(tb, -0.2 * fm::log(tb, err_norm)));
        rejected = true;
        if (leader) ctl.n_rej += 1;
        if (!begin_attempt()) retire(MARLPDE_STATUS_STEP_TOO_SMALL);
      }
    }
  }
}

}  // namespace quad

int rk45_quad_columns_per_cta(int n_cells, int smem_budget) { return quad::columns_per_cta(n_cells, smem_budget); }

#ifndef MARLPDE_HOST_EMU
cudaError_t launch_rk45_quad(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                             int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                             double* d_snap, int32_t* d_ev_counts, double* d_ev_times, int32_t* d_queue, int sm_count,
                             int smem_budget, cudaStream_t stream) {
  quad::Args a;
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
  a.C = quad::columns_per_cta(n_cells, smem_budget);
  if (a.C <= 0) return cudaErrorInvalidValue;
  a.logG = quad::group_log2(n_cells / quad::Q);
  a.opt = opt;
  const size_t smem = quad::Smem::total(a.C);
  cudaError_t e = cudaFuncSetAttribute(quad::rk45_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = MARLPDE_TAIL_SPREAD ? n_columns : (n_columns + a.C - 1) / a.C;
  if (grid > sm_count) grid = sm_count;
  if (grid < 1) grid = 1;
  quad::rk45_quad_kernel<<<grid, quad::TP, smem, stream>>>(a);
  return cudaGetLastError();
}
#endif

}  // namespace marlpde
