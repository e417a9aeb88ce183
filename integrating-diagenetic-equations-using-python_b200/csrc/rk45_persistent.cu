// rk45_persistent.cu — batched adaptive Dormand-Prince 5(4) for sediment columns, sm_100a.
//
// Replaces the loop `scipy.integrate.solve_ivp(eq.fun_numba, ..., method="RK45")`
// (reference call site marlpde/Evolve_scenario.py:104-109; algorithm: scipy/integrate/_ivp/rk.py
// `RungeKutta._step_impl`, `rk_step`, `RkDenseOutput`; ivp.py main loop for t_eval sampling).
//
// Design (B200-first):
//   * persistent CTAs, one per SM; each CTA integrates C columns side by side ("slots"),
//     thread <-> (slot, depth cell), so N=200 gives 600 of 608 lanes busy;
//   * the whole integration of a column happens on-chip: state y and the running
//     y_new / error accumulators live in registers, stage derivatives K1..K6 and a
//     double-buffered stage-input tile live in shared memory (64 kB per column at N=200);
//     HBM is touched only to load y0 and to store snapshots / the final state;
//   * every column runs its own adaptive controller (own t, h, accept/reject); the error
//     norm is a warp-shuffle butterfly over aligned cell groups plus one shared-memory hop, summed
//     in a fixed, slot-independent order: all threads of a column take bit-identical decisions
//     and a column's trajectory does not depend on what else shares the CTA;
//   * one block barrier per RHS evaluation (the stage tile is double buffered) and one for
//     the norm: 7 barriers per step attempt;
//   * finished slots claim the next column from a global atomic queue, so columns with
//     different step counts do not leave SMs idle.
#include <cuda_runtime.h>

#include <cstdint>

#include "lheureux_device.cuh"
#include "rk45_persistent.cuh"

namespace marlpde {

namespace dp {  // Dormand-Prince coefficients, as scipy RK45.{A,B,E,P}
constexpr double a21 = 1.0 / 5.0;
constexpr double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
constexpr double a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0;
constexpr double a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0,
                 a54 = -212.0 / 729.0;
constexpr double a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0,
                 a64 = 49.0 / 176.0, a65 = -5103.0 / 18656.0;
constexpr double b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0,
                 b5 = -2187.0 / 6784.0, b6 = 11.0 / 84.0;
constexpr double e1 = -71.0 / 57600.0, e3 = 71.0 / 16695.0, e4 = -71.0 / 1920.0,
                 e5 = 17253.0 / 339200.0, e6 = -22.0 / 525.0, e7 = 1.0 / 40.0;
// dense output P[s][j], s = stage 1..7 (row 2 is zero), j = 0..3
__constant__ double P[7][4] = {
    {1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0.0, 0.0, 0.0, 0.0},
    {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};
constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
}  // namespace dp

// ---- shared memory carve-up -------------------------------------------------------------
// K[6][5][T] | tile[2][5][T] | y[5][T] | grp[T>>logG] | consts[C] | ctl[C] | log/exp tables | slot_col[C] | svc flag
struct SlotCtl {          // per-slot counters, touched by the slot's leader thread only
  long long n_acc, n_rej, nfev;
};

struct SmemLayout {
  int T;        // C * N, padded to a multiple of 32
  int C;
  size_t off_K, off_tile, off_y, off_grp, off_consts, off_ctl, off_tab, off_slot, total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) / 16 * 16; }

__host__ __device__ inline SmemLayout smem_layout(int C, int N, int logG) {
  SmemLayout L;
  L.C = C;
  L.T = ((C * N + 31) / 32) * 32;
  size_t o = 0;
  L.off_K = o;      o += sizeof(double) * 6 * 5 * (size_t)L.T;
  L.off_tile = o;   o += sizeof(double) * 2 * 5 * (size_t)L.T;
  L.off_y = o;      o += sizeof(double) * 5 * (size_t)L.T;
  L.off_grp = o;    o += align16(sizeof(double) * (size_t)((L.T >> logG) + 1));
  L.off_consts = o; o += align16(sizeof(ColumnConsts)) * (size_t)C;
  L.off_ctl = o;    o += align16(sizeof(SlotCtl)) * (size_t)C;
  L.off_tab = o;    o += (size_t)fm::kTableBytes;
  L.off_slot = o;   o += sizeof(int) * (size_t)(C + 4);
  L.total = align16(o);
  return L;
}

static int group_log2(int n_cells) {
  int logG = 0;
  while (logG < 5 && (n_cells % (2 << logG)) == 0) ++logG;
  return logG;
}

int rk45_columns_per_cta(int n_cells, int smem_budget) {
  if (n_cells < 32 || n_cells > kRk45MaxThreads) return 0;
  int C = kRk45MaxThreads / n_cells;
  while (C > 0 && smem_layout(C, n_cells, group_log2(n_cells)).total > (size_t)smem_budget) --C;
  return C;
}

size_t rk45_smem_bytes(int C, int n_cells) { return smem_layout(C, n_cells, group_log2(n_cells)).total; }

// Stage table: after K_{i+1} has been evaluated (i = 1..5) the next stage input is
// y + h * sum_{j=0..i} kStage[i-1][j] * K_{j+1}; row 4 holds b (the 5th-order weights, FSAL).
__constant__ double kStage[5][6] = {
    {dp::a31, dp::a32, 0, 0, 0, 0},
    {dp::a41, dp::a42, dp::a43, 0, 0, 0},
    {dp::a51, dp::a52, dp::a53, dp::a54, 0, 0},
    {dp::a61, dp::a62, dp::a63, dp::a64, dp::a65, 0},
    {dp::b1, 0.0, dp::b3, dp::b4, dp::b5, dp::b6}};
__constant__ double kErr[7] = {dp::e1, 0.0, dp::e3, dp::e4, dp::e5, dp::e6, dp::e7};

__global__ void __launch_bounds__(kRk45MaxThreads, 1)
rk45_persistent_kernel(double* __restrict__ g_y, const marlpde_column_params* __restrict__ g_params,
                       marlpde_column_state* __restrict__ g_state, int n_columns, int N, int C, int logG,
                       marlpde_rk45_options opt, const double* __restrict__ g_t_eval,
                       double* __restrict__ g_snap, int32_t* __restrict__ g_queue) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmemLayout L = smem_layout(C, N, logG);
  const int T = L.T;
  double* const sK = reinterpret_cast<double*>(smem_raw + L.off_K);        // [6][5][T]
  double* const sTile = reinterpret_cast<double*>(smem_raw + L.off_tile);  // [2][5][T]
  double* const sY = reinterpret_cast<double*>(smem_raw + L.off_y);        // [5][T]
  double* const sGrp = reinterpret_cast<double*>(smem_raw + L.off_grp);    // [T >> logG]
  int* const sSlotCol = reinterpret_cast<int*>(smem_raw + L.off_slot);     // [C]
  int* const sSvc = sSlotCol + C;

  const int tid = threadIdx.x;
  const fm::Tables tb = fm::stage_tables(smem_raw + L.off_tab, tid, blockDim.x);
  const bool active = tid < C * N;
  const int slot = active ? tid / N : C;          // C = "no slot" for the padding lanes
  const int cell = active ? tid - slot * N : 0;
  const int base = slot * N;                       // first tile index of my column
  const ColumnConsts& kc = *reinterpret_cast<const ColumnConsts*>(
      smem_raw + L.off_consts + align16(sizeof(ColumnConsts)) * (active ? slot : 0));
  SlotCtl& ctl = *reinterpret_cast<SlotCtl*>(smem_raw + L.off_ctl + align16(sizeof(SlotCtl)) * (active ? slot : 0));
  const bool leader = active && cell == 0;
  // Error-norm reduction tree, identical for every slot so that a column's trajectory does not
  // depend on where it is scheduled: cells are summed in aligned groups of G = 2^k lanes
  // (G = largest power of two <= 32 dividing N, hence dividing every slot base) by an xor
  // butterfly, then the N/G group sums are added in cell order from shared memory.
  const int G = 1 << logG;
  const int nGroups = N >> logG;
  const int grpBase = base >> logG;

  // per-thread column state (everything else lives in shared memory)
  int col = -1;                 // column index being integrated by my slot, -1 = idle
  bool exhausted = false;       // the work queue ran dry
  bool rejected = false;
  double t = 0.0, h_abs = 0.0, h = 0.0, t_new = 0.0;
  int next_eval = 0;
  int attempts = 0;             // step attempts made for this column in this launch

  auto tileAt = [&](int b, int f, int i) -> double& { return sTile[(b * 5 + f) * T + i]; };
  auto KAt = [&](int s, int f) -> double& { return sK[(s * 5 + f) * T + tid]; };
  auto yAt = [&](int f) -> double& { return sY[f * T + tid]; };

  // RHS of my cell from stage tile `b` (one code instance, see the stage loop)
  auto eval_rhs = [&](int b, CellRates& out) {
    double c[5], m[5], p[5];
    load_triple(kc, cell, [&](int f, int i) { return tileAt(b, f, base + i); }, c, m, p);
    cell_rhs(kc, tb, c, m, p, cell >= kc.mask_lo && cell < kc.mask_hi, out);
  };

  // scipy _step_impl: min_step = 10 * |nextafter(t, inf) - t| ; clamp h_abs at the start of a step
  auto min_step_at = [&](double tt) { return 10.0 * fabs(nextafter(tt, (double)INFINITY) - tt); };
  auto begin_step = [&]() {
    const double ms = min_step_at(t);
    if (h_abs > opt.max_step) h_abs = opt.max_step;
    else if (h_abs < ms) h_abs = ms;
    rejected = false;
  };
  auto begin_attempt = [&]() -> bool {   // false: TOO_SMALL_STEP
    if (h_abs < min_step_at(t)) return false;
    h = h_abs;
    t_new = t + h;
    if (t_new - opt.t_bound > 0.0) t_new = opt.t_bound;
    h = t_new - t;
    h_abs = fabs(h);
    return true;
  };
  auto retire = [&](int status) {  // store the column's end point and free the slot
#pragma unroll
    for (int f = 0; f < 5; ++f) g_y[((size_t)col * 5 + f) * N + cell] = yAt(f);
    if (leader) {
      marlpde_column_state st;
      st.t = t;
      st.h_abs = h_abs;
      st.n_accepted = ctl.n_acc;
      st.n_rejected = ctl.n_rej;
      st.nfev = ctl.nfev;
      st.status = status;
      st.next_eval = next_eval;
      g_state[col] = st;
    }
    col = -1;
  };
  auto write_stage2 = [&]() {
#pragma unroll
    for (int f = 0; f < 5; ++f) tileAt(1, f, tid) = fma(h * dp::a21, KAt(0, f), yAt(f));
  };

  if (tid == 0) *sSvc = 0;
  __syncthreads();

  bool fresh = false;           // column just loaded: K1 = f(y) still to be evaluated (stage i = 0)
  for (;;) {
    // ================= stage 2 input + slot service =======================================
    if (col >= 0) write_stage2();
    if (leader && col < 0 && !exhausted) *sSvc = 1;
    int nlive = __syncthreads_count(col >= 0);
    int i0 = 1;
    if (*sSvc) {
      // -- claim columns for idle slots
      if (leader && col < 0 && !exhausted) {
        const int c = atomicAdd(g_queue, 1);
        sSlotCol[slot] = c < n_columns ? c : -1;
      }
      __syncthreads();
      if (tid == 0) *sSvc = 0;
      if (active && col < 0 && !exhausted) {
        col = sSlotCol[slot];
        if (col < 0) {
          exhausted = true;
        } else {
          const marlpde_column_state st = g_state[col];
          if (leader) {
            ColumnConsts tmp;
            make_consts(g_params[col], N, tmp);
            *const_cast<ColumnConsts*>(&kc) = tmp;
            ctl.n_acc = st.n_accepted;
            ctl.n_rej = st.n_rejected;
            ctl.nfev = st.nfev;
          }
          attempts = 0;
          t = st.t;
          h_abs = st.h_abs;
          next_eval = st.next_eval;
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            const double v = g_y[((size_t)col * 5 + f) * N + cell];
            yAt(f) = v;
            tileAt(0, f, tid) = v;
          }
          if (t >= opt.t_bound) {            // nothing to integrate
            retire(MARLPDE_STATUS_FINISHED);
          } else {
            begin_step();
            if (begin_attempt()) fresh = true;
            else retire(MARLPDE_STATUS_STEP_TOO_SMALL);
          }
        }
      }
      nlive = __syncthreads_count(col >= 0);   // also publishes consts, y and tile 0 of the new columns
      if (nlive == 0) continue;                // everything claimed retired at once: look again
      i0 = 0;
    }
    if (nlive == 0) break;

    // ================= stages: i = 1..6 evaluates K_{i+1} from tile (i & 1) =================
    // (i = 0, only after a slot service: K1 = f(y) of freshly loaded columns from tile 0.)
    // ONE code instance of the RHS serves every stage: the kernel stays inside the instruction
    // cache and K1 of a resumed column is bit-identical to the FSAL K7 it replaces.
    const bool live = col >= 0;
    CellRates r;
#pragma unroll 1
    for (int i = i0; i <= 6; ++i) {
      if (i > i0) __syncthreads();
      if (live && (i > 0 || fresh)) {
        eval_rhs(i & 1, r);
        if (i == 0) {
#pragma unroll
          for (int f = 0; f < 5; ++f) KAt(0, f) = r.r[f];
          if (leader) ctl.nfev += 1;
          write_stage2();
          fresh = false;
        } else if (i < 6) {
          // store K_{i+1}; next stage input (row 4 = y_new) into the other tile
          const double* row = kStage[i - 1];
#pragma unroll
          for (int f = 0; f < 5; ++f) {
            const double k = r.r[f];
            KAt(i, f) = k;
            double acc = row[i] * k;
            for (int j = i - 1; j >= 0; --j) acc = fma(row[j], KAt(j, f), acc);
            tileAt((i + 1) & 1, f, tid) = fma(h, acc, yAt(f));
          }
        }
      }
    }
    // ---- K7 = f(y_new) is in r; error estimate and its norm
    double part = 0.0;
    if (live) {
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        double e = kErr[6] * r.r[f];
#pragma unroll
        for (int j = 5; j >= 0; --j)
          if (j != 1) e = fma(kErr[j], KAt(j, f), e);
        const double ynew = tileAt(0, f, tid);
        const double scale = fma(fmax(fabs(yAt(f)), fabs(ynew)), opt.rtol, opt.atol);
        const double q = (h * e) * fm::rcp(scale);
        part = fma(q, q, part);
      }
    }
    {
      double a = part;
      for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (live && (cell & (G - 1)) == 0) sGrp[tid >> logG] = a;
    }
    __syncthreads();
    if (live) {
      // group sums in cell order (4 interleaved partial sums, fixed association)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double* gp = sGrp + grpBase;
      int gi = 0;
      for (; gi + 4 <= nGroups; gi += 4) {
        s0 += gp[gi];
        s1 += gp[gi + 1];
        s2 += gp[gi + 2];
        s3 += gp[gi + 3];
      }
      for (; gi < nGroups; ++gi) s0 += gp[gi];
      const double sum = (s0 + s1) + (s2 + s3);
      const double err_norm = sqrt(sum / (double)(5 * N));
      if (leader) ctl.nfev += 6;
      attempts += 1;
      if (err_norm < 1.0) {
        double factor = dp::MAX_FACTOR;
        if (err_norm != 0.0) factor = fmin(dp::MAX_FACTOR, dp::SAFETY * fm::exp(tb, -0.2 * fm::log(tb, err_norm)));
        if (rejected) factor = fmin(1.0, factor);
        // ---- dense output for t_eval points in (t, t_new] (ivp.py: searchsorted side='right')
        while (next_eval < opt.n_eval) {
          const double te = g_t_eval[next_eval];
          if (!(te <= t_new)) break;
          const double x = (te - t) / h;
#pragma unroll 1
          for (int f = 0; f < 5; ++f) {
            double q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              double s = dp::P[0][j] * KAt(0, f);
#pragma unroll
              for (int st = 2; st < 6; ++st) s = fma(dp::P[st][j], KAt(st, f), s);
              q[j] = fma(dp::P[6][j], r.r[f], s);
            }
            const double poly = x * (q[0] + x * (q[1] + x * (q[2] + x * q[3])));
            g_snap[(((size_t)col * opt.n_eval + next_eval) * 5 + f) * N + cell] = fma(h, poly, yAt(f));
          }
          ++next_eval;
        }
        // ---- accept
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          yAt(f) = tileAt(0, f, tid);
          KAt(0, f) = r.r[f];           // FSAL
        }
        t = t_new;
        h_abs *= factor;
        if (leader) ctl.n_acc += 1;
        if (t >= opt.t_bound) {
          retire(MARLPDE_STATUS_FINISHED);
        } else if (opt.max_steps > 0 && (long long)attempts >= opt.max_steps) {
          retire(MARLPDE_STATUS_STEP_BUDGET);
        } else {
          begin_step();
          if (!begin_attempt()) retire(MARLPDE_STATUS_STEP_TOO_SMALL);
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

cudaError_t launch_rk45(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                        int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                        double* d_snap, int32_t* d_queue, int sm_count, int smem_budget, cudaStream_t stream) {
  const int C = rk45_columns_per_cta(n_cells, smem_budget);
  if (C <= 0) return cudaErrorInvalidValue;
  const int logG = group_log2(n_cells);
  const SmemLayout L = smem_layout(C, n_cells, logG);
  cudaError_t e = cudaFuncSetAttribute(rk45_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)L.total);
  if (e != cudaSuccess) return e;
  int grid = (n_columns + C - 1) / C;
  if (grid > sm_count) grid = sm_count;
  if (grid < 1) grid = 1;
  rk45_persistent_kernel<<<grid, L.T, L.total, stream>>>(d_y, d_params, d_state, n_columns, n_cells, C, logG, opt,
                                                         d_t_eval, d_snap, d_queue);
  return cudaGetLastError();
}

}  // namespace marlpde
