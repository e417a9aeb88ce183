// radau_batch.cu — batched implicit integrator (3-stage Radau IIA, order 5) for sediment columns.
//
// Replaces `scipy.integrate.solve_ivp(eq.fun_numba, ..., method="Radau", jac_sparsity=...)`, the
// reference's DEFAULT solver (marlpde/parameters.py:213, call site Evolve_scenario.py:104-109;
// algorithm: scipy/integrate/_ivp/radau.py `Radau._step_impl`, `solve_collocation_system`,
// `predict_factor`, `RadauDenseOutput`).  Same constants, same simplified-Newton iteration with the
// same convergence-rate tests, same error-estimate formula (re-filtered once after a rejection), same
// Gustafsson step-size prediction, same Jacobian/LU reuse policy.  One deliberate difference: the linear
// solves — Newton updates AND the error estimate LU_real.solve(f + Z^T E / h) — apply fp32 inverse Schur
// complements without refinement, so the error estimate carries ~cond x 1e-7 of relative error and
// accept / reject decisions are not bit-for-bit SciPy's.  Measured against SciPy Radau on 31 columns of the
// benchmark lattice to T* (profiles/r02i_lattice_radau_vs_scipy.log, tests/test_gpu_lattice.py): factorisation
// and Jacobian counts within 2-4 % column by column, end states within 0.01 tolerance units for 24 of 31
// columns (the rest: columns with a sharp porosity feature, tolerance-sensitive in either code).  A NaN error
// norm rejects the step here (SciPy would accept it and fail later).
//
// What is B200-native about it:
//   * one WARP per sediment column, 12 columns per SM in flight, columns claimed from a global queue; all
//     control flow of a column is warp-uniform, so there is no block barrier anywhere;
//   * the Jacobian is what it is for this PDE: block-TRIDIAGONAL in cell-major order with dense 5x5
//     blocks (the reference hands SciPy a 27-diagonal field-major pattern, parameters.py:150-199,
//     and SciPy then runs a general sparse LU).  Off-diagonal blocks are analytic, the diagonal blocks come
//     from 5 finite-difference evaluations (jac_columns) of the same rhs_pair code as the explicit kernel; it
//     is kept in HBM as [cell][L|D|U][5][5];
//   * the two linear systems of a Radau step, (mu_real/h I - J) and (mu_complex/h I - J), are
//     factorised TOGETHER by one two-ended block-Thomas pass (top-down and bottom-up chains meeting in the
//     middle cell): one entry of [S | I] of both systems per lane, Gauss-Jordan with partial pivoting inside
//     the 5x5 block, pivot search by warp REDUX.  What is stored per cell is S_i^{-1} (fp32: it only
//     preconditions the simplified Newton iteration), so a solve is two 5x5 mat-vecs per cell and direction;
//   * in a solve the two chains of both systems run side by side in one warp: N/2 sequential steps per sweep.
// State vectors are fp64 and live in a per-column HBM workspace; when all 12 x 148 columns are in flight the
// kernel is bound by the DRAM traffic of those vectors and of the fp32 records.
#include "implicit_common.cuh"
#include "radau_batch.cuh"

namespace marlpde {
namespace rd {
using namespace imp;

constexpr double kMuReal = 3.637834252744496;
constexpr double kMuCRe = 2.6810828736277523, kMuCIm = -3.050430199247411;
__constant__ double kC[3] = {0.15505102572168222, 0.6449489742783178, 1.0};
__constant__ double kE[3] = {-10.048809399827414, 1.382142733160748, -0.3333333333333333};
__constant__ double kT[3][3] = {{0.09443876248897524, -0.1412552950209542, 0.03002919410514742},
                                {0.2502131229653333, 0.20412935229379994, -0.3829421127572619},
                                {1.0, 1.0, 0.0}};
__constant__ double kTI[3][3] = {{4.178718591551904, 0.32768282076106237, 0.5233764454994495},
                                 {-4.178718591551904, -0.32768282076106237, 0.47662355450055044},
                                 {0.5028726349457868, -2.571926949855605, 0.5960392048282249}};
__constant__ double kP[3][3] = {{10.048809399827414, -25.62959144707664, 15.580782047249224},
                                {-1.382142733160748, 10.296258113743303, -8.914115380582556},
                                {0.3333333333333333, -2.6666666666666665, 3.3333333333333335}};
constexpr int kNewtonMaxIter = 6;
constexpr double kMinFactor = 0.2, kMaxFactor = 10.0;

// per-column workspace, in doubles (n = 5 N): see radau_workspace_doubles()
struct Work {
  double *y, *yold, *f, *Z, *W, *B, *Q, *err, *tmp;   // n, n, n, 3n, 3n, 3n, 3n, n, n  — all CELL-major [cell][field]
  double* J;                                       // [N][3][5][5]  (L, D, U blocks of the Jacobian, each column-major)
  float* Rec;                                      // [N][128] what the sweeps read, FP32, one 512-byte record per cell:
                                                   // [0,25) L_i | [26,51) S_i^-1 of the real system (real) | [52,102)
                                                   // S_i^-1 of the complex system (float2) | [102,127) U_i, blocks column-
                                                   // major.  They only precondition the simplified Newton iteration (the
                                                   // residual is fp64) and the sweeps are bound by their bytes.
};

__host__ __device__ inline size_t work_doubles(int N) {
  const size_t n = 5 * (size_t)N;
  return 18 * n + 76 * (size_t)N + 64 * (size_t)N;   // 76: keeps the double2 array 16-byte aligned
}



// brentq on the dense output between t_old and t for every monitor in `act` (ivp.py handle_events ->
// solve_event_equation, xtol = rtol = 4 eps).  Rare, so it lives outside the step loop's instruction footprint.
__device__ __noinline__ void locate_events(const Args& A, const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                           int col, unsigned act, double t_old, double t, double h_old,
                                           const double* yold, const double* Q, bool writer) {
#pragma unroll 1
  for (int k = 0; k < 7; ++k) {
    if (!((act >> k) & 1u)) continue;
    BrentState bs;
    bs.init(t_old, t);
    double xeval = t_old, root = t;
    for (;;) {
      const double xx = (xeval - t_old) / h_old;
      double gv[7];
      monitors(kc, tb, N, lane, yold, Q, xx, gv);
      double gk = gv[0];
#pragma unroll
      for (int kk = 1; kk < 7; ++kk) gk = (k == kk) ? gv[kk] : gk;
      if (bs.feed(gk, xeval, root)) break;
    }
    if (lane == 0 && writer) {   // (a team runs the search redundantly on both warps; warp 0 records)
      int32_t* cnt = A.g_ev_counts + (size_t)col * MARLPDE_NEVENTS + k;
      const int have = *cnt;
      if (have < A.opt.event_capacity)
        A.g_ev_times[((size_t)col * MARLPDE_NEVENTS + k) * A.opt.event_capacity + have] = root;
      *cnt = have + 1;
    }
  }
}

__device__ __noinline__ double predict_factor(double h_abs, double h_abs_old, double err, double err_old) {
  // radau.py predict_factor; "None" is encoded as a negative value
  double mult = 1.0;
  if (!(err_old < 0.0 || h_abs_old < 0.0 || err == 0.0)) mult = h_abs / h_abs_old * sqrt(sqrt(err_old / err));
  return fmin(1.0, mult) / sqrt(sqrt(err));      // x**0.25 as two square roots; err == 0 -> inf, as numpy
}

// kTeam = 1: CTAs of kWarpsPerCta warps, one column per WARP (the throughput shape).  kTeam = 2: CTAs of two warps, one
// column per CTA (the latency shape for the longest columns of a sweep, implicit_common.cuh "TEAMS"): six of them are
// resident per SM, the same 12 warps.  Every element-wise loop runs over team lanes (tl, stride ts), every barrier is
// the team barrier, every reduction is team-wide, so both warps of a team take identical decisions.
template <int kTeam>
__device__ __forceinline__ double team_sum(double v, double (*red)[2], int& phase) {
  v = warp_sum(v);
  if (kTeam == 1) return v;
  if ((threadIdx.x & 31) == 0) red[phase][threadIdx.x >> 5] = v;
  __syncthreads();
  const double r = red[phase][0] + red[phase][1];          // fixed order: identical in both warps
  phase ^= 1;                                              // (double-buffered: one barrier per reduction is enough)
  return r;
}

template <bool VD, bool kJacFD, int kTeam>
__global__ void __launch_bounds__((kTeam == 1 ? kWarpsPerCta : 2) * 32, (kTeam == 1 ? 1 : 2) * MARLPDE_RADAU_MINBLOCKS)
radau_kernel(const Args A) {
  constexpr int kWarps = kTeam == 1 ? kWarpsPerCta : 2;
  __shared__ __align__(16) unsigned char tab_raw[fm::kTableBytes];
  __shared__ WarpScratch scratch[kWarps];
  __shared__ double red[2][2];
  __shared__ int s_col;
  const fm::Tables tb = fm::stage_tables(tab_raw, threadIdx.x, blockDim.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wt = kTeam == 1 ? 0 : (int)(threadIdx.x >> 5);          // warp within the team
  const int tl = kTeam == 1 ? lane : (int)threadIdx.x;             // team lane and team width
  constexpr int ts = 32 * kTeam;
  int red_phase = 0;
  WarpScratch& ws = scratch[threadIdx.x >> 5];
  const int N = A.N, n = 5 * N;
  const double rtol = A.opt.rtol, atol = A.opt.atol;
  const double newton_tol = fmax(10.0 * kEps / rtol, fmin(0.03, sqrt(rtol)));

  for (;;) {
    int col = 0;
    if (kTeam == 1) {
      if (lane == 0) col = atomicAdd(A.g_queue, 1);
      col = __shfl_sync(0xffffffffu, col, 0);
    } else {
      if (threadIdx.x == 0) s_col = atomicAdd(A.g_queue, 1);
      __syncthreads();
      col = s_col;
      __syncthreads();                                     // (s_col is rewritten for the next column)
    }
    if (col >= A.n_columns) break;

    // ---- column set-up
    // Inside the kernel every per-column vector is CELL-major ([cell][field]): the block-Thomas sweeps read
    // the five entries of a cell from one 40-byte run instead of five sectors, the RHS reads a thread's two
    // cells as one 80-byte run, element-wise loops are unit stride.  The caller's y / snapshots stay
    // field-major (the reference's layout); they are transposed on the way in and out.
    double* const gy = A.g_y + (size_t)col * n;
    double* wbase = A.g_work + (size_t)col * work_doubles(N);
    Work w;
    w.y = wbase;                  wbase += 2 * n;          // (+ n of padding keeps the offsets below 16-byte aligned)
    double* const y = w.y;
    w.yold = wbase;               wbase += n;
    w.f = wbase;                  wbase += n;
    w.Z = wbase;                  wbase += 3 * n;
    w.W = wbase;                  wbase += 3 * n;
    w.B = wbase;                  wbase += 3 * n;
    w.Q = wbase;                  wbase += 3 * n;
    w.err = wbase;                wbase += n;
    w.tmp = wbase;                wbase += n;
    w.J = wbase;                  wbase += 76 * (size_t)N;
    w.Rec = reinterpret_cast<float*>(wbase);
    if (lane == 0) make_consts(A.g_params[col], N, ws.kc);
#pragma unroll 1
    for (int idx = tl; idx < n; idx += ts) y[idx] = gy[(idx % 5) * N + idx / 5];
    team_sync<kTeam>();
    const ColumnConsts& kc = ws.kc;
    marlpde_column_state st = A.g_state[col];
    double t = st.t, h_attr = st.h_abs;
    long long n_acc = st.n_accepted, n_rej = st.n_rejected, nfev = st.nfev;
    long long njev = 0, nlu = 0, n_newton = 0, n_newton_fail = 0;
    int next_eval = st.next_eval;
    int status = MARLPDE_STATUS_FINISHED;
    long long steps_done = 0;

    auto eval_to = [&](const double* yy, const double* add, double* out) {   // out = rhs(yy + add)
      rhs_eval<VD, kTeam>(&kc, &tb, N, lane, yy, add, out, ws.stage);
    };

    if (t < A.opt.t_bound) {
      eval_to(y, nullptr, w.f);
      nfev += 1;
      jacobian<VD, kJacFD, kTeam>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
      njev += 1;
      nfev += jac_rhs_evals<kJacFD>();   // (finite-difference diagonal blocks: one evaluation per field)
    }
    const bool ev_on = (A.opt.flags & MARLPDE_FLAG_EVENTS) != 0;
    unsigned ev_prev = 0u;
    if (ev_on && t < A.opt.t_bound)                        // ivp.py: g = [event(t0, y0) for event in events]
      ev_prev = monitor_bits(kc, tb, N, lane, y);
    bool current_jac = true, lu_valid = false, have_sol = false;
    double h_abs_old = -1.0, err_old = -1.0;     // "None"
    double t_old = t, h_old = 0.0;

    while (t < A.opt.t_bound) {
      if (A.opt.max_steps > 0 && steps_done >= A.opt.max_steps) {
        status = MARLPDE_STATUS_STEP_BUDGET;
        break;
      }
      // ------------------------------------------------------------------ radau.py _step_impl
      // 10 * |nextafter(t, inf) - t| for t >= 0 (forward integration from t0 >= 0 is validated on the host)
      const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
      double h_abs = h_attr, hao = h_abs_old, eo = err_old;
      if (h_attr > A.opt.max_step) {
        h_abs = A.opt.max_step;
        hao = -1.0;
        eo = -1.0;
      } else if (h_attr < min_step) {
        h_abs = min_step;
        hao = -1.0;
        eo = -1.0;
      }
      bool rejected = false, accepted = false, failed = false;
      double h = 0.0, t_new = t, error_norm = 0.0, safety = 1.0, rate = -1.0;
      int n_iter = 0;
      while (!accepted) {
        if (h_abs < min_step) {
          failed = true;
          break;
        }
        h = h_abs;
        t_new = t + h;
        if (t_new - A.opt.t_bound > 0.0) t_new = A.opt.t_bound;
        h = t_new - t;
        h_abs = fabs(h);

        bool converged = false;
        for (;;) {
          if (!lu_valid) {                 // (radau.py keeps the factors while the step-size factor is 1)
            if (kTeam == 1) {
              factorise<true>(ws, N, lane, kMuReal / h, make_double2(kMuCRe / h, kMuCIm / h), w.J, w.Rec);
            } else {                       // the two chains side by side, then the meeting cell on warp 0
              const int mid = N / 2;
              factorise<true>(ws, N, lane, kMuReal / h, make_double2(kMuCRe / h, kMuCIm / h), w.J, w.Rec,
                              wt == 0 ? 0 : mid, wt == 0 ? mid : N - 1);
              __syncthreads();
              if (wt == 0)
                factorise<true>(ws, N, lane, kMuReal / h, make_double2(kMuCRe / h, kMuCIm / h), w.J, w.Rec, N - 1, N,
                                &scratch[1]);
              __syncthreads();
            }
            nlu += 2;
            lu_valid = true;
          }
          // ---- Z0 from the previous step's dense output (radau.py: Z0 = sol(t + h C).T - y), W = TI Z0
          _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts) {
            double z[3] = {0.0, 0.0, 0.0};
            if (have_sol) {
              const double q0 = w.Q[idx], q1 = w.Q[n + idx], q2 = w.Q[2 * n + idx];
              const double base = w.yold[idx] - y[idx];
#pragma unroll
              for (int sgi = 0; sgi < 3; ++sgi) {
                const double xx = (t + h * kC[sgi] - t_old) / h_old;
                z[sgi] = base + xx * (q0 + xx * (q1 + xx * q2));
              }
            }
#pragma unroll
            for (int sgi = 0; sgi < 3; ++sgi) {
              w.Z[sgi * n + idx] = z[sgi];
              w.W[sgi * n + idx] = kTI[sgi][0] * z[0] + kTI[sgi][1] * z[1] + kTI[sgi][2] * z[2];
            }
          }
          team_sync<kTeam>();
          // ---- solve_collocation_system
          const double Mr = kMuReal / h, Mcr = kMuCRe / h, Mci = kMuCIm / h;
          double dW_norm_old = -1.0;
          rate = -1.0;
          converged = false;
          int k = 0;
          for (k = 0; k < kNewtonMaxIter; ++k) {
            eval_to(y, w.Z, w.B);
            eval_to(y, w.Z + n, w.B + n);
            eval_to(y, w.Z + 2 * n, w.B + 2 * n);
            nfev += 3;
            bool finite = true;
            // (element-wise passes; RADAU_BATCH4 issues 4 strides of loads per trip)
#pragma unroll 1
            for (int base = tl; base < n; base += (RADAU_BATCH4 ? 4 : 1) * ts) {
              double F0[4], F1[4], F2[4], W0[4], W1[4], W2[4];
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + ts * u;
                const bool ok = idx < n;
                F0[u] = ok ? w.B[idx] : 0.0;
                F1[u] = ok ? w.B[n + idx] : 0.0;
                F2[u] = ok ? w.B[2 * n + idx] : 0.0;
                W0[u] = ok ? w.W[idx] : 0.0;
                W1[u] = ok ? w.W[n + idx] : 0.0;
                W2[u] = ok ? w.W[2 * n + idx] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + ts * u;
                if (idx >= n) continue;
                finite = finite && isfinite(F0[u]) && isfinite(F1[u]) && isfinite(F2[u]);
                // f_real = F^T TI_REAL - M_real W0 ; f_complex = F^T TI_COMPLEX - M_complex (W1 + i W2)
                w.B[idx] = kTI[0][0] * F0[u] + kTI[0][1] * F1[u] + kTI[0][2] * F2[u] - Mr * W0[u];
                w.B[n + idx] = kTI[1][0] * F0[u] + kTI[1][1] * F1[u] + kTI[1][2] * F2[u] - (Mcr * W1[u] - Mci * W2[u]);
                w.B[2 * n + idx] = kTI[2][0] * F0[u] + kTI[2][1] * F1[u] + kTI[2][2] * F2[u] - (Mcr * W2[u] + Mci * W1[u]);
              }
            }
            team_sync<kTeam>();
            if (!team_all<kTeam>(finite)) break;
            if (wt == 0) solve<false>(ws, N, lane, w.Rec, w.B, w.B + n, w.B + 2 * n, true);
            team_sync<kTeam>();
            // norm(dW / scale) and, in the same pass, W += dW, Z = T W.  (radau.py leaves W and Z untouched when
            // the rate test below breaks; they are dead then — every continuation re-initialises them from Z0.)
            double ss = 0.0;
#pragma unroll 1
            for (int base = tl; base < n; base += (RADAU_BATCH4 ? 4 : 1) * ts) {
              double d0[4], d1[4], d2[4], yy4[4], W0[4], W1[4], W2[4];
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + ts * u;
                const bool ok = idx < n;
                d0[u] = ok ? w.B[idx] : 0.0;
                d1[u] = ok ? w.B[n + idx] : 0.0;
                d2[u] = ok ? w.B[2 * n + idx] : 0.0;
                yy4[u] = ok ? y[idx] : 0.0;
                W0[u] = ok ? w.W[idx] : 0.0;
                W1[u] = ok ? w.W[n + idx] : 0.0;
                W2[u] = ok ? w.W[2 * n + idx] : 0.0;
              }
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + ts * u;
                if (idx >= n) continue;
                const double sc = fma(fabs(yy4[u]), rtol, atol);
                const double a0 = d0[u] / sc, a1 = d1[u] / sc, a2 = d2[u] / sc;
                ss += a0 * a0 + a1 * a1 + a2 * a2;
                const double V0 = W0[u] + d0[u], V1 = W1[u] + d1[u], V2 = W2[u] + d2[u];
                w.W[idx] = V0;
                w.W[n + idx] = V1;
                w.W[2 * n + idx] = V2;
#pragma unroll
                for (int sgi = 0; sgi < 3; ++sgi) w.Z[sgi * n + idx] = kT[sgi][0] * V0 + kT[sgi][1] * V1 + kT[sgi][2] * V2;
              }
            }
            team_sync<kTeam>();
            const double dW_norm = sqrt(team_sum<kTeam>(ss, red, red_phase) / (double)(3 * n));
            if (dW_norm_old >= 0.0) rate = dW_norm / dW_norm_old;
            double rate_pow = rate;                     // rate ** (NEWTON_MAXITER - k)
            for (int e = 1; e < kNewtonMaxIter - k; ++e) rate_pow *= rate;
            if (rate >= 0.0 && (rate >= 1.0 || rate_pow / (1.0 - rate) * dW_norm > newton_tol)) break;
            if (dW_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dW_norm < newton_tol)) {
              converged = true;
              break;
            }
            dW_norm_old = dW_norm;
          }
          n_iter = (k < kNewtonMaxIter ? k : kNewtonMaxIter - 1) + 1;
          n_newton += n_iter;
          if (converged) break;
          n_newton_fail += 1;
          if (current_jac) break;
          jacobian<VD, kJacFD, kTeam>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
          njev += 1;
          nfev += jac_rhs_evals<kJacFD>();   // (finite-difference diagonal blocks: one evaluation per field)
          current_jac = true;
          lu_valid = false;
        }
        if (!converged) {
          h_abs *= 0.5;
          lu_valid = false;
          continue;
        }
        // ---- error estimate: error = LU_real.solve(f + Z^T E / h)
        _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts) {
          const double ze = (w.Z[idx] * kE[0] + w.Z[n + idx] * kE[1] + w.Z[2 * n + idx] * kE[2]) / h;
          w.tmp[idx] = ze;
          w.err[idx] = w.f[idx] + ze;
        }
        team_sync<kTeam>();
        if (wt == 0) solve<false>(ws, N, lane, w.Rec, w.err, nullptr, nullptr, false);
        team_sync<kTeam>();
        auto err_norm_of = [&]() {
          double ss = 0.0;
          _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts) {
            const double yn = y[idx] + w.Z[2 * n + idx];
            const double sc = fma(fmax(fabs(y[idx]), fabs(yn)), rtol, atol);
            const double a = w.err[idx] / sc;
            ss += a * a;
          }
          return sqrt(team_sum<kTeam>(ss, red, red_phase) / (double)n);
        };
        error_norm = err_norm_of();
        safety = 0.9 * (2 * kNewtonMaxIter + 1) / (double)(2 * kNewtonMaxIter + n_iter);
        if (rejected && error_norm > 1.0) {
          // error = LU_real.solve(fun(t, y + error) + ZE)
          eval_to(y, w.err, w.B);
          nfev += 1;
          _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts) w.err[idx] = w.B[idx] + w.tmp[idx];
          team_sync<kTeam>();
          if (wt == 0) solve<false>(ws, N, lane, w.Rec, w.err, nullptr, nullptr, false);
        team_sync<kTeam>();
          error_norm = err_norm_of();
        }
        if (error_norm > 1.0 || !(error_norm == error_norm)) {
          const double factor = predict_factor(h_abs, hao, error_norm, eo);
          h_abs *= fmax(kMinFactor, safety * factor);
          lu_valid = false;
          rejected = true;
          n_rej += 1;
        } else {
          accepted = true;
        }
      }
      if (failed) {
        status = MARLPDE_STATUS_STEP_TOO_SMALL;
        break;
      }
      // ------------------------------------------------------------------ accepted step
      const bool recompute_jac = n_iter > 2 && rate > 1e-3;
      double factor = fmin(kMaxFactor, safety * predict_factor(h_abs, hao, error_norm, eo));
      if (!recompute_jac && factor < 1.2) factor = 1.0;
      else lu_valid = false;
      // y_old <- y, y <- y + Z[2], Q = Z^T P (dense output of this step)
      _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts) {
        const double z0 = w.Z[idx], z1 = w.Z[n + idx], z2 = w.Z[2 * n + idx];
        const double yo = y[idx];
        w.yold[idx] = yo;
        y[idx] = yo + z2;
#pragma unroll
        for (int j = 0; j < 3; ++j) w.Q[j * n + idx] = z0 * kP[0][j] + z1 * kP[1][j] + z2 * kP[2][j];
      }
      team_sync<kTeam>();
      eval_to(y, nullptr, w.f);
      nfev += 1;
      if (recompute_jac) {
        jacobian<VD, kJacFD, kTeam>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
        njev += 1;
        nfev += jac_rhs_evals<kJacFD>();   // (finite-difference diagonal blocks: one evaluation per field)
        current_jac = true;
      } else {
        current_jac = false;
      }
      h_abs_old = h_attr;
      err_old = error_norm;
      h_attr = h_abs * factor;
      t_old = t;
      h_old = h;
      t = t_new;
      have_sol = true;
      n_acc += 1;
      steps_done += 1;
      // ---- events (ivp.py main loop: after every accepted step, before the t_eval samples)
      if (ev_on) {
        const unsigned ev_new = monitor_bits(kc, tb, N, lane, y);
        unsigned act = 0u;
        if (ev_new != ev_prev || (ev_new & kEqBitsMask) != 0u) act = active_events(event_classes(ev_prev), event_classes(ev_new));
        ev_prev = ev_new;
        if (act) locate_events(A, kc, tb, N, lane, col, act, t_old, t, h_old, w.yold, w.Q, wt == 0);   // rare: out of line
      }
      // ---- t_eval samples in (t_old, t] (t_eval[0] == t0 belongs to the first step): y_old + Q p(x)
      while (next_eval < A.opt.n_eval) {
        const double te = A.g_t_eval[next_eval];
        if (!(te <= t)) break;
        const double xx = (te - t_old) / h_old;
        double* snap = A.g_snap + ((size_t)col * A.opt.n_eval + next_eval) * n;
        _Pragma("unroll 1") for (int idx = tl; idx < n; idx += ts)
          snap[(idx % 5) * N + idx / 5] = w.yold[idx] + xx * (w.Q[idx] + xx * (w.Q[n + idx] + xx * w.Q[2 * n + idx]));
        ++next_eval;
      }
    }
#pragma unroll 1
    for (int idx = tl; idx < n; idx += ts) gy[(idx % 5) * N + idx / 5] = y[idx];
    if (tl == 0) {
      st.t = t;
      st.h_abs = h_attr;
      st.n_accepted = n_acc;
      st.n_rejected = n_rej;
      st.nfev = nfev;
      st.status = status;
      st.next_eval = next_eval;
      A.g_state[col] = st;
      int64_t* s4 = A.g_stats + (size_t)col * 4;
      s4[0] += njev;
      s4[1] += nlu;
      s4[2] += n_newton;
      s4[3] += n_newton_fail;
    }
    team_sync<kTeam>();
  }
}

}  // namespace rd

size_t radau_workspace_bytes(int n_columns, int n_cells) {
  return sizeof(double) * rd::work_doubles(n_cells) * (size_t)n_columns;
}

#ifndef MARLPDE_HOST_EMU
cudaError_t launch_radau(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                         int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                         double* d_snap, int64_t* d_stats, int32_t* d_ev_counts, double* d_ev_times, double* d_work,
                         int32_t* d_queue, int sm_count, cudaStream_t stream) {
  rd::Args a;
  a.g_y = d_y;
  a.g_params = d_params;
  a.g_state = d_state;
  a.g_t_eval = d_t_eval;
  a.g_snap = d_snap;
  a.g_stats = d_stats;
  a.g_ev_counts = d_ev_counts;
  a.g_ev_times = d_ev_times;
  a.g_work = d_work;
  a.g_queue = d_queue;
  a.n_columns = n_columns;
  a.N = n_cells;
  a.opt = opt;
  // opt.quantum (implicit integrators): the first `quantum` columns of the batch — the longest ones of a cost-ordered
  // sweep — run in TEAM mode (two warps per column, a CTA each) on a second stream, side by side with the one-warp-per-
  // column launch of the others; their queue counter is d_queue[1].  Needs the analytic Jacobian.
  const bool vd = (opt.flags & MARLPDE_FLAG_VAR_DPHI) != 0, fd = (opt.flags & MARLPDE_FLAG_JAC_FD) != 0;
  int n_team = fd ? 0 : opt.quantum;
  if (n_team < 0) n_team = 0;
  if (n_team > n_columns) n_team = n_columns;
  cudaStream_t s2 = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (n_team > 0) {
    cudaError_t e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&e0, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&e1, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, e0, 0);
    if (e != cudaSuccess) return e;
    rd::Args t = a;
    t.n_columns = n_team;
    t.g_queue = d_queue + 1;
    int ctas2 = n_team < sm_count * 2 * MARLPDE_RADAU_MINBLOCKS ? n_team : sm_count * 2 * MARLPDE_RADAU_MINBLOCKS;
    if (vd) rd::radau_kernel<true, false, 2><<<ctas2, 64, 0, s2>>>(t);
    else rd::radau_kernel<false, false, 2><<<ctas2, 64, 0, s2>>>(t);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t n = 5 * (size_t)n_cells;
    a.g_y += (size_t)n_team * n;
    a.g_params += n_team;
    a.g_state += n_team;
    if (a.g_snap) a.g_snap += (size_t)n_team * (size_t)opt.n_eval * n;
    a.g_stats += (size_t)n_team * 4;
    if (a.g_ev_counts) a.g_ev_counts += (size_t)n_team * MARLPDE_NEVENTS;
    if (a.g_ev_times) a.g_ev_times += (size_t)n_team * MARLPDE_NEVENTS * (size_t)(opt.event_capacity > 0 ? opt.event_capacity : 0);
    a.g_work += (size_t)n_team * rd::work_doubles(n_cells);
    a.n_columns = n_columns - n_team;
  }
  if (a.n_columns > 0) {
    int ctas = (a.n_columns + rd::kWarpsPerCta - 1) / rd::kWarpsPerCta;
    const int max_ctas = sm_count * MARLPDE_RADAU_MINBLOCKS;
    if (ctas > max_ctas) ctas = max_ctas;
    const int threads = rd::kWarpsPerCta * 32;
    if (vd && fd) rd::radau_kernel<true, true, 1><<<ctas, threads, 0, stream>>>(a);
    else if (vd) rd::radau_kernel<true, false, 1><<<ctas, threads, 0, stream>>>(a);
    else if (fd) rd::radau_kernel<false, true, 1><<<ctas, threads, 0, stream>>>(a);
    else rd::radau_kernel<false, false, 1><<<ctas, threads, 0, stream>>>(a);
  }
  cudaError_t le = cudaGetLastError();
  if (n_team > 0) {
    if (le == cudaSuccess) le = cudaEventRecord(e1, s2);
    if (le == cudaSuccess) le = cudaStreamWaitEvent(stream, e1, 0);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s2);
  }
  return le;
}

#endif

}  // namespace marlpde
