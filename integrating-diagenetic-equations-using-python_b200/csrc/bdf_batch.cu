// bdf_batch.cu — batched variable-order BDF integrator (orders 1-5, quasi-constant step, NDF coefficients) for sediment
// columns: the device counterpart of the reference's multistep implicit solvers.
//
// Replaces `scipy.integrate.solve_ivp(eq.fun_numba, ..., method="BDF", jac_sparsity=...)` (marlpde/parameters.py:213-216,
// :235-236; call site Evolve_scenario.py:104-109) step for step — algorithm: scipy/integrate/_ivp/bdf.py `BDF._step_impl`,
// `solve_bdf_system`, `change_D`, `BdfDenseOutput` — and is the device path for `method="LSODA"` batches (parameters.py:
// 214-219): on this stiff, diffusion-dominated system LSODA leaves its Adams mode within the first few steps and then IS
// a variable-order BDF code (measured with odeint's `mused` on the oracle: tests/test_oracle_golden.py::
// test_lsoda_runs_in_bdf_mode); its banded Jacobian with lband = uband = 1 on the field-major vector drops every
// coupling between fields, the block-tridiagonal Jacobian used here is the exact structure.
//
// Same semantics as bdf.py: differences array D[0..order+2], predictor sum(D), psi, simplified Newton (NEWTON_MAXITER = 4)
// with the same convergence-rate tests, error = error_const[order] d, the stale-LU-after-rejection policy, order selection
// from the three error norms after order+1 equal steps, step-size factors in [0.2, 10], Jacobian refreshed only after a
// Newton failure, dense output from D for t_eval and event location.  What differs is the linear algebra (shared with the
// Radau kernel, implicit_common.cuh): one warp per column, block-tridiagonal Jacobian with analytic off-diagonal blocks,
// two-ended block-Thomas factorisation of (I/c - J) in fp64, fp32 inverse Schur complements applied in the Newton sweeps
// (they only precondition the iteration: the residual c f(y) - psi - d is fp64, and the error estimate is the converged
// correction d itself, so — unlike Radau's — it carries no fp32 error).
// A resumed column (step budget) restarts at order 1 from its stored state: the differences array is not kept.
#include "implicit_common.cuh"
#include "bdf_batch.cuh"

namespace marlpde {
namespace bd {
using namespace imp;

constexpr int kMaxOrder = 5;
constexpr int kNewtonMaxIter = 4;
constexpr double kMinFactor = 0.2, kMaxFactor = 10.0;
// bdf.py __init__: kappa = [0, -0.1850, -1/9, -0.0823, -0.0415, 0], gamma = [0, cumsum(1/j)], alpha = (1 - kappa) gamma,
// error_const = kappa gamma + 1/(j + 1)
__constant__ double kGamma[6] = {0.0, 1.0, 1.5, 1.8333333333333333, 2.083333333333333, 2.283333333333333};
__constant__ double kAlpha[6] = {0.0, 1.185, 1.6666666666666667, 1.9842166666666667, 2.1697916666666663, 2.283333333333333};
__constant__ double kErrConst[6] = {1.0, 0.315, 0.16666666666666666, 0.09911666666666669, 0.11354166666666668,
                                    0.16666666666666666};

// per-column workspace in doubles (n = 5 N), all vectors CELL-major [cell][field]:
//   D[8][n] | y[n] | psi[n] | d[n] | f[n] | b[n] | tmp[n] | Jacobian scratch [2n] | J [N][3][5][5] (+N) | Rec [N][128] fp32
__host__ __device__ inline size_t work_doubles(int N) {
  const size_t n = 5 * (size_t)N;
  return 16 * n + 76 * (size_t)N + 64 * (size_t)N;
}

// change_D (bdf.py): D[:order+1] <- (R U)^T D[:order+1] with R = compute_R(order, factor), U = compute_R(order, 1);
// R[i][j] = prod_{k=1..i} (k - 1 - factor j) / k.  The 6 x 6 matrices are formed by the warp in shared memory.
__device__ __noinline__ void change_D(WarpScratch& ws, int n, int lane, double* D, int order, double factor) {
  double* const R = ws.cd[0];
  double* const U = ws.cd[1];
  double* const RU = ws.cd[2];
#pragma unroll 1
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e - 6 * i;
    double r = 1.0, u = 1.0;
    for (int k = 1; k <= i; ++k) {
      r *= ((double)(k - 1) - factor * (double)j) / (double)k;
      u *= (double)(k - 1 - j) / (double)k;
    }
    R[e] = r;
    U[e] = u;
  }
  __syncwarp();
#pragma unroll 1
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e - 6 * i;
    double s = 0.0;
    for (int k = 0; k <= order; ++k) s = fma(R[i * 6 + k], U[k * 6 + j], s);
    RU[e] = s;
  }
  __syncwarp();
#pragma unroll 1
  for (int idx = lane; idx < n; idx += 32) {
    double dk[6], out[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) dk[k] = k <= order ? D[(size_t)k * n + idx] : 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fma(RU[k * 6 + i], dk[k], s);   // (rows k > order of dk are zero)
      out[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i <= order) D[(size_t)i * n + idx] = out[i];
  }
  __syncwarp();
}

// BdfDenseOutput._call_impl at one time te: out = D[0] + sum_j D[j] prod_{m<j} (te - (t - h m)) / (h (1 + m))
__device__ __forceinline__ void dense_coeffs(double te, double t, double h, int order, double (&p)[kMaxOrder]) {
  double acc = 1.0;
#pragma unroll
  for (int m = 0; m < kMaxOrder; ++m) {
    acc *= (te - (t - h * (double)m)) / (h * (double)(1 + m));
    p[m] = m < order ? acc : 0.0;
  }
}
__device__ __forceinline__ double dense_value(const double* D, int n, int idx, int order, const double (&p)[kMaxOrder]) {
  double v = 0.0;
#pragma unroll
  for (int j = 1; j <= kMaxOrder; ++j)
    if (j <= order) v = fma(D[(size_t)j * n + idx], p[j - 1], v);
  return D[idx] + v;
}

// brentq on the dense output between t_old and t for every monitor in `act` (ivp.py handle_events): the interpolated
// state goes through a scratch vector, events are rare.
__device__ __noinline__ void locate_events(const Args& A, const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                           int col, unsigned act, double t_old, double t, double h, int order,
                                           const double* D, double* tmp) {
  const int n = 5 * N;
#pragma unroll 1
  for (int k = 0; k < 7; ++k) {
    if (!((act >> k) & 1u)) continue;
    BrentState bs;
    bs.init(t_old, t);
    double xeval = t_old, root = t;
    for (;;) {
      double p[kMaxOrder];
      dense_coeffs(xeval, t, h, order, p);
#pragma unroll 1
      for (int idx = lane; idx < n; idx += 32) tmp[idx] = dense_value(D, n, idx, order, p);
      __syncwarp();
      double gv[7];
      monitors(kc, tb, N, lane, tmp, nullptr, 0.0, gv);
      double gk = gv[0];
#pragma unroll
      for (int kk = 1; kk < 7; ++kk) gk = (k == kk) ? gv[kk] : gk;
      __syncwarp();
      if (bs.feed(gk, xeval, root)) break;
    }
    if (lane == 0) {
      int32_t* cnt = A.g_ev_counts + (size_t)col * MARLPDE_NEVENTS + k;
      const int have = *cnt;
      if (have < A.opt.event_capacity)
        A.g_ev_times[((size_t)col * MARLPDE_NEVENTS + k) * A.opt.event_capacity + have] = root;
      *cnt = have + 1;
    }
  }
}

// THE one instance of the RHS in this kernel, fused with the Newton residual: out = Mc f(yy) - Ml (psi + d) — the right-
// hand side M_lu (c f - psi - d) of solve_bdf_system's linear system in the scaling of factorise() — so f never travels
// through memory (Mc = 1, Ml = 0 gives the plain f(yy)).  Returns whether every lane's rates were finite.
template <bool VD>
__device__ __noinline__ bool rhs_residual(const ColumnConsts* kc, const fm::Tables* tb, int N, int lane, const double* yy,
                                          const double* psi, const double* d, double Mc, double Ml, double* out,
                                          double (*stage)[2][kStageDoubles + 6]) {
  bool finite = true;
  auto sink = [&](int i, const double (&r5)[5]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      finite = finite && isfinite(r5[f]);
      out[i * 5 + f] = fma(Mc, r5[f], -Ml * (psi[i * 5 + f] + d[i * 5 + f]));
    }
  };
  rhs_column<VD>(*kc, *tb, N, lane, yy, nullptr, stage, sink);
  __syncwarp();
  return __all_sync(0xffffffffu, finite);
}

// resident CTAs per SM (r02v, 4096 / 64 lattice columns to t = 0.05): 2 -> 2.40 / 0.73 s, 3 -> 1.90 / 0.75 s, 4 -> 1.83 / 0.84 s
#ifndef MARLPDE_BDF_MINBLOCKS
#define MARLPDE_BDF_MINBLOCKS 3
#endif

template <bool VD, bool kJacFD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MARLPDE_BDF_MINBLOCKS) bdf_kernel(const Args A) {
  __shared__ __align__(16) unsigned char tab_raw[fm::kTableBytes];
  __shared__ WarpScratch scratch[kWarpsPerCta];
  const fm::Tables tb = fm::stage_tables(tab_raw, threadIdx.x, blockDim.x);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WarpScratch& ws = scratch[threadIdx.x >> 5];
  const int N = A.N, n = 5 * N;
  const double rtol = A.opt.rtol, atol = A.opt.atol;
  const double newton_tol = fmax(10.0 * kEps / rtol, fmin(0.03, sqrt(rtol)));

  for (;;) {
    int col = 0;
    if (lane == 0) col = atomicAdd(A.g_queue, 1);
    col = __shfl_sync(0xffffffffu, col, 0);
    if (col >= A.n_columns) break;

    double* const gy = A.g_y + (size_t)col * n;
    double* wbase = A.g_work + (size_t)col * work_doubles(N);
    double* const D = wbase;        wbase += 8 * (size_t)n;
    double* const y = wbase;        wbase += n;
    double* const psi = wbase;      wbase += n;
    double* const d = wbase;        wbase += n;
    double* const f = wbase;        wbase += n;
    double* const b = wbase;        wbase += n;
    double* const tmp = wbase;      wbase += n;
    double* const jscr = wbase;     wbase += 2 * (size_t)n;
    double* const J = wbase;        wbase += 76 * (size_t)N;
    float* const Rec = reinterpret_cast<float*>(wbase);
    if (lane == 0) make_consts(A.g_params[col], N, ws.kc);
#pragma unroll 1
    for (int idx = lane; idx < n; idx += 32) {
      const double v = gy[(idx % 5) * N + idx / 5];
      y[idx] = v;
      D[idx] = v;
      psi[idx] = 0.0;
      d[idx] = 0.0;
    }
    __syncwarp();
    const ColumnConsts& kc = ws.kc;
    marlpde_column_state st = A.g_state[col];
    double t = st.t, h_abs = st.h_abs;
    long long n_acc = st.n_accepted, n_rej = st.n_rejected, nfev = st.nfev;
    long long njev = 0, nlu = 0, n_newton = 0, n_newton_fail = 0;
    int next_eval = st.next_eval;
    int status = MARLPDE_STATUS_FINISHED;
    long long steps_done = 0;
    int order = 1, n_equal_steps = 0;
    bool lu_valid = false;
    double M_lu = 0.0;                                   // 1/c of the factors in Rec

    auto eval_to = [&](const double* yy, double* out) {   // out = f(yy)   (psi, d are finite: zeroed at column set-up)
      rhs_residual<VD>(&kc, &tb, N, lane, yy, psi, d, 1.0, 0.0, out, ws.stage);
    };
    auto jacobian = [&](const double* yy) {              // J at yy (finite-difference build: num_jac with its own f0)
      if (kJacFD) {
        eval_to(yy, tmp);
        nfev += 1 + jac_rhs_evals<kJacFD>();
      }
      imp::jacobian<VD, kJacFD>(kc, tb, N, lane, yy, tmp, atol, J, jscr, ws.stage);
      njev += 1;
    };

    if (t < A.opt.t_bound) {
      eval_to(y, f);
      nfev += 1;
      jacobian(y);
#pragma unroll 1
      for (int idx = lane; idx < n; idx += 32) D[(size_t)n + idx] = f[idx] * h_abs;   // D[1] = f h
      __syncwarp();
    }
    const bool ev_on = (A.opt.flags & MARLPDE_FLAG_EVENTS) != 0;
    unsigned ev_prev = 0u;
    if (ev_on && t < A.opt.t_bound) ev_prev = monitor_bits(kc, tb, N, lane, y);

    while (t < A.opt.t_bound) {
      if (A.opt.max_steps > 0 && steps_done >= A.opt.max_steps) {
        status = MARLPDE_STATUS_STEP_BUDGET;
        break;
      }
      // ------------------------------------------------------------------ bdf.py _step_impl
      const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
      if (h_abs > A.opt.max_step) {
        change_D(ws, n, lane, D, order, A.opt.max_step / h_abs);
        h_abs = A.opt.max_step;
        n_equal_steps = 0;
      } else if (h_abs < min_step) {
        change_D(ws, n, lane, D, order, min_step / h_abs);
        h_abs = min_step;
        n_equal_steps = 0;
      }
      bool current_jac = false, accepted = false, failed = false;
      double t_new = t, safety = 1.0, error_norm = 0.0;
      int n_iter = 0;
      while (!accepted) {
        if (h_abs < min_step) {
          failed = true;
          break;
        }
        double h = h_abs;
        t_new = t + h;
        if (t_new - A.opt.t_bound > 0.0) {
          t_new = A.opt.t_bound;
          change_D(ws, n, lane, D, order, fabs(t_new - t) / h_abs);
          n_equal_steps = 0;
          lu_valid = false;
        }
        h = t_new - t;
        h_abs = fabs(h);
        const double c = h / kAlpha[order];

        bool converged = false;
        for (;;) {
          if (!lu_valid) {
            M_lu = 1.0 / c;
            factorise<false>(ws, N, lane, M_lu, make_double2(0.0, 0.0), J, Rec);
            nlu += 1;
            lu_valid = true;
          }
          // ---- y_predict = sum(D[:order+1]), psi = D[1:order+1]^T gamma[1:order+1] / alpha[order], d = 0
          {
            const double ia = 1.0 / kAlpha[order];
            _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
              double s = D[idx], ps = 0.0;
#pragma unroll
              for (int k = 1; k <= kMaxOrder; ++k)
                if (k <= order) {
                  const double v = D[(size_t)k * n + idx];
                  s += v;
                  ps = fma(v, kGamma[k], ps);
                }
              y[idx] = s;
              psi[idx] = ps * ia;
              d[idx] = 0.0;
            }
            __syncwarp();
          }
          // ---- solve_bdf_system
          const double Mc = M_lu * c;                    // (I - c_lu J)^{-1} v = M_lu (M_lu I - J)^{-1} v: stale factors keep their c
          double dy_norm_old = -1.0, rate = -1.0;
          int k = 0;
          bool predict_invalid = false;                  // f(y_predict) itself is not finite
          for (k = 0; k < kNewtonMaxIter; ++k) {
            const bool finite = rhs_residual<VD>(&kc, &tb, N, lane, y, psi, d, Mc, M_lu, b, ws.stage);   // b = M_lu (c f - psi - d)
            nfev += 1;
            if (!finite) {
              predict_invalid = k == 0;
              break;
            }
            solve<true>(ws, N, lane, Rec, b, nullptr, nullptr, false);
            // norm(dy / scale) with scale = atol + rtol |y_predict| (y_predict = y - d), and in the same pass y += dy,
            // d += dy (bdf.py leaves them untouched when the rate test below breaks; they are dead then)
            double ss = 0.0;
            _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
              const double dy = b[idx], yv = y[idx], dv = d[idx];
              const double a = dy / fma(fabs(yv - dv), rtol, atol);
              ss = fma(a, a, ss);
              y[idx] = yv + dy;
              d[idx] = dv + dy;
            }
            __syncwarp();
            const double dy_norm = sqrt(warp_sum(ss) / (double)n);
            rate = dy_norm_old >= 0.0 ? dy_norm / dy_norm_old : -1.0;
            double rate_pow = rate;                      // rate ** (NEWTON_MAXITER - k)
            for (int e = 1; e < kNewtonMaxIter - k; ++e) rate_pow *= rate;
            if (rate >= 0.0 && (rate >= 1.0 || rate_pow / (1.0 - rate) * dy_norm > newton_tol)) break;
            if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < newton_tol)) {
              converged = true;
              break;
            }
            dy_norm_old = dy_norm;
          }
          n_iter = (k < kNewtonMaxIter ? k : kNewtonMaxIter - 1) + 1;
          n_newton += n_iter;
          if (converged) break;
          n_newton_fail += 1;
          if (current_jac) break;
          // bdf.py evaluates J = self.jac(t_new, y_predict) here.  ONE deliberate difference: when the predictor itself has
          // left the model's domain (extrapolated porosity <= 0: log(Phi) is NaN) that Jacobian is NaN, stays "current"
          // for the rest of the step and every smaller step size fails on it until the step size underflows (r02t: 87 of
          // 4096 lattice columns stalled this way; SciPy's own BDF has the same trap).  The step size is halved instead and
          // the Jacobian is refreshed at a predictor that can be evaluated.
          if (predict_invalid) break;
          _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
            double s = D[idx];
#pragma unroll
            for (int kk = 1; kk <= kMaxOrder; ++kk)
              if (kk <= order) s += D[(size_t)kk * n + idx];
            y[idx] = s;
          }
          __syncwarp();
          jacobian(y);
          current_jac = true;
          lu_valid = false;
        }
        if (!converged) {
          h_abs *= 0.5;
          change_D(ws, n, lane, D, order, 0.5);
          n_equal_steps = 0;
          lu_valid = false;
          continue;
        }
        safety = 0.9 * (2 * kNewtonMaxIter + 1) / (double)(2 * kNewtonMaxIter + n_iter);
        {
          const double ec = kErrConst[order];
          double ss = 0.0;
          _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
            const double a = ec * d[idx] / fma(fabs(y[idx]), rtol, atol);
            ss = fma(a, a, ss);
          }
          error_norm = sqrt(warp_sum(ss) / (double)n);
        }
        if (!(error_norm <= 1.0)) {                      // (a NaN norm rejects the step; bdf.py would accept it)
          const double factor = fmax(kMinFactor, safety * pow(error_norm, -1.0 / (double)(order + 1)));
          h_abs *= factor;
          change_D(ws, n, lane, D, order, factor);
          n_equal_steps = 0;
          n_rej += 1;                                    // (no trouble with convergence: the factors are kept)
        } else {
          accepted = true;
        }
      }
      if (failed) {
        status = MARLPDE_STATUS_STEP_TOO_SMALL;
        break;
      }
      // ------------------------------------------------------------------ accepted step
      n_equal_steps += 1;
      const double t_old = t;
      t = t_new;
      const bool adapt = n_equal_steps >= order + 1;
      // D[order+2] = d - D[order+1]; D[order+1] = d; D[i] += D[i+1] for i = order .. 0; and, when the order is
      // up for selection, the norms of error_const[order -+ 1] D[order], D[order+2] against scale(y_new)
      double ss_m = 0.0, ss_p = 0.0;
      {
        const double ecm = order > 1 ? kErrConst[order - 1] : 0.0, ecp = order < kMaxOrder ? kErrConst[order + 1] : 0.0;
        _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
          const double dv = d[idx];
          const double top = dv - D[(size_t)(order + 1) * n + idx];
          D[(size_t)(order + 2) * n + idx] = top;
          D[(size_t)(order + 1) * n + idx] = dv;
          double run = dv, d_order = 0.0;
#pragma unroll
          for (int i = kMaxOrder; i >= 0; --i)
            if (i <= order) {
              run += D[(size_t)i * n + idx];
              D[(size_t)i * n + idx] = run;
              if (i == order) d_order = run;
            }
          const double isc = 1.0 / fma(fabs(y[idx]), rtol, atol);
          const double am = ecm * d_order * isc, ap = ecp * top * isc;
          ss_m = fma(am, am, ss_m);
          ss_p = fma(ap, ap, ss_p);
        }
        __syncwarp();
      }
      if (adapt) {
        const double inf = (double)INFINITY;
        const double em = order > 1 ? sqrt(warp_sum(ss_m) / (double)n) : inf;
        const double ep = order < kMaxOrder ? sqrt(warp_sum(ss_p) / (double)n) : inf;
        // factors = error_norms ** (-1 / [order, order + 1, order + 2]); argmax takes the first maximum
        const double fm_ = pow(em, -1.0 / (double)order), f0 = pow(error_norm, -1.0 / (double)(order + 1)),
                     fp = pow(ep, -1.0 / (double)(order + 2));
        int delta = -1;
        double best = fm_;
        if (f0 > best) {
          best = f0;
          delta = 0;
        }
        if (fp > best) {
          best = fp;
          delta = 1;
        }
        order += delta;
        const double factor = fmin(kMaxFactor, safety * best);
        h_abs *= factor;
        change_D(ws, n, lane, D, order, factor);
        n_equal_steps = 0;
        lu_valid = false;
      }
      n_acc += 1;
      steps_done += 1;
      // ---- events (ivp.py: after every accepted step, located on the step's dense output)
      if (ev_on) {
        const unsigned ev_new = monitor_bits(kc, tb, N, lane, y);
        unsigned act = 0u;
        if (ev_new != ev_prev || (ev_new & kEqBitsMask) != 0u) act = active_events(event_classes(ev_prev), event_classes(ev_new));
        ev_prev = ev_new;
        if (act) locate_events(A, kc, tb, N, lane, col, act, t_old, t, h_abs, order, D, tmp);
      }
      // ---- t_eval samples in (t_old, t] (t_eval[0] == t0 belongs to the first step)
      while (next_eval < A.opt.n_eval) {
        const double te = A.g_t_eval[next_eval];
        if (!(te <= t)) break;
        double p[kMaxOrder];
        dense_coeffs(te, t, h_abs, order, p);
        double* snap = A.g_snap + ((size_t)col * A.opt.n_eval + next_eval) * n;
        _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32)
          snap[(idx % 5) * N + idx / 5] = dense_value(D, n, idx, order, p);
        ++next_eval;
      }
    }
    // the state at time t: y (= y_new of the last accepted step); after a failed step y holds a Newton iterate, the
    // last accepted state is then D[0] (change_D leaves row 0 alone)
    const double* const y_out = status == MARLPDE_STATUS_STEP_TOO_SMALL ? D : y;
#pragma unroll 1
    for (int idx = lane; idx < n; idx += 32) gy[(idx % 5) * N + idx / 5] = y_out[idx];
    if (lane == 0) {
      st.t = t;
      st.h_abs = h_abs;
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
    __syncwarp();
  }
}

}  // namespace bd

size_t bdf_workspace_bytes(int n_columns, int n_cells) {
  return sizeof(double) * bd::work_doubles(n_cells) * (size_t)n_columns;
}

#ifndef MARLPDE_HOST_EMU
cudaError_t launch_bdf(double* d_y, const marlpde_column_params* d_params, marlpde_column_state* d_state,
                       int n_columns, int n_cells, const marlpde_rk45_options& opt, const double* d_t_eval,
                       double* d_snap, int64_t* d_stats, int32_t* d_ev_counts, double* d_ev_times, double* d_work,
                       int32_t* d_queue, int sm_count, cudaStream_t stream) {
  imp::Args a;
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
  int ctas = (n_columns + imp::kWarpsPerCta - 1) / imp::kWarpsPerCta;
  const int max_ctas = sm_count * MARLPDE_BDF_MINBLOCKS;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  const bool vd = (opt.flags & MARLPDE_FLAG_VAR_DPHI) != 0, fd = (opt.flags & MARLPDE_FLAG_JAC_FD) != 0;
  const int threads = imp::kWarpsPerCta * 32;
  if (vd && fd) bd::bdf_kernel<true, true><<<ctas, threads, 0, stream>>>(a);
  else if (vd) bd::bdf_kernel<true, false><<<ctas, threads, 0, stream>>>(a);
  else if (fd) bd::bdf_kernel<false, true><<<ctas, threads, 0, stream>>>(a);
  else bd::bdf_kernel<false, false><<<ctas, threads, 0, stream>>>(a);
  return cudaGetLastError();
}
#endif

}  // namespace marlpde
