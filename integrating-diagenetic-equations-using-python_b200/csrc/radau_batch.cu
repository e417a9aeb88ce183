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
#include <cuda_runtime.h>

#include <cstdint>

#include "brent.cuh"
#include "events.cuh"
#include "lheureux_device.cuh"
#include "radau_batch.cuh"

namespace marlpde {
namespace rd {

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
constexpr double kEps = 2.220446049250313e-16;
constexpr double kSqrtEps = 1.4901161193847656e-08;

constexpr int kWarpsPerCta = 4;
// resident CTAs per SM the kernel is compiled for: 3 x 4 warps = 12 columns in flight per SM at 168 registers per
// thread.  Measured on 4096 columns to t = 0.05: r01b 4 -> 3.07 s, 6 -> 3.12 s, 8 -> 3.42 s (more residency only adds
// spills); r01g 3 / 4 / 5 CTAs 1.52-1.56 / 1.55-1.61 / 1.69 s; r01h (two-ended sweeps) 2 / 3 / 4 CTAs 1.556 / 1.490 /
// 1.531 s.  Ring depth MARLPDE_RADAU_DEPTH 2 / 3 / 4: 1.57 / 1.55 / 1.62 s.
#ifndef MARLPDE_RADAU_MINBLOCKS
#define MARLPDE_RADAU_MINBLOCKS 3
#endif

// ---- complex helpers (double2 = re, im) -------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {   // a*b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 crfma(double a, double2 b, double2 c) {   // real a * b + c
  return make_double2(fma(a, b.x, c.x), fma(a, b.y, c.y));
}
__device__ __forceinline__ double2 cinv(double2 a) {   // pivots are finite and non-zero unless the state is not
  const double d = fm::rcp3(fma(a.x, a.x, a.y * a.y));
  return make_double2(a.x * d, -a.y * d);
}

// ---- asynchronous global -> shared copies (LDGSTS): the block-Thomas sweeps are sequential in the cell
// index, so the matrices of the next kDepth cells are kept in flight while the current cell is processed
#ifdef MARLPDE_HOST_EMU   // tests/emu/: kernel control logic on the host (test infrastructure only): copy at once
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  *reinterpret_cast<double*>(smem) = *reinterpret_cast<const double*>(gmem);
}
__device__ __forceinline__ void cp_async_commit() {}
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
#endif
// Tuning switches, all measured on 4096 columns (r01c, baseline 2.48 s to t = 0.05) and left OFF: with 16 columns
// per SM in flight the kernel is bound by DRAM traffic (1.6-1.8 TB/s of 1-kB runs), so more loads in flight lose:
//   RADAU_BATCH4 (four strides of loads per trip of the element-wise passes) 2.68 s,
//   (L1 prefetch of the next 32 cell pairs of an RHS evaluation 2.55 s — replaced in r02m by cp.async staging),
//   RADAU_PF_SOLVE (L1 prefetch of the right-hand side 8 cells ahead in the sweeps) 2.57 s.
#ifndef RADAU_PF_SOLVE
#define RADAU_PF_SOLVE 0
#endif
#ifndef RADAU_BATCH4
#define RADAU_BATCH4 0
#endif
#ifdef MARLPDE_HOST_EMU
__device__ __forceinline__ void prefetch_l1(const void*) {}
#else
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif
#ifndef MARLPDE_RADAU_DEPTH
#define MARLPDE_RADAU_DEPTH 3
#endif
constexpr int kDepth = MARLPDE_RADAU_DEPTH;   // cells in flight ahead of the one being processed
constexpr int kSlots = kDepth + 1;

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

#ifndef MARLPDE_RADAU_RHS_STAGE
#define MARLPDE_RADAU_RHS_STAGE 1
#endif
constexpr int kStageDoubles = 66 * 5;       // cells 2 base - 1 .. 2 base + 64 of one pass of an RHS evaluation

struct __align__(16) WarpScratch {          // shared memory per warp
  ColumnConsts kc;
  union {                       // an RHS evaluation never overlaps a factorisation or a solve
  double stage[2][2][kStageDoubles + 6];   // RHS: double-buffered windows of yy and add for the pass in flight (cp.async)
  struct {
  double2 vec[2][2][2][8];    // solve: two broadcast buffers x two chains x two systems x 5 entries (padded)
  // factorise (lane = 5 r + c holds entry (r, c) of [S | I] of both systems):
  double2 g0[2][32];          //   double-buffered exchange stage, real system: (S entry, I entry)
  double2 g1a[2][32];         //   complex system, S entries
  double2 g1b[2][32];         //   complex system, I entries
  double2 sinv1[26];          //   S_i^{-1} of the complex system (row major)
  double2 x1[26];             //   X_{i-1} = S_{i-1}^{-1} U_{i-1}, complex system (row major)
  double sinv0[26];           //   S_i^{-1} of the real system
  double x0[26];              //   X_{i-1}, real system
  double2 xs1[26];            //   X of the last cell of the top chain, kept for the meeting cell (complex system)
  double xs0[26];             //   same, real system
  union {                     // factorise and solve never overlap (each drains its cp.async groups before it returns)
    double jst[kSlots][80];      // factorise: staged Jacobian blocks [L|D|U] of the next cells of the schedule (ring)
    double mst[2][kSlots][52];   // solve: per chain, 51 eight-byte words of a cell's fp32 record (ring)
  };
  };
  };
};
static_assert(sizeof(WarpScratch) * kWarpsPerCta + fm::kTableBytes <= 48 * 1024, "static shared memory of the Radau kernel");

struct Args {
  double* g_y;
  const marlpde_column_params* g_params;
  marlpde_column_state* g_state;
  const double* g_t_eval;
  double* g_snap;
  int64_t* g_stats;     // [n_columns][4]: njev, nlu, newton iterations, newton failures
  int32_t* g_ev_counts; // [n_columns][7]              (MARLPDE_FLAG_EVENTS)
  double* g_ev_times;   // [n_columns][7][event_capacity]
  double* g_work;
  int32_t* g_queue;
  int n_columns, N;
  marlpde_rk45_options opt;
};

// One RHS evaluation of the whole column by one warp: state = yy (+ add, may be NULL), both cell-major
// [cell][field]; sink(i, r5) receives the five rates of cell i.  All 32 lanes run every iteration
// (rhs_pair votes), lanes without a pair work on benign values.
// (r02m) The inputs of a pass — the 40-byte runs of cells 2 base - 1 .. 2 base + 64 of yy and add — are staged in
// shared memory by cp.async one pass AHEAD: the ncu source page attributed 15 % of the kernel's stall samples to these
// loads.  Measured: 64 columns to t = 0.05 0.486 -> 0.467 s (a lone warp is latency bound); 4096 columns unchanged
// (1.42 s: with 1 776 columns in flight the kernel is bound by DRAM throughput, the stalls just move).
template <bool VD, class Sink>
__device__ __forceinline__ void rhs_column(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                           const double* yy, const double* add, double (*stage)[2][kStageDoubles + 6],
                                           Sink&& sink) {
  const int Hc = (N + 1) >> 1;
#if MARLPDE_RADAU_RHS_STAGE
  auto issue = [&](int base, int b) {          // window of the pass that starts at pair `base` -> stage[b]
    const int c_lo = base > 0 ? 2 * base - 1 : 0;
    const int c_hi = 2 * base + 65 < N ? 2 * base + 65 : N;
    const int d_lo = c_lo * 5, cnt = (c_hi - c_lo) * 5;
    for (int k = lane; k < cnt; k += 32) {
      cp_async8(&stage[b][0][k], yy + d_lo + k);
      if (add) cp_async8(&stage[b][1][k], add + d_lo + k);
    }
    cp_async_commit();
  };
  issue(0, 0);
  int buf = 0;
#else
  auto ld = [&](int ff, int i) -> double { return add ? yy[i * 5 + ff] + add[i * 5 + ff] : yy[i * 5 + ff]; };
#endif
#pragma unroll 1
  for (int base = 0; base < Hc; base += 32) {
    const int p = base + lane;
    const int cell0 = 2 * p;
#if MARLPDE_RADAU_RHS_STAGE
    if (base + 32 < Hc) {
      issue(base + 32, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();                                 // every lane's copies of this pass have landed
    const int c_lo = base > 0 ? 2 * base - 1 : 0;
    const double* const sy = stage[buf][0];
    const double* const sa = stage[buf][1];
    auto ld = [&](int ff, int i) -> double {
      const int k = (i - c_lo) * 5 + ff;
      return add ? sy[k] + sa[k] : sy[k];
    };
#endif
    const bool v0 = cell0 < N, v1 = cell0 + 1 < N;
    double c[5][2], mlo[5], phi[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      c[f][0] = v0 ? ld(f, cell0) : 0.5;
      c[f][1] = v1 ? ld(f, cell0 + 1) : 0.5;
      mlo[f] = (v0 && cell0 > 0) ? ld(f, cell0 - 1) : top_ghost(kc, f, c[f][0]);
      if (cell0 + 2 < N) {
        phi[f] = ld(f, cell0 + 2);
      } else if (v1) {
        phi[f] = bottom_ghost(f, c[f][1], c[f][0]);
      } else {
        c[f][1] = bottom_ghost(f, c[f][0], mlo[f]);
        phi[f] = c[f][1];
      }
    }
    const bool in_mask[2] = {cell0 >= kc.mask_lo && cell0 < kc.mask_hi,
                             cell0 + 1 >= kc.mask_lo && cell0 + 1 < kc.mask_hi};
    double r[5][2], U[2], Wv[2];
    PairFlags fl = rhs_pair<rhs_schedule(kSchedAll), VD>(kc, tb, c, mlo, phi, in_mask, r, U, Wv);
    fl.bad[0] = fl.bad[0] && v0;
    fl.bad[1] = fl.bad[1] && v1;
    if (fl.bad[0] || fl.bad[1]) rhs_pair_fixup(kc, tb, fl, c, mlo, phi, in_mask, r, U, Wv);
    if (v0) {
      const double r5[5] = {r[0][0], r[1][0], r[2][0], r[3][0], r[4][0]};
      sink(cell0, r5);
    }
    if (v1) {
      const double r5[5] = {r[0][1], r[1][1], r[2][1], r[3][1], r[4][1]};
      sink(cell0 + 1, r5);
    }
#if MARLPDE_RADAU_RHS_STAGE
    __syncwarp();                                 // this buffer is the target of the copies issued in the next trip
    buf ^= 1;
#endif
  }
}

// finite-difference step of num_jac (scipy/integrate/_ivp/common.py): h = (y + factor*y_scale) - y,
// y_scale = sign(f) * max(threshold, |y|), factor = sqrt(eps) (not adapted here), threshold = atol
__device__ __forceinline__ double fd_step(double y, double f, double atol) {
  const double ys = (f >= 0.0 ? 1.0 : -1.0) * fmax(atol, fabs(y));
  return (y + kSqrtEps * ys) - y;
}

// THE one instance of the RHS in this kernel (instruction-cache footprint matters: 12 warps per SM sit
// in different phases of their columns): out = rhs(yy + add) (add may be NULL), all cell-major [N][5].
// (VD: the kernel build for batches with MARLPDE_MODEL_VAR_DPHI columns.  This kernel competes for the instruction
//  cache — 12 warps per SM in different phases, stall_no_instruction 0.9 per issue — so the default build carries none
//  of the variant's code: always compiling it in measured 1.5 per issue.)
template <bool VD>
__device__ __noinline__ void rhs_eval(const ColumnConsts* kc, const fm::Tables* tb, int N, int lane, const double* yy,
                                      const double* add, double* out, double (*stage)[2][kStageDoubles + 6]) {
  auto sink = [&](int i, const double (&r5)[5]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) out[i * 5 + f] = r5[f];
  };
  rhs_column<VD>(*kc, *tb, N, lane, yy, add, stage, sink);
  __syncwarp();
}

// (Measured and dropped, r02a: the three stage evaluations of a Newton iteration fused with B = TI F - M W, one RHS
// instance in a rolled loop with the accumulators in local memory — 6 x 5N fewer doubles through DRAM per iteration, but
// 4096 columns to t = 0.05 took 1.645 s instead of 1.495 s; profiles/r02a_ab_candidates.log.)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column `fld` of all three Jacobian blocks of every cell, after ONE evaluation Fp = rhs(pert) with field
// fld of EVERY cell perturbed (pert = y0 + d, d_j = fd_step):
//   * L_i e_fld and U_i e_fld analytically.  Given a cell's own values the RHS is linear in each neighbour value
//     (the stencils of rhs_pair_finish; the coefficients U, W, sigma, h1, h2c, dWc depend on the cell's own Phi
//     only), so these are exact.  Structure: (CA,CA), (CC,CC) on the upwind side, (cCa,cCa),
//     (cCO3,cCO3), (Phi,Phi), (cCa,Phi), (cCO3,Phi) on both sides.  The bottom ghost of CA and CC is
//     2 a_{N-1} - a_{N-2}: the last cell's dependence on it folds into its L block.
//   * D_i e_fld = (Fp_i - f0_i - d_{i-1} L_i e_fld - d_{i+1} U_i e_fld) / d_i: with the coefficients taken at the
//     cell's PERTURBED own values this is exactly the one-cell difference quotient num_jac forms.
// So the Jacobian costs 5 RHS evaluations + 5 of these cheap passes instead of the 15 colours a block-
// tridiagonal pattern needs (the reference's 27-diagonal pattern: 21).  One cell per lane, generic (IEEE) math
// as in cell_rhs: no restriction on the state.  The formulas are checked against central differences of the
// oracle RHS in tests/test_host_side.py::test_offdiagonal_jacobian_block_formulas.
template <bool VD>
__device__ __noinline__ void jac_columns(const ColumnConsts* kp, const fm::Tables* tbp, int N, int lane, int fld,
                                         const double* pert, const double* Fp, const double* y0, const double* f0,
                                         double atol, double* J) {
  const ColumnConsts& k = *kp;
  const fm::Tables& tb = *tbp;
  const double hdx = 0.5 * k.inv_dx;
#pragma unroll 1
  for (int i = lane; i < N; i += 32) {
    const bool first = i == 0, last = i == N - 1;
    const double Phi = pert[i * 5 + 4];
    // ---- the cell's own coefficients (LHeureux_model.py:414-462, arithmetic of cell_rhs)
    const double rPhi = fm::rcp(Phi);
    const double F = 1.0 - fm::exp(tb, fma(-10.0, rPhi, 10.0));
    const double Phi2 = Phi * Phi;
    const double FoP = fm::div(F, 1.0 - Phi);
    const double U = fma(k.rhorat * (Phi2 * Phi), FoP, k.presum);
    const double W = fma(-k.rhorat * Phi2, F, k.presum);
    const double den = fma(-2.0, fm::log(tb, Phi), 1.0);
    const double rden = fm::rcp(den);
    double dPhi = k.dPhi, kPePhi = k.kPePhi;                // MARLPDE_MODEL_VAR_DPHI: the cell's own coefficient (cell_rhs)
    if (VD && k.var_dphi) {
      dPhi = k.auxcon * (Phi2 * Phi) * FoP;
      kPePhi = fm::div(k.half_dx, dPhi);
    }
    double Lc[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, Uc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    // weighted gradient of field f (2, 3 or 4) at this cell: 0.5 ((1 - s) forward + (1 + s) backward)
    auto grad = [&](int f, double sg) {
      const double ce = pert[i * 5 + f];
      const double pv = first ? fma(2.0, k.bc_top[f], -ce) : pert[(i - 1) * 5 + f];
      const double nx = last ? ce : pert[(i + 1) * 5 + f];
      return ((1.0 - sg) * (nx - ce) + (1.0 + sg) * (ce - pv)) * hdx;
    };
    auto sigma = [&](double Pe) { return k.FV_switch ? fv_sigma(tb, Pe, W, k.Pe_min, k.Pe_max) : 0.0; };
    if (fld < 2) {
      const bool back = U > 0.0;
      double lv = back ? U * k.inv_dx : 0.0;
      const double uv = back ? 0.0 : -U * k.inv_dx;
      if (last) lv -= uv;                                   // bottom ghost 2 a_{N-1} - a_{N-2}
      Lc[0] = fld == 0 ? lv : 0.0;
      Lc[1] = fld == 1 ? lv : 0.0;
      Uc[0] = fld == 0 ? uv : 0.0;
      Uc[1] = fld == 1 ? uv : 0.0;
    } else {
      const double sPhi = sigma(W * kPePhi);
      const double h2c = (2.0 + den) * (rden * rden);
      if (fld < 4) {
        const double sg = sigma(W * den * (fld == 2 ? k.kPeCa : k.kPeCO3)), d = fld == 2 ? k.dCa : k.dCO3;
        const double h1 = Phi * rden, h2 = grad(4, sPhi) * h2c;
        const double dgp = -(1.0 + sg) * hdx, dgn = (1.0 - sg) * hdx;
        const double lv = rPhi * d * fma(h2, dgp, h1 * k.inv_dx2) - W * dgp;
        const double uv = rPhi * d * fma(h2, dgn, h1 * k.inv_dx2) - W * dgn;
        Lc[2] = fld == 2 ? lv : 0.0;
        Lc[3] = fld == 3 ? lv : 0.0;
        Uc[2] = fld == 2 ? uv : 0.0;
        Uc[3] = fld == 3 ? uv : 0.0;
      } else {
        const double dgp = -(1.0 + sPhi) * hdx, dgn = (1.0 - sPhi) * hdx;
        const double Wden = W * den;
        const double t2 = rPhi * k.dCa * grad(2, sigma(Wden * k.kPeCa)) * h2c;
        const double t3 = rPhi * k.dCO3 * grad(3, sigma(Wden * k.kPeCO3)) * h2c;
        const double dWc = -k.rhorat * fma(2.0 * Phi, F, 10.0 * (F - 1.0));
        const double t4 = fma(dWc, Phi, W);
        Lc[2] = t2 * dgp;
        Uc[2] = t2 * dgn;
        Lc[3] = t3 * dgp;
        Uc[3] = t3 * dgn;
        Lc[4] = fma(dPhi, k.inv_dx2, -t4 * dgp);
        Uc[4] = fma(dPhi, k.inv_dx2, -t4 * dgn);
      }
    }
    const double di = fd_step(y0[i * 5 + fld], f0[i * 5 + fld], atol);
    double dp = 0.0, dn = 0.0;
    if (first) {
#pragma unroll
      for (int r = 0; r < 5; ++r) Lc[r] = 0.0;
    } else {
      dp = fd_step(y0[(i - 1) * 5 + fld], f0[(i - 1) * 5 + fld], atol);
    }
    if (last) {
#pragma unroll
      for (int r = 0; r < 5; ++r) Uc[r] = 0.0;
    } else {
      dn = fd_step(y0[(i + 1) * 5 + fld], f0[(i + 1) * 5 + fld], atol);
    }
    const double inv = 1.0 / di;
    double* blk = J + (size_t)i * 75 + fld * 5;             // blocks [L|D|U] are COLUMN-major: 40-byte runs
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      blk[r] = Lc[r];
      blk[25 + r] = fma(-dn, Uc[r], fma(-dp, Lc[r], Fp[i * 5 + r] - f0[i * 5 + r])) * inv;
      blk[50 + r] = Uc[r];
    }
  }
  __syncwarp();
}

// J = d rhs / d y: per field one RHS evaluation with that field perturbed in every cell, then jac_columns.
// `scratch` holds 2 x 5N doubles (perturbed state, its RHS).
template <bool VD>
__device__ __noinline__ void fd_jacobian(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane, const double* y,
                                         const double* f, double atol, double* J, double* scratch,
                                         double (*stage)[2][kStageDoubles + 6]) {
  const int n = 5 * N;
  double* const pert = scratch;
  double* const Fp = scratch + n;
#pragma unroll 1
  for (int idx = lane; idx < n; idx += 32) pert[idx] = y[idx];
  __syncwarp();
#pragma unroll 1
  for (int fld = 0; fld < 5; ++fld) {
#pragma unroll 1
    for (int i = lane; i < N; i += 32) {
      const double v = y[i * 5 + fld];
      pert[i * 5 + fld] = v + fd_step(v, f[i * 5 + fld], atol);
    }
    __syncwarp();
    rhs_eval<VD>(&kc, &tb, N, lane, pert, nullptr, Fp, stage);
    jac_columns<VD>(&kc, &tb, N, lane, fld, pert, Fp, y, f, atol, J);
#pragma unroll 1
    for (int i = lane; i < N; i += 32) pert[i * 5 + fld] = y[i * 5 + fld];
    __syncwarp();
  }
}

// Block-Thomas factorisation of (M I - J) for both systems at once.
//   lane = 5 r + c (25 lanes): ONE entry (r, c) of [S | I] of both systems per lane — the real system
//   (M = mu_real/h) in real arithmetic, the complex one (M = mu_complex/h) in double2.
// Per cell: S = M I - D_i - L_i X_{i-1}; Gauss-Jordan on [S | I] with partial pivoting (rows are never
// swapped: a lane remembers at which step its row was the pivot row); X_i = S^{-1} U_i for the next cell.
// Each elimination step exchanges the pivot row and the pivot column through a double-buffered shared
// stage (one __syncwarp per step) and finds the pivot with one warp REDUX per system on a packed
// (magnitude, row) key.  ~3x fewer instructions per cell than the column-per-lane version of r01d, whose
// pivot search and multipliers ran on 2 of 32 lanes (ncu r01e: 1480 warp instructions per cell, 36 % of the
// kernel's samples).
// The recurrence is sequential in the cell index, so memory latency is taken off its critical path:
// all 32 lanes fetch the Jacobian blocks of cell i+kDepth (75 doubles, coalesced) while cell i is being
// eliminated, and hand them over through a ring of shared-memory stages.
__device__ __forceinline__ unsigned pivot_key(double mag, int r, bool candidate) {
  // float magnitude in the high bits (non-negative floats order like their bit patterns), 7 - r in the low
  // three: the warp maximum is the largest magnitude, the smallest row among (float-)equal ones
  return candidate ? ((__float_as_uint((float)mag) & ~7u) | (unsigned)(7 - r)) : 0u;
}

// TWO-ENDED ("twisted") elimination: cells 0 .. mid-1 are eliminated top-down (S_i = M I - D_i - L_i X_{i-1},
// X_i = S_i^{-1} U_i), cells N-1 .. mid+1 bottom-up with the roles of L and U swapped (T_i = M I - D_i - U_i Y_{i+1},
// Y_i = T_i^{-1} L_i), and the two chains meet in cell mid = N/2: S_mid = M I - D_mid - L_mid X_{mid-1} - U_mid Y_{mid+1}.
// Same flops as the one-directional sweep; what it buys is that the SOLVES run both chains side by side in the two
// halves of the warp (half as many sequential steps, see solve()).  One loop over a schedule of N steps — top chain,
// bottom chain, meeting cell — keeps a single instance of the elimination code.
__device__ __noinline__ void factorise(WarpScratch& ws, int N, int lane, double h, const double* J, float* Rec) {
  const int l = lane < 25 ? lane : 0;     // lanes 25-31 shadow lane 0 (same values to the same shared words)
  const int r = l / 5, c = l - 5 * r;
  const bool diag = r == c;
  const double M0 = kMuReal / h;
  const double2 M1 = make_double2(kMuCRe / h, kMuCIm / h);
  const int mid = N / 2;
  auto cell_of = [&](int j) { return j < mid ? j : (j < N - 1 ? N - 1 - (j - mid) : mid); };
  // ring of kSlots staged cells: steps 0 .. kDepth-1 are requested up front, step j+kDepth at iteration j
  auto request = [&](int j) {
    if (j < N) {
      const double* Jn = J + (size_t)cell_of(j) * 75;
      double* dst = ws.jst[j % kSlots];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (lane + 32 * k < 75) cp_async8(dst + lane + 32 * k, Jn + lane + 32 * k);
    }
    cp_async_commit();                                       // (empty groups keep the group count uniform)
  };
  for (int c0 = 0; c0 < kDepth; ++c0) request(c0);
  ws.x0[l] = 0.0;                                            // X_{-1} = 0
  ws.x1[l] = make_double2(0.0, 0.0);
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    request(j + kDepth);                  // slot (j + kDepth) % kSlots was released at the end of iteration j-1
    cp_async_wait<kDepth>();              // all but the kDepth newest groups have landed: step j is in shared memory
    const int i = cell_of(j);
    const bool middle = j == N - 1, bottom = !middle && j >= mid;
    if (j == mid) {                       // the top chain is complete: keep its X for the meeting cell, restart from 0
      ws.xs0[l] = ws.x0[l];
      ws.xs1[l] = ws.x1[l];
      ws.x0[l] = 0.0;
      ws.x1[l] = make_double2(0.0, 0.0);
    }
    __syncwarp();
    const double* Ji = ws.jst[j % kSlots];
    float* const rec = Rec + (size_t)i * 128;
    if (lane < 25) {                                         // fp32 copies of L_i and U_i for the sweeps
      rec[lane] = (float)Ji[lane];
      rec[102 + lane] = (float)Ji[50 + lane];
    }
    const int offP = bottom ? 50 : 0;     // block that couples to the PREVIOUS cell of the chain (L top-down, U bottom-up)
    const int offN = bottom ? 0 : 50;     // block that couples to the NEXT cell of the chain
    // entry (r, c) of S = M I - D_i - P_i X_prev (- U_i Y_{i+1} in the meeting cell)  and of the identity
    double A0 = (diag ? M0 : 0.0) - Ji[25 + c * 5 + r];
    double2 A1 = make_double2((diag ? M1.x : 0.0) - Ji[25 + c * 5 + r], diag ? M1.y : 0.0);
    {
      const double* xp0 = middle ? ws.xs0 : ws.x0;
      const double2* xp1 = middle ? ws.xs1 : ws.x1;
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const double nl = -Ji[offP + m * 5 + r];
        A0 = fma(nl, xp0[m * 5 + c], A0);
        A1 = crfma(nl, xp1[m * 5 + c], A1);
      }
    }
    if (middle) {
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const double nu = -Ji[50 + m * 5 + r];
        A0 = fma(nu, ws.x0[m * 5 + c], A0);
        A1 = crfma(nu, ws.x1[m * 5 + c], A1);
      }
    }
    double B0 = diag ? 1.0 : 0.0;
    double2 B1 = make_double2(diag ? 1.0 : 0.0, 0.0);
    bool used0 = false, used1 = false;
    int step0 = 0, step1 = 0;                                // elimination step at which my row was the pivot row
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
      const int buf = k & 1;
      ws.g0[buf][l] = make_double2(A0, B0);
      ws.g1a[buf][l] = A1;
      ws.g1b[buf][l] = B1;
      const int p0 = 7 - (int)(__reduce_max_sync(0xffffffffu, pivot_key(fabs(A0), r, c == k && !used0)) & 7u);
      const int p1 =
          7 - (int)(__reduce_max_sync(0xffffffffu, pivot_key(fabs(A1.x) + fabs(A1.y), r, c == k && !used1)) & 7u);
      __syncwarp();
      {   // real system
        const double2 prow = ws.g0[buf][5 * p0 + c];         // (S, I) entries of the pivot row in my column
        const double ark = ws.g0[buf][5 * r + k].x;          // my row's entry in the pivot column
        const double piv = ws.g0[buf][5 * p0 + k].x;
        const double inv = piv * fm::rcp3(piv * piv);
        const bool mine = r == p0;
        const double m = mine ? inv : ark * inv;
        A0 = mine ? prow.x * m : fma(-m, prow.x, A0);
        B0 = mine ? prow.y * m : fma(-m, prow.y, B0);
        if (mine) {
          used0 = true;
          step0 = k;
        }
      }
      {   // complex system
        const double2 pa = ws.g1a[buf][5 * p1 + c], pb = ws.g1b[buf][5 * p1 + c];
        const double2 ark = ws.g1a[buf][5 * r + k];
        const double2 inv = cinv(ws.g1a[buf][5 * p1 + k]);
        const bool mine = r == p1;
        const double2 m = mine ? inv : cmul(ark, inv);
        const double2 nm = make_double2(-m.x, -m.y);
        const double2 a_s = cmul(pa, m), a_e = cfma(nm, pa, A1);
        const double2 b_s = cmul(pb, m), b_e = cfma(nm, pb, B1);
        A1 = mine ? a_s : a_e;
        B1 = mine ? b_s : b_e;
        if (mine) {
          used1 = true;
          step1 = k;
        }
      }
      // (the stage written at step k is rewritten at step k+2, behind the __syncwarp of step k+1)
    }
    // ---- the identity part now holds S^{-1}: row step_s of it sits in the lanes of row r
    ws.sinv0[step0 * 5 + c] = B0;
    ws.sinv1[step1 * 5 + c] = B1;
    if (lane < 25) {
      rec[26 + step0 * 5 + c] = (float)B0;
      reinterpret_cast<float2*>(rec + 52)[step1 * 5 + c] = make_float2((float)B1.x, (float)B1.y);
    }
    __syncwarp();
    // ---- X_i = S_i^{-1} N_i for the next cell of the chain, while N_i is staged
    double X0 = 0.0;
    double2 X1 = make_double2(0.0, 0.0);
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const double u = Ji[offN + c * 5 + m];
      X0 = fma(u, ws.sinv0[r * 5 + m], X0);
      X1 = crfma(u, ws.sinv1[r * 5 + m], X1);
    }
    ws.x0[l] = X0;
    ws.x1[l] = X1;
    __syncwarp();                         // everyone is done with slot j % kSlots before it is requested again
  }
  cp_async_wait<0>();
  __syncwarp();
}

// Solve (M I - J) x = b for both systems with the two-ended factors of factorise().  lane = 16 ch + 8 s + r (r < 5):
// chain ch (0: cells 0 .. mid-1 top-down, 1: cells N-1 .. mid+1 bottom-up), system s, row r.  b0: real right-hand
// side of system 0; (b1, b2): real and imaginary part of the right-hand side of system 1; all CELL-major [N][5],
// overwritten with the solution.  `both` = false solves system 0 only (error estimate).
//   inward :  top    p_i = S_i^{-1} (b_i + L_i p_{i-1}),   bottom q_i = T_i^{-1} (b_i + U_i q_{i+1})      (lock-step)
//   meeting:  x_mid  = S_mid^{-1} (b_mid + L_mid p_{mid-1} + U_mid q_{mid+1})
//   outward:  top    x_i = p_i + S_i^{-1} (U_i x_{i+1}),    bottom x_i = q_i + T_i^{-1} (L_i x_{i-1})     (lock-step)
// (Measured and dropped, r02n: outward sweep x_i = p_i + X_i x_next with X_i = S_i^{-1} N_i stored by the factorisation —
//  one dependent mat-vec and one warp barrier per cell instead of two, 304 instead of 408 record bytes per cell, but
//  832 instead of 512 bytes written per cell and factorisation: 1.42 vs 1.46 s at t = 0.05, 0.469 vs 0.469 s for 64
//  columns, 21.6 vs 21.2 s to T*: neutral, not worth 60 % more workspace.)
// Both chains of both systems run side by side in one warp, so a solve takes N/2 sequential steps per direction
// instead of N.  As in factorise(), the records of the next kDepth cells of each chain are in flight (cp.async)
// while the current cell is processed; a sweep reads 51 eight-byte words of a cell's record: words [0,51) =
// {L, S0, S1} (top inward, bottom outward) or words [13,64) = {S0, S1, U} (bottom inward, top outward).
__device__ __noinline__ void solve(WarpScratch& ws, int N, int lane, const float* Rec, double* b0,
                                   double* b1, double* b2, bool both) {
  const int ch = lane >> 4, s = (lane >> 3) & 1, r = lane & 7, l16 = lane & 15;
  const bool valid = r < 5 && (s == 0 || both);
  double* const bre = s == 0 ? b0 : b1;
  double* const bim = s == 0 ? nullptr : b2;
  const int mid = N / 2;
  const int nch = ch == 0 ? mid : N - 1 - mid;               // cells of my chain
  const int nmax = mid;                                      // (the top chain is never the shorter one)
  auto ldb = [&](int i) { return make_double2(bre[i * 5 + r], bim ? bim[i * 5 + r] : 0.0); };
  auto stb = [&](int i, double2 v) {
    bre[i * 5 + r] = v.x;
    if (bim) bim[i * 5 + r] = v.y;
  };
  // 5x5 block (fp32, row stride / column stride given) times the 5-vector published in vec[buf][chv][s][.]
  auto s_times = [&](const float* S0, const float* S1, int buf, int chv, double2 acc) {
    const float* S0r = S0 + r * 5;
    const float2* S1r = reinterpret_cast<const float2*>(S1) + r * 5;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const double2 sv = s == 0 ? make_double2((double)S0r[m], 0.0) : make_double2((double)S1r[m].x, (double)S1r[m].y);
      acc = cfma(sv, ws.vec[buf][chv][s][m], acc);
    }
    return acc;
  };
  int buf = 0;
  double2 v = make_double2(0.0, 0.0);     // inward: p / q of the chain's previous cell; outward: x of it

  // one lock-step pass over both chains; `outward` = false: cells 0.. and N-1.. towards the meeting cell
  auto sweep = [&](const bool outward) {
    auto cell = [&](int j) { return outward ? (ch == 0 ? mid - 1 - j : mid + 1 + j) : (ch == 0 ? j : N - 1 - j); };
    const bool low_window = (ch == 0) != outward;            // words [0,51): {L, S0, S1}; else [13,64): {S0, S1, U}
    const int oC = low_window ? 0 : 76, oS0 = low_window ? 26 : 0, oS1 = low_window ? 52 : 26;   // float offsets
    auto request = [&](int j) {
      if (j < nch) {
        const double* src = reinterpret_cast<const double*>(Rec + (size_t)cell(j) * 128) + (low_window ? 0 : 13);
        double* dst = ws.mst[ch][j % kSlots];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (l16 + 16 * k < 51) cp_async8(dst + l16 + 16 * k, src + l16 + 16 * k);
      }
      cp_async_commit();                                     // (empty groups keep the group count uniform)
    };
    for (int c0 = 0; c0 < kDepth; ++c0) request(c0);
    double2 bnext = (valid && nch > 0) ? ldb(cell(0)) : make_double2(0.0, 0.0);
#pragma unroll 1
    for (int j = 0; j < nmax; ++j) {
      request(j + kDepth);
      cp_async_wait<kDepth>();
      __syncwarp();
      const float* M = reinterpret_cast<const float*>(ws.mst[ch][j % kSlots]);
      const bool act = valid && j < nch;
      const double2 bi = bnext;
      if (valid && j + 1 < nch) bnext = ldb(cell(j + 1));
      if (valid) ws.vec[buf][ch][s][r] = v;
      __syncwarp();
      // coupling block times the previous cell's vector: inward it joins the right-hand side, outward it is what
      // S^{-1} is applied to
      double2 g = outward ? make_double2(0.0, 0.0) : bi;
      if (act && (outward || j > 0)) {
#pragma unroll
        for (int m = 0; m < 5; ++m) g = crfma((double)M[oC + m * 5 + r], ws.vec[buf][ch][s][m], g);
      }
      buf ^= 1;
      if (valid) ws.vec[buf][ch][s][r] = g;
      __syncwarp();
      if (act) {
        v = s_times(M + oS0, M + oS1, buf, ch, outward ? bi : make_double2(0.0, 0.0));
        stb(cell(j), v);
      }
      buf ^= 1;
      __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
  };

  sweep(false);
  // ---- meeting cell: g = b_mid + L_mid p_{mid-1} + U_mid q_{mid+1}, x_mid = S_mid^{-1} g.  Its record is read
  // straight from global memory (once per solve); both halves of the warp compute the same x_mid.
  {
    const float* Rm = Rec + (size_t)mid * 128;
    if (valid) ws.vec[buf][ch][s][r] = v;                    // p_{mid-1} (chain 0), q_{mid+1} (chain 1)
    __syncwarp();
    double2 g = make_double2(0.0, 0.0);
    if (valid) {
      g = ldb(mid);
      if (mid > 0) {
#pragma unroll
        for (int m = 0; m < 5; ++m) g = crfma((double)Rm[m * 5 + r], ws.vec[buf][0][s][m], g);
      }
      if (N - 1 - mid > 0) {
#pragma unroll
        for (int m = 0; m < 5; ++m) g = crfma((double)Rm[102 + m * 5 + r], ws.vec[buf][1][s][m], g);
      }
    }
    buf ^= 1;
    if (valid && ch == 0) ws.vec[buf][0][s][r] = g;
    __syncwarp();
    if (valid) {
      v = s_times(Rm + 26, Rm + 52, buf, 0, make_double2(0.0, 0.0));
      if (ch == 0) stb(mid, v);
    }
    buf ^= 1;
    __syncwarp();
  }
  sweep(true);
}

// The seven event monitors (LHeureux_model.py:524-593) of the state val(f, i), by one warp:
// g = {min y, min CA, min CC, max(CA+CC)-1, max Phi - 1, min U(Phi), max W(Phi)}.  NaNs propagate like
// np.amin / np.amax.  U and W use the arithmetic of rhs_pair.
// State = y0[idx] (q == NULL) or the dense output y0[idx] + x (q0 + x (q1 + x q2)) with q = Q [3][n].
__device__ __noinline__ void monitors(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane, const double* y0,
                                      const double* q, double x, double* g) {
  const int n = 5 * N;
  auto val = [&](int f, int i) -> double {
    const int idx = i * 5 + f;
    return q ? y0[idx] + x * (q[idx] + x * (q[n + idx] + x * q[2 * n + idx])) : y0[idx];
  };
  const double inf = (double)INFINITY;
  double m[7] = {inf, inf, inf, -inf, -inf, inf, -inf};
  bool nan5 = false, nanS = false, nanPhi = false, nanCA = false, nanCC = false;
#pragma unroll 1
  for (int i = lane; i < N; i += 32) {
    const double CA = val(0, i), CC = val(1, i), cCa = val(2, i), cCO3 = val(3, i), Phi = val(4, i);
    nanCA |= CA != CA;
    nanCC |= CC != CC;
    nanPhi |= Phi != Phi;
    nan5 |= (cCa != cCa) || (cCO3 != cCO3);
    const double F = 1.0 - fm::exp(tb, fma(-10.0, fm::rcp3(Phi), 10.0));
    const double Phi2 = Phi * Phi;
    const double U = fma(kc.rhorat * (Phi2 * Phi), F * fm::rcp3(1.0 - Phi), kc.presum);
    const double W = fma(-kc.rhorat * Phi2, F, kc.presum);
    m[0] = fmin(m[0], fmin(fmin(fmin(CA, CC), fmin(cCa, cCO3)), Phi));
    m[1] = fmin(m[1], CA);
    m[2] = fmin(m[2], CC);
    m[3] = fmax(m[3], CA + CC);
    m[4] = fmax(m[4], Phi);
    m[5] = fmin(m[5], U);
    m[6] = fmax(m[6], W);
    nanS |= (U != U) || (W != W);
  }
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const double other = __shfl_xor_sync(0xffffffffu, m[k], o);
      m[k] = (k == 3 || k == 4 || k == 6) ? fmax(m[k], other) : fmin(m[k], other);
    }
  }
  const unsigned bCA = __ballot_sync(0xffffffffu, nanCA), bCC = __ballot_sync(0xffffffffu, nanCC),
                 bPhi = __ballot_sync(0xffffffffu, nanPhi), b5 = __ballot_sync(0xffffffffu, nan5),
                 bS = __ballot_sync(0xffffffffu, nanS);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  g[0] = (bCA | bCC | bPhi | b5) ? qnan : m[0];
  g[1] = bCA ? qnan : m[1];
  g[2] = bCC ? qnan : m[2];
  g[3] = (bCA | bCC) ? qnan : m[3] - 1.0;
  g[4] = bPhi ? qnan : m[4] - 1.0;
  g[5] = (bPhi | bS) ? qnan : m[5];
  g[6] = (bPhi | bS) ? qnan : m[6];
}

// Detection only needs the SIGN of each monitor (ivp.py find_active_events): 21 predicate bits per cell, OR-reduced
// over the column — the same bookkeeping as the RK45 kernels (csrc/events.cuh).  This is what runs after every
// accepted step; monitors() above (five times the code) only runs inside Brent when a sign change has to be located.
__device__ __noinline__ unsigned monitor_bits(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                              const double* y) {
  unsigned b = 0u;
#pragma unroll 1
  for (int i = lane; i < N; i += 32) {
    double v[5][2], U[2], W[2];
#pragma unroll
    for (int f = 0; f < 5; ++f) v[f][0] = v[f][1] = y[i * 5 + f];
    const double Phi = v[4][0];
    const double F = 1.0 - fm::exp(tb, fma(-10.0, fm::rcp3(Phi), 10.0));
    const double Phi2 = Phi * Phi;
    U[0] = U[1] = fma(kc.rhorat * (Phi2 * Phi), F * fm::rcp3(1.0 - Phi), kc.presum);
    W[0] = W[1] = fma(-kc.rhorat * Phi2, F, kc.presum);
    b |= event_bits(v, U, W, false);
  }
  return __reduce_or_sync(0xffffffffu, b);
}

// brentq on the dense output between t_old and t for every monitor in `act` (ivp.py handle_events ->
// solve_event_equation, xtol = rtol = 4 eps).  Rare, so it lives outside the step loop's instruction footprint.
__device__ __noinline__ void locate_events(const Args& A, const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                           int col, unsigned act, double t_old, double t, double h_old,
                                           const double* yold, const double* Q) {
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
    if (lane == 0) {
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

template <bool VD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MARLPDE_RADAU_MINBLOCKS) radau_kernel(const Args A) {
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
    for (int idx = lane; idx < n; idx += 32) y[idx] = gy[(idx % 5) * N + idx / 5];
    __syncwarp();
    const ColumnConsts& kc = ws.kc;
    marlpde_column_state st = A.g_state[col];
    double t = st.t, h_attr = st.h_abs;
    long long n_acc = st.n_accepted, n_rej = st.n_rejected, nfev = st.nfev;
    long long njev = 0, nlu = 0, n_newton = 0, n_newton_fail = 0;
    int next_eval = st.next_eval;
    int status = MARLPDE_STATUS_FINISHED;
    long long steps_done = 0;

    auto eval_to = [&](const double* yy, const double* add, double* out) {   // out = rhs(yy + add)
      rhs_eval<VD>(&kc, &tb, N, lane, yy, add, out, ws.stage);
    };

    if (t < A.opt.t_bound) {
      eval_to(y, nullptr, w.f);
      nfev += 1;
      fd_jacobian<VD>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
      njev += 1;
      nfev += 5;   // Jacobian: one evaluation per field (fd_jacobian)
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
            factorise(ws, N, lane, h, w.J, w.Rec);
            nlu += 2;
            lu_valid = true;
          }
          // ---- Z0 from the previous step's dense output (radau.py: Z0 = sol(t + h C).T - y), W = TI Z0
          _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
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
          __syncwarp();
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
            for (int base = lane; base < n; base += (RADAU_BATCH4 ? 128 : 32)) {
              double F0[4], F1[4], F2[4], W0[4], W1[4], W2[4];
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + 32 * u;
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
                const int idx = base + 32 * u;
                if (idx >= n) continue;
                finite = finite && isfinite(F0[u]) && isfinite(F1[u]) && isfinite(F2[u]);
                // f_real = F^T TI_REAL - M_real W0 ; f_complex = F^T TI_COMPLEX - M_complex (W1 + i W2)
                w.B[idx] = kTI[0][0] * F0[u] + kTI[0][1] * F1[u] + kTI[0][2] * F2[u] - Mr * W0[u];
                w.B[n + idx] = kTI[1][0] * F0[u] + kTI[1][1] * F1[u] + kTI[1][2] * F2[u] - (Mcr * W1[u] - Mci * W2[u]);
                w.B[2 * n + idx] = kTI[2][0] * F0[u] + kTI[2][1] * F1[u] + kTI[2][2] * F2[u] - (Mcr * W2[u] + Mci * W1[u]);
              }
            }
            __syncwarp();
            if (!__all_sync(0xffffffffu, finite)) break;
            solve(ws, N, lane, w.Rec, w.B, w.B + n, w.B + 2 * n, true);
            // norm(dW / scale) and, in the same pass, W += dW, Z = T W.  (radau.py leaves W and Z untouched when
            // the rate test below breaks; they are dead then — every continuation re-initialises them from Z0.)
            double ss = 0.0;
#pragma unroll 1
            for (int base = lane; base < n; base += (RADAU_BATCH4 ? 128 : 32)) {
              double d0[4], d1[4], d2[4], yy4[4], W0[4], W1[4], W2[4];
#pragma unroll
              for (int u = 0; u < (RADAU_BATCH4 ? 4 : 1); ++u) {
                const int idx = base + 32 * u;
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
                const int idx = base + 32 * u;
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
            __syncwarp();
            const double dW_norm = sqrt(warp_sum(ss) / (double)(3 * n));
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
          fd_jacobian<VD>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
          njev += 1;
          nfev += 5;   // Jacobian: one evaluation per field (fd_jacobian)
          current_jac = true;
          lu_valid = false;
        }
        if (!converged) {
          h_abs *= 0.5;
          lu_valid = false;
          continue;
        }
        // ---- error estimate: error = LU_real.solve(f + Z^T E / h)
        _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
          const double ze = (w.Z[idx] * kE[0] + w.Z[n + idx] * kE[1] + w.Z[2 * n + idx] * kE[2]) / h;
          w.tmp[idx] = ze;
          w.err[idx] = w.f[idx] + ze;
        }
        __syncwarp();
        solve(ws, N, lane, w.Rec, w.err, nullptr, nullptr, false);
        auto err_norm_of = [&]() {
          double ss = 0.0;
          _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
            const double yn = y[idx] + w.Z[2 * n + idx];
            const double sc = fma(fmax(fabs(y[idx]), fabs(yn)), rtol, atol);
            const double a = w.err[idx] / sc;
            ss += a * a;
          }
          return sqrt(warp_sum(ss) / (double)n);
        };
        error_norm = err_norm_of();
        safety = 0.9 * (2 * kNewtonMaxIter + 1) / (double)(2 * kNewtonMaxIter + n_iter);
        if (rejected && error_norm > 1.0) {
          // error = LU_real.solve(fun(t, y + error) + ZE)
          eval_to(y, w.err, w.B);
          nfev += 1;
          _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) w.err[idx] = w.B[idx] + w.tmp[idx];
          __syncwarp();
          solve(ws, N, lane, w.Rec, w.err, nullptr, nullptr, false);
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
      _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32) {
        const double z0 = w.Z[idx], z1 = w.Z[n + idx], z2 = w.Z[2 * n + idx];
        const double yo = y[idx];
        w.yold[idx] = yo;
        y[idx] = yo + z2;
#pragma unroll
        for (int j = 0; j < 3; ++j) w.Q[j * n + idx] = z0 * kP[0][j] + z1 * kP[1][j] + z2 * kP[2][j];
      }
      __syncwarp();
      eval_to(y, nullptr, w.f);
      nfev += 1;
      if (recompute_jac) {
        fd_jacobian<VD>(kc, tb, N, lane, y, w.f, atol, w.J, w.B, ws.stage);
        njev += 1;
        nfev += 5;   // Jacobian: one evaluation per field (fd_jacobian)
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
        if (act) locate_events(A, kc, tb, N, lane, col, act, t_old, t, h_old, w.yold, w.Q);   // rare: kept out of line
      }
      // ---- t_eval samples in (t_old, t] (t_eval[0] == t0 belongs to the first step): y_old + Q p(x)
      while (next_eval < A.opt.n_eval) {
        const double te = A.g_t_eval[next_eval];
        if (!(te <= t)) break;
        const double xx = (te - t_old) / h_old;
        double* snap = A.g_snap + ((size_t)col * A.opt.n_eval + next_eval) * n;
        _Pragma("unroll 1") for (int idx = lane; idx < n; idx += 32)
          snap[(idx % 5) * N + idx / 5] = w.yold[idx] + xx * (w.Q[idx] + xx * (w.Q[n + idx] + xx * w.Q[2 * n + idx]));
        ++next_eval;
      }
    }
#pragma unroll 1
    for (int idx = lane; idx < n; idx += 32) gy[(idx % 5) * N + idx / 5] = y[idx];
    if (lane == 0) {
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
    __syncwarp();
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
  int ctas = (n_columns + rd::kWarpsPerCta - 1) / rd::kWarpsPerCta;
  const int max_ctas = sm_count * MARLPDE_RADAU_MINBLOCKS;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  if (opt.flags & MARLPDE_FLAG_VAR_DPHI) rd::radau_kernel<true><<<ctas, rd::kWarpsPerCta * 32, 0, stream>>>(a);
  else rd::radau_kernel<false><<<ctas, rd::kWarpsPerCta * 32, 0, stream>>>(a);
  return cudaGetLastError();
}

#endif

}  // namespace marlpde
