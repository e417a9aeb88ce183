// implicit_common.cuh — building blocks shared by the batched IMPLICIT integrators (radau_batch.cu: 3-stage Radau IIA,
// bdf_batch.cu: variable-order BDF): one warp per sediment column, the RHS of a whole column evaluated by that warp
// (rhs_pair, the explicit kernels' code), the block-tridiagonal Jacobian (analytic off-diagonal 5x5 blocks, diagonal
// blocks by finite differences), its two-ended block-Thomas factorisation with one matrix entry per lane, the two-ended
// triangular solves on fp32 records, and the seven event monitors evaluated by one warp.
// Reference: the structure handed to SciPy's implicit methods (marlpde/parameters.py:150-199, :213-219) and what
// SciPy does with it (scipy/integrate/_ivp/common.py num_jac, radau.py / bdf.py LU + solve_lu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "brent.cuh"
#include "events.cuh"
#include "lheureux_device.cuh"

namespace marlpde {
namespace imp {

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

// ---- TEAMS.  kTeam = 1: one warp per column (the throughput shape: no barrier anywhere).  kTeam = 2: a CTA of two warps
// works on ONE column — the latency shape for the few longest columns of a sweep, whose sequential time bounds the whole
// sweep: the RHS passes, the element-wise passes and the Jacobian are split between the warps, the two
// chains of the two-ended factorisation run on one warp each, the solves stay on warp 0.  Every decision is taken from
// team-wide reductions, so both warps follow the same control flow; the team barrier is the block barrier.
template <int kTeam>
__device__ __forceinline__ void team_sync() {
  if (kTeam == 1) __syncwarp();
  else __syncthreads();
}
template <int kTeam>
__device__ __forceinline__ bool team_all(bool pred) {
  if (kTeam == 1) return __all_sync(0xffffffffu, pred);
  return __syncthreads_or(pred ? 0 : 1) == 0;
}

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


#ifndef MARLPDE_RADAU_RHS_STAGE
#define MARLPDE_RADAU_RHS_STAGE 1
#endif
constexpr int kStageDoubles = 66 * 5;       // cells 2 base - 1 .. 2 base + 64 of one pass of an RHS evaluation

struct __align__(16) WarpScratch {          // shared memory per warp
  ColumnConsts kc;
  union {                       // an RHS evaluation never overlaps a factorisation or a solve
  double stage[2][2][kStageDoubles + 6];   // RHS: double-buffered windows of yy and add for the pass in flight (cp.async)
  double cd[3][36];             // BDF change_D: R, U and R U of the difference-array rescaling (6 x 6 each)
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

// kernel arguments of both implicit integrators (stats: [n_columns][4] = njev, nlu, newton iterations, newton failures)
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
// (pass0, pstride): this warp evaluates the passes pass0, pass0 + pstride, ... of 32 cell pairs each (a team of two warps
// interleaves them; one warp alone: 0, 1).
template <bool VD, class Sink>
__device__ __forceinline__ void rhs_column(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
                                           const double* yy, const double* add, double (*stage)[2][kStageDoubles + 6],
                                           Sink&& sink, const int pass0 = 0, const int pstride = 1) {
  const int Hc = (N + 1) >> 1;
  const int b0 = 32 * pass0, bs = 32 * pstride;
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
  if (b0 < Hc) issue(b0, 0);
  int buf = 0;
#else
  auto ld = [&](int ff, int i) -> double { return add ? yy[i * 5 + ff] + add[i * 5 + ff] : yy[i * 5 + ff]; };
#endif
#pragma unroll 1
  for (int base = b0; base < Hc; base += bs) {
    const int p = base + lane;
    const int cell0 = 2 * p;
#if MARLPDE_RADAU_RHS_STAGE
    if (base + bs < Hc) {
      issue(base + bs, buf ^ 1);
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
template <bool VD, int kTeam = 1>
__device__ __noinline__ void rhs_eval(const ColumnConsts* kc, const fm::Tables* tb, int N, int lane, const double* yy,
                                      const double* add, double* out, double (*stage)[2][kStageDoubles + 6]) {
  auto sink = [&](int i, const double (&r5)[5]) {
#pragma unroll
    for (int f = 0; f < 5; ++f) out[i * 5 + f] = r5[f];
  };
  rhs_column<VD>(*kc, *tb, N, lane, yy, add, stage, sink, kTeam == 1 ? 0 : (int)(threadIdx.x >> 5), kTeam);
  team_sync<kTeam>();
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

// ANALYTIC Jacobian (the default; MARLPDE_FLAG_JAC_FD selects the finite-difference diagonal blocks above, see jacobian()):
// all three 5x5 blocks of every cell in ONE pass, one cell per lane, no RHS evaluation.  L_i and U_i as in jac_columns;
// D_i = d rhs_i / d y_i differentiates the cell's own dependence: the porosity functions F, U, W, 1 - 2 ln Phi and the
// time-varying dPhi, the Fiadeiro-Veronis weights through their Peclet numbers (the Langevin function's derivative
// 1/Pe^2 - 1/sinh^2 Pe in the mid range, 0 in the two clamped regimes), the centre weights of the stencils, the
// tortuosity factors, and the clamped saturation powers (m x^(m-1) inside the active regime).  Piecewise pieces — upwind
// direction, weight regimes, saturation clamps, dissolution mask — are differentiated inside their current regime, which
// is what num_jac's one-sided difference sees away from a switch.  The ghost cells fold into the own-cell block: top
// ghost 2 bc - y_0 (D_0 -= L_0), bottom ghost y_{N-1} for the solutes and the porosity (D += U) and 2 y_{N-1} - y_{N-2} for
// CA, CC (D += 2 U, L -= U).  Restated in numpy (oracle/jacobian_blocks.py) and checked there against central differences
// of the oracle RHS on evolved states of the reference's fixtures, all cells, both model variants
// (tests/test_host_side.py::test_analytic_jacobian_blocks): agreement to the accuracy of the differences.
// Generic (IEEE) maths as in cell_rhs: no restriction on the state.
// SWITCHING SURFACES.  The porosity column of the own-cell block is the only place the model's switches enter the Jacobian
// (the Peclet numbers and U depend on the cell's own Phi only).  A cell whose state sits ON a switch — |Pe| within a
// relative kSwitchTol of Pe_min or Pe_max, U within kSwitchTol of 0 — gets that column from the one-sided difference
// quotient num_jac would form (its step rule, sign(f) included; the ghost cells follow the perturbed cell): a trajectory
// that slides along a switching surface is held there by exactly this straddling difference (jacobian() below).
constexpr double kSwitchTol = 1e-5;

template <bool VD, int kTeam = 1>
__device__ __noinline__ void jac_analytic(const ColumnConsts* kp, const fm::Tables* tbp, int N, int lane, const double* y,
                                          double atol, double* J) {
  const ColumnConsts& k = *kp;
  const fm::Tables& tb = *tbp;
  const double hdx = 0.5 * k.inv_dx;
#pragma unroll 1
  for (int i = kTeam == 1 ? lane : (int)threadIdx.x; i < N; i += 32 * kTeam) {
    const bool first = i == 0, last = i == N - 1;
    double c[5], m[5], pl[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      c[f] = y[i * 5 + f];
      m[f] = first ? fma(2.0, k.bc_top[f], -c[f]) : y[(i - 1) * 5 + f];
      pl[f] = last ? c[f] : y[(i + 1) * 5 + f];
    }
    if (last) {
      pl[0] = 2.0 * c[0] - m[0];
      pl[1] = 2.0 * c[1] - m[1];
    }
    const double CA = c[0], CC = c[1], a = c[2], o = c[3], P = c[4];
    // ---- porosity functions and their derivatives
    const double rP = fm::rcp(P);
    const double E = fm::exp(tb, fma(-10.0, rP, 10.0));
    const double F = 1.0 - E, dF = -10.0 * E * rP * rP;
    const double omP = 1.0 - P, romP = fm::rcp(omP);
    const double FoP = F * romP, dFoP = fma(dF, omP, F) * romP * romP;
    const double P2 = P * P, P3 = P2 * P;
    const double U = fma(k.rhorat * P3, FoP, k.presum);
    const double dU = k.rhorat * fma(3.0 * P2, FoP, P3 * dFoP);
    const double W = fma(-k.rhorat * P2, F, k.presum);
    const double dW = -k.rhorat * fma(2.0 * P, F, -10.0 * E);
    const double ddWP = -k.rhorat * (2.0 * F * P - 20.0 * E - 100.0 * E * rP);      // P d2W/dPhi2
    const double den = fma(-2.0, fm::log(tb, P), 1.0), dden = -2.0 * rP;
    const double rden = fm::rcp(den);
    double dPhi = k.dPhi, ddPhi = 0.0, kPePhi = k.kPePhi, dkPePhi = 0.0;
    if (VD && k.var_dphi) {
      dPhi = k.auxcon * P3 * FoP;
      ddPhi = k.auxcon * fma(3.0 * P2, FoP, P3 * dFoP);
      kPePhi = fm::div(k.half_dx, dPhi);
      dkPePhi = -kPePhi * fm::div(ddPhi, dPhi);
    }
    // ---- Fiadeiro-Veronis weights and their derivatives with respect to Phi
    auto weight = [&](double Pe, double dPe, double& sg, double& dsg) {
      const double ab = fabs(Pe);
      sg = 0.0;
      dsg = 0.0;
      if (!k.FV_switch || ab < k.Pe_min) return;
      if (ab > k.Pe_max) {
        sg = (W > 0.0) ? 1.0 : ((W < 0.0) ? -1.0 : W);
        return;
      }
      if (!(ab <= k.Pe_max)) {
        sg = Pe;                                           // NaN Peclet number
        return;
      }
      const double em = fm::expm1(tb, 2.0 * Pe), rem = fm::rcp(em), rPe = fm::rcp(Pe);
      sg = fm::div(fma(Pe, em + 2.0, -em), Pe * em);
      dsg = fma(rPe, rPe, -4.0 * (em + 1.0) * rem * rem) * dPe;
    };
    const double dWden = fma(dW, den, W * dden);
    double s2, ds2, s3, ds3, s4, ds4;
    weight(W * den * k.kPeCa, k.kPeCa * dWden, s2, ds2);
    weight(W * den * k.kPeCO3, k.kPeCO3 * dWden, s3, ds3);
    weight(W * kPePhi, fma(dW, kPePhi, W * dkPePhi), s4, ds4);
    bool on_switch = fabs(U) <= kSwitchTol * fmax(1.0, fabs(k.presum));
    if (k.FV_switch) {
      const double pe[3] = {fabs(W * den * k.kPeCa), fabs(W * den * k.kPeCO3), fabs(W * kPePhi)};
#pragma unroll
      for (int q = 0; q < 3; ++q)
        on_switch = on_switch || fabs(pe[q] - k.Pe_min) <= kSwitchTol * k.Pe_min || fabs(pe[q] - k.Pe_max) <= kSwitchTol * k.Pe_max;
    }
    const double sg[3] = {s2, s3, s4}, dsg[3] = {ds2, ds3, ds4};
    double g[3], dg_own[3], dg_P[3], dg_m[3], dg_p[3], lap[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double fw = pl[q + 2] - c[q + 2], bw = c[q + 2] - m[q + 2];
      g[q] = ((1.0 - sg[q]) * fw + (1.0 + sg[q]) * bw) * hdx;
      dg_own[q] = 2.0 * sg[q] * hdx;
      dg_P[q] = (bw - fw) * hdx * dsg[q];
      dg_m[q] = -(1.0 + sg[q]) * hdx;
      dg_p[q] = (1.0 - sg[q]) * hdx;
      lap[q] = (fw - bw) * k.inv_dx2;
    }
    const double h1 = P * rden, dh1 = fma(2.0 * rden, rden, rden);
    const double h2c = (2.0 + den) * (rden * rden);
    const double dh2c = dden * rden * rden * fma(-2.0 * (2.0 + den), rden, 1.0);
    const double gP_P = dg_own[2] + dg_P[2];
    const double h2 = g[2] * h2c, dh2 = fma(gP_P, h2c, g[2] * dh2c);
    const double dif[2] = {k.dCa, k.dCO3};
    // ---- reaction terms: coA = CA A(three), coC = CC C(two)
    const double two = a * o, three = two * k.KRat;
    double A, dA, Cc, dC;
    if (three < 1.0) {
      const bool in_mask = i >= k.mask_lo && i < k.mask_hi;
      const double x = 1.0 - three;
      A = in_mask ? pow_nonneg(tb, x, k.m2) : 0.0;
      dA = (in_mask && x > 0.0) ? -k.m2 * pow_nonneg(tb, x, k.m2 - 1.0) : 0.0;
    } else {
      const double x = three - 1.0;
      A = -k.nu1 * pow_nonneg(tb, x, k.m1);
      dA = x > 0.0 ? -k.nu1 * k.m1 * pow_nonneg(tb, x, k.m1 - 1.0) : 0.0;
    }
    if (two < 1.0) {
      const double x = 1.0 - two;
      Cc = -k.nu2 * pow_nonneg(tb, x, k.n2);
      dC = x > 0.0 ? k.nu2 * k.n2 * pow_nonneg(tb, x, k.n2 - 1.0) : 0.0;
    } else {
      const double x = two - 1.0;
      Cc = pow_nonneg(tb, x, k.n1);
      dC = x > 0.0 ? k.n1 * pow_nonneg(tb, x, k.n1 - 1.0) : 0.0;
    }
    const double coA = CA * A, coC = CC * Cc;
    const double dcoA[5] = {A, 0.0, CA * dA * k.KRat * o, CA * dA * k.KRat * a, 0.0};
    const double dcoC[5] = {0.0, Cc, CC * dC * o, CC * dC * a, 0.0};
    const double h3 = fma(-k.lambda_, coC, coA);
    const double react = k.Da * omP * h3;
    double dreact[5];
#pragma unroll
    for (int f = 0; f < 4; ++f) dreact[f] = k.Da * omP * fma(-k.lambda_, dcoC[f], dcoA[f]);
    dreact[4] = -k.Da * h3;
    const bool back = U > 0.0;
    const double gA = (back ? (CA - m[0]) : (pl[0] - CA)) * k.inv_dx;
    const double gC = (back ? (CC - m[1]) : (pl[1] - CC)) * k.inv_dx;
    const double dgs_own = back ? k.inv_dx : -k.inv_dx;

    double D[5][5], Lb[5][5], Ub[5][5];                      // [row][column]
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int f = 0; f < 5; ++f) Lb[r][f] = Ub[r][f] = 0.0;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      D[0][f] = -k.Da * fma(1.0 - CA, dcoA[f], k.lambda_ * CA * dcoC[f]);
      D[1][f] = k.Da * fma(k.lambda_ * (1.0 - CC), dcoC[f], CC * dcoA[f]);
      D[2][f] = dreact[f] * (k.delta - a) * rP;
      D[3][f] = dreact[f] * (k.delta - o) * rP;
      D[4][f] = dreact[f];
    }
    D[0][0] += -U * dgs_own - k.Da * (k.lambda_ * coC - coA);
    D[0][4] += -dU * gA;
    D[1][1] += -U * dgs_own + k.Da * (coA - k.lambda_ * coC);
    D[1][4] += -dU * gC;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = q + 2;
      const double H = dif[q] * fma(h2, g[q], h1 * lap[q]);
      const double own = k.delta - c[r];
      D[r][r] += (dif[q] * fma(h2, dg_own[q], -2.0 * h1 * k.inv_dx2) - react) * rP - W * dg_own[q];
      D[r][4] += dif[q] * (dh2 * g[q] + h2 * dg_P[q] + dh1 * lap[q]) * rP - fma(react, own, H) * rP * rP - dW * g[q] -
                 W * dg_P[q];
    }
    const double t4 = fma(dW, P, W);
    D[4][4] += -gP_P * t4 - g[2] * fma(2.0, dW, ddWP) + ddPhi * lap[2] - 2.0 * dPhi * k.inv_dx2;
    {
      const double lA = back ? U * k.inv_dx : 0.0, uA = back ? 0.0 : -U * k.inv_dx;
      Lb[0][0] = Lb[1][1] = lA;
      Ub[0][0] = Ub[1][1] = uA;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int r = q + 2;
        Lb[r][r] = dif[q] * fma(h2, dg_m[q], h1 * k.inv_dx2) * rP - W * dg_m[q];
        Ub[r][r] = dif[q] * fma(h2, dg_p[q], h1 * k.inv_dx2) * rP - W * dg_p[q];
        Lb[r][4] = dif[q] * g[q] * h2c * dg_m[2] * rP;
        Ub[r][4] = dif[q] * g[q] * h2c * dg_p[2] * rP;
      }
      Lb[4][4] = fma(dPhi, k.inv_dx2, -t4 * dg_m[2]);
      Ub[4][4] = fma(dPhi, k.inv_dx2, -t4 * dg_p[2]);
    }
    // ---- ghost cells fold into the own-cell block
    if (first) {
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          D[r][f] -= Lb[r][f];
          Lb[r][f] = 0.0;
        }
    }
    if (last) {
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          D[r][f] += (f < 2 ? 2.0 : 1.0) * Ub[r][f];
          if (f < 2) Lb[r][f] -= Ub[r][f];
          Ub[r][f] = 0.0;
        }
    }
    if (on_switch) {                                         // rare: one-sided difference quotient for d / d Phi (see above)
      CellRates r0, r1;
      const bool in_mask = i >= k.mask_lo && i < k.mask_hi;
      cell_rhs(k, tb, c, m, pl, in_mask, r0);
      const double hP = fd_step(P, r0.r[4], atol);
      double c2[5], m2[5], p2[5];
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        c2[f] = c[f];
        m2[f] = m[f];
        p2[f] = pl[f];
      }
      c2[4] = P + hP;
      if (first) m2[4] = fma(2.0, k.bc_top[4], -c2[4]);
      if (last) p2[4] = c2[4];
      cell_rhs(k, tb, c2, m2, p2, in_mask, r1);
      const double inv = 1.0 / hP;
#pragma unroll
      for (int r = 0; r < 5; ++r) D[r][4] = (r1.r[r] - r0.r[r]) * inv;
    }
    double* blk = J + (size_t)i * 75;                        // blocks [L|D|U], each COLUMN-major
#pragma unroll
    for (int f = 0; f < 5; ++f)
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        blk[f * 5 + r] = Lb[r][f];
        blk[25 + f * 5 + r] = D[r][f];
        blk[50 + f * 5 + r] = Ub[r][f];
      }
  }
  team_sync<kTeam>();
}

// The Jacobian the integrators use.  DEFAULT (kJacFD = false): all blocks analytic in one pass (jac_analytic), with the
// switching-surface rule above.  MARLPDE_FLAG_JAC_FD (kJacFD = true): off-diagonal blocks analytic, diagonal blocks from 5
// finite-difference evaluations with num_jac's step rule — what SciPy's own Jacobian is.
// Measured on the 4096-column lattice, default base (r02s, r02v, r02w; sweeps to T* in predicted-cost order):
//                                          t = 0.05 (Radau / BDF)     to T*: Radau              BDF
//   finite-difference diagonal blocks        1.40 / 1.92 s              17.7 s, 4085 finish      22.3 s, 4072 finish
//   analytic, no switching-surface rule      1.27 / 1.84 s              38.4 s, 4045 finish      21.4 s, 4014 finish
//   analytic + one FD retry after a failure  1.71 / 2.23 s              22.1 s, 4085             27.3 s, 4073        (removed)
//   analytic + switching-surface rule        1.27 / 1.86 s              17.0 s, 4085 finish      21.9 s, 4071 finish
// The analytic blocks change nothing in the columns that finish either way (work ratio 1.00).  Without the rule 40 more
// columns stall late in the integration: the Fiadeiro-Veronis weight jumps from 0 to Pe/3 at |Pe| = Pe_min
// (LHeureux_model.py:437-442), W of a cell hovers around 0 while a 2-cell sawtooth in Phi develops, and the trajectory
// slides along that surface.  A one-sided difference straddles the jump and hands Newton a steep slope that holds the
// state on the surface; the in-regime derivative does not, and the step size collapses (SciPy Radau stalls on those
// columns with either Jacobian; the stall state found with SciPy sat at Pe_Phi = -0.0100000).  Taking just the porosity
// column of just the cells on a surface from the difference quotient restores exactly the finite-difference behaviour
// (the same 11 columns stay unfinished — they stop in SciPy too) at the analytic price.
template <bool kJacFD>
__host__ __device__ constexpr int jac_rhs_evals() { return kJacFD ? 5 : 0; }
template <bool VD, bool kJacFD, int kTeam = 1>
__device__ __forceinline__ void jacobian(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane, const double* y,
                                         const double* f, double atol, double* J, double* scratch,
                                         double (*stage)[2][kStageDoubles + 6]) {
  static_assert(kTeam == 1 || !kJacFD, "teams use the analytic Jacobian");
  if (kJacFD) fd_jacobian<VD>(kc, tb, N, lane, y, f, atol, J, scratch, stage);
  else jac_analytic<VD, kTeam>(&kc, &tb, N, lane, y, atol, J);
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
// kComplex = false (BDF: one real system) leaves all double2 work out; the record's complex part is then not written.
// (j0, j1, peer): the steps [j0, j1) of the schedule run here (default: all N).  A team of two warps runs the top chain
// [0, mid) on warp 0 and the bottom chain [mid, N-1) on warp 1 side by side, and after a team barrier warp 0 runs the
// meeting step [N-1, N) with its own X (top chain) and `peer`'s X (bottom chain).
template <bool kComplex>
__device__ __noinline__ void factorise(WarpScratch& ws, int N, int lane, const double M0, const double2 M1,
                                       const double* J, float* Rec, const int j0 = 0, int j1 = -1,
                                       const WarpScratch* peer = nullptr) {
  if (j1 < 0) j1 = N;
  const int l = lane < 25 ? lane : 0;     // lanes 25-31 shadow lane 0 (same values to the same shared words)
  const int r = l / 5, c = l - 5 * r;
  const bool diag = r == c;
  const int mid = N / 2;
  auto cell_of = [&](int j) { return j < mid ? j : (j < N - 1 ? N - 1 - (j - mid) : mid); };
  // ring of kSlots staged cells: steps 0 .. kDepth-1 are requested up front, step j+kDepth at iteration j
  auto request = [&](int j) {
    if (j < j1) {
      const double* Jn = J + (size_t)cell_of(j) * 75;
      double* dst = ws.jst[j % kSlots];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (lane + 32 * k < 75) cp_async8(dst + lane + 32 * k, Jn + lane + 32 * k);
    }
    cp_async_commit();                                       // (empty groups keep the group count uniform)
  };
  for (int c0 = 0; c0 < kDepth; ++c0) request(j0 + c0);
  if (!peer) {
    ws.x0[l] = 0.0;                                          // X_{-1} = 0
    if (kComplex) ws.x1[l] = make_double2(0.0, 0.0);
  }
  // meeting step: X of the top chain / of the bottom chain
  const double* const top0 = peer ? ws.x0 : ws.xs0;
  const double2* const top1 = peer ? ws.x1 : ws.xs1;
  const double* const bot0 = peer ? peer->x0 : ws.x0;
  const double2* const bot1 = peer ? peer->x1 : ws.x1;
#pragma unroll 1
  for (int j = j0; j < j1; ++j) {
    request(j + kDepth);                  // slot (j + kDepth) % kSlots was released at the end of iteration j-1
    cp_async_wait<kDepth>();              // all but the kDepth newest groups have landed: step j is in shared memory
    const int i = cell_of(j);
    const bool middle = j == N - 1, bottom = !middle && j >= mid;
    if (j == mid) {                       // the top chain is complete: keep its X for the meeting cell, restart from 0
      ws.xs0[l] = ws.x0[l];
      ws.x0[l] = 0.0;
      if (kComplex) {
        ws.xs1[l] = ws.x1[l];
        ws.x1[l] = make_double2(0.0, 0.0);
      }
    }
    __syncwarp();
    const double* Ji = ws.jst[j % kSlots];
    float* const rec = Rec + (size_t)i * 128;
    if (lane < 25) {                                         // fp32 copies of L_i and U_i for the sweeps
      rec[lane] = (float)Ji[lane];
      rec[(kComplex ? 102 : 52) + lane] = (float)Ji[50 + lane];
    }
    const int offP = bottom ? 50 : 0;     // block that couples to the PREVIOUS cell of the chain (L top-down, U bottom-up)
    const int offN = bottom ? 0 : 50;     // block that couples to the NEXT cell of the chain
    // entry (r, c) of S = M I - D_i - P_i X_prev (- U_i Y_{i+1} in the meeting cell)  and of the identity
    double A0 = (diag ? M0 : 0.0) - Ji[25 + c * 5 + r];
    double2 A1 = make_double2((diag ? M1.x : 0.0) - Ji[25 + c * 5 + r], diag ? M1.y : 0.0);
    {
      const double* xp0 = middle ? top0 : ws.x0;
      const double2* xp1 = middle ? top1 : ws.x1;
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const double nl = -Ji[offP + m * 5 + r];
        A0 = fma(nl, xp0[m * 5 + c], A0);
        if (kComplex) A1 = crfma(nl, xp1[m * 5 + c], A1);
      }
    }
    if (middle) {
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const double nu = -Ji[50 + m * 5 + r];
        A0 = fma(nu, bot0[m * 5 + c], A0);
        if (kComplex) A1 = crfma(nu, bot1[m * 5 + c], A1);
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
      if (kComplex) {
        ws.g1a[buf][l] = A1;
        ws.g1b[buf][l] = B1;
      }
      const int p0 = 7 - (int)(__reduce_max_sync(0xffffffffu, pivot_key(fabs(A0), r, c == k && !used0)) & 7u);
      int p1 = 0;
      if (kComplex)
        p1 = 7 - (int)(__reduce_max_sync(0xffffffffu, pivot_key(fabs(A1.x) + fabs(A1.y), r, c == k && !used1)) & 7u);
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
      if (kComplex) {   // complex system
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
    if (kComplex) ws.sinv1[step1 * 5 + c] = B1;
    if (lane < 25) {
      rec[26 + step0 * 5 + c] = (float)B0;
      if (kComplex) reinterpret_cast<float2*>(rec + 52)[step1 * 5 + c] = make_float2((float)B1.x, (float)B1.y);
    }
    __syncwarp();
    // ---- X_i = S_i^{-1} N_i for the next cell of the chain, while N_i is staged
    double X0 = 0.0;
    double2 X1 = make_double2(0.0, 0.0);
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const double u = Ji[offN + c * 5 + m];
      X0 = fma(u, ws.sinv0[r * 5 + m], X0);
      if (kComplex) X1 = crfma(u, ws.sinv1[r * 5 + m], X1);
    }
    ws.x0[l] = X0;
    if (kComplex) ws.x1[l] = X1;
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
// kCompact (BDF: one real system): records are [L | S^-1 | U] at floats 0 / 26 / 52 (factorise<false>), a sweep reads
// 26 eight-byte words of a cell's record — words [0,26) = {L, S} or [13,39) = {S, U} — instead of 51; system 0 only.
template <bool kCompact>
static __device__ __noinline__ void solve(WarpScratch& ws, int N, int lane, const float* Rec, double* b0,
                                   double* b1, double* b2, bool both) {
  constexpr int kWords = kCompact ? 26 : 51;                 // words of a record a sweep reads
  constexpr int kOffU = kCompact ? 52 : 102;                 // float offset of U_i in a record
  if (kCompact) both = false;
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
    const int oC = low_window ? 0 : kOffU - 26, oS0 = low_window ? 26 : 0, oS1 = low_window ? 52 : 26;   // float offsets
    auto request = [&](int j) {
      if (j < nch) {
        const double* src = reinterpret_cast<const double*>(Rec + (size_t)cell(j) * 128) + (low_window ? 0 : 13);
        double* dst = ws.mst[ch][j % kSlots];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (l16 + 16 * k < kWords) cp_async8(dst + l16 + 16 * k, src + l16 + 16 * k);
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
        for (int m = 0; m < 5; ++m) g = crfma((double)Rm[kOffU + m * 5 + r], ws.vec[buf][1][s][m], g);
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
static __device__ __noinline__ void monitors(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane, const double* y0,
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
static __device__ __noinline__ unsigned monitor_bits(const ColumnConsts& kc, const fm::Tables& tb, int N, int lane,
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

}  // namespace imp
}  // namespace marlpde
