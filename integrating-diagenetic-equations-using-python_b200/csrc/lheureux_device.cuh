// lheureux_device.cuh — per-cell right-hand side of L'Heureux's five diagenetic equations,
// written for sm_100a fp64.  One call = one depth cell of one sediment column.
//
// Computes what the reference computes in LMAHeureuxPorosityDiff.pde_rhs
// (marlpde/LHeureux_model.py:413-520) after its 13 py-pde stencil operators (:372-384):
//   upwinded CA/CC advection by the sign of U (:418-423), Fiadeiro-Veronis weighted
//   gradients for cCa, cCO3, Phi (:433-469), thresholded dissolution/precipitation
//   rates coA, coC (:479-491) and the five rate formulas (:497-520).
//
// B200-first choices (not a transliteration):
//   * the three stencils per field are formed from (minus, centre, plus) register values;
//     ghost cells are synthesised by the edge threads, never stored;
//   * of each pair of real powers ((1-min(x,1))^a, (max(x,1)-1)^b) at most one is non-zero,
//     so one exp(e*log(.)) replaces two pow() calls, and it is skipped altogether where the
//     dissolution mask zeroes it;
//   * coth(Pe)-1/Pe is evaluated from one expm1 as (Pe(em+2)-em)/(Pe em): one division;
//   * 1/Phi, 1/den are formed once and reused (the reference divides 27 times);
//   * log/exp/expm1/reciprocals are the table-driven versions of fp64_math.cuh (about a third of
//     libdevice's fp64-pipe instructions, coefficients in constant memory instead of UMOV pairs).
// Rounding differs from the CPU path at the 1e-16 level (FMA contraction, reciprocal reuse);
// the parity gate is |delta| <= 1e-12 * term-magnitude (tests/test_rhs_parity.py).
#pragma once

#include <cmath>
#include <cstdint>

#include "../../include/marlpde_b200.h"
#include "fp64_math.cuh"

namespace marlpde {

// Derived per-column constants kept in shared memory by the kernels (broadcast reads).
struct ColumnConsts {
  double bc_top[5];
  double inv_dx, inv_dx2;
  double presum, rhorat, Da, lambda_, Da_lambda;
  double dCa, dCO3, delta, KRat;
  double nu1, nu2, m1, m2, n1, n2;
  double dPhi;
  double kPeCa, kPeCO3, kPePhi;   // delta_x/(2 dCa), delta_x/(2 dCO3), delta_x/(2 dPhi)
  double Pe_min, Pe_max;
  int32_t FV_switch, mask_lo, mask_hi, n_cells;
  // model variant MARLPDE_MODEL_VAR_DPHI (only looked at by kVarDPhi code and the generic path):
  double auxcon, half_dx;         // dPhi = auxcon F Phi^3 / (1 - Phi), Pe_Phi = W half_dx / dPhi
  int32_t var_dphi, pad_[3];      // != 0: this column uses the time-varying dPhi
};

__host__ __device__ inline void make_consts(const marlpde_column_params& p, int n_cells, ColumnConsts& c) {
  for (int f = 0; f < 5; ++f) c.bc_top[f] = p.bc_top[f];
  c.inv_dx = p.inv_dx;
  c.inv_dx2 = p.inv_dx2;
  c.presum = p.presum;
  c.rhorat = p.rhorat;
  c.Da = p.Da;
  c.lambda_ = p.lambda_;
  c.Da_lambda = p.Da * p.lambda_;
  c.dCa = p.dCa;
  c.dCO3 = p.dCO3;
  c.delta = p.delta;
  c.KRat = p.KRat;
  c.nu1 = p.nu1;
  c.nu2 = p.nu2;
  c.m1 = p.m1;
  c.m2 = p.m2;
  c.n1 = p.n1;
  c.n2 = p.n2;
  c.dPhi = p.dPhi_fixed;
  c.kPeCa = p.delta_x / (2.0 * p.dCa);
  c.kPeCO3 = p.delta_x / (2.0 * p.dCO3);
  c.kPePhi = p.delta_x / (2.0 * p.dPhi_fixed);
  c.Pe_min = p.Peclet_min;
  c.Pe_max = p.Peclet_max;
  c.auxcon = p.auxcon;
  c.half_dx = p.delta_x / 2.0;
  c.var_dphi = (p.model_flags & MARLPDE_MODEL_VAR_DPHI) != 0;
  c.pad_[0] = c.pad_[1] = c.pad_[2] = 0;
  c.FV_switch = p.FV_switch;
  c.mask_lo = p.mask_lo;
  c.mask_hi = p.mask_hi;
  c.n_cells = n_cells;
}

// x^e for x >= 0 (the clamps of LHeureux_model.py:480-484 guarantee it) and e > 0 (validated on
// the host): exp(e log x); log(0) = -inf -> exp(-inf) = 0 ; NaN propagates.
__device__ __forceinline__ double pow_nonneg(const fm::Tables& tb, double x, double e) {
  return fm::exp(tb, e * fm::log(tb, x));
}

// Fiadeiro-Veronis weight, LHeureux_model.py:437-442 (calculate_sigma :147-160).
__device__ __forceinline__ double fv_sigma(const fm::Tables& tb, double Pe, double W, double Pe_min, double Pe_max) {
  const double a = fabs(Pe);
  if (a < Pe_min) return 0.0;
  if (a > Pe_max) return (W > 0.0) ? 1.0 : ((W < 0.0) ? -1.0 : W);  // np.sign (keeps NaN / 0)
  if (!(a <= Pe_max)) return Pe;                                   // NaN Peclet -> NaN weight
  const double em = fm::expm1(tb, 2.0 * Pe);                         // coth = 1 + 2/em
  return fm::div(fma(Pe, em + 2.0, -em), Pe * em);
}

struct CellRates {
  double r[5];   // dCA/dt, dCC/dt, dcCa/dt, dcCO3/dt, dPhi/dt
  double U, W;   // advection velocities at this cell (feed the event monitors e5, e6)
};

// c = centre values, m = cell i-1 (or top ghost), p = cell i+1 (or bottom ghost), field order
// CA, CC, cCa, cCO3, Phi.  `in_mask` = not_too_deep*not_too_shallow != 0 for this cell.
__device__ __forceinline__ void cell_rhs(const ColumnConsts& k, const fm::Tables& tb, const double c[5],
                                         const double m[5], const double p[5], bool in_mask, CellRates& out) {
  const double CA = c[0], CC = c[1], cCa = c[2], cCO3 = c[3], Phi = c[4];

  // ---- porosity-dependent velocities (:414-431)
  const double rPhi = fm::rcp(Phi);
  const double F = 1.0 - fm::exp(tb, fma(-10.0, rPhi, 10.0));
  const double omP = 1.0 - Phi;
  const double Phi2 = Phi * Phi;
  const double FoP = fm::div(F, omP);
  const double U = fma(k.rhorat * (Phi2 * Phi), FoP, k.presum);
  const double W = fma(-k.rhorat * Phi2, F, k.presum);
  const double den = fma(-2.0, fm::log(tb, Phi), 1.0);
  const double rden = fm::rcp(den);
  // porosity diffusion coefficient: fixed (:431), or the commented-out per-cell value (:430) for flagged columns
  double dPhi = k.dPhi, kPePhi = k.kPePhi;
  if (k.var_dphi) {
    dPhi = k.auxcon * (Phi2 * Phi) * FoP;
    kPePhi = fm::div(k.half_dx, dPhi);
  }

  // ---- solids: first-order upwind by the sign of U (:418-423)
  const bool back = U > 0.0;
  const double gCA = (back ? (CA - m[0]) : (p[0] - CA)) * k.inv_dx;
  const double gCC = (back ? (CC - m[1]) : (p[1] - CC)) * k.inv_dx;

  // ---- Fiadeiro-Veronis weights (:433-462)
  double sCa = 0.0, sCO3 = 0.0, sPhi = 0.0;
  if (k.FV_switch) {
    const double Wden = W * den;
    sCa = fv_sigma(tb, Wden * k.kPeCa, W, k.Pe_min, k.Pe_max);
    sCO3 = fv_sigma(tb, Wden * k.kPeCO3, W, k.Pe_min, k.Pe_max);
    sPhi = fv_sigma(tb, W * kPePhi, W, k.Pe_min, k.Pe_max);
  }
  // g = 0.5((1-s) gf + (1+s) gb) = 0.5 (gf + gb) - 0.5 s (gf - gb)   (:464-469)
  const double hdx = 0.5 * k.inv_dx;
  const double gCa = ((1.0 - sCa) * (p[2] - cCa) + (1.0 + sCa) * (cCa - m[2])) * hdx;
  const double gCO3 = ((1.0 - sCO3) * (p[3] - cCO3) + (1.0 + sCO3) * (cCO3 - m[3])) * hdx;
  const double gPhi = ((1.0 - sPhi) * (p[4] - Phi) + (1.0 + sPhi) * (Phi - m[4])) * hdx;
  const double lapCa = (m[2] - 2.0 * cCa + p[2]) * k.inv_dx2;
  const double lapCO3 = (m[3] - 2.0 * cCO3 + p[3]) * k.inv_dx2;
  const double lapPhi = (m[4] - 2.0 * Phi + p[4]) * k.inv_dx2;

  // ---- tortuosity-corrected diffusion (:471-477)
  const double h1 = Phi * rden;
  const double h2 = gPhi * (2.0 + den) * (rden * rden);
  const double HCa = k.dCa * (h2 * gCa + h1 * lapCa);
  const double HCO3 = k.dCO3 * (h2 * gCO3 + h1 * lapCO3);

  // ---- reaction terms (:479-493): one real power per saturation product
  const double two = cCa * cCO3;
  const double three = two * k.KRat;
  double coA;
  if (three < 1.0) {
    coA = in_mask ? CA * pow_nonneg(tb, 1.0 - three, k.m2) : CA * 0.0;
  } else {
    coA = -CA * k.nu1 * pow_nonneg(tb, three - 1.0, k.m1);      // also carries NaN
  }
  double coC;
  if (two < 1.0) {
    coC = -CC * k.nu2 * pow_nonneg(tb, 1.0 - two, k.n2);
  } else {
    coC = CC * pow_nonneg(tb, two - 1.0, k.n1);
  }
  const double h3 = coA - k.lambda_ * coC;
  const double dWdx = -k.rhorat * gPhi * (2.0 * Phi * F + 10.0 * (F - 1.0));
  const double react = k.Da * omP * h3;

  out.r[0] = -U * gCA - k.Da * ((1.0 - CA) * coA + k.lambda_ * CA * coC);
  out.r[1] = -U * gCC + k.Da * (k.lambda_ * (1.0 - CC) * coC + CC * coA);
  out.r[2] = (HCa + react * (k.delta - cCa)) * rPhi - W * gCa;
  out.r[3] = (HCO3 + react * (k.delta - cCO3)) * rPhi - W * gCO3;
  out.r[4] = -(dWdx * Phi + W * gPhi) + dPhi * lapPhi + react;
  out.U = U;
  out.W = W;
}

// Load the (minus, centre, plus) triple for one cell from a field-major column in memory
// (global or shared), synthesising the py-pde ghost cells (LHeureux_model.py:26-30):
//   top, all fields:        value v        -> ghost = 2 v - a_0
//   bottom, CA and CC:      curvature 0    -> ghost = 2 a_{N-1} - a_{N-2}
//   bottom, cCa cCO3 Phi:   derivative 0   -> ghost = a_{N-1}
template <typename Load>
__device__ __forceinline__ void load_triple(const ColumnConsts& k, int cell, Load&& at, double c[5],
                                            double m[5], double p[5]) {
  const int n = k.n_cells;
  const int im = cell > 0 ? cell - 1 : 0;
  const int ip = cell < n - 1 ? cell + 1 : n - 1;
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    c[f] = at(f, cell);
    m[f] = at(f, im);
    p[f] = at(f, ip);
  }
  if (cell == 0) {
#pragma unroll
    for (int f = 0; f < 5; ++f) m[f] = 2.0 * k.bc_top[f] - c[f];
  }
  if (cell == n - 1) {
    p[0] = 2.0 * c[0] - m[0];
    p[1] = 2.0 * c[1] - m[1];
    p[2] = c[2];
    p[3] = c[3];
    p[4] = c[4];
  }
}


// =================================================================================================
// Two depth cells per thread (cells 2k and 2k+1 of one column) — the form both kernels use.
//
// c[f][q]  centre values of the two cells; mlo[f] = value of cell 2k-1 (or the top ghost),
// phi[f] = value of cell 2k+2 (or the bottom ghost).  Same formulas as cell_rhs, arranged so that
//   * every phase is straight-line code over q = 0, 1: the two cells' dependency chains are
//     independent, ptxas interleaves them (the fp64 pipe has ~8 cycles of latency to hide and
//     only ~3 resident warps per scheduler fit the register file);
//   * the rarely needed blocks (real power behind the dissolution mask, coth for a mid-range
//     Peclet number) are skipped by WARP-UNIFORM votes, so they cost nothing where no lane needs
//     them and never split a warp;
//   * log/exp arguments outside the table-driven range only raise a flag; the caller then
//     re-evaluates that cell once with cell_rhs (generic path, IEEE special values preserved).
// Must be called by all 32 lanes of a warp (full-mask votes).  Returns the per-cell flags.
// Gradients: with a = p - c, b = c - m:  gf + gb = (a + b)/dx, gf - gb = (a - b)/dx, so
//   g = ((a + b) - sigma (a - b)) / (2 dx)   and   laplace = (a - b) dx^-2.
// =================================================================================================
struct PairFlags {
  bool bad[2];
};

// How rhs_pair_own lays out its transcendental chains in basic blocks (the values are identical, only the
// instruction schedule differs; `kSched` template parameter of rhs_pair_own / rhs_pair):
//   kSchedSplit   one block per function behind its own warp vote (3 x coth, aragonite power, calcite power);
//   kSchedMerged  coth of the porosity Peclet number + calcite power unconditional in ONE block (4 chains in
//                 flight), coth of the two solute Peclet numbers behind one vote, aragonite power behind one;
//   kSchedAll     as kSchedMerged with the aragonite power unconditional too (6 chains, no vote);
//   kSchedTwoArm  two instances of the merged block, with (6 chains) and without (4 chains) the aragonite
//                 power, chosen by one vote;
//   kSchedLean    as kSchedTwoArm, and the weight selection itself is voted: the solute weights are 0 for the whole
//                 warp unless some lane's Peclet number reaches Pe_min; the porosity weight is the coth value for
//                 the whole warp when every lane is mid-range (saves ~50 compare / select instructions).
// Measured on B200 (r01g, 4096 columns): on-chip RK45 18.7 / 19.5 / 19.9 / 20.4 M column-steps/s; Radau to t = 0.05
// 1.61 / 1.58 / 1.55 / 1.60 s; overlapped tiles (N = 20 000 x 64) 224 / 220 / 211 / 218 k column-steps/s.  Each
// kernel instantiates the one that suits it.  -DMARLPDE_RHS_MERGE=n forces one schedule everywhere (A/B builds).
constexpr int kSchedSplit = 0, kSchedMerged = 1, kSchedAll = 2, kSchedTwoArm = 3, kSchedLean = 4;
__host__ __device__ constexpr int rhs_schedule(int preferred) {
#ifdef MARLPDE_RHS_MERGE
  return MARLPDE_RHS_MERGE;
#else
  return preferred;
#endif
}

// The three pieces of fv_sigma_pair for one value (same arithmetic, so the weights are bit-identical however the
// pieces are scheduled): range test, coth(Pe) - 1/Pe from one expm1, selection (np.sign keeps NaN / 0).
__device__ __forceinline__ bool fv_mid(bool fv_on, double Pe, double Pe_min, double Pe_max) {
  const double a = fabs(Pe);
  return fv_on && (a >= Pe_min) && (a <= Pe_max);
}
__device__ __forceinline__ double fv_langevin(const fm::Tables& tb, double Pe) {
  const double em = fm::expm1_nb(tb, 2.0 * Pe);                    // coth = 1 + 2/em
  return fma(Pe, em + 2.0, -em) * fm::rcp3(Pe * em);
}
__device__ __forceinline__ double fv_select(bool fv_on, double Pe, double W, double Pe_min, double Pe_max, bool mid,
                                            double sm) {
  const double a = fabs(Pe);
  const double sgn = (W > 0.0) ? 1.0 : ((W < 0.0) ? -1.0 : W);
  return (!fv_on || a < Pe_min) ? 0.0 : ((a > Pe_max) ? sgn : (mid ? sm : Pe));
}

__device__ __forceinline__ void fv_sigma_pair(const fm::Tables& tb, bool fv_on, const double (&Pe)[2],
                                              const double (&W)[2], double Pe_min, double Pe_max,
                                              double (&s)[2]) {
  bool mid[2];
  double a[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    a[q] = fabs(Pe[q]);
    mid[q] = fv_on && (a[q] >= Pe_min) && (a[q] <= Pe_max);
  }
  double sm[2] = {0.0, 0.0};
  if (__any_sync(0xffffffffu, mid[0] || mid[1])) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double em = fm::expm1_nb(tb, 2.0 * Pe[q]);             // coth = 1 + 2/em
      sm[q] = fma(Pe[q], em + 2.0, -em) * fm::rcp3(Pe[q] * em);
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double sgn = (W[q] > 0.0) ? 1.0 : ((W[q] < 0.0) ? -1.0 : W[q]);   // np.sign (keeps NaN / 0)
    s[q] = (!fv_on || a[q] < Pe_min) ? 0.0 : ((a[q] > Pe_max) ? sgn : (mid[q] ? sm[q] : Pe[q]));
  }
}

// Everything of the RHS that needs only the thread's OWN two cells (no neighbour): porosity functions,
// velocities, Fiadeiro-Veronis weights, reaction terms — about two thirds of the arithmetic.  The
// integrators evaluate this part while the block barrier that publishes the neighbours' values is
// still pending (split arrive / wait), then finish with the stencil part below.
struct OwnTerms {
  double U[2], W[2], rPhi[2];
  double sCa[2], sCO3[2], sPhi[2];
  double h1[2];      // Phi / den
  double h2c[2];     // (2 + den) / den^2            (h2 = h2c * grad Phi)
  double dWc[2];     // -rhorat (2 Phi F + 10 (F - 1))   (dW/dx = dWc * grad Phi)
  double react[2];   // Da (1 - Phi) (coA - lambda coC)
  double rA[2];      // Da ((1 - CA) coA + lambda CA coC)
  double rC[2];      // Da (lambda (1 - CC) coC + CC coA)
  double dPhi[2];    // porosity diffusion coefficient of the cell (kVarDPhi instantiations only)
};

// kVarDPhi: the instantiation for batches that may contain MARLPDE_MODEL_VAR_DPHI columns (LHeureux_model.py:222-223,
// :430-431); the per-column flag k.var_dphi then selects the per-cell or the fixed coefficient (bit-identical to the
// kVarDPhi = false instantiation for columns without the flag).  The default instantiation carries none of this.
template <int kSched, bool kVarDPhi = false>
__device__ __forceinline__ PairFlags rhs_pair_own(const ColumnConsts& k, const fm::Tables& tb, const double (&c)[5][2],
                                                  const bool (&in_mask)[2], OwnTerms& o) {
  PairFlags fl;
  fl.bad[0] = fl.bad[1] = false;

  // ---- porosity-dependent velocities (:414-431)
  double F[2], omP[2], den[2];
  double kPeV[2];   // delta_x / (2 dPhi) of the two cells (kVarDPhi only)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double Phi = c[4][q];
    o.rPhi[q] = fm::rcp3(Phi);
    F[q] = 1.0 - fm::exp_nb(tb, fma(-10.0, o.rPhi[q], 10.0), fl.bad[q]);
    omP[q] = 1.0 - Phi;
    const double Phi2 = Phi * Phi;
    const double FoP = F[q] * fm::rcp3(omP[q]);
    o.U[q] = fma(k.rhorat * (Phi2 * Phi), FoP, k.presum);
    if constexpr (kVarDPhi) {
      const double dv = k.auxcon * (Phi2 * Phi) * FoP;
      o.dPhi[q] = k.var_dphi ? dv : k.dPhi;
      kPeV[q] = k.var_dphi ? k.half_dx * fm::rcp3(dv) : k.kPePhi;
    }
    o.W[q] = fma(-k.rhorat * Phi2, F[q], k.presum);
    den[q] = fma(-2.0, fm::log_nb(tb, Phi, fl.bad[q]), 1.0);
    const double rden = fm::rcp3(den[q]);
    o.h1[q] = Phi * rden;
    o.h2c[q] = (2.0 + den[q]) * (rden * rden);
    o.dWc[q] = -k.rhorat * fma(2.0 * Phi, F[q], 10.0 * (F[q] - 1.0));
  }

  // ---- Fiadeiro-Veronis weights (:433-462) and the real powers of the rate laws (:479-491)
  // (FV_switch is a per-column value and warps straddle columns: it must not guard the votes, so it is
  //  folded into the lane predicates instead of branching here.)
  // Every transcendental below is a long dependent chain (expm1 + reciprocal, log + exp: ~300 cycles for ~70
  // instructions when a basic block holds only the two cells' chains); kSched decides which chains share a block.
  bool ltA[2], ltC[2], needA[2];
  double xA[2], xC[2], eA[2], eC[2], pA[2] = {0.0, 0.0}, pC[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double two = c[2][q] * c[3][q];
    const double three = two * k.KRat;
    ltA[q] = three < 1.0;
    ltC[q] = two < 1.0;
    xA[q] = fabs(1.0 - three);
    xC[q] = fabs(1.0 - two);
    eA[q] = ltA[q] ? k.m2 : k.m1;
    eC[q] = ltC[q] ? k.n2 : k.n1;
    needA[q] = !(ltA[q] && !in_mask[q]);
  }
  auto powA = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      bool bq = false;
      pA[q] = fm::exp_nb(tb, eA[q] * fm::log_nb(tb, xA[q], bq), bq);
      fl.bad[q] |= bq && needA[q];
    }
  };
  auto powC = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q) pC[q] = fm::exp_nb(tb, eC[q] * fm::log_nb(tb, xC[q], fl.bad[q]), fl.bad[q]);
  };
  {
    const bool fv_on = k.FV_switch != 0;
    double PeCa[2], PeCO3[2], PePhi[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double Wden = o.W[q] * den[q];
      PeCa[q] = Wden * k.kPeCa;
      PeCO3[q] = Wden * k.kPeCO3;
      if constexpr (kVarDPhi) PePhi[q] = o.W[q] * kPeV[q];
      else PePhi[q] = o.W[q] * k.kPePhi;
    }
    if constexpr (kSched == kSchedSplit) {
      fv_sigma_pair(tb, fv_on, PeCa, o.W, k.Pe_min, k.Pe_max, o.sCa);
      fv_sigma_pair(tb, fv_on, PeCO3, o.W, k.Pe_min, k.Pe_max, o.sCO3);
      fv_sigma_pair(tb, fv_on, PePhi, o.W, k.Pe_min, k.Pe_max, o.sPhi);
      if (__any_sync(0xffffffffu, needA[0] || needA[1])) powA();
      powC();
    } else {
      constexpr bool kLean = kSched == kSchedLean;
      bool midCa[2], midCO3[2], midPhi[2];
      double smCa[2] = {0.0, 0.0}, smCO3[2] = {0.0, 0.0}, smPhi[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) midPhi[q] = fv_mid(fv_on, PePhi[q], k.Pe_min, k.Pe_max);
      if constexpr (kLean) {
        // weight 0 <=> !fv_on || |Pe| < Pe_min (fv_select); anything else (mid-range, beyond Pe_max, NaN) is rare
        bool low[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          low[q] = !fv_on || (fabs(PeCa[q]) < k.Pe_min && fabs(PeCO3[q]) < k.Pe_min);
          o.sCa[q] = 0.0;
          o.sCO3[q] = 0.0;
        }
        if (__any_sync(0xffffffffu, !low[0] || !low[1])) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            o.sCa[q] = fv_select(fv_on, PeCa[q], o.W[q], k.Pe_min, k.Pe_max, fv_mid(fv_on, PeCa[q], k.Pe_min, k.Pe_max),
                                 fv_langevin(tb, PeCa[q]));
            o.sCO3[q] = fv_select(fv_on, PeCO3[q], o.W[q], k.Pe_min, k.Pe_max,
                                  fv_mid(fv_on, PeCO3[q], k.Pe_min, k.Pe_max), fv_langevin(tb, PeCO3[q]));
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          midCa[q] = fv_mid(fv_on, PeCa[q], k.Pe_min, k.Pe_max);
          midCO3[q] = fv_mid(fv_on, PeCO3[q], k.Pe_min, k.Pe_max);
        }
        if (__any_sync(0xffffffffu, midCa[0] || midCa[1] || midCO3[0] || midCO3[1])) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            smCa[q] = fv_langevin(tb, PeCa[q]);
            smCO3[q] = fv_langevin(tb, PeCO3[q]);
          }
        }
      }
      // coth of the porosity Peclet number: unconditional (the value is only selected where the number is
      // mid-range; expm1_nb on any other argument yields a discarded number, its table index is masked)
      auto langPhi = [&]() {
#pragma unroll
        for (int q = 0; q < 2; ++q) smPhi[q] = fv_langevin(tb, PePhi[q]);
      };
      if constexpr (kSched == kSchedTwoArm || kLean) {
        // The two arms are written in different orders on purpose: identical leading / trailing code would be
        // hoisted / sunk out of them and the aragonite chains would sit alone in their block again.
        if (__any_sync(0xffffffffu, needA[0] || needA[1])) {
          langPhi();
          powC();
          powA();
        } else {
          powC();
          langPhi();
        }
      } else {
        langPhi();
        if constexpr (kSched == kSchedAll) powA();
        powC();
      }
      if constexpr (kLean) {
        if (__all_sync(0xffffffffu, midPhi[0] && midPhi[1])) {
          o.sPhi[0] = smPhi[0];
          o.sPhi[1] = smPhi[1];
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q)
            o.sPhi[q] = fv_select(fv_on, PePhi[q], o.W[q], k.Pe_min, k.Pe_max, midPhi[q], smPhi[q]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          o.sCa[q] = fv_select(fv_on, PeCa[q], o.W[q], k.Pe_min, k.Pe_max, midCa[q], smCa[q]);
          o.sCO3[q] = fv_select(fv_on, PeCO3[q], o.W[q], k.Pe_min, k.Pe_max, midCO3[q], smCO3[q]);
          o.sPhi[q] = fv_select(fv_on, PePhi[q], o.W[q], k.Pe_min, k.Pe_max, midPhi[q], smPhi[q]);
        }
      }
      if constexpr (kSched == kSchedMerged) {
        if (__any_sync(0xffffffffu, needA[0] || needA[1])) powA();
      }
    }
  }

  // ---- reaction terms (:486-493, the Da(...) parts of :498-520)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double CA = c[0][q], CC = c[1][q];
    const double coA = ltA[q] ? (in_mask[q] ? CA * pA[q] : CA * 0.0) : -CA * k.nu1 * pA[q];
    const double coC = ltC[q] ? -CC * k.nu2 * pC[q] : CC * pC[q];
    const double h3 = fma(-k.lambda_, coC, coA);
    o.react[q] = k.Da * omP[q] * h3;
    o.rA[q] = k.Da * fma(1.0 - CA, coA, k.lambda_ * CA * coC);
    o.rC[q] = k.Da * fma(k.lambda_ * (1.0 - CC), coC, CC * coA);
  }
  return fl;
}

// The stencil part: first differences of the five fields, upwinded / Fiadeiro-Veronis weighted gradients,
// Laplacians and the five rates (:418-423, :464-477, :495-520).
template <bool kVarDPhi = false>
__device__ __forceinline__ void rhs_pair_finish(const ColumnConsts& k, const double (&c)[5][2], const double (&mlo)[5],
                                                const double (&phi)[5], const OwnTerms& o, double (&out)[5][2]) {
  // a = (next - centre), b = (centre - previous) per field and cell
  double a[5][2], b[5][2];
#pragma unroll
  for (int f = 0; f < 5; ++f) {
    b[f][0] = c[f][0] - mlo[f];
    a[f][0] = c[f][1] - c[f][0];
    b[f][1] = a[f][0];
    a[f][1] = phi[f] - c[f][1];
  }
  const double hdx = 0.5 * k.inv_dx;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double cCa = c[2][q], cCO3 = c[3][q], Phi = c[4][q];
    const bool back = o.U[q] > 0.0;
    const double gCA = (back ? b[0][q] : a[0][q]) * k.inv_dx;
    const double gCC = (back ? b[1][q] : a[1][q]) * k.inv_dx;
    const double d2Ca = a[2][q] - b[2][q], d2CO3 = a[3][q] - b[3][q], d2Phi = a[4][q] - b[4][q];
    const double gCa = fma(-o.sCa[q], d2Ca, a[2][q] + b[2][q]) * hdx;
    const double gCO3 = fma(-o.sCO3[q], d2CO3, a[3][q] + b[3][q]) * hdx;
    const double gPhi = fma(-o.sPhi[q], d2Phi, a[4][q] + b[4][q]) * hdx;
    const double lapCa = d2Ca * k.inv_dx2, lapCO3 = d2CO3 * k.inv_dx2, lapPhi = d2Phi * k.inv_dx2;
    const double h2 = gPhi * o.h2c[q];
    const double HCa = k.dCa * fma(h2, gCa, o.h1[q] * lapCa);
    const double HCO3 = k.dCO3 * fma(h2, gCO3, o.h1[q] * lapCO3);
    const double dWdx = o.dWc[q] * gPhi;
    out[0][q] = -o.U[q] * gCA - o.rA[q];
    out[1][q] = fma(-o.U[q], gCC, o.rC[q]);
    out[2][q] = fma(fma(o.react[q], k.delta - cCa, HCa), o.rPhi[q], -o.W[q] * gCa);
    out[3][q] = fma(fma(o.react[q], k.delta - cCO3, HCO3), o.rPhi[q], -o.W[q] * gCO3);
    out[4][q] = fma(kVarDPhi ? o.dPhi[q] : k.dPhi, lapPhi, o.react[q]) - fma(dWdx, Phi, o.W[q] * gPhi);
  }
}

template <int kSched, bool kVarDPhi = false>
__device__ __forceinline__ PairFlags rhs_pair(const ColumnConsts& k, const fm::Tables& tb, const double (&c)[5][2],
                                              const double (&mlo)[5], const double (&phi)[5],
                                              const bool (&in_mask)[2], double (&out)[5][2], double (&Uo)[2],
                                              double (&Wo)[2]) {
  OwnTerms o;
  const PairFlags fl = rhs_pair_own<kSched, kVarDPhi>(k, tb, c, in_mask, o);
  rhs_pair_finish<kVarDPhi>(k, c, mlo, phi, o, out);
  Uo[0] = o.U[0];
  Uo[1] = o.U[1];
  Wo[0] = o.W[0];
  Wo[1] = o.W[1];
  return fl;
}

// Generic-path re-evaluation of ONE cell of a pair (q = 0 or 1) after rhs_pair raised its flag:
// never inlined — it is rare and the hot loop must stay small.
static __device__ __noinline__ void rhs_pair_slow_cell(const ColumnConsts* k, const fm::Tables* tb, const double* c5,
                                                       const double* m5, const double* p5, int in_mask,
                                                       double* out7) {
  CellRates r;
  cell_rhs(*k, *tb, c5, m5, p5, in_mask != 0, r);
#pragma unroll
  for (int f = 0; f < 5; ++f) out7[f] = r.r[f];
  out7[5] = r.U;
  out7[6] = r.W;
}

__device__ __forceinline__ void rhs_pair_fixup(const ColumnConsts& k, const fm::Tables& tb, const PairFlags& fl,
                                               const double (&c)[5][2], const double (&mlo)[5],
                                               const double (&phi)[5], const bool (&in_mask)[2],
                                               double (&out)[5][2], double (&Uo)[2], double (&Wo)[2]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!fl.bad[q]) continue;
    double cc[5], mm[5], pp[5], o[7];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      cc[f] = c[f][q];
      mm[f] = q == 0 ? mlo[f] : c[f][0];
      pp[f] = q == 0 ? c[f][1] : phi[f];
    }
    rhs_pair_slow_cell(&k, &tb, cc, mm, pp, in_mask[q] ? 1 : 0, o);
#pragma unroll
    for (int f = 0; f < 5; ++f) out[f][q] = o[f];
    Uo[q] = o[5];
    Wo[q] = o[6];
  }
}

// Ghost values for the pair that touches a column end (LHeureux_model.py:26-30, py-pde rules):
//   top, all fields: value v -> 2 v - a_0 ;  bottom CA, CC: curvature 0 -> 2 a_{N-1} - a_{N-2} ;
//   bottom cCa, cCO3, Phi: derivative 0 -> a_{N-1}.
__device__ __forceinline__ double top_ghost(const ColumnConsts& k, int f, double a0) {
  return fma(2.0, k.bc_top[f], -a0);
}
__device__ __forceinline__ double bottom_ghost(int f, double a_last, double a_prev) {
  return f < 2 ? fma(2.0, a_last, -a_prev) : a_last;
}

}  // namespace marlpde
