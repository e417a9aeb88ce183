// fp64_math.cuh — table-driven fp64 log / exp / expm1 and Newton reciprocals for sm_100a.
//
// Why not libdevice: the RHS needs 3 log, 3-4 exp, 1-3 coth and ~6 divisions per cell per
// evaluation; CUDA's generic implementations cost 31 / 18 / 45 / 13 fp64-pipe instructions each
// and materialise every polynomial coefficient with two UMOVs (ncu, profiles/): the kernel was
// issue-bound and its 200 KB of SASS thrashed the instruction cache.  These versions use
// 128- and 64-entry tables (staged in shared memory by the caller) and short polynomials whose
// coefficients sit in constant memory:  log 12, exp 10, expm1 13, reciprocal 4 fp64 instructions.
// Accuracy (tests/test_gpu_math.py): log  |err| <= 2e-16 * max(1, |log x|)   (absolute, not ulp:
// it feeds 1-2ln(Phi), m*log(x) and the step controller, never a result near zero),
// exp <= 3e-16 relative, expm1 <= 4e-16 relative for |x| >= 1e-2, rcp/div <= 2.3e-16 relative.
// Arguments outside the fast range (zero, negative, denormal, inf, NaN, |x| >= 690 for exp)
// branch to libdevice, so IEEE special values are preserved.
#pragma once

#include <cmath>
#include <cstdint>

#ifndef MARLPDE_EXP_INTCHECK
#define MARLPDE_EXP_INTCHECK 0
#endif

namespace marlpde {
namespace fm {

#include "fp64_tables.inc"

// polynomial coefficients (constant bank -> uniform registers, two per LDCU.128)
__constant__ double kLog1pC[6] = {-0.5, 1.0 / 3.0, -0.25, 0.2, -1.0 / 6.0, 1.0 / 7.0};
__constant__ double kExpC[5] = {0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0};

// (Measured and dropped, r02a: constants with a zero low word as instruction immediates — 53 of 160 LDC gone from the RK45
// kernel, no change in its run time: 22.53 vs 22.58 M column-steps/s; profiles/r02a_ab_candidates.log.)
#define FM_LN2HI kLn2Hi
#define FM_LN2LO kLn2Lo
#define FM_L64HI kLn2_64Hi
#define FM_L64LO kLn2_64Lo
#define FM_K64 k64_Ln2
#define FM_LOG(i) kLog1pC[i]
#define FM_EXP(i) kExpC[i]

// Shared-memory copies of the tables (random-index reads would serialise in the constant cache).
// A look-up is a 16-byte read at a data-dependent row, and the rows of the 32 lanes are effectively random (the index
// comes from mantissa bits): a warp-wide look-up takes 6-8 shared-memory wavefronts instead of the 4 its 512 bytes need.
// The ncu source page of the RK45 kernel attributes ALL of its excessive wavefronts (695 M of 4.6 G per launch, the
// `l1tex__data_bank_conflicts_pipe_lsu_mem_shared` counter of ~0.9 G) to these two tables: +18 % shared-memory
// wavefronts on an LSU pipe that is ~25 % busy.
// MARLPDE_TABLE_REP=8 (per translation unit) removes them: eight interleaved copies, row j of copy c at 128 j + 16 c
// bytes, lane l reads copy l & 7, so the eight lanes of a quarter warp always hit eight different 16-byte bank groups
// (24 kB instead of 3 kB per CTA).  Measured on the B200 (r02k): the conflict counter halves (0.90 -> 0.46 G; the 8-byte
// reads of the exp table still pair up), but the per-lane copy offset is one more live value in a kernel that sits on
// the 168-register cliff — ptxas spills more in the stage loop (long-scoreboard stalls 0.22 -> 0.57 per issue) and the
// on-chip RK45 kernel gets 5.6 % SLOWER (25.1 -> 23.7 M column-steps/s); the tile kernel +1 % (noise).  So the default
// stays 1: the conflicts are explained, and cheaper than the registers their removal costs.
#ifndef MARLPDE_TABLE_REP
#define MARLPDE_TABLE_REP 1
#endif
constexpr int kTableRep = MARLPDE_TABLE_REP;
static_assert(kTableRep == 1 || kTableRep == 8, "MARLPDE_TABLE_REP must be 1 or 8");
struct Tables {
  const double2* logtab;   // row j at logtab[j * kTableRep]: [128] {ic, L}      (already offset to the lane's copy)
  const double2* exptab;   // row j at exptab[j * kTableRep]: [64]  {Thi, Tlo}
};
constexpr int kTableBytes = (128 + 64) * 16 * kTableRep;

// cooperative copy constant -> shared; call with all threads of the CTA, then __syncthreads()
__device__ __forceinline__ Tables stage_tables(void* smem, int tid, int nthreads) {
  double2* lt = reinterpret_cast<double2*>(smem);
  double2* et = lt + 128 * kTableRep;
  for (int i = tid; i < 128 * kTableRep; i += nthreads) lt[i] = make_double2(kLogTab[i / kTableRep][0], kLogTab[i / kTableRep][1]);
  for (int i = tid; i < 64 * kTableRep; i += nthreads) et[i] = make_double2(kExpTab[i / kTableRep][0], kExpTab[i / kTableRep][1]);
  const int copy = tid & (kTableRep - 1);
  return Tables{lt + copy, et + copy};
}

// Out-of-range arguments are rare: one shared, never-inlined copy of each libdevice fallback
// keeps the hot code small (the instruction cache matters more than these calls).
static __device__ __noinline__ double slow_log(double x) { return ::log(x); }
static __device__ __noinline__ double slow_exp(double x) { return ::exp(x); }
static __device__ __noinline__ double slow_expm1(double x) { return ::expm1(x); }

// 1/x: MUFU.RCP64H seed (~2^-22) + two Newton steps. x must be a finite normal non-zero number;
// 0, inf, NaN give NaN/inf garbage that stays non-finite.
__device__ __forceinline__ double rcp(double x) {
  double r;
#ifdef MARLPDE_HOST_EMU   // tests/emu/: a ~2^-22 seed from the high word, like MUFU.RCP64H
  r = __hiloint2double(__double2hiint(1.0 / __hiloint2double(__double2hiint(x), 0)), 0);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// a/b with one residual correction (<= 1 ulp for normal operands)
__device__ __forceinline__ double div(double a, double b) {
  const double r = rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

__device__ __forceinline__ double log(const Tables& tb, double x) {
  const int hi = __double2hiint(x);
  // fast range: positive normal finite  <=>  0x00100000 <= hi < 0x7ff00000
  if (__builtin_expect((unsigned)(hi - 0x00100000) >= 0x7fe00000u, 0)) return slow_log(x);
  const int e = (hi >> 20) - 1023;
  const int j = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = tb.logtab[j * kTableRep];
  const double r = fma(m, t.x, -1.0);                    // |r| <= 2^-8, exact up to 2^-61
  double p = fma(FM_LOG(5), r, FM_LOG(4));
  p = fma(p, r, FM_LOG(3));
  p = fma(p, r, FM_LOG(2));
  p = fma(p, r, FM_LOG(1));
  p = fma(p, r, FM_LOG(0));
  const double l1p = fma(r * r, p, r);                   // log1p(r)
  const double ed = (double)e;
  return fma(ed, FM_LN2HI, t.y) + fma(ed, FM_LN2LO, l1p);
}

// shared core of exp/expm1: x = (64 k + j) ln2/64 + r ; returns p = expm1(r), sets scale 2^k T_j.
// Degree 5 leaves r^6/720 <= 3.5e-17 relative to exp(r) ~ 1; expm1 needs it relative to r, so it
// takes the r^6 term as well (kExtraTerm).
template <bool kExtraTerm>
__device__ __forceinline__ double exp_core(const Tables& tb, double x, double2& T, int& k) {
  const double magic = 6755399441055744.0;              // 1.5 * 2^52: round-to-nearest-int trick
  const double tn = fma(x, FM_K64, magic);
  const int n = __double2loint(tn);
  const double nf = tn - magic;
  double r = fma(nf, -FM_L64HI, x);
  r = fma(nf, -FM_L64LO, r);                             // |r| <= ln2/128
  T = tb.exptab[(n & 63) * kTableRep];
  k = n >> 6;
  double p = kExtraTerm ? fma(FM_EXP(4), r, FM_EXP(3)) : FM_EXP(3);
  p = fma(p, r, FM_EXP(2));
  p = fma(p, r, FM_EXP(1));
  p = fma(p, r, FM_EXP(0));
  return fma(r * r, p, r);
}

__device__ __forceinline__ double scale2(double v, int k) {  // v * 2^k for results that stay normal
  return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

__device__ __forceinline__ double exp(const Tables& tb, double x) {
  if (__builtin_expect(!(fabs(x) < 690.0), 0)) return slow_exp(x);
  double2 T;
  int k;
  const double p = exp_core<false>(tb, x, T, k);
  return scale2(fma(T.x, p, T.x), k);
}

__device__ __forceinline__ double expm1(const Tables& tb, double x) {
  if (__builtin_expect(!(fabs(x) < 600.0), 0)) return slow_expm1(x);
  double2 T;
  int k;
  const double p = exp_core<true>(tb, x, T, k);
  const double sc = __hiloint2double((k + 1023) << 20, 0);   // 2^k, |k| <= 866 here
  const double s = T.x * sc;                               // exact
  const double sl = T.y * sc;                              // exact (Tlo may be 0)
  return fma(s, p, s - 1.0) + fma(sl, p, sl);
}


// ---- branch-free variants for the two-cells-per-thread RHS ------------------------------------
// Same arithmetic as log/exp above, but an argument outside the fast range only raises `bad`
// (the caller re-evaluates the whole cell on the generic path once, at the end) instead of
// branching per call: the RHS stays one basic block per phase, which is what lets ptxas
// interleave the independent dependency chains of the two cells a thread owns.
// 1/x: MUFU.RCP64H seed + one cubic step r(1 + e + e^2): |rel err| <= e^3 + 1 ulp.
__device__ __forceinline__ double rcp3(double x) {
  double r;
#ifdef MARLPDE_HOST_EMU   // tests/emu/: a ~2^-22 seed from the high word, like MUFU.RCP64H
  r = __hiloint2double(__double2hiint(1.0 / __hiloint2double(__double2hiint(x), 0)), 0);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}

__device__ __forceinline__ double log_nb(const Tables& tb, double x, bool& bad) {
  const int hi = __double2hiint(x);
  bad |= (unsigned)(hi - 0x00100000) >= 0x7fe00000u;
  const int e = (hi >> 20) - 1023;
  const int j = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = tb.logtab[j * kTableRep];
  const double r = fma(m, t.x, -1.0);
  double p = fma(FM_LOG(5), r, FM_LOG(4));
  p = fma(p, r, FM_LOG(3));
  p = fma(p, r, FM_LOG(2));
  p = fma(p, r, FM_LOG(1));
  p = fma(p, r, FM_LOG(0));
  const double l1p = fma(r * r, p, r);
  const double ed = (double)e;
  return fma(ed, FM_LN2HI, t.y) + fma(ed, FM_LN2LO, l1p);
}

__device__ __forceinline__ double exp_nb(const Tables& tb, double x, bool& bad) {
#if MARLPDE_EXP_INTCHECK
  // !(|x| < 690) on the high word: 690.0 = 0x4085900000000000 has a zero low word, NaN / inf compare above it
  bad |= (unsigned)(__double2hiint(x) & 0x7fffffff) >= 0x40859000u;
#else
  bad |= !(fabs(x) < 690.0);
#endif
  double2 T;
  int k;
  const double p = exp_core<false>(tb, x, T, k);
  return scale2(fma(T.x, p, T.x), k);
}

// expm1 for arguments the caller guarantees to be in (-600, 600) on every lane whose result is used
__device__ __forceinline__ double expm1_nb(const Tables& tb, double x) {
  double2 T;
  int k;
  const double p = exp_core<true>(tb, x, T, k);
  const double sc = __hiloint2double((k + 1023) << 20, 0);
  const double s = T.x * sc;
  const double sl = T.y * sc;
  return fma(s, p, s - 1.0) + fma(sl, p, sl);
}

}  // namespace fm
}  // namespace marlpde
