// brent.cuh — the root finder solve_ivp uses to locate events, shared by the RK45 and Radau kernels.
#pragma once

namespace marlpde {

// scipy.optimize.brentq (Zeros/brentq.c) as a resumable state machine: every function value needs
// a block-wide reduction, so the caller feeds values one at a time.  Transliteration of
// oracle/lheureux_oracle.py::brentq_restated, which tests pin bit-for-bit to the installed SciPy;
// explicit round-to-nearest intrinsics keep nvcc from contracting a*b+c (brentq.c is not).
struct BrentState {
  double xpre, xcur, xblk, fpre, fcur, fblk, spre, scur;
  int stage, iter;
  __device__ __forceinline__ void init(double xa, double xb) {
    xpre = xa;
    xcur = xb;
    xblk = fblk = spre = scur = 0.0;
    fpre = fcur = 0.0;
    stage = 0;
    iter = 0;
  }
  // feed f(xeval); returns true when finished (root set), else xeval = next abscissa
  __device__ bool feed(double g, double& xeval, double& root) {
    const double tol = 4.0 * 2.220446049250313e-16;
    if (stage == 0) {
      fpre = g;
      if (fpre == 0.0) { root = xpre; return true; }
      stage = 1;
      xeval = xcur;
      return false;
    }
    fcur = g;
    if (stage == 1) {
      if (fcur == 0.0) { root = xcur; return true; }
      if (signbit(fpre) == signbit(fcur)) {      // SciPy raises ValueError here; keep the closer end
        root = fabs(fpre) < fabs(fcur) ? xpre : xcur;
        return true;
      }
      stage = 2;
    }
    if (iter >= 100) { root = xcur; return true; }
    ++iter;
    if (fpre != 0.0 && fcur != 0.0 && signbit(fpre) != signbit(fcur)) {
      xblk = xpre;
      fblk = fpre;
      spre = scur = __dsub_rn(xcur, xpre);
    }
    if (fabs(fblk) < fabs(fcur)) {
      xpre = xcur; xcur = xblk; xblk = xpre;
      fpre = fcur; fcur = fblk; fblk = fpre;
    }
    const double delta = __dmul_rn(__dadd_rn(tol, __dmul_rn(tol, fabs(xcur))), 0.5);
    const double sbis = __dmul_rn(__dsub_rn(xblk, xcur), 0.5);
    if (fcur == 0.0 || fabs(sbis) < delta) { root = xcur; return true; }
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      double stry;
      if (xpre == xblk) {
        stry = __ddiv_rn(__dmul_rn(-fcur, __dsub_rn(xcur, xpre)), __dsub_rn(fcur, fpre));
      } else {
        const double dpre = __ddiv_rn(__dsub_rn(fpre, fcur), __dsub_rn(xpre, xcur));
        const double dblk = __ddiv_rn(__dsub_rn(fblk, fcur), __dsub_rn(xblk, xcur));
        stry = __ddiv_rn(__dmul_rn(-fcur, __dsub_rn(__dmul_rn(fblk, dblk), __dmul_rn(fpre, dpre))),
                         __dmul_rn(__dmul_rn(dblk, dpre), __dsub_rn(fblk, fpre)));
      }
      if (__dmul_rn(2.0, fabs(stry)) < fmin(fabs(spre), __dsub_rn(__dmul_rn(3.0, fabs(sbis)), delta))) {
        spre = scur;
        scur = stry;
      } else {
        spre = sbis;
        scur = sbis;
      }
    } else {
      spre = sbis;
      scur = sbis;
    }
    xpre = xcur;
    fpre = fcur;
    if (fabs(scur) > delta) xcur = __dadd_rn(xcur, scur);
    else xcur = __dadd_rn(xcur, sbis > 0.0 ? delta : -delta);
    xeval = xcur;
    return false;
  }
};

// ivp.py find_active_events for one monitor with direction 0, on the monitor VALUES
__device__ __forceinline__ bool event_active(double g_old, double g_new) {
  return (g_old <= 0.0 && g_new >= 0.0) || (g_old >= 0.0 && g_new <= 0.0);
}

}  // namespace marlpde
