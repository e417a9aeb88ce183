// events.cuh — sign bookkeeping of the 7 event monitors, shared by the on-chip and the streaming RK45 kernels.
#pragma once

#include "../../include/marlpde_b200.h"

namespace marlpde {

// ---- event monitors (LHeureux_model.py:524-593, all non-terminal, direction 0) -----------------
// k: 0 min(y)  1 min(CA)  2 min(CC)  3 max(CA+CC)-1  4 max(Phi)-1  5 min(U)  6 max(W).
// solve_ivp only needs the SIGN of each monitor after every accepted step (ivp.py
// find_active_events); signs of a min/max follow from per-cell predicates, so detection is an
// OR-reduction of 21 bits per thread (3 per monitor: beyond the threshold / on it / NaN) that
// rides on the error-norm barrier — no fp64 reduction unless a sign change has to be located.
constexpr unsigned kMaxTypeMask = (1u << 3) | (1u << 4) | (1u << 6);   // monitors that are a max
constexpr unsigned kEqBitsMask = 0x92492u;                              // the "on the threshold" bit of every monitor

__device__ __forceinline__ unsigned event_bits(const double (&v)[5][2], const double (&U)[2],
                                               const double (&W)[2], bool has1) {
  unsigned b = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q == 1 && !has1) break;
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

// the same bits for cells that are selected one by one (streaming kernels: window overlap, column ends)
__device__ __forceinline__ unsigned event_bits_sel(const double (&v)[5][2], const double (&U)[2],
                                                   const double (&W)[2], bool use0, bool use1) {
  double w[5][2], u2[2], w2[2];
  unsigned b = 0;
  if (use0) b |= event_bits(v, U, W, false);
  if (use1) {
#pragma unroll
    for (int f = 0; f < 5; ++f) w[f][0] = w[f][1] = v[f][1];
    u2[0] = u2[1] = U[1];
    w2[0] = w2[1] = W[1];
    b |= event_bits(w, u2, w2, false);
  }
  return b;
}

// 21 predicate bits -> 7 sign classes, 2 bits each: 0 negative, 1 zero, 2 positive, 3 NaN
__device__ __forceinline__ unsigned event_classes(unsigned bits) {
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

// ivp.py find_active_events with direction 0: (g <= 0 & g_new >= 0) | (g >= 0 & g_new <= 0)
__device__ __forceinline__ unsigned active_events(unsigned cls_old, unsigned cls_new) {
  unsigned act = 0;
#pragma unroll
  for (int k = 0; k < MARLPDE_NEVENTS; ++k) {
    const unsigned a = (cls_old >> (2 * k)) & 3u, b = (cls_new >> (2 * k)) & 3u;
    const bool a_le = a <= 1u, a_ge = a == 1u || a == 2u, b_le = b <= 1u, b_ge = b == 1u || b == 2u;
    if ((a_le && b_ge) || (a_ge && b_le)) act |= 1u << k;
  }
  return act;
}

}  // namespace marlpde
