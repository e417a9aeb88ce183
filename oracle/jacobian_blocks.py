"""TEST INFRASTRUCTURE ONLY — numpy restatement of the ANALYTIC 5x5 Jacobian blocks of the RHS
(marlpde/LHeureux_model.py:361-522) as the implicit kernels form them (csrc/implicit_common.cuh jac_analytic):
L_i = d rhs_i / d y_{i-1}, D_i = d rhs_i / d y_i, U_i = d rhs_i / d y_{i+1}, rows = rates, columns = fields
(CA, CC, cCa, cCO3, Phi).  The piecewise pieces of the model — upwind direction by the sign of U, the three regimes of
the Fiadeiro-Veronis weight, the clamped saturation products, the dissolution mask — are differentiated inside their
current regime (what a one-sided finite difference sees away from a switch).  Checked against central differences of the
oracle RHS in tests/test_host_side.py.  Parameter vector `p`: oracle.kernel_params (P_NAMES order)."""
import numpy as np


SWITCH_TOL = 1e-5          # csrc/implicit_common.cuh kSwitchTol


def blocks(y, p, i, N, atol=None):
    """`atol` given: the kernels' rule for cells ON a switching surface (|Pe| within SWITCH_TOL of Pe_min / Pe_max, U within
    SWITCH_TOL of 0) — the porosity column of D is num_jac's one-sided difference quotient (step sqrt(eps) sign(f)
    max(atol, |Phi|)) instead of the in-regime derivative."""
    Y = y.reshape(5, N)
    c = Y[:, i]
    CA, CC, a, o, P = c
    bc = p[0:5]
    dx, inv_dx2, delta_x, presum, rhorat, Da, lam = p[5], p[6], p[7], p[8], p[9], p[10], p[11]
    dCa, dCO3, delta, KRat, nu1, nu2, m1, m2, n1, n2 = p[12:22]
    dPhi_fixed, Pe_min, Pe_max, fv_on = p[22], p[23], p[24], p[25] != 0
    in_mask = int(p[26]) <= i < int(p[27])
    auxcon, var_dphi = p[28], p[29] != 0
    first, last = i == 0, i == N - 1
    m = Y[:, i - 1] if not first else 2 * bc - c
    pl = Y[:, i + 1] if not last else np.array([2 * c[0] - m[0], 2 * c[1] - m[1], c[2], c[3], c[4]])
    inv_dx, hdx = 1 / dx, 0.5 / dx

    # ---- porosity functions and their derivatives
    E = np.exp(10 - 10 / P)
    F, dF = 1 - E, -E * 10 / P ** 2
    omP = 1 - P
    FoP = F / omP
    dFoP = (dF * omP + F) / omP ** 2
    U = presum + rhorat * P ** 3 * FoP
    dU = rhorat * (3 * P ** 2 * FoP + P ** 3 * dFoP)
    W = presum - rhorat * P ** 2 * F
    dW = -rhorat * (2 * P * F + P ** 2 * dF)
    ddW = -rhorat * (2 * F + 4 * P * dF + P ** 2 * (-dF * 10 / P ** 2 * 0 + _d2F(P)))
    den, dden = 1 - 2 * np.log(P), -2 / P
    if var_dphi:
        dPhi, ddPhi = auxcon * P ** 3 * FoP, auxcon * (3 * P ** 2 * FoP + P ** 3 * dFoP)
    else:
        dPhi, ddPhi = dPhi_fixed, 0.0
    kCa, kCO3, kPhi = delta_x / (2 * dCa), delta_x / (2 * dCO3), delta_x / (2 * dPhi)
    dkPhi = -delta_x / (2 * dPhi ** 2) * ddPhi

    def sig(Pe, dPe):
        """Fiadeiro-Veronis weight and its derivative with respect to Phi."""
        if not fv_on or abs(Pe) < Pe_min:
            return 0.0, 0.0
        if abs(Pe) > Pe_max:
            return np.sign(W), 0.0
        em = np.expm1(2 * Pe)
        return 1 + 2 / em - 1 / Pe, (1 / Pe ** 2 - 4 * (em + 1) / em ** 2) * dPe

    s2, ds2 = sig(W * den * kCa, kCa * (dW * den + W * dden))
    s3, ds3 = sig(W * den * kCO3, kCO3 * (dW * den + W * dden))
    s4, ds4 = sig(W * kPhi, dW * kPhi + W * dkPhi)
    s = {2: s2, 3: s3, 4: s4}
    ds = {2: ds2, 3: ds3, 4: ds4}
    g = {f: ((1 - s[f]) * (pl[f] - c[f]) + (1 + s[f]) * (c[f] - m[f])) * hdx for f in (2, 3, 4)}
    dg_own = {f: 2 * s[f] * hdx for f in (2, 3, 4)}                      # d g_f / d (own value of field f)
    dg_P = {f: ((c[f] - m[f]) - (pl[f] - c[f])) * hdx * ds[f] for f in (2, 3, 4)}   # through the weight
    dg_m = {f: -(1 + s[f]) * hdx for f in (2, 3, 4)}
    dg_p = {f: (1 - s[f]) * hdx for f in (2, 3, 4)}
    lap = {f: (m[f] - 2 * c[f] + pl[f]) * inv_dx2 for f in (2, 3, 4)}
    rden = 1 / den
    h1, dh1 = P * rden, rden + 2 * rden ** 2
    h2c = (2 + den) * rden ** 2
    dh2c = dden * rden ** 2 * (1 - 2 * (2 + den) * rden)
    gP_P = dg_own[4] + dg_P[4]                                            # d gPhi / d Phi (own)
    h2, dh2 = g[4] * h2c, gP_P * h2c + g[4] * dh2c
    dif = {2: dCa, 3: dCO3}
    H = {f: dif[f] * (h2 * g[f] + h1 * lap[f]) for f in (2, 3)}

    # ---- reaction terms
    two = a * o
    three = two * KRat
    if three < 1:
        A = (1 - three) ** m2 if in_mask else 0.0
        dA = (-m2 * (1 - three) ** (m2 - 1)) if in_mask else 0.0
    else:
        A, dA = -nu1 * (three - 1) ** m1, -nu1 * m1 * (three - 1) ** (m1 - 1) if three > 1 else 0.0
    if two < 1:
        Cc, dC = -nu2 * (1 - two) ** n2, nu2 * n2 * (1 - two) ** (n2 - 1)
    else:
        Cc, dC = (two - 1) ** n1, n1 * (two - 1) ** (n1 - 1) if two > 1 else 0.0
    coA, coC = CA * A, CC * Cc
    dcoA = np.array([A, 0.0, CA * dA * KRat * o, CA * dA * KRat * a, 0.0])
    dcoC = np.array([0.0, Cc, CC * dC * o, CC * dC * a, 0.0])
    h3 = coA - lam * coC
    dh3 = dcoA - lam * dcoC
    react = Da * omP * h3
    dreact = Da * omP * dh3
    dreact[4] = -Da * h3

    back = U > 0
    gA = ((CA - m[0]) if back else (pl[0] - CA)) * inv_dx
    gC = ((CC - m[1]) if back else (pl[1] - CC)) * inv_dx
    dgs_own = inv_dx if back else -inv_dx

    D = np.zeros((5, 5))
    D[0] = -Da * ((1 - CA) * dcoA + lam * CA * dcoC)
    D[0, 0] += -U * dgs_own - Da * (-coA + lam * coC)
    D[0, 4] += -dU * gA
    D[1] = Da * (lam * (1 - CC) * dcoC + CC * dcoA)
    D[1, 1] += -U * dgs_own + Da * (-lam * coC + coA)
    D[1, 4] += -dU * gC
    for r, f in ((2, 2), (3, 3)):
        D[r] = dreact * (delta - c[f]) / P
        D[r, f] += (dif[f] * (h2 * dg_own[f] - 2 * h1 * inv_dx2) - react) / P - W * dg_own[f]
        D[r, 4] += dif[f] * (dh2 * g[f] + h2 * dg_P[f] + dh1 * lap[f]) / P - (H[f] + react * (delta - c[f])) / P ** 2 \
            - dW * g[f] - W * dg_P[f]
    t4 = dW * P + W
    D[4] = dreact
    D[4, 4] += -gP_P * t4 - g[4] * (ddW * P + 2 * dW) + ddPhi * lap[4] - 2 * dPhi * inv_dx2

    L, Ub = np.zeros((5, 5)), np.zeros((5, 5))
    for blk, dgx, dgs in ((L, dg_m, (-inv_dx if back else 0.0)), (Ub, dg_p, (0.0 if back else inv_dx))):
        blk[0, 0] = blk[1, 1] = -U * dgs
        for r, f in ((2, 2), (3, 3)):
            blk[r, f] = dif[f] * (h2 * dgx[f] + h1 * inv_dx2) / P - W * dgx[f]
            blk[r, 4] = dif[f] * g[f] * h2c * dgx[4] / P
        blk[4, 4] = dPhi * inv_dx2 - t4 * dgx[4]
    # ---- ghost cells fold into the own-cell block (LHeureux_model.py:26-30): top ghost 2 bc - c for every field;
    # bottom ghost 2 c - m for CA, CC (curvature 0), c for the solutes and the porosity (derivative 0)
    pes = [abs(W * den * kCa), abs(W * den * kCO3), abs(W * kPhi)] if fv_on else []
    on_switch = abs(U) <= SWITCH_TOL * max(1.0, abs(presum)) or any(
        abs(pe - Pe_min) <= SWITCH_TOL * Pe_min or abs(pe - Pe_max) <= SWITCH_TOL * Pe_max for pe in pes)
    if first:
        D -= L
        L[:] = 0
    if last:
        D[:, :2] += 2 * Ub[:, :2]
        L[:, :2] -= Ub[:, :2]
        D[:, 2:] += Ub[:, 2:]
        Ub[:] = 0
    if atol is not None and on_switch:
        import lheureux_oracle as oracle
        rows = [f * N + i for f in range(5)]
        f0 = oracle.rhs(np.ascontiguousarray(y, dtype=np.float64), p, np.empty(5 * N))[rows].copy()
        ys = (1.0 if f0[4] >= 0 else -1.0) * max(atol, abs(P))
        h = (P + np.sqrt(np.finfo(float).eps) * ys) - P
        y2 = np.array(y, dtype=np.float64)
        y2[4 * N + i] = P + h
        D[:, 4] = (oracle.rhs(y2, p, np.empty(5 * N))[rows] - f0) / h
    return L, D, Ub


def _d2F(P):
    """Second derivative of F = 1 - exp(10 - 10/Phi)."""
    E = np.exp(10 - 10 / P)
    return -E * (100 / P ** 4 - 20 / P ** 3)
