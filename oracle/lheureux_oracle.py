"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

A plain NumPy/Numba restatement of the reference's hot path so the CUDA kernels
can be checked against it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module.

What it restates (reference = /root/reference, astro-turing/
Integrating-diagenetic-equations-using-Python):

* derived constants      marlpde/LHeureux_model.py:31-72, :87-88, :130-133
* grid / masks / y0      marlpde/Evolve_scenario.py:40, :51-54, :64-65, :76-86
* boundary conditions    marlpde/LHeureux_model.py:26-30 with py-pde 0.32.2 ghost-cell
                         rules (py-pde is an un-vendored dependency pinned in
                         poetry.lock:955-956; its rules are restated from the
                         published library: value -> 2v-a0, derivative 0 -> a_{N-1},
                         curvature 0 -> 2a_{N-1}-a_{N-2}; forward/backward
                         differences divide by dx, the Laplacian multiplies by dx**-2)
* per-cell maths         marlpde/LHeureux_model.py:413-520 (the numba variant's
                         operation order is followed literally)
* events                 marlpde/LHeureux_model.py:524-593
* time stepping          the reference calls scipy.integrate.solve_ivp
                         (marlpde/Evolve_scenario.py:104-109); SciPy is installed in this
                         image, so the oracle drives the *real* SciPy steppers.

Parity pinning: see tests/test_oracle_golden.py — (i) the reference's own HDF5
regression fixtures (tests/Regression_test/data, extracted to tests/golden/ by
tests/golden/make_golden.py) are reproduced by this oracle with Radau at tight
tolerance; (ii) the reference's own `pde_rhs` / `fun` code, imported from
/root/reference with stand-ins for the absent py-pde/pint packages
(oracle/refstubs), produced tests/golden/rhs_reference_*.npz, which this oracle
reproduces bit-for-bit.  The py-pde `curvature` ghost rule (reached only when
U <= 0) is specification-by-recall: no reference test exercises it.
"""
from __future__ import annotations

import math

import numpy as np
from numba import njit

FIELDS = ("CA", "CC", "cCa", "cCO3", "Phi")

# Parameter vector layout shared by the numba RHS and oracle_rhs.c (index -> meaning)
P_NAMES = (
    "CA0", "CC0", "cCa0", "cCO30", "Phi0",        # 0-4  Dirichlet values at the top (LHeureux_model.py:26-30)
    "dx", "dx_m2", "delta_x",                      # 5-7  grid spacing, dx**-2, x[1]-x[0]
    "presum", "rhorat", "Da", "lambda_",           # 8-11
    "dCa", "dCO3", "delta", "KRat",                # 12-15
    "nu1", "nu2", "m1", "m2", "n1", "n2",          # 16-21
    "dPhi_fixed", "Peclet_min", "Peclet_max",      # 22-24
    "FV_switch", "mask_lo", "mask_hi",             # 25-27 (mask = 1 on cells [mask_lo, mask_hi))
    "auxcon", "var_dPhi",                          # 28-29 model variant: the line upstream keeps commented out at
)                                                  #       LHeureux_model.py:222-223 / :430-431 (dPhi per cell)
NP = len(P_NAMES)
P_IDX = {n: i for i, n in enumerate(P_NAMES)}


def default_scenario() -> dict:
    """Values of asdict(Map_Scenario()) (marlpde/parameters.py:16-48, :60-143).

    Restated with the same floating-point expressions as `post_init`
    (parameters.py:120-143); checked against the reference's own parameters.py
    (imported with a stand-in pint) in tests/golden/make_golden.py.
    """
    p = dict(
        muA=100.09, rhoa=2.95, rhoc=2.71, rhot=2.8, rhow=1.023, D0Ca=131.9, DCO3=272.6,
        KA=10 ** (-6.19), KC=10 ** (-6.37), beta=0.1, b=5.0, k1=1.0, k2=1.0, k3=0.1, k4=0.1,
        n1=2.8, m1=2.48, sedimentationrate=0.1, PhiInfty=0.01, Phi0=0.8, ca0=0.326e-3,
        co30=0.326e-3, CC0=0.3, CA0=0.6, ShallowLimit=50.0, max_depth=500.0, Th=100.0,
        PhiIni=0.8, ca00=0.326e-3, co300=0.326e-3, CCIni=0.3, CAIni=0.6,
    )
    return derive_scenario(p)


def derive_scenario(p: dict) -> dict:
    """parameters.py:120-143 (`post_init`) applied to raw Scenario magnitudes."""
    p = dict(p)
    p["cCa0"] = p["ca0"] / np.sqrt(p["KC"])
    p["cCaIni"] = p["ca00"] / np.sqrt(p["KC"])
    p["cCO30"] = p["co30"] / np.sqrt(p["KC"])
    p["cCO3Ini"] = p["co300"] / np.sqrt(p["KC"])
    p["DeepLimit"] = p["ShallowLimit"] + p["Th"]
    p["rhos0"] = p["rhoa"] * p["CA0"] + p["rhoc"] * p["CC0"] + p["rhot"] * (1 - (p["CA0"] + p["CC0"]))
    p["rhos"] = p["rhos0"]
    p["Xstar"] = p["D0Ca"] / p["sedimentationrate"]
    p["Tstar"] = p["Xstar"] / p["sedimentationrate"]
    p["b"] = p["b"] / 1e4
    p["m2"] = p["m1"]
    p["n2"] = p["n1"]
    p["DCa"] = p["D0Ca"]
    p["PhiNR"] = p["PhiIni"]
    p["N"] = 200
    p["FV_switch"] = 1
    return p


def grid_coords(pde: dict) -> tuple[np.ndarray, float]:
    """Cell-centred grid of Evolve_scenario.py:40 (py-pde CartesianGrid)."""
    n = int(pde["N"])
    length = pde["max_depth"] / pde["Xstar"]
    dx = (length - 0.0) / n
    x = 0.0 + (np.arange(n) + 0.5) * dx
    return x, dx


def masks(pde: dict) -> tuple[np.ndarray, np.ndarray]:
    """Heaviside masks of Evolve_scenario.py:51-54 (H(0) = 0)."""
    x, _ = grid_coords(pde)
    sh = pde["ShallowLimit"] / pde["Xstar"]
    dp = pde["DeepLimit"] / pde["Xstar"]
    not_too_shallow = np.where(x - sh > 0, 1.0, 0.0)
    not_too_deep = np.where(dp - x > 0, 1.0, 0.0)
    return not_too_shallow, not_too_deep


def kernel_params(pde: dict) -> np.ndarray:
    """Derived constants of LMAHeureuxPorosityDiff.__init__ (LHeureux_model.py:31-72,
    :87-88, :130-133), same expression order, packed per P_NAMES."""
    g = 100 * 9.81
    k1, k2, k3, k4 = pde["k1"], pde["k2"], pde["k3"], pde["k4"]
    nu1 = k1 / k2
    nu2 = k4 / k3
    KRat = pde["KC"] / pde["KA"]
    dCa = pde["DCa"] / pde["D0Ca"]
    dCO3 = pde["DCO3"] / pde["D0Ca"]
    delta = pde["rhos"] / (pde["muA"] * np.sqrt(pde["KC"]))
    Da = k2 * pde["Tstar"]
    lambda_ = k3 / k2
    auxcon = pde["beta"] / (pde["D0Ca"] * pde["b"] * g * pde["rhow"] * (pde["PhiNR"] - pde["PhiInfty"]))
    rhorat0 = (pde["rhos0"] / pde["rhow"] - 1) * pde["beta"] / pde["sedimentationrate"]
    rhorat = (pde["rhos"] / pde["rhow"] - 1) * pde["beta"] / pde["sedimentationrate"]
    Phi0 = pde["Phi0"]
    presum = 1 - rhorat0 * Phi0 ** 3 * (1 - np.exp(10 - 10 / Phi0)) / (1 - Phi0)
    PhiIni = pde["PhiIni"]
    F_fixed = 1 - np.exp(10 - 10 / PhiIni)
    dPhi_fixed = auxcon * F_fixed * PhiIni ** 3 / (1 - PhiIni)
    x, dx = grid_coords(pde)
    delta_x = x[1] - x[0]
    nts, ntd = masks(pde)
    m = nts * ntd
    nz = np.nonzero(m)[0]
    if nz.size:
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        assert np.all(m[lo:hi] == 1.0)
    else:
        lo = hi = 0
    vals = dict(
        CA0=pde["CA0"], CC0=pde["CC0"], cCa0=pde["cCa0"], cCO30=pde["cCO30"], Phi0=Phi0,
        dx=dx, dx_m2=dx ** -2, delta_x=delta_x, presum=presum, rhorat=rhorat, Da=Da,
        lambda_=lambda_, dCa=dCa, dCO3=dCO3, delta=delta, KRat=KRat, nu1=nu1, nu2=nu2,
        m1=pde["m1"], m2=pde["m2"], n1=pde["n1"], n2=pde["n2"], dPhi_fixed=dPhi_fixed,
        Peclet_min=1e-2, Peclet_max=1 / 1e-2, FV_switch=float(pde["FV_switch"]),
        mask_lo=float(lo), mask_hi=float(hi), auxcon=auxcon,
        var_dPhi=float(bool(pde.get("time_varying_dPhi", False))),
    )
    return np.array([vals[n] for n in P_NAMES], dtype=np.float64)


def initial_state(pde: dict) -> np.ndarray:
    """Field-major y0 (Evolve_scenario.py:64-65, :76-86)."""
    n = int(pde["N"])
    return np.concatenate([np.full(n, pde[k], dtype=np.float64)
                           for k in ("CAIni", "CCIni", "cCaIni", "cCO3Ini", "PhiIni")])


@njit(cache=True)
def _sigma(Pe, W, Pmin, Pmax):
    # LHeureux_model.py:437-442 (same for :445-450, :453-458)
    if abs(Pe) < Pmin:
        return 0.0
    elif abs(Pe) > Pmax:
        return np.sign(W)
    return np.cosh(Pe) / np.sinh(Pe) - 1 / Pe


@njit(cache=True)
def rhs(y, p, out):
    """One RHS evaluation, y and out field-major f64[5N] (LHeureux_model.py:361-522)."""
    N = y.size // 5
    CA = y[0:N]
    CC = y[N:2 * N]
    cCa = y[2 * N:3 * N]
    cCO3 = y[3 * N:4 * N]
    Phi = y[4 * N:5 * N]
    CA0, CC0, cCa0, cCO30, Phi0 = p[0], p[1], p[2], p[3], p[4]
    dx, dx_m2, delta_x = p[5], p[6], p[7]
    presum, rhorat, Da, lambda_ = p[8], p[9], p[10], p[11]
    dCa, dCO3, delta, KRat = p[12], p[13], p[14], p[15]
    nu1, nu2, m1, m2, n1, n2 = p[16], p[17], p[18], p[19], p[20], p[21]
    dPhi_fixed, Peclet_min, Peclet_max = p[22], p[23], p[24]
    FV_switch = p[25] != 0.0
    mask_lo, mask_hi = int(p[26]), int(p[27])
    auxcon, var_dPhi = p[28], p[29] != 0.0
    for i in range(N):
        # ---- ghost cells + stencils (py-pde rules; LHeureux_model.py:26-30, :372-384)
        if i > 0:
            CA_m, CC_m, cCa_m, cCO3_m, Phi_m = CA[i - 1], CC[i - 1], cCa[i - 1], cCO3[i - 1], Phi[i - 1]
        else:
            CA_m = 2 * CA0 - CA[0]
            CC_m = 2 * CC0 - CC[0]
            cCa_m = 2 * cCa0 - cCa[0]
            cCO3_m = 2 * cCO30 - cCO3[0]
            Phi_m = 2 * Phi0 - Phi[0]
        if i < N - 1:
            CA_p, CC_p, cCa_p, cCO3_p, Phi_p = CA[i + 1], CC[i + 1], cCa[i + 1], cCO3[i + 1], Phi[i + 1]
        else:
            CA_p = 2 * CA[N - 1] - CA[N - 2]
            CC_p = 2 * CC[N - 1] - CC[N - 2]
            cCa_p, cCO3_p, Phi_p = cCa[N - 1], cCO3[N - 1], Phi[N - 1]
        CA_grad_back = (CA[i] - CA_m) / dx
        CA_grad_forw = (CA_p - CA[i]) / dx
        CC_grad_back = (CC[i] - CC_m) / dx
        CC_grad_forw = (CC_p - CC[i]) / dx
        cCa_grad_back = (cCa[i] - cCa_m) / dx
        cCa_grad_forw = (cCa_p - cCa[i]) / dx
        cCa_laplace = (cCa_m - 2 * cCa[i] + cCa_p) * dx_m2
        cCO3_grad_back = (cCO3[i] - cCO3_m) / dx
        cCO3_grad_forw = (cCO3_p - cCO3[i]) / dx
        cCO3_laplace = (cCO3_m - 2 * cCO3[i] + cCO3_p) * dx_m2
        Phi_grad_back = (Phi[i] - Phi_m) / dx
        Phi_grad_forw = (Phi_p - Phi[i]) / dx
        Phi_laplace = (Phi_m - 2 * Phi[i] + Phi_p) * dx_m2

        # ---- cell loop, LHeureux_model.py:413-520 ---------------------------------
        F = 1 - np.exp(10 - 10 / Phi[i])
        U = presum + rhorat * Phi[i] ** 3 * F / (1 - Phi[i])
        if U > 0:
            CA_grad = CA_grad_back
            CC_grad = CC_grad_back
        else:
            CA_grad = CA_grad_forw
            CC_grad = CC_grad_forw
        W = presum - rhorat * Phi[i] ** 2 * F
        denominator = 1 - 2 * np.log(Phi[i])
        one_minus_Phi = 1 - Phi[i]
        if var_dPhi:                               # :430, the commented-out time-varying coefficient
            dPhi = auxcon * F * (Phi[i] ** 3) / one_minus_Phi
        else:
            dPhi = dPhi_fixed                      # :431
        if FV_switch:
            Peclet_cCa = W * delta_x * denominator / (2. * dCa)
            sigma_cCa = _sigma(Peclet_cCa, W, Peclet_min, Peclet_max)
            Peclet_cCO3 = W * delta_x * denominator / (2. * dCO3)
            sigma_cCO3 = _sigma(Peclet_cCO3, W, Peclet_min, Peclet_max)
            Peclet_Phi = W * delta_x / (2. * dPhi)
            sigma_Phi = _sigma(Peclet_Phi, W, Peclet_min, Peclet_max)
        else:
            sigma_cCa = 0.0
            sigma_cCO3 = 0.0
            sigma_Phi = 0.0
        cCa_grad = 0.5 * ((1 - sigma_cCa) * cCa_grad_forw + (1 + sigma_cCa) * cCa_grad_back)
        cCO3_grad = 0.5 * ((1 - sigma_cCO3) * cCO3_grad_forw + (1 + sigma_cCO3) * cCO3_grad_back)
        Phi_grad = 0.5 * ((1 - sigma_Phi) * Phi_grad_forw + (1 + sigma_Phi) * Phi_grad_back)

        common_helper1 = Phi[i] / denominator
        common_helper2 = Phi_grad * (2 + denominator) / denominator ** 2
        helper_cCa_grad = dCa * (common_helper2 * cCa_grad + common_helper1 * cCa_laplace)
        helper_cCO3_grad = dCO3 * (common_helper2 * cCO3_grad + common_helper1 * cCO3_laplace)

        two_factors = cCa[i] * cCO3[i]
        two_factors_upp_lim = min(two_factors, 1)
        two_factors_low_lim = max(two_factors, 1)
        three_factors = two_factors * KRat
        three_factors_upp_lim = min(three_factors, 1)
        three_factors_low_lim = max(three_factors, 1)
        mask = 1.0 if (i >= mask_lo and i < mask_hi) else 0.0

        coA = CA[i] * (((1 - three_factors_upp_lim) ** m2) * mask - nu1 *
                       (three_factors_low_lim - 1) ** m1)
        coC = CC[i] * (((two_factors_low_lim - 1) ** n1) - nu2 *
                       (1 - two_factors_upp_lim) ** n2)
        common_helper3 = coA - lambda_ * coC
        dW_dx = -rhorat * Phi_grad * (2 * Phi[i] * F + 10 * (F - 1))

        out[i] = - U * CA_grad - Da * ((1 - CA[i]) * coA + lambda_ * CA[i] * coC)
        out[N + i] = - U * CC_grad + Da * (lambda_ * (1 - CC[i]) * coC + CC[i] * coA)
        out[2 * N + i] = helper_cCa_grad / Phi[i] - W * cCa_grad + Da * one_minus_Phi * \
            (delta - cCa[i]) * common_helper3 / Phi[i]
        out[3 * N + i] = helper_cCO3_grad / Phi[i] - W * cCO3_grad + Da * one_minus_Phi * \
            (delta - cCO3[i]) * common_helper3 / Phi[i]
        out[4 * N + i] = - (dW_dx * Phi[i] + W * Phi_grad) + dPhi * Phi_laplace + \
            Da * one_minus_Phi * common_helper3
    return out


@njit(cache=True)
def term_scale(y, p, out):
    """Term-magnitude scale S_i: the sum of the absolute values of
    the cancelling terms of each RHS entry.  |rhs_a - rhs_b| <= tol * S is the well-posed
    statement of single-call parity on evolved states, where the RHS is a small residual
    of terms ~1e5 times larger."""
    N = y.size // 5
    dx = p[5]
    presum, rhorat, Da, lambda_ = p[8], p[9], p[10], p[11]
    dCa, dCO3, delta, KRat = p[12], p[13], p[14], p[15]
    nu1, nu2, m1, m2, n1, n2 = p[16], p[17], p[18], p[19], p[20], p[21]
    dPhi = p[22]
    mask_lo, mask_hi = int(p[26]), int(p[27])
    for i in range(N):
        CA, CC, cCa, cCO3, Phi = y[i], y[N + i], y[2 * N + i], y[3 * N + i], y[4 * N + i]
        F = 1 - np.exp(10 - 10 / Phi)
        U = presum + rhorat * Phi ** 3 * F / (1 - Phi)
        W = presum - rhorat * Phi ** 2 * F
        den = 1 - 2 * np.log(Phi)
        two = cCa * cCO3
        three = two * KRat
        mask = 1.0 if (i >= mask_lo and i < mask_hi) else 0.0
        coAp = abs(CA) * ((1 - min(three, 1)) ** m2 * mask + nu1 * (max(three, 1) - 1) ** m1)
        coCp = abs(CC) * ((max(two, 1) - 1) ** n1 + nu2 * (1 - min(two, 1)) ** n2)
        h3p = coAp + lambda_ * coCp
        out[i] = abs(U) * 2 * abs(CA) / dx + Da * (abs(1 - CA) * coAp + lambda_ * abs(CA) * coCp)
        out[N + i] = abs(U) * 2 * abs(CC) / dx + Da * (lambda_ * abs(1 - CC) * coCp + abs(CC) * coAp)
        out[2 * N + i] = dCa * 4 * abs(cCa) / (dx * dx * abs(den)) + abs(W) * 2 * abs(cCa) / dx + \
            Da * abs(1 - Phi) * (delta + abs(cCa)) * h3p / Phi
        out[3 * N + i] = dCO3 * 4 * abs(cCO3) / (dx * dx * abs(den)) + abs(W) * 2 * abs(cCO3) / dx + \
            Da * abs(1 - Phi) * (delta + abs(cCO3)) * h3p / Phi
        out[4 * N + i] = 2 * Phi / dx * (abs(W) + rhorat * abs(2 * Phi * F + 10 * (F - 1))) + \
            dPhi * 4 * Phi / (dx * dx) + Da * abs(1 - Phi) * h3p
    return out


def rhs_fn(p: np.ndarray):
    """f(t, y) closure for SciPy (fresh output array each call, like fun_numba)."""
    def f(t, y):
        return rhs(np.ascontiguousarray(y, dtype=np.float64), p, np.empty_like(y))
    return f


def event_fns(p: np.ndarray, N: int):
    """The 7 non-terminal monitors of LHeureux_model.py:524-593 (direction 0)."""
    presum, rhorat = p[P_IDX["presum"]], p[P_IDX["rhorat"]]

    def zeros(t, y):
        return np.amin(y)

    def zeros_CA(t, y):
        return np.amin(y[0:N])

    def zeros_CC(t, y):
        return np.amin(y[N:2 * N])

    def ones_CA_plus_CC(t, y):
        return np.amax(y[0:N] + y[N:2 * N]) - 1

    def ones_Phi(t, y):
        return np.amax(y[4 * N:]) - 1

    def zeros_U(t, y):
        Phi = y[4 * N:]
        F = 1 - np.exp(10 - 10 / Phi)
        return np.amin(presum + rhorat * Phi ** 3 * F / (1 - Phi))

    def zeros_W(t, y):
        Phi = y[4 * N:]
        F = 1 - np.exp(10 - 10 / Phi)
        return np.amax(presum - rhorat * Phi ** 2 * F)

    evs = [zeros, zeros_CA, zeros_CC, ones_CA_plus_CC, ones_Phi, zeros_U, zeros_W]
    for e in evs:
        e.terminal = False
    return evs


def integrate(pde: dict, method: str = "RK45", first_step: float = 1e-6, rtol: float = 1e-3,
              atol: float = 1e-3, t_span=(0.0, 1.0), t_eval=None, events: bool = True, y0=None, **opts):
    """What Evolve_scenario.py:104-109 does: SciPy solve_ivp on the restated RHS
    (`y0` overrides the uniform initial state of Evolve_scenario.py:76-86, for tests)."""
    from scipy.integrate import solve_ivp
    p = kernel_params(pde)
    y0 = initial_state(pde) if y0 is None else np.ascontiguousarray(y0, dtype=np.float64).ravel()
    N = int(pde["N"])
    if t_eval is None:
        t_eval = np.linspace(t_span[0], t_span[1], 2)
    return solve_ivp(rhs_fn(p), t_span, y0, method=method, first_step=first_step, rtol=rtol,
                     atol=atol, t_eval=t_eval, events=event_fns(p, N) if events else None, **opts)


def jacobian_sparsity(N: int = 200):
    """parameters.py:150-199: 27 diagonals at k*N + {-1,0,1}, k=-4..4, minus d(CA,CC)/dPhi."""
    from scipy.sparse import lil_matrix, dia_matrix, csr_matrix
    n = 5 * N
    offsets = []
    for off in range(-n + N, n - N + 1, N):
        offsets += [off - 1, off, off + 1]
    raw = lil_matrix(dia_matrix((np.ones((len(offsets), n)), offsets), shape=(n, n)))
    raw[:2 * N, 4 * N:] = 0
    return csr_matrix(raw)


def brentq_restated(f, xa, xb, xtol=4 * np.finfo(float).eps, rtol=4 * np.finfo(float).eps, maxiter=100):
    """Restatement of `scipy.optimize.brentq` (scipy/optimize/Zeros/brentq.c, Brent's method with
    inverse quadratic extrapolation), the root finder `solve_ivp` uses to locate events
    (scipy/integrate/_ivp/ivp.py `solve_event_equation`, xtol = rtol = 4 eps).  The CUDA kernel's
    event location (csrc/rk45_persistent.cu `BrentState`) is a transliteration of THIS function;
    tests/test_oracle_golden.py pins it bit-for-bit (root, iterations, function calls) to the
    installed SciPy.  Returns (root, iterations, funcalls)."""
    xpre, xcur = float(xa), float(xb)
    xblk = fblk = spre = scur = 0.0
    fpre = f(xpre)
    fcur = f(xcur)
    funcalls = 2
    if fpre == 0:
        return xpre, 0, funcalls
    if fcur == 0:
        return xcur, 0, funcalls
    if math.copysign(1.0, fpre) == math.copysign(1.0, fcur):
        raise ValueError("f(a) and f(b) must have different signs")
    it = 0
    for _ in range(maxiter):
        it += 1
        if fpre != 0 and fcur != 0 and math.copysign(1.0, fpre) != math.copysign(1.0, fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, it, funcalls
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)                 # secant
            else:
                dpre = (fpre - fcur) / (xpre - xcur)                         # inverse quadratic
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = f(xcur)
        funcalls += 1
    return xcur, it, funcalls
