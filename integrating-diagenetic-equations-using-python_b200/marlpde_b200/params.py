"""Per-column kernel constants from the reference's parameter dictionaries.

Mirrors LMAHeureuxPorosityDiff.__init__ (marlpde/LHeureux_model.py:31-72, :87-88, :130-133),
the grid of Evolve_scenario.py:40 and the Heaviside masks of Evolve_scenario.py:51-54, with the
same floating-point expression order, vectorised over columns: every value of `pde` may be a
scalar or an array of shape (B,) (a parameter sweep).
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

from ._cabi import MODEL_VAR_DPHI, PARAMS_DTYPE

_REQUIRED = ("CA0", "CC0", "cCa0", "cCO30", "Phi0", "sedimentationrate", "Xstar", "Tstar", "k1", "k2",
             "k3", "k4", "m1", "m2", "n1", "n2", "b", "beta", "rhos", "rhow", "rhos0", "KA", "KC",
             "muA", "D0Ca", "PhiNR", "PhiInfty", "PhiIni", "DCa", "DCO3", "FV_switch", "max_depth",
             "ShallowLimit", "DeepLimit", "N")


def _n_columns(pde: Mapping) -> int:
    n = 1
    for k in _REQUIRED:
        v = np.asarray(pde[k])
        if v.ndim > 1:
            raise ValueError(f"parameter {k!r} must be a scalar or 1-D")
        if v.ndim == 1:
            if n not in (1, v.shape[0]):
                raise ValueError("inconsistent sweep lengths")
            n = v.shape[0]
    return n


def derive_column_params(pde: Mapping) -> np.ndarray:
    """Return a structured array (B,) laid out as `marlpde_column_params`."""
    missing = [k for k in _REQUIRED if k not in pde]
    if missing:
        raise KeyError(f"missing scenario parameters: {missing}")
    N = pde["N"]
    if np.ndim(N) != 0:
        raise ValueError("all columns of one batch share the same N")
    N = int(N)
    B = _n_columns(pde)
    g = lambda k: np.asarray(pde[k], dtype=np.float64)  # noqa: E731
    for k in ("m1", "m2", "n1", "n2"):
        if np.any(g(k) <= 0):
            raise ValueError(f"reaction exponent {k} must be positive")

    grav = 100 * 9.81                                                   # :59
    nu1 = g("k1") / g("k2")                                             # :36
    nu2 = g("k4") / g("k3")                                             # :39
    KRat = g("KC") / g("KA")                                            # :51
    dCa = g("DCa") / g("D0Ca")                                          # :60
    dCO3 = g("DCO3") / g("D0Ca")                                        # :61
    delta = g("rhos") / (g("muA") * np.sqrt(g("KC")))                   # :62
    Da = g("k2") * g("Tstar")                                           # :63
    lambda_ = g("k3") / g("k2")                                         # :64
    auxcon = g("beta") / (g("D0Ca") * g("b") * grav * g("rhow") * (g("PhiNR") - g("PhiInfty")))  # :65-66
    rhorat0 = (g("rhos0") / g("rhow") - 1) * g("beta") / g("sedimentationrate")  # :67-68
    rhorat = (g("rhos") / g("rhow") - 1) * g("beta") / g("sedimentationrate")    # :69-70
    Phi0 = g("Phi0")
    presum = 1 - rhorat0 * Phi0 ** 3 * (1 - np.exp(10 - 10 / Phi0)) / (1 - Phi0)  # :71-72
    PhiIni = g("PhiIni")
    F_fixed = 1 - np.exp(10 - 10 / PhiIni)                              # :131
    dPhi_fixed = auxcon * F_fixed * PhiIni ** 3 / (1 - PhiIni)          # :132-133

    # grid (Evolve_scenario.py:40; py-pde cell-centred CartesianGrid) and masks (:51-54)
    length = np.broadcast_to(g("max_depth") / g("Xstar"), (B,))
    dx = (length - 0.0) / N
    x = 0.0 + (np.arange(N)[None, :] + 0.5) * dx[:, None]
    delta_x = x[:, 1] - x[:, 0]                                         # LHeureux_model.py:23-24
    shallow = np.broadcast_to(g("ShallowLimit") / g("Xstar"), (B,))[:, None]
    deep = np.broadcast_to(g("DeepLimit") / g("Xstar"), (B,))[:, None]
    mask = ((x - shallow) > 0) & ((deep - x) > 0)                       # heaviside(., 0)
    any_ = mask.any(axis=1)
    lo = np.where(any_, mask.argmax(axis=1), 0)
    hi = np.where(any_, N - mask[:, ::-1].argmax(axis=1), 0)
    if np.any(mask.sum(axis=1) != hi - lo):
        raise ValueError("dissolution mask is not one contiguous interval of cells")

    out = np.zeros(B, dtype=PARAMS_DTYPE)
    for f, k in enumerate(("CA0", "CC0", "cCa0", "cCO30", "Phi0")):
        out["bc_top"][:, f] = g(k)
    out["dx"] = dx
    out["inv_dx"] = 1.0 / dx
    out["inv_dx2"] = dx ** -2
    out["delta_x"] = delta_x
    for name, val in (("presum", presum), ("rhorat", rhorat), ("Da", Da), ("lambda_", lambda_),
                      ("dCa", dCa), ("dCO3", dCO3), ("delta", delta), ("KRat", KRat), ("nu1", nu1),
                      ("nu2", nu2), ("m1", g("m1")), ("m2", g("m2")), ("n1", g("n1")), ("n2", g("n2")),
                      ("dPhi_fixed", dPhi_fixed)):
        out[name] = val
    out["Peclet_min"] = 1e-2                                            # :87
    out["Peclet_max"] = 1 / 1e-2                                        # :88
    out["FV_switch"] = np.asarray(pde["FV_switch"]).astype(np.int32) != 0
    out["mask_lo"] = lo
    out["mask_hi"] = hi
    # model variant upstream toggles by editing its source (:222-223, :430-431): dPhi = auxcon F Phi^3 / (1 - Phi)
    out["auxcon"] = auxcon
    out["model_flags"] = np.where(np.asarray(pde.get("time_varying_dPhi", False)).astype(bool), MODEL_VAR_DPHI, 0)
    return out


def initial_state(pde: Mapping) -> np.ndarray:
    """y0[B, 5, N]: each field uniform at its *Ini value (Evolve_scenario.py:76-86)."""
    N = int(pde["N"])
    B = _n_columns(pde)
    y0 = np.empty((B, 5, N), dtype=np.float64)
    for f, k in enumerate(("CAIni", "CCIni", "cCaIni", "cCO3Ini", "PhiIni")):
        y0[:, f, :] = np.broadcast_to(np.asarray(pde[k], dtype=np.float64), (B,))[:, None]
    return y0


def sweep_lattice(base: Mapping, n_S: int = 16, n_b: int = 16, n_D: int = 16,
                  S=(0.09, 0.11), b=(4.0, 6.0), D0co3=(245.0, 300.0)) -> dict:
    """The synthetic Map_Scenario parameter sweep of BASELINE.json configs[1]/[2]
    (SURVEY.md §8d): a deterministic lattice over sedimentation rate S [cm/a], compaction
    coefficient b [1/kPa] and the CO3 diffusion coefficient [cm2/a]; column index
    c = (i*n_b + j)*n_D + k.  Derived values are recomputed per column exactly as
    Map_Scenario.__post_init__ does (marlpde/parameters.py:120-143): Xstar = D0Ca/S,
    Tstar = Xstar/S, b -> b/1e4.  Every other value comes from `base`.
    """
    i, j, k = np.meshgrid(np.arange(n_S), np.arange(n_b), np.arange(n_D), indexing="ij")
    frac = lambda idx, n: idx.ravel() / max(n - 1, 1)  # noqa: E731
    Sv = S[0] + (S[1] - S[0]) * frac(i, n_S)
    bv = b[0] + (b[1] - b[0]) * frac(j, n_b)
    Dv = D0co3[0] + (D0co3[1] - D0co3[0]) * frac(k, n_D)
    pde = dict(base)
    pde["sedimentationrate"] = Sv
    pde["b"] = bv / 1e4
    pde["DCO3"] = Dv
    pde["Xstar"] = pde["D0Ca"] / Sv
    pde["Tstar"] = pde["Xstar"] / Sv
    return pde
