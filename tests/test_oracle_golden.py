"""The oracle is only trustworthy if it is pinned to the reference.  These CPU tests pin it to
(i) golden vectors produced by the reference's OWN pde_rhs/fun code (tests/golden/make_golden.py),
(ii) the reference's own regression fixtures and tolerances (tests/Regression_test/test_regression.py)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import lheureux_oracle as oracle
from conftest import rhs_states

np.seterr(all="ignore")


def test_default_scenario_matches_reference_map_scenario(scenario_reference):
    ref = scenario_reference["Map_Scenario"]
    mine = oracle.default_scenario()
    assert set(ref) == set(mine)
    for k, v in ref.items():
        assert mine[k] == v, k
    assert scenario_reference["Solver_first"]["method"] == "Radau"
    assert scenario_reference["Solver_first"]["first_step"] == 1e-6
    assert scenario_reference["Tracker"]["t_eval"] == [0.0, 1.0]


def test_jacobian_sparsity_structure(scenario_reference):
    js = oracle.jacobian_sparsity(200)
    ref = scenario_reference["jac_sparsity"]
    assert list(js.shape) == ref["shape"] and js.nnz == ref["nnz"] == 13799
    assert js[0].indices.tolist() == ref["row0_cols"]
    assert js[500].indices.tolist() == ref["row500_cols"]


@pytest.mark.parametrize("name", ["default", "scenario_A", "matlab", "fv_off", "exponents", "lattice_corner"])
def test_oracle_rhs_is_bit_identical_to_reference_pde_rhs(rhs_golden, name):
    g, meta = rhs_golden
    pde = meta[name]
    p = oracle.kernel_params(pde)
    derived = g[f"{name}/derived"]
    mine = [p[oracle.P_IDX[k]] for k in ("presum", "rhorat", "Da", "lambda_", "dCa", "dCO3", "delta", "KRat",
                                         "nu1", "nu2", "dPhi_fixed", "delta_x")]
    assert np.array_equal(np.array(mine), derived[:12])
    n_states = 0
    for sname, y, r_numba, r_numpy, ev in rhs_states(g, name):
        out = oracle.rhs(y, p, np.empty_like(y))
        assert np.array_equal(out, r_numba, equal_nan=True), sname          # numba backend: bit-for-bit
        S = oracle.term_scale(y, p, np.empty_like(y))
        assert np.nanmax(np.abs(out - r_numpy) / S) < 2e-15, sname          # numpy backend: re-association only
        evs = [f(0.0, y) for f in oracle.event_fns(p, pde["N"])]
        assert_allclose(evs, ev, rtol=0, atol=0)
        n_states += 1
    assert n_states >= 4


@pytest.mark.parametrize("name", ["default", "scenario_A", "fv_off"])
def test_oracle_time_varying_dPhi_is_bit_identical_to_reference_variant(rhs_golden_vardphi, name):
    """dPhi = auxcon F Phi^3 / (1 - Phi) per cell (LHeureux_model.py:222-223, :430-431, commented out
    upstream).  The golden outputs come from the reference's own source with exactly those two lines switched."""
    g, meta = rhs_golden_vardphi
    pde = meta[name] | {"time_varying_dPhi": True}
    p = oracle.kernel_params(pde)
    p_fixed = oracle.kernel_params(meta[name])
    n = 0
    for key in sorted(k for k in g.files if k.startswith(name + "/") and k.endswith("/y")):
        stem = key[:-2]
        y = g[key]
        out = oracle.rhs(y, p, np.empty_like(y))
        assert np.array_equal(out, g[stem + "/rhs_numba"], equal_nan=True), stem
        S = oracle.term_scale(y, p, np.empty_like(y))
        assert np.nanmax(np.abs(out - g[stem + "/rhs_numpy"]) / S) < 2e-15, stem
        n += 1
    assert n >= 6
    # the variant is a different model: same state, fixed coefficient -> different porosity rates
    y = g[f"{name}/noise/y"]
    fixed = oracle.rhs(y, p_fixed, np.empty_like(y))
    assert np.array_equal(fixed, g[f"{name}/noise/rhs_numba_fixed_dPhi"])
    assert np.max(np.abs(fixed[800:] - g[f"{name}/noise/rhs_numba"][800:])) > 1e-4 * np.max(np.abs(fixed[800:]))
    assert np.array_equal(fixed[:400], g[f"{name}/noise/rhs_numba"][:400])      # CA, CC do not see dPhi


def test_known_answer_values_default_y0():
    """Sanity values quoted in SURVEY.md §8c for the default scenario at y0."""
    pde = oracle.default_scenario()
    p = oracle.kernel_params(pde)
    assert p[oracle.P_IDX["presum"]] == pytest.approx(-3.2265364833646046, rel=1e-15)
    assert p[oracle.P_IDX["dPhi_fixed"]] == pytest.approx(0.004494236272775526, rel=1e-14)
    assert (int(p[oracle.P_IDX["mask_lo"]]), int(p[oracle.P_IDX["mask_hi"]])) == (20, 60)
    r = oracle.rhs(oracle.initial_state(pde), p, np.empty(1000)).reshape(5, 200)
    assert_allclose(r[:, 100], [106.43637819973569, -124.1757745663583, 1920.1337034124347,
                                1920.1337034124347, 35.47879273324522], rtol=1e-11)
    assert_allclose(r[:, 30], [-1920.087388900739, 1395.7170507589974, 56758.479179718976,
                               56758.479179718976, 1048.7406762834821], rtol=1e-11)


# ---- the reference's regression tests, run through the oracle (Radau, like upstream) ------------
def _regression_case(name, cases):
    c = cases[name]
    return oracle.default_scenario() | c["overrides"], c["first_step"]


def test_regression_scenario_A_radau(fixtures_reference, stepper_golden):
    """tests/Regression_test/test_regression.py:21-53 (rtol=0.1, atol=0.01), live SciPy Radau."""
    _, cases = stepper_golden
    pde, fs = _regression_case("scenario_A", cases)
    sol = oracle.integrate(pde, method="Radau", first_step=fs, jac_sparsity=oracle.jacobian_sparsity(200))
    assert sol.status == 0
    assert_allclose(sol.y[:, -1].reshape(5, 200), fixtures_reference["scenario_A"][-1], rtol=0.1, atol=0.01)


def test_regression_matlab_radau(fixtures_reference, stepper_golden):
    """test_regression.py:90-148: Matlab output interpolated to the cell centres, atol=0.05, cells 2.."""
    _, cases = stepper_golden
    pde, fs = _regression_case("matlab", cases)
    sol = oracle.integrate(pde, method="Radau", first_step=fs, jac_sparsity=oracle.jacobian_sparsity(200))
    last = sol.y[:, -1].reshape(5, 200)
    m = fixtures_reference["matlab"]
    x, _ = oracle.grid_coords(pde)
    depths = x * pde["Xstar"]
    interp = np.stack([np.interp(depths, np.linspace(0, 500, m.shape[1]), m[f, :, 0]) for f in range(5)])
    assert_allclose(last[:, 2:], interp[:, 2:], atol=0.05)


def test_committed_stepper_goldens_pass_reference_tolerances(fixtures_reference, stepper_golden):
    """RK45 and Radau end states stored by make_stepper_golden.py, against the reference's fixtures."""
    g, _ = stepper_golden
    for case, fix in (("scenario_A", "scenario_A"), ("high_porosity", "high_porosity")):
        for method in ("RK45", "Radau"):
            y = g[f"{case}/{method}/t1/tol0.001/y"][:, -1].reshape(5, 200)
            assert_allclose(y, fixtures_reference[fix][-1], rtol=0.1, atol=0.01)
            assert g[f"{case}/{method}/t1/tol0.001/counts"][3] == 0
    # tight-tolerance Radau pins grid, ghost rules, stencils and rate terms to the fixture at ~1e-7
    tight = g["scenario_A/Radau/t1/tol1e-08/y"].reshape(5, 200, -1)
    fixA = fixtures_reference["scenario_A"]
    idx = fixtures_reference["snapshot_index"]
    # per-field bounds: CA, CC, cCa, Phi agree to <1e-6; cCO3 carries the fixture's own time-stepping
    # error in the top boundary-layer cells (the fixture came from py-pde's adaptive stepper)
    bound = np.array([1e-7, 1e-7, 1e-6, 2e-4, 1e-7])[:, None]
    assert np.all(np.abs(tight[:, :, -1] - fixA[-1]) < bound)
    assert np.abs(tight[3, 8:, -1] - fixA[-1][3, 8:]).max() < 1e-7     # cCO3 below the boundary layer
    # intermediate snapshots t = 0.1 ... 0.9 (tight run sampled on linspace(0,1,11))
    for j in range(1, 11):
        k = int(np.where(idx == 10 * j)[0][0])
        assert np.all(np.abs(tight[:, :, j] - fixA[k]) < bound), j
    tightB = g["high_porosity/Radau/t1/tol1e-06/y"].reshape(5, 200, -1)
    fixB = fixtures_reference["high_porosity"]
    for j in (1, 3, 5, 10):
        k = int(np.where(idx == 10 * j)[0][0])
        assert np.abs(tightB[:, :, j] - fixB[k]).max() < 5e-4, j


def test_live_scipy_rk45_reproduces_committed_short_golden(stepper_golden):
    g, cases = stepper_golden
    pde, fs = _regression_case("scenario_A", cases)
    key = "scenario_A/RK45/t0.002/tol0.001"
    sol = oracle.integrate(pde, method="RK45", first_step=fs, t_span=(0, 0.002), t_eval=g[key + "/t"])
    assert sol.nfev == g[key + "/counts"][0]
    assert_allclose(sol.y, g[key + "/y"], rtol=0, atol=1e-12)


def test_brentq_restatement_is_bit_identical_to_scipy():
    """The event locator restated in the oracle (and transliterated in the kernel) against the
    installed scipy.optimize.brentq: same root, iteration count and function calls."""
    import math
    from scipy.optimize import brentq
    rng = np.random.default_rng(0)
    eps4 = 4 * np.finfo(float).eps
    checked = 0
    for trial in range(1200):
        a, b = sorted(rng.uniform(-2, 2, 2))
        r0 = rng.uniform(a, b)
        f = [lambda x: (x - r0) * (1 + 0.3 * np.sin(5 * x)), lambda x: math.expm1(3 * (x - r0)),
             lambda x: (x - r0) ** 3 + 1e-3 * (x - r0), lambda x: min(x - r0, 0.3 * (x - r0))][trial % 4]
        if f(a) * f(b) > 0:
            continue
        root, res = brentq(f, a, b, xtol=eps4, rtol=eps4, full_output=True)
        assert oracle.brentq_restated(f, a, b) == (root, res.iterations, res.function_calls)
        checked += 1
    assert checked > 1000
    assert oracle.brentq_restated(lambda x: x, 0.0, 1.0)[0] == 0.0          # f(a) == 0 returns a
    with pytest.raises(ValueError):
        oracle.brentq_restated(lambda x: x * x + 1, -1.0, 1.0)


def test_lsoda_runs_in_bdf_mode():
    """What `method="LSODA"` (parameters.py:214-219) does on this system, measured on the oracle RHS with ODEPACK's own
    method log (odeint `mused`: 1 = Adams, 2 = BDF): it leaves the Adams mode within the first few dozen steps (t ~ 2e-5 T*)
    and never returns — from then on LSODA is a variable-order BDF code, which is what the device path for LSODA batches
    runs (csrc/bdf_batch.cu).  Both scenario bases, the reference's lband = uband = 1."""
    from scipy.integrate import odeint
    for over in ({"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}, {}):
        pde = oracle.default_scenario() | over
        f = oracle.rhs_fn(oracle.kernel_params(pde))
        ts = np.concatenate([[0.0], np.geomspace(1e-6, 0.05, 60)])
        _, info = odeint(lambda y, t: f(t, y), oracle.initial_state(pde), ts, rtol=1e-3, atol=1e-3, h0=1e-6, ml=1, mu=1,
                         full_output=True, mxstep=500000)
        mused, nst = info["mused"], info["nst"]
        first_bdf = int(np.argmax(mused == 2))
        assert mused[first_bdf] == 2 and np.all(mused[first_bdf:] == 2)          # switches once, stays
        assert ts[1 + first_bdf] <= 1e-4 and nst[first_bdf] <= 100               # ... within the first steps
        assert nst[-1] >= 20 * nst[first_bdf]                                    # already at t = 0.05 T*: > 95 % BDF steps
        # (to T*, scenario A: 24 Adams steps of 15 059)
