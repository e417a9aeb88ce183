"""GPU parity of the batched variable-order BDF integrator (marlpde_bdf_integrate through the C ABI) against SciPy's BDF
on the CPU oracle — `solve_ivp(method="BDF", jac_sparsity=...)`, one of the reference's implicit solvers
(parameters.py:213-216, :235-236; call site Evolve_scenario.py:104-109) and the algorithm LSODA (:214-219) runs in its
stiff mode.

The kernel restates scipy/integrate/_ivp/bdf.py.  Given the SAME Jacobian structure — the block-tridiagonal pattern
instead of the reference's 27 diagonals, which drop d(CA,CC)/dPhi — SciPy's BDF takes the same steps: the gate is
step-for-step (Jacobian / LU / Newton-iteration counts equal, up to a decision flipped by fp64 rounding) and a small
fraction of the tolerance on the states.  Against SciPy with the reference's own (incomplete) pattern the two codes are
two BDF integrations at rtol = 1e-3: compared through their distance to a tight-tolerance solution."""
import numpy as np
import pytest

import lheureux_oracle as oracle
import marlpde_b200 as mb

pytestmark = pytest.mark.gpu
np.seterr(all="ignore")
SCEN_A = {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}


def exact_sparsity(N):
    import scipy.sparse as sp
    cells = np.arange(5 * N) % N
    return sp.csr_matrix((np.abs(cells[:, None] - cells[None, :]) <= 1).astype(float))


def _scipy_bdf(pde, t_end, t_eval, tol=1e-3, first_step=1e-6, events=False, sparsity=None):
    N = int(pde["N"])
    return oracle.integrate(pde, method="BDF", t_span=(0, t_end), t_eval=t_eval, events=events, first_step=first_step,
                            rtol=tol, atol=tol, jac_sparsity=exact_sparsity(N) if sparsity is None else sparsity)


@pytest.mark.parametrize("tol,t_end,jac", [(1e-3, 1.0, "fd"), (1e-3, 1.0, "analytic"), (1e-6, 0.2, "fd")])
def test_scenario_A_step_for_step_with_scipy_bdf(tol, t_end, jac):
    """Both Jacobians of the kernel: finite-difference diagonal blocks (default; SciPy's num_jac rule) and all-analytic."""
    pde = oracle.default_scenario() | SCEN_A
    te = np.linspace(0, t_end, 6)
    res = mb.integrate_bdf_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, t_end), first_step=1e-6,
                                 rtol=tol, atol=tol, t_eval=te, jac=jac)
    sol = _scipy_bdf(pde, t_end, te, tol)
    assert res.status[0] == 0 and res.t[0] == t_end and res.next_eval[0] == te.size
    assert abs(int(res.njev[0]) - sol.njev) <= 1 and abs(int(res.nlu[0]) - sol.nlu) <= 2
    assert abs(int(res.newton_iterations[0]) - (sol.nfev - 1)) <= 0.02 * sol.nfev
    got, want = res.solutions(0), sol.y.reshape(5, 200, -1)
    assert np.max(np.abs(got - want) / (tol + tol * np.abs(want))) <= 0.05
    assert np.max(np.abs(got[:, :, 0] - mb.initial_state(pde)[0])) <= 1e-12    # t_eval[0] = t0: dense output of step 1, as SciPy


def test_against_scipy_bdf_with_the_reference_sparsity():
    """The reference hands SciPy its 27-diagonal pattern (parameters.py:150-199): a slightly different Newton matrix, so a
    slightly different step sequence.  Both integrations sit ~15 tolerance units from the converged solution at
    rtol = 1e-3 (SciPy BDF's own accuracy on this problem); the kernel must not be further away than SciPy is."""
    pde = oracle.default_scenario() | SCEN_A
    te = np.linspace(0, 1, 6)
    res = mb.integrate_bdf_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 1), t_eval=te)
    sol = _scipy_bdf(pde, 1.0, te, sparsity=oracle.jacobian_sparsity(200))
    ref = oracle.integrate(pde, method="Radau", t_span=(0, 1), t_eval=te, events=False, rtol=1e-8, atol=1e-8,
                           jac_sparsity=oracle.jacobian_sparsity(200)).y.reshape(5, 200, -1)
    unit = 1e-3 + 1e-3 * np.abs(ref)
    err_gpu = np.max(np.abs(res.solutions(0) - ref) / unit)
    err_scipy = np.max(np.abs(sol.y.reshape(5, 200, -1) - ref) / unit)
    assert err_gpu <= 1.25 * err_scipy, (err_gpu, err_scipy)
    assert 0.8 * sol.nlu <= res.nlu[0] <= 1.25 * sol.nlu and abs(int(res.njev[0]) - sol.njev) <= 3


def test_default_scenario_events_match_scipy_bdf():
    """Default Map_Scenario up to the porosity excursion (Phi crosses 1 near t = 0.026, then max W changes sign): event
    detection on the sign classes, Brent on the BDF dense output — against SciPy BDF's t_events.  Monitoring does not
    change the trajectory."""
    pde = oracle.default_scenario()
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    res = mb.integrate_bdf_batch(y0, P, t_span=(0, 0.0275), first_step=5e-7, t_eval=[0, 0.0275], events=True,
                                 event_capacity=64)
    sol = _scipy_bdf(pde, 0.0275, [0, 0.0275], first_step=5e-7, events=True)
    assert res.status[0] == 0
    assert [len(e) for e in sol.t_events] == list(res.event_counts[0])
    for k in (4, 6):
        assert len(sol.t_events[k]) >= 1
        assert np.max(np.abs(np.sort(res.event_times[0, k, :len(sol.t_events[k])]) - sol.t_events[k])) <= 1e-6
    want = sol.y.reshape(5, 200, -1)[:, :, -1]
    assert np.max(np.abs(res.y[0] - want) / (1e-3 + 1e-3 * np.abs(want))) <= 0.05
    plain = mb.integrate_bdf_batch(y0, P, t_span=(0, 0.0275), first_step=5e-7)
    assert np.array_equal(plain.y, res.y)


def test_lattice_columns_independent():
    base = oracle.default_scenario() | SCEN_A
    pde = mb.sweep_lattice(base, 2, 2, 2)
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    res = mb.integrate_bdf_batch(y0, P, t_span=(0, 0.2), first_step=1e-6, rtol=1e-6, atol=1e-6, t_eval=[0.0, 0.1, 0.2])
    assert np.all(res.status == 0) and np.all(res.next_eval == 3)
    for c in (0, 5, 7):
        one = {k: (float(v[c]) if np.ndim(v) else v) for k, v in pde.items()}
        sol = _scipy_bdf(one, 0.2, [0.0, 0.1, 0.2], tol=1e-6)
        want = sol.y.reshape(5, 200, -1)
        assert np.max(np.abs(res.solutions(c) - want) / (1e-6 + 1e-6 * np.abs(want))) <= 0.05
        assert abs(int(res.nlu[c]) - sol.nlu) <= 2
    single = mb.integrate_bdf_batch(y0[5:6], P[5:6], t_span=(0, 0.2), first_step=1e-6, rtol=1e-6, atol=1e-6)
    assert np.array_equal(single.y[0], res.y[5])                       # a column does not depend on its neighbours


def test_grid_sizes_budget_and_device_path():
    import torch
    for n_cells in (3, 37, 500):
        pde = oracle.default_scenario() | {"N": n_cells} | SCEN_A
        res = mb.integrate_bdf_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 0.05),
                                     first_step=1e-6, t_eval=[0.05])
        sol = _scipy_bdf(pde, 0.05, [0.05])
        assert res.status[0] == 0
        want = sol.y.reshape(5, n_cells, -1)
        assert np.max(np.abs(res.solutions(0) - want) / (1e-3 + 1e-3 * np.abs(want))) <= 0.05, n_cells
    pde = oracle.default_scenario() | SCEN_A
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    part = mb.integrate_bdf_batch(y0, P, t_span=(0, 1), first_step=1e-6, max_steps=30)
    assert part.status[0] == 1 and part.n_accepted[0] == 30 and 0 < part.t[0] < 1
    rest = mb.integrate_bdf_batch(part.y, P, t_span=(0, 1), state=part.state)     # resumes at order 1
    whole = mb.integrate_bdf_batch(y0, P, t_span=(0, 1), first_step=1e-6)
    ref = oracle.integrate(pde, method="Radau", t_span=(0, 1), events=False, rtol=1e-8, atol=1e-8,
                           jac_sparsity=oracle.jacobian_sparsity(200)).y[:, -1].reshape(5, 200)
    assert rest.status[0] == 0 and rest.t[0] == 1.0
    unit = 1e-3 + 1e-3 * np.abs(ref)
    assert np.max(np.abs(rest.y[0] - ref) / unit) <= 1.5 * np.max(np.abs(whole.y[0] - ref) / unit) + 2.0
    dev = mb.integrate_bdf_batch(torch.from_numpy(y0).cuda(), P, t_span=(0, 1), first_step=1e-6)
    assert np.array_equal(dev.y.cpu().numpy(), whole.y)
    assert mb.integrate_bdf_batch(y0[:0], P[:0]).y.shape == (0, 5, 200)


def test_bdf_time_varying_dPhi_model_variant():
    base = oracle.default_scenario() | SCEN_A
    var = base | {"time_varying_dPhi": True}
    P = np.concatenate([mb.derive_column_params(var), mb.derive_column_params(base)])
    Y = np.repeat(mb.initial_state(base), 2, 0)
    res = mb.integrate_bdf_batch(Y, P, t_span=(0, 0.2), first_step=1e-6, t_eval=[0.2])
    plain = mb.integrate_bdf_batch(Y[1:], P[1:], t_span=(0, 0.2), first_step=1e-6, t_eval=[0.2])
    assert np.all(res.status == 0) and np.array_equal(res.y[1], plain.y[0])
    sol = _scipy_bdf(var, 0.2, [0.2])
    want = sol.y[:, -1].reshape(5, 200)
    assert np.max(np.abs(res.y[0] - want) / (1e-3 + 1e-3 * np.abs(want))) <= 0.05
    assert np.max(np.abs(res.y[0][4] - res.y[1][4])) > 1e-6


def test_dropin_bdf_and_lsoda_device_routes(tmp_path, monkeypatch):
    """integrate_equations(method="BDF") runs on the BDF kernel; method="LSODA" is SciPy's own stepper over the CUDA RHS
    unless MARLPDE_LSODA_DEVICE=1, which runs the column on the BDF kernel (`lband` / `uband` accepted and ignored).
    Gates: the device BDF against SciPy BDF on the oracle (end state), and every route within the reference's regression
    tolerance band (test_regression.py:29-30, scaled by BDF's own accuracy: SciPy BDF itself sits at 1.5x that band on
    this case) of a tight-tolerance solution."""
    import os
    from dataclasses import asdict
    from marlpde.Evolve_scenario import integrate_equations
    from marlpde.parameters import Map_Scenario, Solver, Tracker
    (tmp_path / "run").mkdir()
    monkeypatch.chdir(tmp_path / "run")
    pde = asdict(Map_Scenario()) | SCEN_A
    ref = oracle.integrate(oracle.default_scenario() | SCEN_A, method="Radau", t_span=(0, 1), events=False, rtol=1e-8,
                           atol=1e-8, jac_sparsity=oracle.jacobian_sparsity(200)).y[:, -1].reshape(5, 200)
    sol = _scipy_bdf(oracle.default_scenario() | SCEN_A, 1.0, [0, 1.0]).y[:, -1].reshape(5, 200)
    last, covered, _, _, folder = integrate_equations(asdict(Solver(method="BDF")), asdict(Tracker()), dict(pde))
    assert covered == pde["Tstar"] and os.path.exists(os.path.join(folder, "LMAHeureuxPorosityDiff.hdf5"))
    assert np.max(np.abs(last - sol) / (1e-3 + 1e-3 * np.abs(sol))) <= 0.05
    monkeypatch.setenv("MARLPDE_LSODA_DEVICE", "1")
    last2, covered2, *_ = integrate_equations(asdict(Solver(method="LSODA")), asdict(Tracker()), dict(pde))
    assert covered2 == pde["Tstar"] and np.array_equal(last2, last)
    assert np.max(np.abs(last - ref) / (0.01 + 0.1 * np.abs(ref))) <= 2.0


def test_predictor_outside_the_model_domain_does_not_poison_the_step():
    """Lattice column 1305 (default base): near t = 0.165 the order-3 predictor extrapolates the porosity of one cell below
    zero (log(Phi) = NaN).  bdf.py evaluates the Jacobian AT that predictor and keeps it for the rest of the step, so every
    smaller step size fails too and the integration ends with "step size too small"; the kernel halves the step size
    first (csrc/bdf_batch.cu, the one deliberate difference from bdf.py).  The column — and its neighbours that stalled
    the same way in r02s — must get through the stiff phase."""
    from dataclasses import asdict
    from marlpde.parameters import Map_Scenario
    lat = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
    cols = np.array([1305, 1360, 3932])
    for jac in ("analytic", "fd"):
        res = mb.integrate_bdf_batch(mb.initial_state(lat)[cols], mb.derive_column_params(lat)[cols], t_span=(0, 0.3),
                                     first_step=1e-6, jac=jac)
        assert np.all(res.status == 0) and np.all(res.t == 0.3), (jac, res.status, res.t)
