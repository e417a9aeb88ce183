"""GPU parity of the batched implicit integrator (marlpde_radau_integrate through the C ABI) against
SciPy's Radau on the CPU oracle with the reference's Jacobian sparsity — the reference's DEFAULT
solver (parameters.py:201-221; call site Evolve_scenario.py:104-109).

Two correct Radau IIA implementations with different linear algebra (block-tridiagonal vs SuperLU,
15- vs 21-colour finite-difference Jacobian) do not take bit-identical step sequences; the gate is the
solver's own tolerance, |gpu - scipy| <= atol + rtol*|y| at every output time,
plus the reference's regression tolerances against its HDF5 fixtures."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import lheureux_oracle as oracle
import marlpde_b200 as mb

pytestmark = pytest.mark.gpu
np.seterr(all="ignore")
RTOL = ATOL = 1e-3


def _scipy_radau(pde, t_end, t_eval, first_step=1e-6):
    return oracle.integrate(pde, method="Radau", t_span=(0, t_end), t_eval=t_eval, events=False, first_step=first_step,
                            jac_sparsity=oracle.jacobian_sparsity(int(pde["N"])))


def test_scenario_A_to_Tstar_matches_scipy_and_fixture(fixtures_reference):
    pde = oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    te = np.linspace(0, 1, 11)
    res = mb.integrate_radau_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 1), first_step=1e-6,
                                   rtol=RTOL, atol=ATOL, t_eval=te)
    sol = _scipy_radau(pde, 1.0, te)
    assert res.status[0] == 0 and res.t[0] == 1.0 and res.next_eval[0] == te.size
    got, want = res.solutions(0), sol.y.reshape(5, 200, -1)
    assert np.all(np.abs(got - want) <= ATOL + RTOL * np.abs(want))
    assert np.array_equal(got[:, :, 0], mb.initial_state(pde)[0])
    # same order of work as SciPy (41 steps, 76 LU there)
    assert 25 <= res.n_accepted[0] <= 80 and res.nlu[0] <= 300 and res.njev[0] <= 60
    # the reference's regression test (test_regression.py:21-53): rtol=0.1, atol=0.01 against its fixture
    assert_allclose(res.y[0], fixtures_reference["scenario_A"][-1], rtol=0.1, atol=0.01)


def test_default_scenario_short_horizon_with_porosity_excursion():
    """Default Map_Scenario: Phi transiently exceeds 1 near t = 0.027 and W changes sign — the stiff,
    step-rejecting part of the trajectory.  There two correct solvers at rtol = 1e-3 sit several tolerance
    units away from the converged solution (SciPy itself: 3.7 units at t = 0.03), so the gate is the
    distance to a tight-tolerance reference, compared with SciPy's own, plus the work counters."""
    pde = oracle.default_scenario()
    te = np.array([0.0, 0.01, 0.02, 0.03, 0.05])
    res = mb.integrate_radau_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 0.05),
                                   first_step=5e-7, rtol=RTOL, atol=ATOL, t_eval=te)
    sol = _scipy_radau(pde, 0.05, te, first_step=5e-7)
    ref = oracle.integrate(pde, method="Radau", t_span=(0, 0.05), t_eval=te, events=False, first_step=5e-7, rtol=1e-8,
                           atol=1e-8, jac_sparsity=oracle.jacobian_sparsity(200)).y.reshape(5, 200, -1)
    assert res.status[0] == 0
    got, want = res.solutions(0), sol.y.reshape(5, 200, -1)
    unit = ATOL + RTOL * np.abs(ref)
    err_gpu, err_scipy = np.max(np.abs(got - ref) / unit), np.max(np.abs(want - ref) / unit)
    assert err_gpu <= max(8.0, 2.0 * err_scipy), (err_gpu, err_scipy)
    assert np.max(np.abs(got[:, :, :3] - want[:, :, :3]) / unit[:, :, :3]) <= 1.0        # before the excursion
    # same amount of work as SciPy (there: 774 LU, 152 Jacobians)
    assert 0.5 * sol.nlu <= res.nlu[0] <= 2 * sol.nlu and 0.5 * sol.njev <= res.njev[0] <= 2 * sol.njev


def test_lattice_columns_independent_and_tight_tolerance():
    """Columns with different parameters in one launch; at rtol=atol=1e-6 GPU and SciPy agree to ~1e-5."""
    base = oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    pde = mb.sweep_lattice(base, 2, 2, 2)
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    res = mb.integrate_radau_batch(y0, P, t_span=(0, 0.2), first_step=1e-6, rtol=1e-6, atol=1e-6, t_eval=[0.0, 0.1, 0.2])
    assert np.all(res.status == 0) and np.all(res.next_eval == 3)
    for c in (0, 5, 7):
        one = {k: (float(v[c]) if np.ndim(v) else v) for k, v in pde.items()}
        sol = oracle.integrate(one, method="Radau", t_span=(0, 0.2), t_eval=[0.0, 0.1, 0.2], events=False, rtol=1e-6,
                               atol=1e-6, jac_sparsity=oracle.jacobian_sparsity(200))
        assert_allclose(res.solutions(c), sol.y.reshape(5, 200, -1), rtol=0, atol=2e-5)
    single = mb.integrate_radau_batch(y0[5:6], P[5:6], t_span=(0, 0.2), first_step=1e-6, rtol=1e-6, atol=1e-6)
    assert np.array_equal(single.y[0], res.y[5])                       # a column does not depend on its neighbours


def test_grid_sizes_budget_and_device_path():
    import torch
    for n_cells in (3, 37, 500):
        pde = oracle.default_scenario() | {"N": n_cells, "Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
        res = mb.integrate_radau_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 0.05),
                                       first_step=1e-6, t_eval=[0.05])
        sol = oracle.integrate(pde, method="Radau", t_span=(0, 0.05), t_eval=[0.05], events=False)
        assert res.status[0] == 0
        want = sol.y.reshape(5, n_cells, -1)
        assert np.all(np.abs(res.solutions(0) - want) <= 2 * (ATOL + RTOL * np.abs(want))), n_cells
    pde = oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    part = mb.integrate_radau_batch(y0, P, t_span=(0, 1), first_step=1e-6, max_steps=10)
    assert part.status[0] == 1 and part.n_accepted[0] == 10 and 0 < part.t[0] < 1
    rest = mb.integrate_radau_batch(part.y, P, t_span=(0, 1), state=part.state)
    whole = mb.integrate_radau_batch(y0, P, t_span=(0, 1), first_step=1e-6)
    assert rest.status[0] == 0 and rest.t[0] == 1.0
    assert np.all(np.abs(rest.y - whole.y) <= 2 * (ATOL + RTOL * np.abs(whole.y)))
    dev = mb.integrate_radau_batch(torch.from_numpy(y0).cuda(), P, t_span=(0, 1), first_step=1e-6)
    assert np.array_equal(dev.y.cpu().numpy(), whole.y)
    assert mb.integrate_radau_batch(y0[:0], P[:0]).y.shape == (0, 5, 200)


def test_radau_events_match_scipy():
    """Event monitors on the implicit path: same monitors, same sign-change rule, Brent on the cubic dense
    output.  Default Map_Scenario: the porosity crosses 1 near t = 0.027 (event 4) and max W changes sign
    right after (event 6) — crossings of the solution, not of round-off, so their first occurrence must
    agree with SciPy Radau to the accuracy of two rtol = 1e-3 trajectories.  (Scenario A's events 0/1 are not
    used here: CA decays to 1e-17 and the sign of min(CA) is noise — SciPy reports 0.2488 at rtol 1e-3 and
    0.2939 at 1e-8.)"""
    pde = oracle.default_scenario()
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    res = mb.integrate_radau_batch(y0, P, t_span=(0, 0.05), first_step=5e-7, t_eval=[0, 0.05], events=True,
                                   event_capacity=64)
    sol = oracle.integrate(pde, method="Radau", t_span=(0, 0.05), t_eval=[0, 0.05], events=True, first_step=5e-7,
                           jac_sparsity=oracle.jacobian_sparsity(200))
    got = [np.sort(res.event_times[0, k, :min(int(res.event_counts[0, k]), 64)]) for k in range(7)]
    for k in (0, 1, 2, 3, 5):
        assert res.event_counts[0, k] == 0 == len(sol.t_events[k])
    assert len(sol.t_events[4]) >= 1 and len(got[4]) >= 1 and abs(got[4][0] - sol.t_events[4][0]) <= 5e-4
    assert len(sol.t_events[6]) >= 1 and len(got[6]) >= 1 and abs(got[6][0] - sol.t_events[6][0]) <= 1e-3
    assert np.all(np.diff(got[6]) > 0) and np.all((got[6] > 0) & (got[6] <= 0.05))
    plain = mb.integrate_radau_batch(y0, P, t_span=(0, 0.05), first_step=5e-7)
    assert np.array_equal(plain.y, res.y)                              # monitoring does not change the trajectory


def test_radau_time_varying_dPhi_model_variant():
    """MARLPDE_MODEL_VAR_DPHI on the implicit path (per-cell dPhi in the RHS and in the analytic off-diagonal Jacobian
    blocks) against SciPy Radau on the oracle variant; a plain column in the same launch is unaffected."""
    base = oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    var = base | {"time_varying_dPhi": True}
    P = np.concatenate([mb.derive_column_params(var), mb.derive_column_params(base)])
    Y = np.repeat(mb.initial_state(base), 2, 0)
    res = mb.integrate_radau_batch(Y, P, t_span=(0, 0.2), first_step=1e-6, t_eval=[0.2])
    plain = mb.integrate_radau_batch(Y[1:], P[1:], t_span=(0, 0.2), first_step=1e-6, t_eval=[0.2])
    assert np.all(res.status == 0) and np.array_equal(res.y[1], plain.y[0])
    sol = oracle.integrate(var, method="Radau", t_span=(0, 0.2), t_eval=[0.2], events=False, first_step=1e-6,
                           jac_sparsity=oracle.jacobian_sparsity(200))
    want = sol.y[:, -1].reshape(5, 200)
    assert np.max(np.abs(res.y[0] - want) / (ATOL + RTOL * np.abs(want))) <= 2.0
    assert np.max(np.abs(res.y[0][4] - res.y[1][4])) > 1e-6
    # Newton converges as well as for the plain model: the analytic off-diagonal blocks carry the cell's own dPhi
    assert res.newton_failures[0] <= plain.newton_failures[0] + 2


@pytest.mark.parametrize("name,over,snap", [("scenario_A", {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}, 5),
                                            ("high_porosity", {}, 13),
                                            ("high_porosity", {"time_varying_dPhi": True}, 9)])
def test_device_jacobian_blocks_against_the_numpy_restatement(fixtures_reference, name, over, snap):
    """The analytic block-tridiagonal Jacobian as the implicit kernels form it on the device (marlpde_probe_jacobian ->
    jac_analytic) against its numpy restatement (oracle/jacobian_blocks.py, which tests/test_host_side.py pins to
    central differences of the oracle RHS): every entry of L, D, U of all 200 cells to 1e-11 of the block's largest."""
    import ctypes as C
    import jacobian_blocks as jb
    from marlpde_b200 import _cabi
    pde = oracle.derive_scenario({**oracle.default_scenario(), **over})
    p, N = oracle.kernel_params(pde), 200
    P = mb.derive_column_params(pde)
    y = np.ascontiguousarray(fixtures_reference[name][snap], dtype=np.float64).reshape(-1)
    J = np.zeros((N, 3, 5, 5))
    _cabi.check(_cabi.lib().marlpde_probe_jacobian(_cabi.ptr(y), _cabi.ptr(P), N, _cabi.ptr(J), 0))
    for i in range(N):
        want = jb.blocks(y, p, i, N)
        for b in range(3):
            got = J[i, b].T                                   # stored [column][row]
            assert np.max(np.abs(got - want[b])) <= 1e-11 * max(1.0, np.abs(want[b]).max()), (i, b)
    # ---- a cell ON a switching surface: the threshold Pe_min is moved onto |Pe_Phi| of cell 150 (x (1 -+ 1e-9)); the kernel
    # then takes d / d Phi of that cell from num_jac's one-sided difference quotient, which straddles the jump of the
    # weight when the porosity moves towards it (csrc/implicit_common.cuh kSwitchTol) — same rule in the restatement
    pde = oracle.default_scenario() | over                 # (Peclet numbers of O(1), as in the benchmark sweeps)
    p, P = oracle.kernel_params(pde), mb.derive_column_params(pde)
    i = 150
    y = y.copy()
    y[4 * N + i] += 0.01                                   # a kink in the porosity profile: the weight's jump then matters
    Phi_i = y[4 * N + i]
    F = 1 - np.exp(10 - 10 / Phi_i)
    W = p[8] - p[9] * Phi_i ** 2 * F
    dPhi_c = p[28] * Phi_i ** 3 * F / (1 - Phi_i) if over.get("time_varying_dPhi") else p[22]
    pe = abs(W * p[7] / (2 * dPhi_c))
    jumps = 0
    for side in (1 - 1e-9, 1 + 1e-9):
        p2, P2 = p.copy(), P.copy()
        p2[23] = pe * side
        P2["Peclet_min"] = pe * side
        _cabi.check(_cabi.lib().marlpde_probe_jacobian(_cabi.ptr(y), _cabi.ptr(P2), N, _cabi.ptr(J), 0))
        L, D, Ub = jb.blocks(y, p2, i, N, atol=1e-3)
        plain = jb.blocks(y, p2, i, N)[1]
        got = J[i, 1].T
        assert np.max(np.abs(got[:, :4] - D[:, :4])) <= 1e-11 * np.abs(D).max()
        assert np.max(np.abs(got[:, 4] - D[:, 4])) <= 1e-5 * max(np.abs(D[:, 4]).max(), np.abs(D).max()), side
        jumps += np.max(np.abs(D[:, 4] - plain[:, 4])) > 10 * np.abs(plain[:, 4]).max()   # the straddling quotient is huge
        # the neighbours are not on a switch: untouched
        assert np.max(np.abs(J[i - 1, 1].T - jb.blocks(y, p2, i - 1, N)[1])) <= 1e-11 * np.abs(plain).max()
    assert jumps >= 1                                          # from one of the two sides the difference crosses the switch


def test_analytic_jacobian_option_takes_the_same_steps():
    """The default Jacobian (every 5x5 block analytic, no RHS evaluation) against jac="fd" (MARLPDE_FLAG_JAC_FD:
    finite-difference diagonal blocks, what SciPy forms): same trajectory within the tolerance, same step / LU / Newton
    counts within 5 %, 5 RHS evaluations fewer per Jacobian."""
    pde = mb.sweep_lattice(oracle.default_scenario(), 2, 2, 2)
    P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
    a = mb.integrate_radau_batch(y0, P, t_span=(0, 0.02), first_step=1e-6)
    f = mb.integrate_radau_batch(y0, P, t_span=(0, 0.02), first_step=1e-6, jac="fd")
    assert np.all(a.status == 0) and np.all(f.status == 0)
    assert np.max(np.abs(a.y - f.y) / (ATOL + RTOL * np.abs(f.y))) <= 1.0      # two Newton histories at rtol = 1e-3 (measured 0.63)
    for name in ("n_accepted", "nlu", "newton_iterations"):
        x, y = getattr(a, name).sum(), getattr(f, name).sum()
        assert abs(int(x) - int(y)) <= 0.05 * y, name        # (measured: 3 % in nlu over these 8 columns)
    assert np.all(f.nfev - a.nfev >= 5 * f.njev - 0.05 * f.nfev)
    with pytest.raises(ValueError):
        mb.integrate_radau_batch(y0, P, jac="colored")


def test_team_columns_match_the_one_warp_shape():
    """opts.quantum = K: the first K columns of a batch run in the TEAM shape (two warps per column, second stream), the
    others one warp per column, in the same call.  The two shapes differ only in the order of the norm reductions (1e-16):
    on the smooth scenario-A base they take identical steps (states to 1e-9); through the stiff phase of the default base a
    rounding-level difference can flip a step-size decision (seen: 129 instead of 127 steps in one column), so there the
    gate is the tolerance.  "auto" puts a small batch entirely into teams."""
    smooth = mb.sweep_lattice(oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}, 2, 2, 3)
    P, y0 = mb.derive_column_params(smooth), mb.initial_state(smooth)
    kw = dict(t_span=(0, 0.5), first_step=1e-6, t_eval=[0.0, 0.25, 0.5], events=True, event_capacity=16)
    plain = mb.integrate_radau_batch(y0, P, team_columns=0, **kw)
    for K in (12, 5, "auto"):
        team = mb.integrate_radau_batch(y0, P, team_columns=K, **kw)
        assert np.all(team.status == 0)
        for name in ("n_accepted", "n_rejected", "nlu", "njev", "newton_iterations", "newton_failures", "nfev"):
            assert np.array_equal(getattr(team, name), getattr(plain, name)), (K, name)
        assert np.array_equal(team.event_counts, plain.event_counts)
        assert np.max(np.abs(team.y - plain.y)) <= 1e-9
        assert np.max(np.abs(np.asarray(team.snapshots) - np.asarray(plain.snapshots))) <= 1e-9
    stiff = mb.sweep_lattice(oracle.default_scenario(), 2, 2, 3)
    P, y0 = mb.derive_column_params(stiff), mb.initial_state(stiff)
    kw = dict(t_span=(0, 0.03), first_step=5e-7, t_eval=[0.0, 0.03], events=True, event_capacity=16)
    plain = mb.integrate_radau_batch(y0, P, team_columns=0, **kw)
    team = mb.integrate_radau_batch(y0, P, team_columns=7, **kw)
    assert np.all(team.status == 0) and team.event_counts.sum() > 0
    assert np.array_equal(team.event_counts[:, :6], plain.event_counts[:, :6])       # (6: W oscillates, count is step-dependent)
    for name in ("n_accepted", "nlu", "newton_iterations"):
        assert np.all(np.abs(getattr(team, name) - getattr(plain, name)) <= 0.05 * getattr(plain, name) + 2), name
    assert np.max(np.abs(team.y - plain.y) / (ATOL + RTOL * np.abs(plain.y))) <= 1.0
    # device-tensor path, and a team that is larger than the batch
    import torch
    dev = mb.integrate_radau_batch(torch.from_numpy(y0).cuda(), P, team_columns=100, **kw)
    assert np.max(np.abs(dev.y.cpu().numpy() - plain.y) / (ATOL + RTOL * np.abs(plain.y))) <= 1.0
    # the finite-difference Jacobian has no team shape: the request is ignored
    fd = mb.integrate_radau_batch(y0, P, team_columns=12, jac="fd", **kw)
    assert np.all(fd.status == 0)
