"""GPU parity of the single RHS call (marlpde_rhs_batch through the C ABI) against
 (a) golden outputs of the reference's own pde_rhs (tests/golden/rhs_reference.npz),
 (b) the CPU oracle on seeded random states.
Gate (BASELINE.json north_star: "single-call RHS <= 1e-12 relative"), stated the way SURVEY.md
App. A.6 makes it well-posed: |gpu - ref|_i <= 1e-12 * S_i with S_i the magnitude of the terms
that cancel in entry i, and field-inf-norm relative <= 1e-12 on non-cancelling (random) states."""
import numpy as np
import pytest

import lheureux_oracle as oracle
import marlpde_b200 as mb
from conftest import rhs_states

pytestmark = pytest.mark.gpu
np.seterr(all="ignore")
TOL = 1e-12


def _scaled_err(got, ref, y, po):
    S = oracle.term_scale(y, po, np.empty_like(y))
    return np.nanmax(np.abs(got - ref) / S)


@pytest.mark.parametrize("name", ["default", "scenario_A", "matlab", "fv_off", "exponents", "lattice_corner"])
def test_rhs_matches_reference_golden(rhs_golden, name):
    g, meta = rhs_golden
    pde = meta[name]
    P = mb.derive_column_params(pde)
    po = oracle.kernel_params(pde)
    states = list(rhs_states(g, name))
    Y = np.stack([s[1].reshape(5, -1) for s in states])
    out = mb.rhs_batch(Y, np.repeat(P, len(states)))            # all states of this case in one launch
    for (sname, y, r_numba, _r_numpy, _ev), got in zip(states, out):
        got = got.ravel()
        assert np.array_equal(np.isnan(got), np.isnan(r_numba)), sname
        assert _scaled_err(got, r_numba, y, po) <= TOL, sname
        if sname in ("y0", "noise"):                                # non-cancelling states: plain relative error
            for f in range(5):
                sl = slice(f * 200, (f + 1) * 200)
                assert np.max(np.abs(got[sl] - r_numba[sl])) <= TOL * np.max(np.abs(r_numba[sl])), (sname, f)


@pytest.mark.parametrize("name", ["default", "scenario_A", "fv_off"])
def test_rhs_time_varying_dPhi_matches_reference_variant(rhs_golden_vardphi, name):
    """Model variant MARLPDE_MODEL_VAR_DPHI against golden outputs of the reference's own pde_rhs with its commented-out
    dPhi line switched on (LHeureux_model.py:430-431; tests/golden/make_golden.py vardphi)."""
    g, meta = rhs_golden_vardphi
    pde = meta[name] | {"time_varying_dPhi": True}
    P = mb.derive_column_params(pde)
    po = oracle.kernel_params(pde)
    keys = sorted(k for k in g.files if k.startswith(name + "/") and k.endswith("/y"))
    Y = np.stack([g[k].reshape(5, -1) for k in keys])
    out = mb.rhs_batch(Y, np.repeat(P, len(keys)))
    for key, got in zip(keys, out):
        ref = g[key[:-2] + "/rhs_numba"]
        assert np.array_equal(np.isnan(got.ravel()), np.isnan(ref)), key
        assert _scaled_err(got.ravel(), ref, g[key], po) <= TOL, key
    # flagged and plain columns in one launch: the plain ones are what a launch without flagged columns gives
    P0 = mb.derive_column_params(meta[name])
    mixed = mb.rhs_batch(np.stack([Y[1], Y[1], Y[1]]), np.concatenate([P, P0, P]))
    assert np.array_equal(mixed[1], mb.rhs_batch(Y[1:2], P0)[0]) and np.array_equal(mixed[0], out[1])
    assert not np.array_equal(mixed[0][4], mixed[1][4])


def test_rhs_matches_reference_numpy_backend_golden(rhs_golden):
    """Row a4: `fun` (the reference's NumPy backend, LHeureux_model.py:162-288) computes the same maths re-associated;
    the CUDA RHS is within the same 1e-12 term-magnitude gate of its golden outputs."""
    g, meta = rhs_golden
    for name in ("default", "scenario_A", "lattice_corner"):
        pde = meta[name]
        P, po = mb.derive_column_params(pde), oracle.kernel_params(pde)
        states = list(rhs_states(g, name))
        out = mb.rhs_batch(np.stack([s[1].reshape(5, -1) for s in states]), np.repeat(P, len(states)))
        for (sname, y, _r_numba, r_numpy, _ev), got in zip(states, out):
            assert _scaled_err(got.ravel(), r_numpy, y, po) <= TOL, (name, sname)


def test_rhs_sweep_batch_against_oracle():
    """512 lattice columns with different parameters, each on its own perturbed state."""
    base = oracle.default_scenario() | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    pde = mb.sweep_lattice(base, 8, 8, 8)
    P = mb.derive_column_params(pde)
    rng = np.random.default_rng(7)
    Y = mb.initial_state(pde) * (1 + 0.05 * rng.uniform(-1, 1, (512, 5, 200)))
    out = mb.rhs_batch(Y, P)
    worst = 0.0
    for c in range(0, 512, 7):
        one = {k: (v[c] if np.ndim(v) else v) for k, v in pde.items()}
        po = oracle.kernel_params(one)
        ref = oracle.rhs(Y[c].ravel(), po, np.empty(1000))
        worst = max(worst, _scaled_err(out[c].ravel(), ref, Y[c].ravel(), po))
    assert worst <= TOL


@pytest.mark.parametrize("n_cells", [2, 3, 31, 32, 33, 255, 256, 257, 1000, 2000])
def test_rhs_grid_sizes_and_tile_edges(n_cells):
    """Ragged sizes around the 256-cell CTA tile and the smallest grids py-pde's stencils allow."""
    pde = oracle.default_scenario() | {"N": n_cells, "Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    P = mb.derive_column_params(pde)
    po = oracle.kernel_params(pde)
    rng = np.random.default_rng(n_cells)
    y = mb.initial_state(pde) * (1 + 0.05 * rng.uniform(-1, 1, (1, 5, n_cells)))
    y[0, 4] = np.linspace(0.3, 0.9, n_cells)
    got = mb.rhs_batch(np.repeat(y, 3, axis=0), np.repeat(P, 3))
    ref = oracle.rhs(y.ravel(), po, np.empty(5 * n_cells))
    for c in range(3):
        assert _scaled_err(got[c].ravel(), ref, y.ravel(), po) <= TOL
    assert np.array_equal(got[0], got[2])


def test_rhs_empty_batch_and_device_tensor_path():
    import torch
    pde = oracle.default_scenario()
    P = mb.derive_column_params(pde)
    assert mb.rhs_batch(np.empty((0, 5, 200)), P[:0]).shape == (0, 5, 200)
    y = mb.initial_state(pde)
    host = mb.rhs_batch(y, P)
    dev = mb.rhs_batch(torch.from_numpy(y).cuda(), P)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)      # same kernel, same bits


def test_rhs_nonfinite_inputs_follow_ieee():
    """Phi == 1 and Phi == 0 raise ZeroDivisionError inside numba upstream; on the GPU they follow
    IEEE (inf/NaN in that cell only) and never poison neighbouring columns."""
    pde = oracle.default_scenario()
    P = mb.derive_column_params(pde)
    y = np.repeat(mb.initial_state(pde), 2, axis=0)
    y[0, 4, 50] = 1.0
    out = mb.rhs_batch(y, np.repeat(P, 2))
    assert not np.all(np.isfinite(out[0, :, 50]))
    assert np.all(np.isfinite(out[0, :, :49])) and np.all(np.isfinite(out[0, :, 52:]))
    assert np.all(np.isfinite(out[1]))


@pytest.mark.parametrize("sched", [0, 1, 2, 3, 4])
def test_rhs_every_instruction_schedule(rhs_golden, sched, monkeypatch):
    """rhs_pair_own exists in five instruction schedules (csrc/lheureux_device.cuh: kSchedSplit, kSchedMerged,
    kSchedAll, kSchedTwoArm, kSchedLean); each integrator kernel instantiates the one that is fastest for it.  All of them
    must meet the single-call gate on the reference's golden states (default case: dissolution zone, evolved
    fixture snapshots, noise), on a case with FV_switch = 0, and on IEEE special values."""
    monkeypatch.setenv("MARLPDE_RHS_SCHEDULE", str(sched))
    g, meta = rhs_golden
    for name in ("default", "fv_off", "lattice_corner"):
        pde = meta[name]
        P = mb.derive_column_params(pde)
        po = oracle.kernel_params(pde)
        states = list(rhs_states(g, name))
        Y = np.stack([s[1].reshape(5, -1) for s in states])
        out = mb.rhs_batch(Y, np.repeat(P, len(states)))
        for (sname, y, r_numba, _r_numpy, _ev), got in zip(states, out):
            got = got.ravel()
            assert np.array_equal(np.isnan(got), np.isnan(r_numba)), (name, sname)
            assert _scaled_err(got, r_numba, y, po) <= TOL, (name, sname)
    pde = oracle.default_scenario()
    P = mb.derive_column_params(pde)
    y = np.repeat(mb.initial_state(pde), 2, axis=0)
    y[0, 4, 50] = 1.0
    out = mb.rhs_batch(y, np.repeat(P, 2))
    assert not np.all(np.isfinite(out[0, :, 50])) and np.all(np.isfinite(out[1]))
    monkeypatch.delenv("MARLPDE_RHS_SCHEDULE")
    ref = mb.rhs_batch(y, np.repeat(P, 2))                       # default schedule (the RK45 kernel's)
    monkeypatch.setenv("MARLPDE_RHS_SCHEDULE", str(sched))
    assert np.array_equal(np.isfinite(out), np.isfinite(ref))
    assert np.allclose(out[1], ref[1], rtol=1e-12, atol=0.0)
