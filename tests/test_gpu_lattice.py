"""Parity on the BENCHMARK workload itself (BASELINE.json configs[1], default base): columns of the 16x16x16 lattice
against SciPy on the oracle (tests/golden/lattice_reference.npz, written by tests/golden/make_lattice_golden.py).

  * every column the GPU RK45 sweep to T* does not finish (the model runs into a singularity near t = 0.5-0.7 with
    Phi ~ 1.15; bench.py `time_to_Tstar.unfinished_columns`) stops in SciPy RK45 too, with the same status at the same
    time; two healthy neighbours finish in both;
  * 32 columns spread over the lattice + those unfinished ones with the implicit integrator: same status as SciPy Radau,
    end states within the tolerance two Radau codes at rtol = atol = 1e-3 can agree to, stop times of the failing ones
    close to SciPy's."""
import json
import os

import numpy as np
import pytest

import lheureux_oracle as oracle
import marlpde_b200 as mb
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
np.seterr(all="ignore")


@pytest.fixture(scope="module")
def lattice():
    g = np.load(os.path.join(GOLDEN, "lattice_reference.npz"))
    cols = json.loads(str(g["__columns__"]))
    pde = mb.sweep_lattice(oracle.default_scenario(), 16, 16, 16)
    return g, cols, pde


def _columns(pde, idx):
    from marlpde_b200 import sweep
    sub = sweep.shard(pde, np.asarray(idx))
    return mb.derive_column_params(sub), mb.initial_state(sub)


def test_rk45_unfinished_columns_stop_where_scipy_stops(lattice):
    g, cols, pde = lattice
    idx = cols["rk45"]
    P, y0 = _columns(pde, idx)
    # SciPy needs 0.51-1.02 M attempts for these columns; the cap bounds the one column (3070) whose attempt count at
    # the singularity is chaotic (0.68 M in SciPy, up to 6.8 M seen on the GPU: h ~ 1e-13, every attempt a coin toss)
    res = mb.integrate_rk45_batch(y0, P, t_span=(0, 1), first_step=1e-6, rtol=1e-3, atol=1e-3, max_steps=1_300_000,
                                  events=True, event_capacity=4)
    n_stopped = 0
    for k, c in enumerate(idx):
        want_status, want_t = int(g[f"rk45/{c}/status"]), float(g[f"rk45/{c}/t"])
        if want_status == 0:
            assert res.status[k] == 0 and res.t[k] == 1.0, c
            assert np.max(np.abs(res.y[k].ravel() - g[f"rk45/{c}/y"])) <= 1e-4, c
            attempts = int(g[f"rk45/{c}/counts"][2])
            assert abs(int(res.n_attempts[k]) - attempts) <= 1e-3 * attempts, c
        else:
            assert res.status[k] in (-1, 1), (c, res.status[k])          # 1: the cap, inside the singularity
            assert abs(res.t[k] - want_t) <= 1e-6, (c, res.t[k], want_t)
            assert res.h_abs[k] < 1e-10, c
            if res.status[k] == -1:
                n_stopped += 1
            # the end state sits on the singularity (a few cells differ by O(0.1)); most of the column is smooth
            if res.status[k] == -1:
                assert np.median(np.abs(res.y[k].ravel() - g[f"rk45/{c}/y"])) <= 1e-6, c
    assert n_stopped >= 8


def test_radau_lattice_columns_match_scipy(lattice):
    g, cols, pde = lattice
    idx = cols["radau"]
    P, y0 = _columns(pde, idx)
    res = mb.integrate_radau_batch(y0, P, t_span=(0, 1), first_step=1e-6, rtol=1e-3, atol=1e-3, t_eval=[1.0],
                                   events=True, event_capacity=8)
    worsts, n_fail, ratios, mismatched = [], 0, [], []
    for k, c in enumerate(idx):
        want_status, want_t = int(g[f"radau/{c}/status"]), float(g[f"radau/{c}/t"])
        if int(res.status[k]) != want_status:
            # a column next to the singular manifold: whether Newton still converges there depends on the linear
            # algebra (SciPy: sparse LU of the reference's pattern; here: exact block structure, fp32 preconditioner).
            # Seen: column 3831, SciPy gives up at t = 0.948, the kernel reaches T*.
            mismatched.append((c, int(res.status[k]), want_status, float(res.t[k]), want_t))
            continue
        if want_status == 0:
            want = g[f"radau/{c}/y"].reshape(5, 200)
            worsts.append(float(np.max(np.abs(res.y[k] - want) / (1e-3 + 1e-3 * np.abs(want)))))
            # the porosity crosses one in every column (how often is tangency-sensitive: 2-4 times in either code)
            assert (res.event_counts[k][4] > 0) == (g[f"radau/{c}/events"][4] > 0), c
            ratios.append((res.nlu[k] / g[f"radau/{c}/counts"][3], res.njev[k] / g[f"radau/{c}/counts"][2]))
        else:
            n_fail += 1
            assert abs(res.t[k] - want_t) <= 2e-2, (c, res.t[k], want_t)
    assert n_fail >= 8 and len(mismatched) <= 2, mismatched
    assert all(c not in (228, 229, 490, 491, 2545, 2546, 2807, 2808, 3069, 3070) for c, *_ in mismatched), mismatched
    # same algorithm, same decisions up to the tolerance of the linear algebra: factorisations and Jacobians per column
    # within 10 % of SciPy's (median over the columns that finish; measured r02i: within 2-4 % column by column)
    med = np.median(np.asarray(ratios), axis=0)
    assert 0.9 <= med[0] <= 1.1 and 0.9 <= med[1] <= 1.1, med
    # end states in units of atol + rtol |y|.  Measured (scripts/diag_lattice_radau.py, r02i): <= 0.01 units for 24 of the
    # 31 finishing columns, <= 1.2 for 28; three columns next to the singular manifold (2246, 2510, 3302) carry a sharp
    # porosity feature near cell 170-190 and differ by 14-233 units at rtol = 1e-3 — and by 0.03-7 units when the kernel
    # runs at 1e-5 against the same SciPy 1e-3 result: tolerance-level sensitivity of those columns, not a discrepancy.
    worsts = np.sort(np.asarray(worsts))
    assert np.median(worsts) <= 0.1, worsts
    assert worsts[int(0.85 * len(worsts))] <= 2.0, worsts
    assert worsts[-1] <= 500.0, worsts


def test_bdf_lattice_columns_match_scipy_bdf():
    """The BDF kernel on 32 columns spread over the benchmark lattice, to T*, against SciPy's BDF handed the same
    block-tridiagonal structure (tests/golden/lattice_reference_bdf.npz).  Through the smooth phases the two take the same
    steps; in the stiff phase (porosity excursion, W changing sign) a decision flipped by rounding sends them down
    different step sequences, so the gate is statistical like the Radau one: work per column within 10 % (median), end
    states within the distance two BDF runs at rtol = 1e-3 have (BDF's own global error here is ~15 tolerance units),
    and the same verdict on which columns finish — up to the knife-edge columns next to the model's switching surfaces
    (SciPy BDF itself finishes 2298 and stalls on 234)."""
    g = np.load(os.path.join(GOLDEN, "lattice_reference_bdf.npz"))
    idx = json.loads(str(g["__columns__"]))["bdf"]
    pde = mb.sweep_lattice(oracle.default_scenario(), 16, 16, 16)
    P, y0 = _columns(pde, idx)
    res = mb.integrate_bdf_batch(y0, P, t_span=(0, 1), first_step=1e-6, rtol=1e-3, atol=1e-3, t_eval=[1.0], events=True,
                                 event_capacity=8)
    worsts, ratios, mismatched = [], [], []
    for k, c in enumerate(idx):
        want_status = int(g[f"bdf/{c}/status"])
        if int(res.status[k]) != want_status:
            mismatched.append((c, int(res.status[k]), want_status, float(res.t[k]), float(g[f"bdf/{c}/t"])))
            continue
        if want_status == 0:
            want = g[f"bdf/{c}/y"].reshape(5, 200)
            worsts.append(float(np.max(np.abs(res.y[k] - want) / (1e-3 + 1e-3 * np.abs(want)))))
            ratios.append((res.nlu[k] / g[f"bdf/{c}/counts"][3], res.njev[k] / g[f"bdf/{c}/counts"][2],
                           res.newton_iterations[k] / g[f"bdf/{c}/counts"][1]))
            assert (res.event_counts[k][4] > 0) == (g[f"bdf/{c}/events"][4] > 0), c
    assert len(mismatched) <= 3, mismatched
    med = np.median(np.asarray(ratios), axis=0)
    assert np.all((0.9 <= med) & (med <= 1.1)), med
    worsts = np.sort(np.asarray(worsts))
    # measured (r02x): 26 of 32 columns agree to <= 0.26 tolerance units (median 0.01: the same steps), six — whose stiff phase
    # takes a different turn after a rounding-level decision — sit 28-194 units apart, as far as two BDF runs at rtol = 1e-3
    # are from each other there (BDF's own global error on these columns: tens of units)
    assert np.median(worsts) <= 1.0 and worsts[int(0.75 * len(worsts))] <= 1.0 and worsts[-1] <= 500.0, worsts
