#!/usr/bin/env python
"""Golden end-to-end trajectories: the installed SciPy `solve_ivp` (the reference's own
time stepper, marlpde/Evolve_scenario.py:104-109) driving the oracle RHS, for the three
parameter sets of the reference's regression tests (tests/Regression_test/
test_regression.py:39-44, :73-80, :114-117) plus short-horizon runs used by fast tests.

    python tests/golden/make_stepper_golden.py        # ~6 min on 8 cores

Writes tests/golden/stepper_reference.npz.  Full-T* RK45 runs take 90-160 s each in SciPy,
which is why their results are committed instead of recomputed inside the test-suite.
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    "scenario_A": ({"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}, 1e-6),
    "high_porosity": ({"Phi0": 0.8, "PhiIni": 0.8, "PhiNR": 0.8}, 5e-7),
    "matlab": ({"Phi0": 0.5, "PhiIni": 0.5, "PhiNR": 0.5, "k3": 0.01, "k4": 0.01}, 1e-6),
}


def run(job):
    import lheureux_oracle as o
    name, method, t_end, tol, n_eval = job
    over, first_step = CASES[name]
    pde = o.default_scenario() | over
    kw = {}
    if method == "Radau":
        kw["jac_sparsity"] = o.jacobian_sparsity(pde["N"])
    t_eval = np.linspace(0.0, t_end, n_eval)
    sol = o.integrate(pde, method=method, first_step=first_step, rtol=tol, atol=tol,
                      t_span=(0.0, t_end), t_eval=t_eval, **kw)
    key = f"{name}/{method}/t{t_end:g}/tol{tol:g}"
    res = {key + "/y": sol.y, key + "/t": sol.t,
           key + "/counts": np.array([sol.nfev, sol.njev, sol.nlu, sol.status], dtype=np.int64)}
    for k, te in enumerate(sol.t_events):
        res[key + f"/event_{k}"] = np.asarray(te)
    print(key, "nfev", sol.nfev, "status", sol.status, flush=True)
    return res


def main():
    jobs = []
    for name in CASES:
        jobs.append((name, "RK45", 1.0, 1e-3, 2))          # what Solver(method="RK45") runs
        jobs.append((name, "RK45", 0.002, 1e-3, 5))        # short horizon, with dense output
        jobs.append((name, "Radau", 1.0, 1e-3, 2))         # what the regression tests run
    jobs.append(("scenario_A", "Radau", 1.0, 1e-8, 11))    # tight: pins the oracle to fixture A
    jobs.append(("high_porosity", "Radau", 1.0, 1e-6, 11))  # pins the oracle to fixture B
    jobs.sort(key=lambda j: -(j[2] * (100 if j[1] == "RK45" else 1)))
    out = {}
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count())) as ex:
        for res in ex.map(run, jobs):
            out.update(res)
    # CPU-vs-CPU round-off sensitivity of the Matlab case: first_step * (1 + 1e-9).  Its cCO3 boundary
    # layer (top ~15 cells) moves by up to 1.3e-2 and nfev by 354: the noise floor any second
    # implementation of the same stepper (the GPU one) is compared against.
    import lheureux_oracle as o
    pde = o.default_scenario() | CASES["matlab"][0]
    sol = o.integrate(pde, method="RK45", first_step=1e-6 * (1 + 1e-9), events=False)
    out["matlab/RK45/t1/tol0.001/y_end_first_step_times_1p000000001"] = sol.y[:, -1].reshape(5, -1)
    out["matlab/RK45/t1/tol0.001/nfev_first_step_times_1p000000001"] = np.array(sol.nfev)
    out["__cases__"] = np.array(json.dumps({k: {"overrides": v[0], "first_step": v[1]} for k, v in CASES.items()}))
    np.savez_compressed(os.path.join(HERE, "stepper_reference.npz"), **out)


if __name__ == "__main__":
    main()
