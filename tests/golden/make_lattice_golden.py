#!/usr/bin/env python
"""Golden results for columns of the BENCHMARK workload itself (BASELINE.json configs[1]: the 16x16x16 Map_Scenario
lattice, default base): the installed SciPy `solve_ivp` machinery (the reference's own time stepper,
marlpde/Evolve_scenario.py:104-109) driving the oracle RHS.

    python tests/golden/make_lattice_golden.py        # ~10 min on 8 cores

Writes tests/golden/lattice_reference.npz:
  radau/<column>/...   32 columns spread evenly over the lattice + the columns the GPU sweeps do not finish: SciPy Radau
                       (rtol = atol = 1e-3, first_step 1e-6, the reference's jac_sparsity, 7 events) to T*: status,
                       time reached, end state, step / nfev / njev / nlu counts, event counts.
  rk45/<column>/...    every column the GPU RK45 sweep ends with status != 0 (profiles/r02c_bench_default.json,
                       `time_to_Tstar.unfinished_columns`) plus two healthy neighbours: SciPy's RK45 class stepped until
                       it stops (status, time reached, step counts; capped at 8 M attempts).
The GPU tests (tests/test_gpu_lattice.py) assert that the sweeps reproduce status, time and end state.
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))

UNFINISHED_RK45 = [228, 229, 490, 491, 2545, 2546, 2807, 2808, 3069, 3070]      # GPU sweep r02c, status -1 (3070: step cap)
HEALTHY = [227, 3071]
SPREAD = [int(round(c)) for c in np.linspace(0, 4095, 32)]
STEP_CAP = 8_000_000


def column(c):
    import marlpde_b200.params as mp
    import lheureux_oracle as o
    pde = mp.sweep_lattice(o.default_scenario(), 16, 16, 16)
    return {k: (float(v[c]) if np.ndim(v) else v) for k, v in pde.items()}


def run_radau(c):
    import lheureux_oracle as o
    np.seterr(all="ignore")
    pde = column(c)
    try:
        sol = o.integrate(pde, method="Radau", first_step=1e-6, rtol=1e-3, atol=1e-3, t_span=(0.0, 1.0), t_eval=None,
                          events=True, jac_sparsity=o.jacobian_sparsity(200))
        y_end, t_end, status = sol.y[:, -1], float(sol.t[-1]), int(sol.status)
        counts = [len(sol.t) - 1, sol.nfev, sol.njev, sol.nlu]
        ev = [len(e) for e in sol.t_events]
    except Exception as exc:                         # numerical blow-up inside SciPy's linear algebra
        y_end, t_end, status, counts, ev = np.full(1000, np.nan), float("nan"), -9, [0, 0, 0, 0], [0] * 7
        print("radau", c, "raised", type(exc).__name__, exc, flush=True)
    if status != 0:
        # solve_ivp only returns the t_eval samples: step SciPy's Radau class by hand to learn where it stops
        from scipy.integrate import Radau
        solver = Radau(o.rhs_fn(o.kernel_params(pde)), 0.0, o.initial_state(pde), 1.0, first_step=1e-6, rtol=1e-3,
                       atol=1e-3, jac_sparsity=o.jacobian_sparsity(200))
        n = 0
        try:
            while solver.status == "running":
                solver.step()
                n += 1
        except Exception as exc:
            print("radau", c, "manual stepping raised", type(exc).__name__, flush=True)
        y_end, t_end = np.array(solver.y), float(solver.t)
        counts = [n, solver.nfev, solver.njev, solver.nlu]
    print("radau", c, "status", status, "t", t_end, "steps", counts[0], flush=True)
    return ("radau", c, status, t_end, y_end, counts, ev)


def run_rk45(c):
    """SciPy's RK45 class stepped by hand (what solve_ivp's loop does, without storing a million steps)."""
    import lheureux_oracle as o
    from scipy.integrate import RK45
    np.seterr(all="ignore")
    pde = column(c)
    p = o.kernel_params(pde)
    f = o.rhs_fn(p)
    solver = RK45(f, 0.0, o.initial_state(pde), 1.0, first_step=1e-6, rtol=1e-3, atol=1e-3)
    n = 0
    status = 1
    while n < STEP_CAP:
        msg = solver.step()
        n += 1
        if solver.status == "finished":
            status = 0
            break
        if solver.status == "failed":
            status = -1
            break
    attempts = (solver.nfev - 1) // 6
    print("rk45", c, "status", status, "t", solver.t, "accepted", n, "attempts", attempts, flush=True)
    return ("rk45", c, status, float(solver.t), np.array(solver.y), [n, solver.nfev, attempts, 0], [0] * 7)


def main():
    jobs = [(run_rk45, c) for c in UNFINISHED_RK45 + HEALTHY]
    jobs += [(run_radau, c) for c in sorted(set(SPREAD + UNFINISHED_RK45))]
    out = {}
    if len(sys.argv) > 1 and sys.argv[1] == "patch-failed-radau":       # re-run only the columns Radau does not finish
        old = np.load(os.path.join(HERE, "lattice_reference.npz"))
        out = {k: old[k] for k in old.files}
        cols = json.loads(str(old["__columns__"]))["radau"]
        jobs = [(run_radau, c) for c in cols if int(old[f"radau/{c}/status"]) != 0]
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count())) as ex:
        futs = [ex.submit(fn, c) for fn, c in jobs]
        for fu in futs:
            kind, c, status, t_end, y_end, counts, ev = fu.result()
            key = f"{kind}/{c}"
            out[key + "/status"] = np.int64(status)
            out[key + "/t"] = np.float64(t_end)
            out[key + "/y"] = np.asarray(y_end, dtype=np.float64)
            out[key + "/counts"] = np.asarray(counts, dtype=np.int64)
            out[key + "/events"] = np.asarray(ev, dtype=np.int64)
    if "__columns__" not in out:
        out["__columns__"] = np.array(json.dumps({"rk45": UNFINISHED_RK45 + HEALTHY,
                                                  "radau": sorted(set(SPREAD + UNFINISHED_RK45))}))
    np.savez_compressed(os.path.join(HERE, "lattice_reference.npz"), **out)
    print("wrote lattice_reference.npz")


if __name__ == "__main__":
    main()
