#!/usr/bin/env python
"""Golden results of SciPy's BDF for columns of the BENCHMARK workload (the 16x16x16 Map_Scenario lattice, default base):
`solve_ivp(method="BDF")` (marlpde/parameters.py:235-236; call site Evolve_scenario.py:104-109) on the oracle RHS, handed
the block-tridiagonal structure (the kernel's; with it SciPy takes the kernel's steps, tests/test_gpu_bdf.py).

    python tests/golden/make_lattice_golden_bdf.py        # ~6 min on 8 cores

Writes tests/golden/lattice_reference_bdf.npz: bdf/<column>/{status, t, y, counts = [steps, nfev, njev, nlu], events} for the
32 columns of make_lattice_golden.py's spread.  tests/test_gpu_lattice.py compares the BDF kernel with it."""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
SPREAD = [int(round(c)) for c in np.linspace(0, 4095, 32)]


def run_bdf(c):
    import scipy.sparse as sp
    import lheureux_oracle as o
    import marlpde_b200.params as mp
    from scipy.integrate import BDF
    np.seterr(all="ignore")
    lat = mp.sweep_lattice(o.default_scenario(), 16, 16, 16)
    pde = {k: (float(v[c]) if np.ndim(v) else v) for k, v in lat.items()}
    cells = np.arange(1000) % 200
    exact = sp.csr_matrix((np.abs(cells[:, None] - cells[None, :]) <= 1).astype(float))
    ev, status = [0] * 7, -9
    try:
        sol = o.integrate(pde, method="BDF", first_step=1e-6, rtol=1e-3, atol=1e-3, t_span=(0.0, 1.0), t_eval=None,
                          events=True, jac_sparsity=exact)
        y_end, t_end, status = sol.y[:, -1], float(sol.t[-1]), int(sol.status)
        counts = [len(sol.t) - 1, sol.nfev, sol.njev, sol.nlu]
        ev = [len(e) for e in sol.t_events]
    except Exception as exc:
        print("bdf", c, "raised", type(exc).__name__, exc, flush=True)
    if status != 0:                                   # where does SciPy's BDF class stop?
        solver = BDF(o.rhs_fn(o.kernel_params(pde)), 0.0, o.initial_state(pde), 1.0, first_step=1e-6, rtol=1e-3, atol=1e-3,
                     jac_sparsity=exact)
        n = 0
        try:
            while solver.status == "running":
                solver.step()
                n += 1
        except Exception as exc:
            print("bdf", c, "manual stepping raised", type(exc).__name__, flush=True)
        y_end, t_end, counts = np.array(solver.y), float(solver.t), [n, solver.nfev, solver.njev, solver.nlu]
    print("bdf", c, "status", status, "t", t_end, "steps", counts[0], flush=True)
    return (c, status, t_end, y_end, counts, ev)


def main():
    out = {}
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count())) as ex:
        for c, status, t_end, y_end, counts, ev in ex.map(run_bdf, SPREAD):
            key = f"bdf/{c}"
            out[key + "/status"] = np.int64(status)
            out[key + "/t"] = np.float64(t_end)
            out[key + "/y"] = np.asarray(y_end, dtype=np.float64)
            out[key + "/counts"] = np.asarray(counts, dtype=np.int64)
            out[key + "/events"] = np.asarray(ev, dtype=np.int64)
    out["__columns__"] = np.array(json.dumps({"bdf": SPREAD}))
    np.savez_compressed(os.path.join(HERE, "lattice_reference_bdf.npz"), **out)
    print("wrote lattice_reference_bdf.npz")


if __name__ == "__main__":
    main()
