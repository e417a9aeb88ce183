"""Per-column deviation of the GPU Radau end states from the SciPy golden (tests/golden/lattice_reference.npz)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import lheureux_oracle as oracle, marlpde_b200 as mb
from marlpde_b200 import sweep
g = np.load(os.path.join(ROOT, "tests", "golden", "lattice_reference.npz"))
idx = json.loads(str(g["__columns__"]))["radau"]
pde = mb.sweep_lattice(oracle.default_scenario(), 16, 16, 16)
sub = sweep.shard(pde, np.asarray(idx))
P, y0 = mb.derive_column_params(sub), mb.initial_state(sub)
for tol in (1e-3, 1e-5):
    res = mb.integrate_radau_batch(y0, P, t_span=(0, 1), first_step=1e-6, rtol=tol, atol=tol, t_eval=[1.0], events=True, event_capacity=8)
    print("tolerance", tol)
    for k, c in enumerate(idx):
        st = int(g[f"radau/{c}/status"])
        if st != 0 or res.status[k] != 0:
            print(c, "status gpu/scipy", res.status[k], st, "t", res.t[k], float(g[f"radau/{c}/t"]))
            continue
        want = g[f"radau/{c}/y"].reshape(5, 200)
        d = np.abs(res.y[k] - want) / (1e-3 + 1e-3 * np.abs(want))
        f, i = np.unravel_index(np.argmax(d), d.shape)
        print(c, f"worst {d.max():8.2f} at field {f} cell {i}  per-field max {np.round(d.max(axis=1), 1)}  median {np.median(d):.3f}  steps gpu {res.n_accepted[k]} nlu gpu/scipy {res.nlu[k]}/{int(g[f'radau/{c}/counts'][3])} njev {res.njev[k]}/{int(g[f'radau/{c}/counts'][2])}")
