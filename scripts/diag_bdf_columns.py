"""BDF kernel on selected columns of the 4096-column lattice (default base): python scripts/diag_bdf_columns.py t_end col [col ...]
(or `@file.npz` = the unfinished columns of a status dump).  Prints status / time reached / counters per column."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
t_end = float(sys.argv[1])
if sys.argv[2].startswith("@"):
    cols = np.nonzero(np.load(sys.argv[2][1:])["status"] != 0)[0]
else:
    cols = np.array([int(c) for c in sys.argv[2:]])
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P, y0 = mb.derive_column_params(pde)[cols], mb.initial_state(pde)[cols]
r = mb.integrate_bdf_batch(y0, P, t_span=(0, t_end), first_step=1e-6)
print("finished", int((r.status == 0).sum()), "of", cols.size)
for k, c in enumerate(cols[:24]):
    print(f"col {c}: status {r.status[k]} t {r.t[k]:.6f} h {r.h_abs[k]:.3e} acc {r.n_accepted[k]} rej {r.n_rejected[k]} njev {r.njev[k]} nlu {r.nlu[k]} newton {r.newton_iterations[k]} fails {r.newton_failures[k]}")
