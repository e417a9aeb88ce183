"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep it tiny)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
which = sys.argv[1] if len(sys.argv) > 1 else "all"
base = asdict(Map_Scenario())
pde = mb.sweep_lattice(base, 2, 2, 2)
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
if which in ("all", "rhs"):
    print("rhs", np.isfinite(mb.rhs_batch(y0, P)).all())
if which in ("all", "rk45"):
    y0e = y0.copy(); y0e[0, 4, 150] = 1.0005; y0e[1, 0, 50] = -2e-3          # forces event location
    r = mb.integrate_rk45_batch(y0e, P, t_span=(0, 1), first_step=1e-6, max_steps=60, t_eval=[0.0, 2e-5], events=True, event_capacity=4)
    print("rk45", r.status, r.n_attempts, r.event_counts.sum())
if which in ("all", "tiles", "stages"):
    pl = base | {"N": 1300, "Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    if which == "stages":
        os.environ["MARLPDE_RK45_STREAM"] = "stages"
    r = mb.integrate_rk45_batch(np.repeat(mb.initial_state(pl), 2, 0), np.repeat(mb.derive_column_params(pl), 2), t_span=(0, 1),
                                first_step=2e-8, max_steps=12, t_eval=[0.0, 1e-7])
    print(which, r.status, r.n_attempts)
if which in ("all", "radau"):
    r = mb.integrate_radau_batch(y0[:3], P[:3], t_span=(0, 1), first_step=1e-6, max_steps=4, t_eval=[0.0, 1e-5], events=True, event_capacity=4)
    print("radau", r.status, r.n_accepted, r.nlu)
