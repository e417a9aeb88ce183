"""Implicit sweep of the 4096-column benchmark lattice (default base) to T*: per-column status, time reached, work counters
-> npz (which columns an integrator does not finish, for the SciPy cross-check).
    python scripts/dump_implicit_status.py radau|bdf out.npz [fd|analytic]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
kind, out = sys.argv[1], sys.argv[2]
jac = sys.argv[3] if len(sys.argv) > 3 else "analytic"
run = mb.integrate_bdf_batch if kind == "bdf" else mb.integrate_radau_batch
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
od = np.argsort(-mb.sweep.predicted_cost(pde, "Radau"), kind="stable")
y = torch.from_numpy(y0[od]).cuda()
run(y[:2], P[od][:2], t_span=(0, 1e-4), first_step=1e-6)
torch.cuda.synchronize(); t0 = time.time()
r = run(y, P[od], t_span=(0, 1.0), first_step=1e-6, t_eval=[0.0, 1.0], events=True, event_capacity=16, jac=jac)
torch.cuda.synchronize(); dt = time.time() - t0
inv = np.empty_like(od); inv[od] = np.arange(od.size)
g = lambda a: np.asarray(a)[inv]
np.savez_compressed(out, seconds=dt, status=g(r.status), t=g(r.t), n_accepted=g(r.n_accepted), n_rejected=g(r.n_rejected),
                    nlu=g(r.nlu), njev=g(r.njev), newton=g(r.newton_iterations), fails=g(r.newton_failures), nfev=g(r.nfev))
bad = np.nonzero(g(r.status) != 0)[0]
print(f"{kind} jac={jac}: {dt:.2f} s, finished {(g(r.status) == 0).sum()} of {od.size}; unfinished columns {bad.tolist()} at t = {np.round(g(r.t)[bad], 4).tolist()}")
