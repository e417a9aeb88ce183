"""Short run of an implicit kernel for timing / ncu: lattice columns of the default base, t in [0, t_end].
    python scripts/profile_implicit.py radau|bdf [n_lattice] [t_end] [cost-ordered: 0|1] [fd|analytic]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
kind = sys.argv[1] if len(sys.argv) > 1 else "bdf"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
t_end = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
ordered = len(sys.argv) > 4 and sys.argv[4] == "1"
jac = sys.argv[5] if len(sys.argv) > 5 else "analytic"
run = mb.integrate_bdf_batch if kind == "bdf" else mb.integrate_radau_batch
pde = mb.sweep_lattice(asdict(Map_Scenario()), n, n, n)
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
if ordered:
    od = np.argsort(-mb.sweep.predicted_cost(pde, "Radau"), kind="stable")
    P, y0 = P[od], y0[od]
y = torch.from_numpy(y0).cuda()
dP = mb.batch.params_to_device(P, y.device)
run(y[:2], mb.batch.params_to_device(P[:2], y.device), t_span=(0, 1e-4), first_step=1e-6)   # warm-up
torch.cuda.synchronize(); t0 = time.time()
r = run(y, dP, t_span=(0, t_end), first_step=1e-6, jac=jac)
torch.cuda.synchronize(); dt = time.time() - t0
fin = int((r.status == 0).sum())
print(f"{kind} {n**3} columns to t={t_end}: {dt:.3f}s finished {fin} steps {r.n_accepted.min()}-{r.n_accepted.max()} (sum {r.n_accepted.sum()}) rejected {r.n_rejected.sum()} nlu {r.nlu.sum()} newton {r.newton_iterations.sum()} fails {r.newton_failures.sum()} nfev {r.nfev.sum()} njev {r.njev.sum()}")
print(f"newton it/s {r.newton_iterations.sum()/dt:.3e}; steps/s {(r.n_accepted.sum()+r.n_rejected.sum())/dt:.3e}; status histogram {dict(zip(*np.unique(r.status, return_counts=True)))}")
