"""Experiment: how much of the implicit sweep's time-to-T* is tail (a few columns need 3-5x the average number of steps)?
natural column order / longest-first by measured work / two passes with a step budget.
    python scripts/exp_radau_tail.py [n_lattice]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pde = mb.sweep_lattice(asdict(Map_Scenario()), n, n, n)
P = mb.derive_column_params(pde)
y0 = mb.initial_state(pde)
dev = torch.device("cuda")

def run(order, **kw):
    y = torch.from_numpy(np.ascontiguousarray(y0[order])).to(dev)
    dP = mb.batch.params_to_device(np.ascontiguousarray(P[order]), dev)
    torch.cuda.synchronize(); t0 = time.time()
    r = mb.integrate_radau_batch(y, dP, t_span=(0, 1), first_step=1e-6, events=True, event_capacity=16, inplace=True, **kw)
    torch.cuda.synchronize()
    return r, time.time() - t0, dP

nat = np.arange(n ** 3)
run(nat[:64], max_steps=50)                                   # warm-up
r, dt, _ = run(nat)
work = r.n_accepted + r.n_rejected + r.newton_failures
print(f"natural order: {dt:.2f}s finished {(r.status == 0).sum()} steps {r.n_accepted.min()}-{r.n_accepted.max()} mean {r.n_accepted.mean():.0f}")
q = np.quantile(r.n_accepted, [0.5, 0.9, 0.99, 0.999]); print("step quantiles 50/90/99/99.9 %:", q)
lpt = np.argsort(-work, kind="stable")
r2, dt2, _ = run(lpt)
print(f"longest-first (measured work): {dt2:.2f}s")
import marlpde_b200.sweep as sw
pred = np.argsort(-sw.predicted_cost(pde) * np.ones(n ** 3), kind="stable")
r3, dt3, _ = run(pred)
print(f"longest-first (predicted explicit cost): {dt3:.2f}s")
for cap in (2000, 3000):
    ra, dta, dP = run(nat, max_steps=cap)
    left = np.nonzero(ra.status == 1)[0]
    yb = ra.y[torch.from_numpy(left).to(dev)].contiguous()
    dPb = mb.batch.params_to_device(np.ascontiguousarray(P[left]), dev)
    torch.cuda.synchronize(); t0 = time.time()
    rb = mb.integrate_radau_batch(yb, dPb, t_span=(0, 1), state=ra.state[left], events=True, event_capacity=16, inplace=True)
    torch.cuda.synchronize(); dtb = time.time() - t0
    print(f"two passes, budget {cap}: {dta:.2f}s + {dtb:.2f}s = {dta + dtb:.2f}s ({len(left)} columns continue, finished {(ra.status == 0).sum() + (rb.status == 0).sum()})")
