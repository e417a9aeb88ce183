"""TEAM columns of the Radau kernel: latency of small batches (all columns in teams) and the 4096-column sweep to T* with the
K longest columns (predicted cost) in teams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
od = np.argsort(-mb.sweep.predicted_cost(pde, "Radau"), kind="stable")
P, y0 = P[od], y0[od]
def run(n, t_end, K):
    y = torch.from_numpy(y0[:n]).cuda(); dP = mb.batch.params_to_device(P[:n], y.device)
    torch.cuda.synchronize(); t0 = time.time()
    r = mb.integrate_radau_batch(y, dP, t_span=(0, t_end), first_step=1e-6, team_columns=K)
    torch.cuda.synchronize(); return time.time() - t0, r
run(8, 1e-3, 8); run(8, 1e-3, 0)
for n, t_end in ((1, 1.0), (16, 1.0), (148, 0.2), (296, 0.2)):
    a, ra = run(n, t_end, 0); b, rb = run(n, t_end, n)
    print(f"{n} columns (the heaviest) to t={t_end}: one warp each {a:.3f} s, teams {b:.3f} s ({a / b:.2f}x); newton {ra.newton_iterations.sum()} / {rb.newton_iterations.sum()}", flush=True)
for K in (0, 128, 256, 400, 600, 256):
    dt, r = run(4096, 1.0, K)
    print(f"4096 columns to T*, K = {K} team columns: {dt:.2f} s, finished {(r.status == 0).sum()}, newton {r.newton_iterations.sum()}", flush=True)
