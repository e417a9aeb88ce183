"""Per-step device time of the bench workload over 16 consecutive steps (events on / off)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde_b200 import batch
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P = mb.derive_column_params(pde)
for ev in (True, False):
    y = torch.from_numpy(mb.initial_state(pde)).cuda(); dP = batch.params_to_device(P, y.device)
    r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), first_step=1e-6, max_steps=3000, events=ev, event_capacity=16)
    out = []
    for i in range(16):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = mb.integrate_rk45_batch(r.y, dP, t_span=(0, 1), max_steps=3000, state=r.state, events=ev, event_capacity=16, inplace=True)
        torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    print("events", ev, " ".join(f"{x:.0f}" for x in out), "ms; events located so far", int(r.event_counts.sum()) if ev else 0, "t range", r.t.min(), r.t.max())
