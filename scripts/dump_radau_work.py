"""Per-column work of the implicit sweep to T* on the 16^3 lattice (default base) -> gpurun_out/radau_work.npz"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P = mb.derive_column_params(pde)
y = torch.from_numpy(mb.initial_state(pde)).cuda()
r = mb.integrate_radau_batch(y, P, t_span=(0, 1), first_step=1e-6, events=True, event_capacity=16, inplace=True,
                             t_eval=[0.02, 0.05, 0.1, 0.2])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "radau_work.npz"), work=r.work, steps=r.n_accepted, status=r.status,
                    nlu=r.nlu, newton=r.newton_iterations, fails=r.newton_failures, t=r.t,
                    phi_max=np.stack([r.snapshots[:, k, 4, :].max(dim=1).values.cpu().numpy() for k in range(4)], 1))
print("ok", r.work.sum())
