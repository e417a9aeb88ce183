"""Short RK45 run for ncu: 4096-column sweep, a few hundred step attempts per launch.
    python scripts/profile_rk45.py [attempts] [launches] [base]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict

attempts = int(sys.argv[1]) if len(sys.argv) > 1 else 300
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = asdict(Map_Scenario())
if len(sys.argv) > 3 and sys.argv[3] == "scenario_A":
    base |= {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
lat = [int(v) for v in os.environ.get("MARLPDE_PROFILE_LATTICE", "16,16,16").split(",")]   # e.g. 4,4,4: a batch smaller than the machine
pde = mb.sweep_lattice(base, *lat)
P = mb.derive_column_params(pde)
y = torch.from_numpy(mb.initial_state(pde)).cuda()
dP = mb.batch.params_to_device(P, y.device)
# get past the initial transient first (not profiled when ncu uses -s 1)
r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), first_step=1e-6, max_steps=2000)
state, y = r.state, r.y
ev = bool(int(os.environ.get("MARLPDE_PROFILE_EVENTS", "0")))
for i in range(launches):
    torch.cuda.synchronize(); t0 = time.time()
    r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), max_steps=attempts, state=state, events=ev, event_capacity=16,
                                quantum=int(os.environ.get("MARLPDE_PROFILE_QUANTUM", "0")))
    torch.cuda.synchronize(); dt = time.time() - t0
    att = int(r.n_attempts.sum() - (state["n_accepted"].sum() + state["n_rejected"].sum()))
    print(f"launch {i}: {att} attempts in {dt*1e3:.1f} ms -> {att/dt:.3e} col-steps/s")
    state, y = r.state, r.y
