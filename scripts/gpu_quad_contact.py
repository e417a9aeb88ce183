"""First contact of an RK45 build with a GPU, torch-free and short (a few seconds): 4096-column lattice through the
host-pointer C ABI, 2000 warm attempts, then timed launches; counters and the first 64 columns are saved so that a second
process (another MARLPDE_RK45_BUILD) can be compared with the first.
    MARLPDE_RK45_BUILD=450 python scripts/gpu_quad_contact.py <out.npz> [<compare_with.npz>]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict

tag = os.environ.get("MARLPDE_RK45_BUILD", "default")
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P, y = mb.derive_column_params(pde), mb.initial_state(pde)
t0 = time.time()
r = mb.integrate_rk45_batch(y, P, t_span=(0, 1), first_step=1e-6, max_steps=2000)
print(f"[{tag}] warm: 2000 attempts/column in {time.time() - t0:.3f} s (includes context creation), status {np.unique(r.status)}", flush=True)
for i in range(2):
    before = int(r.n_attempts.sum())
    t0 = time.time()
    r = mb.integrate_rk45_batch(r.y, P, t_span=(0, 1), max_steps=1500, state=r.state)
    dt = time.time() - t0
    att = int(r.n_attempts.sum()) - before
    print(f"[{tag}] launch {i}: {att} attempts in {dt * 1e3:.1f} ms (host pointers, copies inside) -> {att / dt:.4e} col-steps/s", flush=True)
print(f"[{tag}] accepted {int(r.state['n_accepted'].sum())} rejected {int(r.state['n_rejected'].sum())} nfev {int(r.nfev.sum())} "
      f"t in [{r.t.min():.6g}, {r.t.max():.6g}] finite {bool(np.isfinite(r.y).all())}", flush=True)
np.savez(sys.argv[1], state=r.state, y=r.y[:64])
if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
    o = np.load(sys.argv[2])
    same = {k: bool(np.array_equal(o["state"][k], r.state[k])) for k in ("n_accepted", "n_rejected", "nfev", "status")}
    print(f"[{tag}] against {os.path.basename(sys.argv[2])}: counters identical {same}; max |dy| first 64 columns "
          f"{np.max(np.abs(o['y'] - r.y[:64])):.3e}; max |dt| {np.max(np.abs(o['state']['t'] - r.state['t'])):.3e}", flush=True)
