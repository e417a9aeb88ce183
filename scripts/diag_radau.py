"""Diagnostics: GPU Radau vs SciPy Radau (1e-3) vs tight SciPy Radau (1e-8), default scenario, t in [0, 0.05]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import lheureux_oracle as oracle, marlpde_b200 as mb
np.seterr(all="ignore")
pde = oracle.default_scenario()
te = np.array([0.0, 0.01, 0.02, 0.03, 0.05])
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
t0 = time.time(); res = mb.integrate_radau_batch(y0, P, t_span=(0, 0.05), first_step=5e-7, t_eval=te); tg = time.time() - t0
sp = oracle.jacobian_sparsity(200)
t0 = time.time(); sol = oracle.integrate(pde, method="Radau", t_span=(0, 0.05), t_eval=te, events=False, first_step=5e-7, jac_sparsity=sp); ts = time.time() - t0
ref = oracle.integrate(pde, method="Radau", t_span=(0, 0.05), t_eval=te, events=False, first_step=5e-7, jac_sparsity=sp, rtol=1e-8, atol=1e-8)
g, s, r = res.solutions(0), sol.y.reshape(5, 200, -1), ref.y.reshape(5, 200, -1)
print("gpu: status", res.status, "acc", res.n_accepted, "rej", res.n_rejected, "nfev", res.nfev, "njev", res.njev, "nlu", res.nlu, "newton", res.newton_iterations, res.newton_failures, f"{tg:.2f}s")
print("scipy: nfev", sol.nfev, "njev", sol.njev, "nlu", sol.nlu, f"{ts:.2f}s", "steps", "n/a")
for k, t in enumerate(te):
    sc = 1e-3 + 1e-3 * np.abs(r[:, :, k])
    print(f"t={t}: |gpu-scipy|/sc {np.max(np.abs(g[:,:,k]-s[:,:,k])/sc):.3f}  |gpu-ref|/sc {np.max(np.abs(g[:,:,k]-r[:,:,k])/sc):.3f}  |scipy-ref|/sc {np.max(np.abs(s[:,:,k]-r[:,:,k])/sc):.3f}")
# 64-column sweep timing to T*
for base_name, base in (("scenario_A", pde | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}), ("default", pde)):
    sw = mb.sweep_lattice(base, 4, 4, 4)
    import torch
    y = torch.from_numpy(mb.initial_state(sw)).cuda()
    torch.cuda.synchronize(); t0 = time.time()
    rr = mb.integrate_radau_batch(y, mb.derive_column_params(sw), t_span=(0, 1), first_step=1e-6)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(base_name, "64 columns to T*:", f"{dt:.2f}s", "status", np.unique(rr.status, return_counts=True), "steps", rr.n_accepted.min(), rr.n_accepted.max(),
          "nlu", rr.nlu.min(), rr.nlu.max(), "newton", rr.newton_iterations.max())
# event times, scenario A to T*
pa = pde | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
ra = mb.integrate_radau_batch(mb.initial_state(pa), mb.derive_column_params(pa), t_span=(0, 1), first_step=1e-6, events=True, event_capacity=8)
sa = oracle.integrate(pa, method="Radau", t_span=(0, 1), t_eval=[0, 1], events=True, jac_sparsity=sp)
print("gpu events", ra.event_counts[0], ra.event_times[0, 0], ra.event_times[0, 1]); print("scipy events", [list(e) for e in sa.t_events])
