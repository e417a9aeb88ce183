"""Scratch GPU check used during development (not part of the test-suite)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import marlpde_b200 as mb
import lheureux_oracle as o

np.seterr(all="ignore")
g = np.load(os.path.join(ROOT, "tests/golden/rhs_reference.npz"))
meta = json.loads(str(g["__meta__"]))
worst = 0
for name, pde in meta.items():
    P = mb.derive_column_params(pde)
    po = o.kernel_params(pde)
    for key in [k for k in g.files if k.startswith(name + "/") and k.endswith("/y")]:
        y = g[key]; ref = g[key[:-2] + "/rhs_numba"]
        out = mb.rhs_batch(y.reshape(1, 5, -1), P).ravel()
        S = o.term_scale(y, po, np.empty_like(y))
        e = np.nanmax(np.abs(out - ref) / S)
        rel = np.nanmax(np.abs(out - ref)) / np.nanmax(np.abs(ref))
        worst = max(worst, e)
        print(f"{key:35s} |d|/S={e:.2e} inf-rel={rel:.2e} nanmismatch={(np.isnan(out)!=np.isnan(ref)).sum()}")
print("worst |d|/S", worst)

gs = np.load(os.path.join(ROOT, "tests/golden/stepper_reference.npz"))
cases = json.loads(str(gs["__cases__"]))
for name, c in cases.items():
    pde = o.default_scenario() | c["overrides"]
    P = mb.derive_column_params(pde); y0 = mb.initial_state(pde)
    key = f"{name}/RK45/t0.002/tol0.001"
    te = gs[key + "/t"]
    t0 = time.time()
    r = mb.integrate_rk45_batch(y0, P, t_span=(0, 0.002), first_step=c["first_step"], t_eval=te)
    dt = time.time() - t0
    ref = gs[key + "/y"].reshape(5, 200, -1)
    mine = r.solutions(0)
    print(name, "status", r.status, "nfev", r.nfev, "ref nfev", gs[key + "/counts"][0], "acc/rej", r.n_accepted, r.n_rejected,
          "max|d| per t", np.abs(mine - ref).max(axis=(0, 1)), f"{dt:.3f}s")

# throughput probe: 444 columns, 3000 attempts each
import torch
base = o.default_scenario() | cases["scenario_A"]["overrides"]
for ncol, (a, b, c_) in ((512, (8, 8, 8)), (4096, (16, 16, 16))):
    pde = mb.sweep_lattice(base, a, b, c_)
    P = mb.derive_column_params(pde); y0 = torch.from_numpy(mb.initial_state(pde)).cuda()
    dP = mb.batch.params_to_device(P, y0.device)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        r = mb.integrate_rk45_batch(y0, dP, t_span=(0, 1), first_step=1e-6, max_steps=3000)
        torch.cuda.synchronize(); dt = time.time() - t0
        att = int(r.n_attempts.sum())
        print(f"cols={ncol} attempts={att} time={dt:.4f}s  col-steps/s={att/dt:.3e} status={np.unique(r.status)} t_range=({r.t.min():.3e},{r.t.max():.3e})")
