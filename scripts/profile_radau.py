"""Short Radau run for ncu: lattice columns of the default base, t in [0, t_end].
    python scripts/profile_radau.py [n_lattice] [t_end]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
base = asdict(Map_Scenario())
pde = mb.sweep_lattice(base, n, n, n)
P = mb.derive_column_params(pde)
y = torch.from_numpy(mb.initial_state(pde)).cuda()
dP = mb.batch.params_to_device(P, y.device)
mb.integrate_radau_batch(y[:2], mb.batch.params_to_device(P[:2], y.device), t_span=(0, 1e-4), first_step=1e-6)   # warm-up
torch.cuda.synchronize(); t0 = time.time()
r = mb.integrate_radau_batch(y, dP, t_span=(0, t_end), first_step=1e-6)
torch.cuda.synchronize(); dt = time.time() - t0
print(f"{n**3} columns to t={t_end}: {dt:.3f}s steps {r.n_accepted.min()}-{r.n_accepted.max()} nlu {r.nlu.sum()} newton {r.newton_iterations.sum()} nfev {r.nfev.sum()} njev {r.njev.sum()}")
print(f"per warp-sequential: {dt / max(1, -(-n**3 // (148*16))):.3f}s ; LU pairs/s {r.nlu.sum()/2/dt:.3e}; newton it/s {r.newton_iterations.sum()/dt:.3e}")
if os.environ.get("MARLPDE_LPT"):
    # longest-processing-time-first: same columns, ordered by the work they turned out to need
    import marlpde_b200.sweep as sw
    order = np.argsort(-(r.nlu + 2 * r.newton_iterations), kind="stable")
    pred = np.argsort(-sw.predicted_cost(pde), kind="stable")
    for name, od in (("oracle-LPT", order), ("predicted-LPT", pred)):
        y2 = torch.from_numpy(mb.initial_state(pde)[od]).cuda()
        dP2 = mb.batch.params_to_device(P[od], y.device)
        torch.cuda.synchronize(); t0 = time.time()
        r2 = mb.integrate_radau_batch(y2, dP2, t_span=(0, t_end), first_step=1e-6)
        torch.cuda.synchronize(); print(name, f"{time.time() - t0:.3f}s")
    # throughput without imbalance: every column = the heaviest one
    hv = int(order[0])
    y3 = torch.from_numpy(np.repeat(mb.initial_state(pde)[hv:hv + 1], n ** 3, 0)).cuda()
    dP3 = mb.batch.params_to_device(np.repeat(P[hv:hv + 1], n ** 3), y.device)
    torch.cuda.synchronize(); t0 = time.time()
    r3 = mb.integrate_radau_batch(y3, dP3, t_span=(0, t_end), first_step=1e-6)
    torch.cuda.synchronize(); print("all-heaviest", f"{time.time() - t0:.3f}s", "nlu each", int(r3.nlu[0]), "newton each", int(r3.newton_iterations[0]))
    print("work spread: nlu min/mean/max", r.nlu.min(), r.nlu.mean(), r.nlu.max())
