"""Where does the host-pointer call spend its time? fixed overhead (max_steps=1) vs the kernel."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde_b200 import _cabi, batch
from marlpde.parameters import Map_Scenario
from dataclasses import asdict
pde = mb.sweep_lattice(asdict(Map_Scenario()), 16, 16, 16)
P = mb.derive_column_params(pde); y0 = mb.initial_state(pde); B, N = 4096, 200
lib = _cabi.lib()
hy = torch.from_numpy(y0.copy()).pin_memory().numpy(); hp = torch.from_numpy(P.view(np.uint8).copy()).pin_memory().numpy()
hs = torch.from_numpy(batch.make_state(B, 0.0, 1e-6).view(np.uint8).copy()).pin_memory().numpy()
hec = torch.zeros((B, 7), dtype=torch.int32).pin_memory().numpy(); het = torch.full((B, 7, 16), float("nan"), dtype=torch.float64).pin_memory().numpy()
for steps in (1, 1, 300, 3000, 3000):
    o = _cabi.RK45Options(t_bound=1.0, rtol=1e-3, atol=1e-3, max_step=float("inf"), max_steps=steps, n_eval=0, event_capacity=16, flags=1, quantum=0)
    t0 = time.perf_counter()
    _cabi.check(lib.marlpde_rk45_integrate(hy.ctypes.data, hp.ctypes.data, hs.ctypes.data, B, N, C.byref(o), None, None, hec.ctypes.data, het.ctypes.data, 0))
    print(f"host call, {steps} attempts/column: {1e3*(time.perf_counter()-t0):.1f} ms")
# device path for comparison
dy = torch.from_numpy(y0).cuda(); dP = batch.params_to_device(P, dy.device)
r = mb.integrate_rk45_batch(dy, dP, t_span=(0, 1), first_step=1e-6, max_steps=3000, events=True, event_capacity=16)
torch.cuda.synchronize(); t0 = time.perf_counter()
r = mb.integrate_rk45_batch(r.y, dP, t_span=(0, 1), max_steps=3000, state=r.state, events=True, event_capacity=16)
torch.cuda.synchronize(); print(f"device call, 3000 attempts/column: {1e3*(time.perf_counter()-t0):.1f} ms")
