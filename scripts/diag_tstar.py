"""Phase-resolved throughput of the RK45 sweep to T*: the columns are integrated in chunks of step attempts
(resumed), and every chunk prints its rate, the time window reached, the events located and the SM clock.
    python scripts/diag_tstar.py [columns] [attempts_per_chunk] [base]
MARLPDE_B200_LIB selects the library (A/B of two builds on the same box)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 888
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
base = asdict(Map_Scenario())
if len(sys.argv) > 3 and sys.argv[3] == "scenario_A":
    base |= {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
pde = mb.sweep_lattice(base, 16, 16, 16)
P = mb.derive_column_params(pde)
y0 = mb.initial_state(pde)
sel = np.linspace(0, len(P) - 1, ncol).astype(int)
P, y0 = P[sel], y0[sel]
y = torch.from_numpy(y0).cuda()
dP = mb.batch.params_to_device(P, y.device)


def clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active",
                               "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return repr(e)


state, done_prev, ev_prev, total_s = None, 0, 0, 0.0
print(f"lib={os.environ.get('MARLPDE_B200_LIB', 'in-tree')} columns={ncol} chunk={chunk}", flush=True)
for it in range(400):
    torch.cuda.synchronize(); t0 = time.time()
    r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), first_step=1e-6, max_steps=chunk, state=state, events=True,
                                event_capacity=16)
    torch.cuda.synchronize(); dt = time.time() - t0
    state, y = r.state, r.y
    done = int(r.n_attempts.sum())
    ev = int(np.asarray(r.event_counts).sum()) if getattr(r, "event_counts", None) is not None else -1
    st = np.asarray(state["status"]); tt = np.asarray(state["t"])
    running = int((st > 0).sum())
    total_s += dt
    print(f"chunk {it:3d}: {done - done_prev:>10d} attempts {dt:7.3f} s  {(done - done_prev) / dt:.3e}/s  running {running:5d} "
          f"t [{tt.min():.4f}, {np.median(tt):.4f}, {tt.max():.4f}] events +{ev - ev_prev}  | {clock()}", flush=True)
    done_prev, ev_prev = done, ev
    if running == 0:
        break
print(f"total {done} attempts in {total_s:.2f} s -> {done / total_s:.3e}/s; status {np.unique(st, return_counts=True)}")
