"""Paired kernel shape of the on-chip RK45 kernel (cluster of two CTAs sharing a 7th column) against the classic shape:
the same resumed bench step (4096 / 8192 / 444 / 64 columns x `attempts` step attempts, events on), bit-for-bit comparison
of every output and CUDA-event timings.   python scripts/exp_rk45_pair.py [attempts] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch
import marlpde_b200 as mb
from marlpde.parameters import Map_Scenario
from dataclasses import asdict

attempts = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = asdict(Map_Scenario())


def run(shape, lat, te=None):
    os.environ["MARLPDE_RK45_SHAPE"] = shape
    pde = mb.sweep_lattice(base, *lat)
    P = mb.derive_column_params(pde)
    y = torch.from_numpy(mb.initial_state(pde)).cuda()
    dP = mb.batch.params_to_device(P, y.device)
    r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), first_step=1e-6, max_steps=2000, events=True, event_capacity=16)
    state, y = r.state, r.y
    times, att = [], 0
    for i in range(launches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = mb.integrate_rk45_batch(y, dP, t_span=(0, 1), max_steps=attempts, state=state, events=True, event_capacity=16,
                                    t_eval=te)
        e1.record()
        torch.cuda.synchronize()
        att = int(r.n_attempts.sum() - (state["n_accepted"].sum() + state["n_rejected"].sum()))
        times.append(e0.elapsed_time(e1))
        state, y = r.state, r.y
    best = min(times)
    print(f"  {shape:5s} {lat}: {att} attempts, best {best:.1f} ms of {['%.1f' % t for t in times]} -> {att / best * 1e-3:.3f} M column-steps/s",
          flush=True)
    return r, att / best * 1e-3


lattices = os.environ.get("MARLPDE_EXP_LATTICES", "16,16,16;16,16,32;4,4,4;6,6,12;7,8,8")
for lat in [tuple(int(v) for v in l.split(",")) for l in lattices.split(";")]:
    a, ra = run("solo", lat)
    b, rb = run("pair", lat)
    # (nfev is left out: the two shapes have different slot counts, hence different quanta, and every resumed quantum
    #  re-evaluates K1 once)
    diff = {"y": int((a.y != b.y).any(dim=2).any(dim=1).sum())}
    for k in a.state.dtype.names:
        if k != "nfev":
            diff[k] = int(np.sum(a.state[k] != b.state[k]))
    diff["event_counts"] = int(np.sum(a.event_counts != b.event_counts))
    diff["event_times"] = int(np.sum(~((a.event_times == b.event_times) | (np.isnan(a.event_times) & np.isnan(b.event_times)))))
    same = not any(diff.values())
    print(f"lattice {lat}: pair / solo = {rb / ra:.4f}, bit-identical: {same}" + ("" if same else f" — differing entries {diff}"), flush=True)
