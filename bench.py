#!/usr/bin/env python
"""bench.py — column RK-steps/s of the batched RK45 integrator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--full]

Workload (BASELINE.json configs[1]): the synthetic 4096-column
Map_Scenario parameter sweep (16 x 16 x 16 lattice over sedimentation rate, b, D0co3), N=200
depth cells, RK45 rtol=atol=1e-3, first_step=1e-6.  One bench "step" = every column of the sweep
advances by `--attempts` adaptive RK45 step attempts (each attempt = 6 RHS evaluations + stage
algebra + error norm) in ONE launch of the persistent kernel, resuming where the previous step
stopped; the concatenation of steps is the integration to T*.  At N>1 GPUs every rank owns
8192 columns of the 32x32x64 lattice (65,536 columns at 8 GPUs, configs[2]); no collective on
the data path, one NCCL all-gather of the end states after the timed steps.

`value`  = step attempts of all columns / device time, inputs resident in HBM (CUDA events).
`e2e`    = same through the host-pointer C-ABI call (marlpde_rk45_integrate) with pinned HOST
           buffers: H2D of state+params, kernel, D2H of state, every step.
`roofline` is fp64-pipe based: achieved = value x 307,600 algorithmic flop per column-step
            against an fp64 FMA peak measured on the same device in this run.
`cpu_baseline` / `--impl reference`: the reference's own path — SciPy solve_ivp(RK45) driving the
           restated numba RHS (oracle/, kind "port": py-pde is not installable here) — on the
           host cores of the same box, one column per process, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200")
sys.path.insert(0, PKG)

FLOP_PER_COLUMN_STEP_PER_CELL = 1538          # N*(6*198 + 5*70)
METRIC = "column RK-steps/sec at N=200"
UNIT = "column-steps/s"


def scenario_base(name: str) -> dict:
    """asdict(Map_Scenario()) values (marlpde/parameters.py:16-143) via the host mirror."""
    from marlpde.parameters import Map_Scenario
    from dataclasses import asdict
    base = asdict(Map_Scenario())
    if name == "scenario_A":
        base |= {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
    return base


def lattice_for(world: int):
    return (16, 16, 16) if world == 1 else (32, 32, 64)


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.1 or ts > t1 + 0.1:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(job):
    """One column, SciPy RK45 on the restated numba RHS (what Evolve_scenario.py:104-109 does)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import lheureux_oracle as oracle
    pde, t_end, warm = job
    if warm:                                     # JIT compile outside the timed call
        oracle.integrate(pde, method="RK45", t_span=(0.0, 1e-5), t_eval=[0.0, 1e-5], events=False)
    t0 = time.perf_counter()
    sol = oracle.integrate(pde, method="RK45", first_step=1e-6, rtol=1e-3, atol=1e-3, t_span=(0.0, t_end),
                           t_eval=np.array([0.0, t_end]), events=True)
    return (sol.nfev - 1) // 6, time.perf_counter() - t0, sol.status


def cpu_reference_pass(base_name: str, t_end: float, cores: int, pool, warm: bool):
    """First `cores` columns of the sweep lattice, one process each, integrated to t_end."""
    import numpy as np
    import marlpde_b200 as mb
    pde = mb.sweep_lattice(scenario_base(base_name), 16, 16, 16)
    jobs = []
    for c in range(cores):
        one = {k: (float(v[c]) if np.ndim(v) else v) for k, v in pde.items()}
        jobs.append((one, t_end, warm))
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    attempts = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return attempts, busy if not warm else busy, wall


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t_end = args.cpu_t_end
    with ctx.Pool(cores) as pool:
        cpu_reference_pass(args.base, 1e-4, cores, pool, warm=True)             # compile + warm-up
        for _ in range(max(args.warmup - 1, 0)):
            cpu_reference_pass(args.base, t_end / 10, cores, pool, warm=False)
        tot_att, tot_t = 0, 0.0
        for _ in range(args.steps):
            att, busy, _wall = cpu_reference_pass(args.base, t_end, cores, pool, warm=False)
            tot_att += att
            tot_t += busy
    value = tot_att / tot_t
    sample = (f"first {cores} columns of the 16x16x16 lattice ({args.base} base), SciPy solve_ivp RK45 + numba RHS "
              f"(oracle port), t in [0,{t_end}] of T*, {tot_att // args.steps} step attempts per pass, one process per core")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world: int) -> dict:
    lat = lattice_for(world)
    per_gpu = 4096 if world == 1 else 8192
    return {"workload": (f"{per_gpu * world}-column Map_Scenario sweep ({'one column in every %d of the ' % (8 // world) if 1 < world < 8 else ''}{lat[0]}x{lat[1]}x{lat[2]} lattice over "
                         f"sedimentationrate, b, D0co3; {args.base} base), N=200, RK45 rtol=atol=1e-3, "
                         f"first_step=1e-6; {args.attempts} step attempts per column per bench step, resumed"),
            "columns": per_gpu * world, "columns_per_gpu": per_gpu, "n_cells": 200,
            "attempts_per_column_per_step": args.attempts, "parallelism": f"columns sharded x{world}, no data-path collective",
            "l2_policy": "256 MiB buffer written between timed steps (L2 flush); state is shared-memory resident"}


# --------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import marlpde_b200 as mb
    from marlpde_b200 import _cabi, batch
    import ctypes as C

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _cabi.lib()

    lat = lattice_for(world)
    pde = mb.sweep_lattice(scenario_base(args.base), *lat)
    if world > 1:
        # weak scaling: 8192 columns per GPU, taken as every (8/world)-th column of the 65,536-column lattice
        # so that every world size covers the full parameter ranges (world = 8: the whole lattice)
        n_all = lat[0] * lat[1] * lat[2]
        stride = max(1, n_all // (8192 * world))
        sel = np.arange(0, n_all, stride)[: 8192 * world]
        pde = {k: (np.asarray(v)[sel] if (k != "N" and np.ndim(v) == 1) else v) for k, v in pde.items()}
    P_all = mb.derive_column_params(pde)
    y_all = mb.initial_state(pde)
    from marlpde_b200 import sweep
    a, b = sweep.partition(P_all.shape[0], world)[rank]
    sl = slice(a, b)                                              # contiguous column block per rank
    P, y0 = P_all[sl], y_all[sl]
    B, N = y0.shape[0], y0.shape[2]

    d_params = batch.params_to_device(P, dev)
    d_y = torch.from_numpy(y0).to(dev)
    state = batch.make_state(B, 0.0, 1e-6)
    d_state = torch.from_numpy(state.view(np.uint8).copy()).to(dev)
    d_queue = torch.zeros(1, dtype=torch.int32, device=dev)
    d_ec = torch.zeros((B, 7), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # the 7 event monitors run inside the timed region, as in the reference's solve_ivp call
    # (Evolve_scenario.py:107-109 passes events=[...]); roots are located and stored (first EVCAP per monitor)
    EVCAP = 16
    d_et = torch.full((B, 7, EVCAP), float("nan"), dtype=torch.float64, device=dev)
    opts = _cabi.RK45Options(t_bound=1.0, rtol=1e-3, atol=1e-3, max_step=float("inf"), max_steps=args.attempts,
                             n_eval=0, event_capacity=EVCAP, flags=_cabi.FLAG_EVENTS, reserved=0)
    stream = torch.cuda.current_stream()

    def attempts_done():
        st = d_state.cpu().numpy().view(_cabi.STATE_DTYPE)
        return int(st["n_accepted"].sum() + st["n_rejected"].sum()), st

    def one_step():
        d_queue.zero_()
        _cabi.check(lib.marlpde_rk45_integrate_dev(d_y.data_ptr(), d_params.data_ptr(), d_state.data_ptr(), B, N,
                                                   C.byref(opts), None, None, d_ec.data_ptr(), d_et.data_ptr(),
                                                   d_queue.data_ptr(), stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
        flush.fill_(1)
    barrier()
    a0, _ = attempts_done()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    w0 = time.time()
    for s in range(args.steps):
        ev[s][0].record(stream)
        one_step()
        ev[s][1].record(stream)
        flush.fill_(s & 1)                                          # L2 flush between timed steps (not timed)
    barrier()
    w1 = time.time()
    a1, st = attempts_done()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = dev_ms / args.steps
    clocks = sampler.summary(w0, w1) if sampler else None

    # ---- e2e: the host-pointer C-ABI call with pinned host buffers, copies inside the timed region
    h_y = torch.from_numpy(d_y.cpu().numpy()).pin_memory()
    h_state = torch.from_numpy(st.view(np.uint8).copy()).pin_memory()
    h_params = torch.from_numpy(P.view(np.uint8).copy()).pin_memory()
    h_ec = torch.zeros((B, 7), dtype=torch.int32).pin_memory()
    h_et = torch.full((B, 7, EVCAP), float("nan"), dtype=torch.float64).pin_memory()
    hy, hs, hp, hec, het = h_y.numpy(), h_state.numpy(), h_params.numpy(), h_ec.numpy(), h_et.numpy()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        _cabi.check(lib.marlpde_rk45_integrate(hy.ctypes.data, hp.ctypes.data, hs.ctypes.data, B, N, C.byref(opts),
                                               None, None, hec.ctypes.data, het.ctypes.data, local_rank))
    e2e_step()                                                       # warm-up (allocations, first touch)
    barrier()
    b0 = int(hs.view(_cabi.STATE_DTYPE)["n_accepted"].sum() + hs.view(_cabi.STATE_DTYPE)["n_rejected"].sum())
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    b1 = int(hs.view(_cabi.STATE_DTYPE)["n_accepted"].sum() + hs.view(_cabi.STATE_DTYPE)["n_rejected"].sum())
    h2d = hy.nbytes + hp.nbytes + hs.nbytes + hec.nbytes + het.nbytes
    d2h = hy.nbytes + hs.nbytes + hec.nbytes + het.nbytes

    # ---- aggregate over ranks: max time, sum of work; one all-gather of end states (the only collective)
    att = torch.tensor([a1 - a0, b1 - b0], dtype=torch.float64, device=dev)
    tms = torch.tensor([dev_ms, e2e_s, w1 - w0], dtype=torch.float64, device=dev)
    gather_ms = None
    if dist is not None:
        dist.all_reduce(att, op=dist.ReduceOp.SUM)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        out = torch.empty((world * B, 5, N), dtype=torch.float64, device=dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        dist.all_gather_into_tensor(out, d_y)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
    total_attempts, e2e_attempts = float(att[0]), float(att[1])
    dev_s, e2e_s, wall_s = float(tms[0]) * 1e-3, float(tms[1]), float(tms[2])
    value = total_attempts / dev_s

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline: fp64 FMA peak measured on this device, now
    peak = C.c_double(0.0)
    _cabi.check(lib.marlpde_probe_fp64_peak(local_rank, 4096, 5, C.byref(peak)))
    flop_per_launch = (total_attempts / world / args.steps) * FLOP_PER_COLUMN_STEP_PER_CELL * N
    achieved = flop_per_launch / (kernel_ms * 1e-3) / 1e12
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
            "frac": achieved / peak.value if peak.value else None, "traffic": None,
            "peak_source": "measured in this run: marlpde_probe_fp64_peak (8 independent DFMA chains/thread, "
                           "2048 threads/SM, best of 5); MEASURED_PEAKS.json has no fp64 entry",
            "algorithmic_flop_per_column_step": FLOP_PER_COLUMN_STEP_PER_CELL * N,
            "kernel": "rk45_persistent_kernel", "kernel_ms_per_launch": kernel_ms,
            "hbm_bytes_per_launch_algorithmic": 2 * B * 5 * N * 8}
    prof = os.path.join(ROOT, "profiles", "roofline_latest.json")
    if os.path.exists(prof):
        try:
            with open(prof) as fh:
                roof["traffic"] = json.load(fh).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_attempts / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "marlpde_rk45_integrate (host pointers, pinned)"},
            "gpu_launches": args.steps, "roofline": roof, "clocks": clocks,
            "wall_s_timed_region": wall_s, "t_reached_min_max": [float(st["t"].min()), float(st["t"].max())],
            "accepted_rejected": [int(st["n_accepted"].sum()), int(st["n_rejected"].sum())]}
    if gather_ms is not None:
        line["allgather_end_states_ms"] = gather_ms

    if args.full and not args.skip_rk45_tstar:
        # time-to-T*: the whole sweep from t=0 to T* (BASELINE.json: "time-to-T* per 4096 columns")
        d_y2 = torch.from_numpy(y0).to(dev)
        d_state2 = torch.from_numpy(batch.make_state(B, 0.0, 1e-6).view(np.uint8).copy()).to(dev)
        # Columns that reach T* need 0.50-1.18 M attempts on this lattice.  A few columns run into a singularity
        # of the model near t = 0.5-0.7 and end with status -1 (step below 10 ulp, as SciPy does); how many
        # attempts they burn there is chaotic (0.7-6.8 M seen for the same column in two builds that differ by
        # FMA contraction only) and a lone column advances at ~50 k attempts/s, so the cap bounds the tail.
        o2 = _cabi.RK45Options(t_bound=1.0, rtol=1e-3, atol=1e-3, max_step=float("inf"), max_steps=args.step_cap, n_eval=0,
                               event_capacity=EVCAP, flags=_cabi.FLAG_EVENTS, reserved=0)
        d_queue.zero_()
        d_ec.zero_()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        _cabi.check(lib.marlpde_rk45_integrate_dev(d_y2.data_ptr(), d_params.data_ptr(), d_state2.data_ptr(), B, N,
                                                   C.byref(o2), None, None, d_ec.data_ptr(), d_et.data_ptr(),
                                                   d_queue.data_ptr(), stream.cuda_stream))
        f1.record(stream)
        torch.cuda.synchronize()
        st2 = d_state2.cpu().numpy().view(_cabi.STATE_DTYPE)
        secs = f0.elapsed_time(f1) * 1e-3
        tot = int(st2["n_accepted"].sum() + st2["n_rejected"].sum())
        line["time_to_Tstar"] = {"seconds": secs, "columns": B, "step_attempts": tot, "column_steps_per_s": tot / secs,
                                 "step_cap_per_column": args.step_cap,
                                 "finished": int((st2["status"] == 0).sum()),
                                 "status_histogram": {str(int(k)): int(v) for k, v in
                                                      zip(*np.unique(st2["status"], return_counts=True))},
                                 "events_located_per_monitor": d_ec.cpu().numpy().sum(axis=0).tolist(),
                                 "unfinished_columns": [[int(c), int(st2["status"][c]), float(st2["t"][c]),
                                                         float(st2["h_abs"][c])]
                                                        for c in np.nonzero(st2["status"] != 0)[0][:32]],
                                 "steps_per_column_min_max": [int((st2["n_accepted"] + st2["n_rejected"]).min()),
                                                              int((st2["n_accepted"] + st2["n_rejected"]).max())]}

    if args.full and world == 1:
        # ---- implicit path (BASELINE.json configs[4]): the same sweep to T* with the batched Radau IIA kernel
        implicit = {}
        for base_name in ("scenario_A", "default"):
            sw = mb.sweep_lattice(scenario_base(base_name), *lat)
            Pi, yi = mb.derive_column_params(sw)[sl], mb.initial_state(sw)[sl]
            d_yi = torch.from_numpy(yi).to(dev)
            d_pi = batch.params_to_device(Pi, dev)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            rr = mb.integrate_radau_batch(d_yi, d_pi, t_span=(0.0, 1.0), first_step=1e-6, rtol=1e-3, atol=1e-3,
                                          events=True, event_capacity=EVCAP, inplace=True)
            r1.record(stream)
            torch.cuda.synchronize()
            secs = r0.elapsed_time(r1) * 1e-3
            implicit[base_name] = {
                "seconds": secs, "columns": int(B), "finished": int((rr.status == 0).sum()),
                "steps_per_column_min_max": [int(rr.n_accepted.min()), int(rr.n_accepted.max())],
                "radau_steps": int(rr.n_accepted.sum()), "rejected": int(rr.n_rejected.sum()),
                "newton_iterations": int(rr.newton_iterations.sum()), "lu_factorisations": int(rr.nlu.sum()),
                "jacobians": int(rr.njev.sum()), "nfev": int(rr.nfev.sum()),
                "radau_steps_per_s": float(rr.n_accepted.sum() + rr.n_rejected.sum()) / secs}
            del d_yi, rr
        line["implicit_time_to_Tstar"] = implicit
        # ---- large depth grids (BASELINE.json configs[3]): streaming RK45, HBM roofline
        large = {}
        for nL, bL, attL in ((20000, 64, 96), (2000, 64, 256), (20000, 1, 256)):
            pL = scenario_base("scenario_A") | {"N": nL}
            PL = np.repeat(mb.derive_column_params(pL), bL)
            yL = torch.from_numpy(np.repeat(mb.initial_state(pL), bL, 0)).to(dev)
            dPL = batch.params_to_device(PL, dev)
            stL = batch.make_state(bL, 0.0, 1e-6 * (200 / nL) ** 2)
            d_stL = torch.from_numpy(stL.view(np.uint8).copy()).to(dev)
            nb = int(lib.marlpde_rk45_stream_workspace_bytes(bL, nL))
            d_wL = torch.empty(nb // 8 + 1, dtype=torch.float64, device=dev)
            oL = _cabi.RK45Options(t_bound=1.0, rtol=1e-3, atol=1e-3, max_step=float("inf"), max_steps=attL, n_eval=0,
                                   event_capacity=0, flags=0, reserved=0)

            def run_large():
                _cabi.check(lib.marlpde_rk45_stream_integrate_dev(yL.data_ptr(), dPL.data_ptr(), d_stL.data_ptr(), bL, nL,
                                                                  C.byref(oL), None, None, d_wL.data_ptr(), nb,
                                                                  stream.cuda_stream))
            run_large()                                            # warm-up (also moves past the first steps)
            torch.cuda.synchronize()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record(stream)
            run_large()
            l1.record(stream)
            torch.cuda.synchronize()
            secs = l0.elapsed_time(l1) * 1e-3
            tiles_mode = os.environ.get("MARLPDE_RK45_STREAM", "tiles")[0] != "s"
            # algorithmic HBM bytes per cell and attempt: overlapped tiles read y, K1 and write y_new, K7
            # (4 vector passes x 40 B; y ping-pongs between two buffers); one launch per stage moves 42 passes
            per_cell = 160.0 if tiles_mode else 1680.0
            alg_bytes = per_cell * nL * bL * attL
            flops = FLOP_PER_COLUMN_STEP_PER_CELL * float(nL) * bL * attL
            large[f"N{nL}_B{bL}"] = {"n_cells": nL, "columns": bL, "attempts_per_column": attL, "seconds": secs,
                                     "mode": "overlapped tiles (1 launch per attempt)" if tiles_mode else "1 launch per stage",
                                     "column_steps_per_s": bL * attL / secs, "cell_steps_per_s": bL * attL * nL / secs,
                                     "algorithmic_GBps": alg_bytes / secs / 1e9,
                                     "frac_of_hbm_peak": alg_bytes / secs / 1e9 / 6550.4,
                                     "algorithmic_TFLOPs": flops / secs / 1e12,
                                     "frac_of_fp64_peak": flops / secs / 1e12 / peak.value if peak.value else None,
                                     "working_set_MB": nb / 1e6, "launches": (2 if tiles_mode else 7) * attL + 4}
            del yL, d_wL
        line["large_n_streaming"] = large

    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_pass(args.base, 1e-4, cores, pool, warm=True)
            catt, cbusy, _ = cpu_reference_pass(args.base, args.cpu_t_end, cores, pool, warm=False)
        line["cpu_baseline"] = {
            "value": catt / cbusy, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"first {cores} columns of the lattice, SciPy solve_ivp RK45 + numba RHS (oracle port of the "
                       f"reference path; py-pde not installable), t in [0,{args.cpu_t_end}] of T*, {catt} step attempts, "
                       f"one process per core, {cbusy:.1f} s")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--base", default="default", choices=["default", "scenario_A"])
    ap.add_argument("--attempts", type=int, default=3000, help="RK45 step attempts per column per bench step")
    ap.add_argument("--full", action="store_true", help="also time the whole sweep from t=0 to T*")
    ap.add_argument("--step-cap", type=int, default=2_000_000,
                    help="--full: step-attempt cap per column of the sweep to T* (SURVEY.md 8d, config 2: 'give every "
                         "column a step cap + status'); 0 = none")
    ap.add_argument("--skip-rk45-tstar", action="store_true",
                    help="--full without the ~170 s explicit sweep to T* (implicit sweep and large-N lines only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-t-end", type=float, default=0.03, help="CPU sample: integrate to this fraction of T*")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
