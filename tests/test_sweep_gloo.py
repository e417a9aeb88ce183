"""Host logic of multi-GPU sweeps (marlpde_b200/sweep.py) on CPU: world_size-2 gloo processes run the
real sharding / all-gather code around a stand-in per-rank integrator (the CUDA integrator itself
is covered by the `-m gpu` tests; columns are independent, so this is all the N>1 path adds)."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest

import lheureux_oracle as oracle
import marlpde_b200 as mb
from marlpde_b200 import sweep


def _fake_integrate(y0, P, t_span=(0, 1), t_eval=None, **kw):
    """Deterministic function of (state, parameters) with the shape of an RK45Result."""
    y0 = np.asarray(y0)
    B = y0.shape[0]
    te = np.zeros(0) if t_eval is None else np.asarray(t_eval, dtype=float)
    y = y0 + P["presum"][:, None, None] * 1e-3 + P["dCO3"][:, None, None]
    snaps = np.stack([y0 + tt * P["Da"][:, None, None] for tt in te], axis=1) if te.size else np.zeros((B, 0) + y0.shape[1:])
    steps = (P["dCO3"] * 1e3).astype(np.int64)
    ec = np.stack([steps % (k + 2) for k in range(7)], axis=1)
    return SimpleNamespace(y=y, snapshots=snaps, t=np.full(B, t_span[1]), status=(steps % 3 == 0).astype(np.int32),
                           n_accepted=steps, n_rejected=steps // 7, nfev=1 + 6 * (steps + steps // 7), event_counts=ec)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, balance, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pde = mb.sweep_lattice(oracle.default_scenario(), *shape)
        res = sweep.sweep_rk45(pde, t_span=(0, 1), t_eval=[0.0, 0.5, 1.0], balance=balance, device="cpu",
                               integrate=_fake_integrate)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **{k: getattr(res, k) for k in
                                                          ("y", "snapshots", "t", "status", "n_accepted",
                                                           "n_rejected", "nfev", "event_counts", "owner")})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,balance", [((3, 2, 2), True), ((3, 1, 3), False), ((1, 1, 1), True)])
def test_two_rank_sweep_equals_serial(tmp_path, shape, balance):
    import torch.multiprocessing as tmp
    world = 2
    tmp.spawn(_worker, args=(world, _free_port(), shape, balance, str(tmp_path)), nprocs=world, join=True)
    pde = mb.sweep_lattice(oracle.default_scenario(), *shape)
    serial = sweep.sweep_rk45(pde, t_span=(0, 1), t_eval=[0.0, 0.5, 1.0], device="cpu", integrate=_fake_integrate)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    B = shape[0] * shape[1] * shape[2]
    for k in ("y", "snapshots", "t", "status", "n_accepted", "n_rejected", "nfev", "event_counts"):
        assert np.array_equal(got[0][k], getattr(serial, k)), k            # sharded == serial, original order
        assert np.array_equal(got[0][k], got[1][k]), k                     # every rank holds the full result
    owner = got[0]["owner"]
    assert owner.shape == (B,) and abs(int((owner == 0).sum()) - int((owner == 1).sum())) <= 1


def test_partition_and_balanced_assignment():
    assert sweep.partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sweep.partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sweep.partition(65536, 8)[7] == (57344, 65536)
    pde = mb.sweep_lattice(oracle.default_scenario(), 32, 4, 8)
    cost = sweep.predicted_cost(pde)
    assert cost.shape == (1024,) and np.all(cost > 0)
    # cost ~ dCO3 * Xstar^2: grows with D0co3, falls with the sedimentation rate
    assert cost[7] > cost[0] and cost[0] > cost[-8]
    parts = sweep.balanced_assignment(cost, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(1024))
    assert {len(p) for p in parts} == {128}
    totals = np.array([cost[p].sum() for p in parts])
    blocks = np.array([cost[a:b].sum() for a, b in sweep.partition(1024, 8)])
    assert totals.max() / totals.min() < 1.001 < blocks.max() / blocks.min()   # contiguous blocks are 1.4x apart
    sub = sweep.shard(pde, parts[3])
    assert sub["Xstar"].shape == (128,) and sub["N"] == 200 and np.isscalar(sub["D0Ca"])


def test_predicted_cost_orders_the_implicit_sweep_by_compaction_coefficient():
    """The a-priori cost estimates only order columns: RK45 by the diffusive stability limit, Radau by the empirical
    b^2.16 S^-0.66 law measured on the benchmark lattice (sweep.predicted_cost)."""
    pde = mb.sweep_lattice(oracle.default_scenario(), 4, 4, 4)
    ce, ci = sweep.predicted_cost(pde), sweep.predicted_cost(pde, "Radau")
    assert ce.shape == ci.shape == (64,)
    b, srate = np.asarray(pde["b"]), np.asarray(pde["sedimentationrate"])
    heavy, light = np.argmax(ci), np.argmin(ci)
    assert b[heavy] == b.max() and srate[heavy] == srate.min() and b[light] == b.min() and srate[light] == srate.max()
    assert np.corrcoef(ce, ci)[0, 1] < 0.9                   # the explicit estimate follows DCO3 and Xstar instead
    assign = sweep.balanced_assignment(ci, 2)
    assert abs(ci[assign[0]].sum() / ci[assign[1]].sum() - 1.0) < 0.02
