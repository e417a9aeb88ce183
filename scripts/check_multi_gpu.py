"""torchrun --nproc-per-node G scripts/check_multi_gpu.py : sharded sweep over G GPUs (NCCL all-gather of the
results) must equal the same sweep integrated by one GPU; prints one line per check on rank 0."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
import numpy as np, torch, torch.distributed as dist
import marlpde_b200 as mb
from marlpde_b200 import sweep
from marlpde.parameters import Map_Scenario
from dataclasses import asdict

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
base = asdict(Map_Scenario()) | {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}
pde = mb.sweep_lattice(base, 6, 5, 4)                       # 120 columns, not a multiple of the world size for G=8... fine
te = [0.0, 5e-4, 1e-3]
for balance in (True, False):
    t0 = time.time()
    res = sweep.sweep_rk45(pde, t_span=(0, 1e-3), first_step=1e-6, t_eval=te, events=True, event_capacity=8, balance=balance)
    dt = time.time() - t0
    if rank == 0:
        ser = mb.integrate_rk45_batch(mb.initial_state(pde), mb.derive_column_params(pde), t_span=(0, 1e-3), first_step=1e-6,
                                      t_eval=te, events=True, event_capacity=8)
        ok = (np.array_equal(res.y, ser.y) and np.array_equal(res.snapshots, ser.snapshots) and np.array_equal(res.nfev, ser.nfev)
              and np.array_equal(res.status, ser.status) and np.array_equal(res.event_counts, ser.event_counts))
        print(f"world {world} balance {balance}: sharded == serial: {ok}; owners {np.bincount(res.owner, minlength=world).tolist()}; {dt:.2f}s", flush=True)
        assert ok
dist.barrier()
dist.destroy_process_group()
