"""bench.py contract on CPU: the reference arm (`--impl reference`) runs the oracle port of the reference's
own path (SciPy RK45 + numba RHS, events on) on the host cores and prints ONE JSON line with the keys the
driver reads; workload descriptions are consistent across world sizes.  (The GPU arm is exercised on the
B200 box by the driver; it needs a CUDA device.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-t-end", "2e-4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "column RK-steps/sec at N=200" and d["unit"] == "column-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "4096-column" in d["config"]["workload"] and d["config"]["n_cells"] == 200


def test_workload_descriptions_per_world_size():
    sys.path.insert(0, ROOT)
    import bench
    from types import SimpleNamespace
    args = SimpleNamespace(base="default", attempts=3000)
    for world, cols in ((1, 4096), (2, 16384), (4, 32768), (8, 65536)):
        cfg = bench.workload_config(args, world)
        assert cfg["columns"] == cols and cfg["columns_per_gpu"] == cols // world and f"{cols}-column" in cfg["workload"]
    assert bench.lattice_for(1) == (16, 16, 16) and bench.lattice_for(8) == (32, 32, 64)
    assert bench.FLOP_PER_COLUMN_STEP_PER_CELL * 200 == 307600
