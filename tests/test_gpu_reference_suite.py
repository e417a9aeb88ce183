"""The reference's OWN files, unmodified, against the GPU path.

`__graft_entry__.build()` copies the reference's regression test, its three HDF5 fixtures and its driver /
parameter modules byte for byte from /root/reference into oracle/_ref/refsrc (git-ignored; it travels to
the GPU box, where /root/reference does not exist).  tests/shims/ makes their `h5py`, `pde`, `pint` and
`matplotlib` imports resolve (none of those packages is installed in this image).

  L-b  tests/Regression_test/test_regression.py (reference, unmodified) + the repo's `marlpde` package:
       `integrate_equations(asdict(Solver()), ...)` requests method="Radau" (parameters.py:213) and gets the
       batched implicit CUDA kernel; all three reference tests must pass (rtol 0.1 / atol 0.01; Matlab atol 0.05).
  L-a  the reference's own marlpde/Evolve_scenario.py + marlpde/parameters.py (unmodified) with the repo's
       LHeureux_model.py dropped in beside them: SciPy's Radau steps (reference jac_sparsity, 21-colour FD
       Jacobian), every RHS call is the CUDA kernel through eq.fun_numba, output written through `h5py.File`.
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200")
SHIMS = os.path.join(ROOT, "tests", "shims")
REFSRC = os.path.join(ROOT, "oracle", "_ref", "refsrc")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REFSRC, "tests", "Regression_test", "test_regression.py")),
                                 reason="oracle/_ref/refsrc missing: run __graft_entry__.build() where /root/reference exists")]


def _run_reference_tests(cwd, pythonpath, extra=()):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-rA",
           os.path.join("tests", "Regression_test", "test_regression.py"), *extra]
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)


def _stage_tests(tmp_path):
    work = tmp_path / "checkout"
    shutil.copytree(os.path.join(REFSRC, "tests"), work / "tests")
    return work


def test_reference_regression_suite_unmodified_on_gpu_radau(tmp_path):
    """L-b: 3/3 of the reference's regression tests through the repo's marlpde.integrate_equations."""
    work = _stage_tests(tmp_path)
    done = _run_reference_tests(str(work), [SHIMS, PKG])
    log = done.stdout + done.stderr
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "reference_suite_Lb.log"), "w") as fh:
            fh.write(log)
    assert done.returncode == 0, log[-4000:]
    assert "3 passed" in done.stdout, log[-4000:]
    # the stepping really was the CUDA Radau kernel, not SciPy: the mirror prints LU counts of the batched kernel
    assert "Number of LU decompositions" in log


def test_reference_driver_unmodified_with_cuda_rhs(tmp_path):
    """L-a: reference Evolve_scenario.py + parameters.py + SciPy Radau, RHS on the GPU (Scenario A test)."""
    work = _stage_tests(tmp_path)
    pkg = work / "marlpde"
    pkg.mkdir()
    for name in ("__init__.py", "Evolve_scenario.py", "parameters.py"):
        shutil.copy2(os.path.join(REFSRC, "marlpde", name), pkg / name)
    shutil.copy2(os.path.join(PKG, "marlpde", "LHeureux_model.py"), pkg / "LHeureux_model.py")
    # `python -m pytest` puts the working directory first on sys.path: `marlpde` is the staged package
    done = _run_reference_tests(str(work), [SHIMS, PKG], extra=("-k", "test_integration_Scenario_A"))
    log = done.stdout + done.stderr
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "reference_suite_La.log"), "w") as fh:
            fh.write(log)
    assert done.returncode == 0, log[-4000:]
    assert "1 passed" in done.stdout, log[-4000:]
