"""Accuracy of the kernels' table-driven fp64 maths (csrc/fp64_math.cuh) against mpmath-free
references (numpy's libm, which is correctly rounded to ~1 ulp for these functions)."""
import numpy as np
import pytest

from marlpde_b200 import _cabi

pytestmark = pytest.mark.gpu
np.seterr(all="ignore")


def _probe(op, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _cabi.check(_cabi.lib().marlpde_probe_math(op, x.ctypes.data, x.size, out.ctypes.data, 0))
    return out


def test_log_absolute_accuracy_and_special_values():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e-3, 3.0, 200000), 10.0 ** rng.uniform(-300, 300, 50000),
                        1 + rng.uniform(-1e-3, 1e-3, 50000), [1.0, 0.5, 2.0, np.nextafter(1, 0), np.nextafter(1, 2)]])
    got, ref = _probe(0, x), np.log(x)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 2.5e-16
    sp = _probe(0, np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3])
    assert abs(sp[4] - np.log(5e-324)) < 1e-12 and abs(sp[5] - np.log(1e-310)) < 1e-12


def test_exp_and_expm1_relative_accuracy():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-40, 40, 200000), rng.uniform(-689, 689, 50000), rng.uniform(-1, 1, 50000)])
    got, ref = _probe(1, x), np.exp(x)
    assert np.max(np.abs(got - ref) / ref) <= 4e-16
    sp = _probe(1, np.array([-np.inf, np.inf, np.nan, -800.0, 710.0, 0.0]))
    assert sp[0] == 0 and sp[1] == np.inf and np.isnan(sp[2]) and sp[3] == 0 and sp[4] == np.inf and sp[5] == 1.0
    x = np.concatenate([rng.uniform(-200, 200, 100000), rng.uniform(-1, 1, 100000), rng.uniform(-0.03, 0.03, 50000)])
    x = x[np.abs(x) >= 1e-3]
    got, ref = _probe(2, x), np.expm1(x)
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 6e-16


def test_reciprocal_and_division():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(1e-3, 10, 100000), -rng.uniform(1e-3, 10, 100000), 10.0 ** rng.uniform(-200, 200, 50000)])
    assert np.max(np.abs(_probe(3, x) * x - 1.0)) <= 3.4e-16
    ref = (1.0 + x) / x
    assert np.max(np.abs(_probe(4, x) - ref) / np.abs(ref)) <= 2.3e-16


def test_fiadeiro_veronis_weight_all_branches():
    """sigma(Pe) of LHeureux_model.py:437-442 with W = Pe: 0 below 1e-2, sign above 1e2, coth-1/Pe between."""
    rng = np.random.default_rng(3)
    pe = np.concatenate([rng.uniform(-0.01, 0.01, 1000), rng.uniform(0.01, 100, 100000), -rng.uniform(0.01, 100, 100000),
                         10.0 ** rng.uniform(-2, 2, 100000), rng.uniform(100, 1e4, 1000), -rng.uniform(100, 1e4, 1000),
                         [0.01, -0.01, 100.0, -100.0, 0.0]])
    got = _probe(5, pe)
    with np.errstate(all="ignore"):
        coth = np.cosh(pe) / np.sinh(pe) - 1 / pe
    ref = np.where(np.abs(pe) < 1e-2, 0.0, np.where(np.abs(pe) > 1e2, np.sign(pe), coth))
    # the reference formula itself carries ~1e-16/|Pe| cancellation error; compare absolutely
    assert np.max(np.abs(got - ref)) <= 5e-14
    mid = (np.abs(pe) > 0.5) & (np.abs(pe) <= 100)
    assert np.max(np.abs(got[mid] - ref[mid])) <= 1e-15
    assert np.isnan(_probe(5, np.array([np.nan]))[0])
