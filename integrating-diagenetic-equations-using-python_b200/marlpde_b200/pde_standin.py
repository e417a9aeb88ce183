"""The sliver of py-pde the reference *driver* touches, without numba or py-pde itself.

The reference returns its depth grid from `integrate_equations` and builds plotting depths with
`ScalarField.from_expression(grid, "x").data` (marlpde/Evolve_scenario.py:40, :51-54, :183, :194;
tests/Regression_test/test_regression.py:129-132).  Stencils are NOT provided here: on this side
of the boundary they live in the CUDA kernels (csrc/lheureux_device.cuh).
"""
from __future__ import annotations

import numpy as np


class CartesianGrid:
    """1-D cell-centred grid: N cells on [lo, hi], centres x_i = lo + (i + 1/2) dx."""

    def __init__(self, bounds, shape, periodic=False):
        if periodic not in (False, [False], (False,)):
            raise NotImplementedError("periodic grids are not part of the diagenetic model")
        if len(bounds) != 1:
            raise NotImplementedError("only 1-D depth grids")
        (lo, hi), = bounds
        n = int(shape if np.ndim(shape) == 0 else shape[0])
        self.shape = (n,)
        self.axes_bounds = ((float(lo), float(hi)),)
        self.discretization = np.array([(float(hi) - float(lo)) / n])
        self._axes_coords = (float(lo) + (np.arange(n) + 0.5) * self.discretization[0],)
        self.periodic = [False]
        self._registered = {}

    @property
    def axes_coords(self):
        return self._axes_coords

    @property
    def num_axes(self):
        return 1

    def register_operator(self, name, factory):
        """Accepted for signature compatibility (Evolve_scenario.py:43-46); the GPU path never
        calls the factories."""
        self._registered[name] = factory

    def make_operator(self, name, bc):
        raise NotImplementedError(
            "py-pde operator closures are replaced by the CUDA stencils; use "
            "LMAHeureuxPorosityDiff.fun / fun_numba or marlpde_b200.rhs_batch")

    def __repr__(self):
        (lo, hi), = self.axes_bounds
        return f"CartesianGrid(bounds=(({lo}, {hi}),), shape=({self.shape[0]},), periodic=[False])"


def _heaviside(x, h0=0.5):
    return np.heaviside(x, h0)


class ScalarField:
    def __init__(self, grid, data=0.0, label=None):
        self.grid = grid
        src = data.data if isinstance(data, ScalarField) else data
        self.data = np.array(np.broadcast_to(np.asarray(src, dtype=np.float64), grid.shape))
        self.label = label

    @classmethod
    def from_expression(cls, grid, expression: str, label=None):
        """Evaluate an expression in `x` (and heaviside(., .)) at the cell centres."""
        x = grid._axes_coords[0]
        env = {"__builtins__": {}, "x": x, "heaviside": _heaviside, "Heaviside": _heaviside,
               "exp": np.exp, "log": np.log, "sin": np.sin, "cos": np.cos, "sqrt": np.sqrt, "pi": np.pi}
        val = eval(compile(expression, "<field expression>", "eval"), env)  # noqa: S307 (closed namespace)
        return cls(grid, val, label=label)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)


class FieldCollection:
    def __init__(self, fields):
        self.fields = list(fields)

    @property
    def data(self):
        return np.stack([f.data for f in self.fields])

    @property
    def labels(self):
        return [f.label for f in self.fields]
