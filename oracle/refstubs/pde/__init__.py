"""Stand-in for py-pde 0.32.2 (test infrastructure; see ../README.md).

Restates: cell-centred CartesianGrid; ghost-cell boundary rules
  {"value": v}      -> ghost = 2 v - a_adjacent              (Dirichlet at the face)
  {"derivative": d} -> ghost = a_adjacent + dx d             (outward normal derivative)
  {"curvature": c}  -> ghost = 2 a_adj - a_adj2 + c dx**2
make_operator(name, bc) -> numba closure arr[N] -> new out[N] that pads, sets the two
ghosts and applies the raw stencil; ScalarField arithmetic; FieldCollection.data.
"""
import numbers

import numpy as np
import sympy
from numba import njit

from .grids.operators import cartesian as _cart


def _bc_coeffs(bc_side, dx):
    """ghost = const + f1 * a_adjacent + f2 * a_second_adjacent"""
    (kind, val), = bc_side.items()
    if kind == "value":
        return 2.0 * val, -1.0, 0.0
    if kind == "derivative":
        return dx * val, 1.0, 0.0
    if kind == "curvature":
        return val * dx ** 2, 2.0, -1.0
    raise ValueError(kind)


class CartesianGrid:
    def __init__(self, bounds, shape, periodic=False):
        assert not periodic and len(bounds) == 1
        (lo, hi), = bounds
        n = int(shape[0]) if not isinstance(shape, numbers.Integral) else int(shape)
        self.shape = (n,)
        self.axes_bounds = ((float(lo), float(hi)),)
        dx = (float(hi) - float(lo)) / n
        self.discretization = np.array([dx])
        self._axes_coords = (float(lo) + (np.arange(n) + 0.5) * dx,)
        self._operators = {"laplace": _cart.make_laplace}

    axes_coords = property(lambda self: self._axes_coords)

    def register_operator(self, name, factory):
        self._operators[name] = factory

    def make_operator(self, name, bc):
        raw = self._operators[name](self)
        n = self.shape[0]
        dx = float(self.discretization[0])
        c_lo, f1_lo, f2_lo = _bc_coeffs(bc[0], dx)
        c_hi, f1_hi, f2_hi = _bc_coeffs(bc[1], dx)

        @njit
        def op(arr):
            full = np.empty(n + 2)
            full[1:n + 1] = arr
            full[0] = c_lo + f1_lo * full[1] + f2_lo * full[2]
            full[n + 1] = c_hi + f1_hi * full[n] + f2_hi * full[n - 1]
            out = np.empty(n)
            raw(full, out)
            return out
        return op


class ScalarField:
    def __init__(self, grid, data=0.0, label=None):
        self.grid = grid
        d = data.data if isinstance(data, ScalarField) else data
        self.data = np.broadcast_to(np.asarray(d, dtype=np.float64), grid.shape).copy()
        self.label = label

    @classmethod
    def from_expression(cls, grid, expression, label=None):
        x = sympy.Symbol("x")
        expr = sympy.sympify(expression, locals={"heaviside": sympy.Heaviside, "x": x})
        f = sympy.lambdify(x, expr, modules=[{"Heaviside": lambda a, h0=0.5: np.heaviside(a, h0)}, "numpy"])
        return cls(grid, f(grid._axes_coords[0]), label=label)

    def to_scalar(self, func):
        return ScalarField(self.grid, func(self.data))

    def apply_operator(self, name, bc):
        return ScalarField(self.grid, self.grid.make_operator(name, bc)(self.data))

    def laplace(self, bc):
        return self.apply_operator("laplace", bc)

    # numpy interop: np.exp(10 - 10 / Phi) etc. must give a ScalarField back
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        args = [i.data if isinstance(i, ScalarField) else i for i in inputs]
        return ScalarField(self.grid, getattr(ufunc, method)(*args, **kwargs))

    def _bin(self, o, fn):
        return ScalarField(self.grid, fn(self.data, o.data if isinstance(o, ScalarField) else o))

    __add__ = lambda s, o: s._bin(o, lambda a, b: a + b)
    __radd__ = lambda s, o: s._bin(o, lambda a, b: b + a)
    __sub__ = lambda s, o: s._bin(o, lambda a, b: a - b)
    __rsub__ = lambda s, o: s._bin(o, lambda a, b: b - a)
    __mul__ = lambda s, o: s._bin(o, lambda a, b: a * b)
    __rmul__ = lambda s, o: s._bin(o, lambda a, b: b * a)
    __truediv__ = lambda s, o: s._bin(o, lambda a, b: a / b)
    __rtruediv__ = lambda s, o: s._bin(o, lambda a, b: b / a)
    __pow__ = lambda s, o: s._bin(o, lambda a, b: a ** b)
    __neg__ = lambda s: ScalarField(s.grid, -s.data)


class FieldCollection:
    def __init__(self, fields):
        self.fields = list(fields)

    @property
    def data(self):
        return np.stack([f.data for f in self.fields])
