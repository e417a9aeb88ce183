"""Scenario / solver / tracker configuration — same names, keys and values as the reference's
marlpde/parameters.py, so `asdict(Map_Scenario())`, `asdict(Solver())`, `asdict(Tracker())` feed
`integrate_equations` unchanged.

Differences, on purpose:
  * pint is not required: quantities are `(magnitude, unit-string)` pairs with a `.magnitude`
    attribute, which is all the reference ever consumes (parameters.py:99);
  * `Solver()` filters its fields per method on EVERY instantiation.  Upstream deletes entries from
    the class-level `__dataclass_fields__` (parameters.py:224-240), so only the first `Solver()` of
    a process gets `jac_sparsity`; later ones silently keep `jac_sparsity=None` (dense Jacobian).
    Here every Radau/BDF solver carries the sparsity pattern, every LSODA solver lband/uband.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields, make_dataclass

import numpy as np
from scipy.sparse import csr_matrix, dia_matrix, lil_matrix


class quantity:  # noqa: N801  (the reference exposes the name `quantity`)
    """A number with a cosmetic unit label."""
    __slots__ = ("magnitude", "units")

    def __init__(self, magnitude, units="dimensionless"):
        self.magnitude = magnitude
        self.units = units

    def __repr__(self):
        return f"<Quantity({self.magnitude}, '{self.units}')>"

    def __eq__(self, other):
        return isinstance(other, quantity) and (self.magnitude, self.units) == (other.magnitude, other.units)

    def __hash__(self):
        return hash((self.magnitude, self.units))


def _q(value, units="dimensionless"):
    return field(default_factory=lambda: quantity(value, units))


@dataclass
class Scenario:
    """Scenario A of L'Heureux (2018) in the FORTRAN code's names and units
    (reference: parameters.py:9-48)."""
    mua: quantity = _q(100.09, "g/mol")
    rhoa: quantity = _q(2.95, "g/cm**3")
    rhoc: quantity = _q(2.71, "g/cm**3")
    rhot: quantity = _q(2.8, "g/cm**3")
    rhow: quantity = _q(1.023, "g/cm**3")
    D0ca: quantity = _q(131.9, "cm**2/a")
    D0co3: quantity = _q(272.6, "cm**2/a")
    Ka: quantity = _q(10 ** (-6.19), "M**2")
    Kc: quantity = _q(10 ** (-6.37), "M**2")
    beta: quantity = _q(0.1, "cm/a")
    b: quantity = _q(5.0, "1/kPa")
    k1: quantity = _q(1.0, "1/a")
    k2: quantity = _q(1.0, "1/a")
    k3: quantity = _q(0.1, "1/a")
    k4: quantity = _q(0.1, "1/a")
    nn: quantity = _q(2.8)
    m: quantity = _q(2.48)
    S: quantity = _q(0.1, "cm/a")
    phiinf: quantity = _q(0.01)
    phi0: quantity = _q(0.8)
    ca0: quantity = _q(0.326e-3, "M")
    co30: quantity = _q(0.326e-3, "M")
    ccal0: quantity = _q(0.3)
    cara0: quantity = _q(0.6)
    xdis: quantity = _q(50.0, "cm")
    length: quantity = _q(500.0, "cm")
    Th: quantity = _q(100.0, "cm")
    phi00: quantity = _q(0.8)
    ca00: quantity = _q(0.326e-3, "M")
    co300: quantity = _q(0.326e-3, "M")
    ccal00: quantity = _q(0.3)
    cara00: quantity = _q(0.6)


# FORTRAN name -> name used by the Python/Matlab codes (reference: parameters.py:60-92)
_FORTRAN_TO_PYTHON = {
    "Ka": "KA", "Kc": "KC", "cara0": "CA0", "cara00": "CAIni", "ccal0": "CC0", "ccal00": "CCIni",
    "ca0": "ca0", "ca00": "ca00", "co30": "co30", "co300": "co300", "phi0": "Phi0", "phi00": "PhiIni",
    "xdis": "ShallowLimit", "Th": "Th", "S": "sedimentationrate", "m": "m1", "nn": "n1", "rhoa": "rhoa",
    "rhoc": "rhoc", "rhot": "rhot", "rhow": "rhow", "beta": "beta", "b": "b", "D0ca": "D0Ca", "k1": "k1",
    "k2": "k2", "k3": "k3", "k4": "k4", "mua": "muA", "D0co3": "DCO3", "phiinf": "PhiInfty",
    "length": "max_depth",
}
_DERIVED = ("cCa0", "cCaIni", "cCO30", "cCO3Ini", "DeepLimit", "rhos0", "rhos", "Xstar", "Tstar", "m2", "n2",
            "DCa", "PhiNR")


def _derive(self):
    """reference: parameters.py:120-143."""
    root_kc = np.sqrt(self.KC)
    self.cCa0 = self.ca0 / root_kc
    self.cCaIni = self.ca00 / root_kc
    self.cCO30 = self.co30 / root_kc
    self.cCO3Ini = self.co300 / root_kc
    self.DeepLimit = self.ShallowLimit + self.Th
    self.rhos0 = self.rhoa * self.CA0 + self.rhoc * self.CC0 + self.rhot * (1 - (self.CA0 + self.CC0))
    self.rhos = self.rhos0
    self.Xstar = self.D0Ca / self.sedimentationrate
    self.Tstar = self.Xstar / self.sedimentationrate
    self.b = self.b / 1e4
    self.m2 = self.m1
    self.n2 = self.n1
    self.DCa = self.D0Ca
    self.PhiNR = self.PhiIni
    self.N = 200            # number of grid cells
    self.FV_switch = 1      # 1: Fiadeiro-Veronis weighting of the pore-water/porosity gradients


def Map_Scenario(scenario: Scenario | None = None):
    """Unit-stripped scenario in the Python names plus derived values; `asdict()` of the result
    is the `pde_parms` dictionary of `integrate_equations` (reference: parameters.py:50-148)."""
    scenario = scenario or Scenario()
    spec = [(_FORTRAN_TO_PYTHON[f.name], float, getattr(scenario, f.name).magnitude)
            for f in fields(Scenario) if f.name in _FORTRAN_TO_PYTHON]
    spec += [(name, float, None) for name in _DERIVED]
    spec += [("N", int, None), ("FV_switch", int, None)]
    cls = make_dataclass("Mapped parameters", spec, namespace={"__post_init__": _derive})
    return cls()


def jacobian_sparsity(no_depths: int | None = None) -> csr_matrix:
    """Jacobian structure handed to Radau/BDF (reference: parameters.py:150-199): for each of the
    9 field-block offsets k*N, k=-4..4, the three diagonals k*N-1, k*N, k*N+1; d(CA,CC)/dPhi zeroed.
    Upstream always uses Map_Scenario().N; an explicit `no_depths` is accepted here."""
    n_cells = Map_Scenario().N if no_depths is None else int(no_depths)
    n = 5 * n_cells
    offsets = [k * n_cells + d for k in range(-4, 5) for d in (-1, 0, 1)]
    pattern = lil_matrix(dia_matrix((np.ones((len(offsets), n)), offsets), shape=(n, n)))
    pattern[:2 * n_cells, 4 * n_cells:] = 0
    return csr_matrix(pattern)


_SOLVER_COMMON = [("first_step", float, 1e-6), ("atol", float, 1e-3), ("rtol", float, 1e-3),
                  ("t_span", tuple, (0, 1)), ("method", str, "Radau")]
_SOLVER_TAIL = [("backend", str, "numba"), ("dense_output", bool, False)]
_solver_classes = {}


class Solver:
    """Solver settings, `asdict(Solver(method=...))` is splatted into the time stepper
    (reference: parameters.py:201-240).  Fields depend on the method:
    LSODA -> lband, uband;  Radau/BDF -> jac_sparsity;  explicit methods -> neither."""

    def __new__(cls, *args, **kwargs):
        if cls is not Solver:
            return object.__new__(cls)
        names = [n for n, _, _ in _SOLVER_COMMON]
        bound = dict(zip(names, args))
        method = kwargs.get("method", bound.get("method", "Radau"))
        return _solver_class(method)(*args, **kwargs)


def _solver_class(method: str):
    kind = "banded" if method == "LSODA" else ("sparse" if method in ("Radau", "BDF") else "plain")
    if kind not in _solver_classes:
        spec = list(_SOLVER_COMMON)
        if kind == "banded":
            spec += [("lband", int, 1), ("uband", int, 1)]
        spec += _SOLVER_TAIL
        ns = {}
        if kind == "sparse":
            spec += [("jac_sparsity", csr_matrix, None)]

            def _post(self):
                if self.jac_sparsity is None:
                    self.jac_sparsity = jacobian_sparsity()
            ns["__post_init__"] = _post
        _solver_classes[kind] = make_dataclass("Solver", spec, bases=(Solver,), namespace=ns)
    return _solver_classes[kind]


@dataclass
class Tracker:
    """Progress/recording settings (reference: parameters.py:243-261)."""
    no_progress_updates: int = 100_000
    no_t_eval: int = 2          # 2 = initial and final state only
    t_eval: np.ndarray = None

    def __post_init__(self):
        self.t_eval = np.linspace(*Solver().t_span, num=self.no_t_eval)
