#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ (run in the authoring container,
where /root/reference exists; the GPU box never runs this).

    python tests/golden/make_golden.py

Produces
  fixtures_reference.npz   snapshots extracted from the reference's own regression fixtures
                           (tests/Regression_test/data/*.hdf5, *.h5) — the end states the
                           reference's tests assert against (test_regression.py:52, :87, :147)
                           plus every 10th intermediate snapshot as evolved RHS inputs.
  scenario_reference.json  asdict(Map_Scenario()) from the reference's own parameters.py
                           (imported with oracle/refstubs/pint), Solver/Tracker defaults and the
                           structure of jacobian_sparsity().
  rhs_reference_vardphi.npz  (`python tests/golden/make_golden.py vardphi`) the same for the model variant the reference
                           keeps as a commented-out line: the source text of /root/reference/marlpde/LHeureux_model.py
                           is loaded, the two lines `dPhi = self.dPhi_fixed` (:223) and `dPhi[i] = dPhi_fixed` (:431)
                           are replaced by the commented formulas one line above them (:222, :430), and the resulting
                           module's `fun_numba` / `fun` are evaluated.  Nothing else of the reference is touched.
  rhs_reference.npz        inputs y and outputs of the reference's OWN `fun_numba` -> `pde_rhs`
                           and `fun` (LHeureux_model.py:162-522), imported unmodified from
                           /root/reference with oracle/refstubs/pde standing in for py-pde,
                           for several parameter sets and states (incl. U<0, every Peclet
                           branch, FV_switch=0, Phi>1).
"""
import json
import os
import sys
from dataclasses import asdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200"))
sys.path.insert(0, os.path.join(REF, "marlpde"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from marlpde_b200 import hdf5lite  # noqa: E402  (reader for the reference's HDF5 fixtures)
import parameters as ref_parameters  # noqa: E402  (the reference's own file)
import LHeureux_model as ref_model  # noqa: E402  (the reference's own file)
from pde import CartesianGrid, ScalarField  # noqa: E402  (stand-in)
import lheureux_oracle as oracle  # noqa: E402

np.seterr(all="ignore")  # the reference sets raise-on-invalid globally (LHeureux_model.py:6)


class _NoBar:
    def update(self, n):
        pass


def build_reference_model(pde_parms):
    """Mirror of Evolve_scenario.py:40-68 using the reference's own class."""
    import inspect
    N = pde_parms["N"]
    depths = CartesianGrid([[0, pde_parms["max_depth"] / pde_parms["Xstar"]]], [N], periodic=False)
    nts = ScalarField.from_expression(depths, f"heaviside(x-{pde_parms['ShallowLimit'] / pde_parms['Xstar']}, 0)")
    ntd = ScalarField.from_expression(depths, f"heaviside({pde_parms['DeepLimit'] / pde_parms['Xstar']}-x, 0)")
    names = [p.name for p in inspect.signature(ref_model.LMAHeureuxPorosityDiff).parameters.values()]
    filtered = {k: v for k, v in pde_parms.items() if k in names}
    slices = [slice(i * N, (i + 1) * N) for i in range(5)]
    return ref_model.LMAHeureuxPorosityDiff(depths, slices, nts, ntd, **filtered)


def main():
    data_dir = os.path.join(REF, "tests", "Regression_test", "data")
    # ------------------------------------------------------------------ fixtures
    fa = hdf5lite.File(os.path.join(data_dir, "LMAHeureuxPorosityDiff_Phi0_0.6_PhiIni_0.5.hdf5"))
    fb = hdf5lite.File(os.path.join(data_dir, "LMAHeureuxPorosityDiff_Phi0_PhiIni_0.8.hdf5"))
    fm = hdf5lite.File(os.path.join(data_dir, "Matlab_output_Scenario_A_Phi0_PhiIni_0.5_k3_k4_0.01.h5"))
    A, B, M = np.asarray(fa["data"]), np.asarray(fb["data"]), np.asarray(fm["Solutions after_T*"])
    ks = np.array([0, 1, 2, 5] + list(range(10, 101, 10)))
    np.savez_compressed(os.path.join(HERE, "fixtures_reference.npz"),
                        snapshot_index=ks, times=np.asarray(fa["times"])[ks],
                        scenario_A=A[ks], high_porosity=B[ks], matlab=M)
    # ------------------------------------------------------------------ parameters
    base = asdict(ref_parameters.Map_Scenario())
    mine = oracle.default_scenario()
    assert set(base) == set(mine), set(base) ^ set(mine)
    for k in base:
        assert base[k] == mine[k], (k, base[k], mine[k])
    solver_first = asdict(ref_parameters.Solver())          # 1st instantiation: sparse Jacobian
    js = solver_first.pop("jac_sparsity")
    solver_second = asdict(ref_parameters.Solver())         # quirk: class fields were deleted
    tracker = asdict(ref_parameters.Tracker())
    mine_js = oracle.jacobian_sparsity(200)
    assert (js != mine_js).nnz == 0
    scen = {
        "Map_Scenario": {k: (float(v) if not isinstance(v, int) else v) for k, v in base.items()},
        "Solver_first": {k: (list(v) if isinstance(v, tuple) else v) for k, v in solver_first.items()},
        "Solver_second_keys": sorted(solver_second.keys()),
        "Solver_second_jac_sparsity_is_None": solver_second.get("jac_sparsity", "absent") is None,
        "Tracker": {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in tracker.items()},
        "jac_sparsity": {"shape": list(js.shape), "nnz": int(js.nnz),
                         "row0_cols": js[0].indices.tolist(), "row500_cols": js[500].indices.tolist()},
    }
    with open(os.path.join(HERE, "scenario_reference.json"), "w") as fh:
        json.dump(scen, fh, indent=1, sort_keys=True)

    # ------------------------------------------------------------------ RHS golden vectors
    rng = np.random.default_rng(20261018)
    cases = {
        "default": {},
        "scenario_A": {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6},
        "matlab": {"Phi0": 0.5, "PhiIni": 0.5, "PhiNR": 0.5, "k3": 0.01, "k4": 0.01},
        "fv_off": {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6, "FV_switch": 0},
        "exponents": {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6, "m1": 2.1, "m2": 2.7, "n1": 3.1, "n2": 2.2,
                      "k1": 0.7, "k4": 0.3},
    }
    # a sweep-lattice corner: S=0.11, b=6, D0co3=300, derived values recomputed
    raw = dict(muA=100.09, rhoa=2.95, rhoc=2.71, rhot=2.8, rhow=1.023, D0Ca=131.9, DCO3=300.0,
               KA=10 ** (-6.19), KC=10 ** (-6.37), beta=0.1, b=6.0, k1=1.0, k2=1.0, k3=0.1, k4=0.1,
               n1=2.8, m1=2.48, sedimentationrate=0.11, PhiInfty=0.01, Phi0=0.8, ca0=0.326e-3,
               co30=0.326e-3, CC0=0.3, CA0=0.6, ShallowLimit=50.0, max_depth=500.0, Th=100.0,
               PhiIni=0.8, ca00=0.326e-3, co300=0.326e-3, CCIni=0.3, CAIni=0.6)
    lattice_corner = oracle.derive_scenario(raw)
    out = {}
    meta = {}
    for name, over in list(cases.items()) + [("lattice_corner", None)]:
        pde = lattice_corner if over is None else (base | over)
        eq = build_reference_model(pde)
        N = pde["N"]
        y0 = oracle.initial_state(pde)
        states = {"y0": y0,
                  "noise": y0 * (1 + 0.05 * rng.uniform(-1, 1, y0.size))}
        if name == "scenario_A":
            for k in (1, 10, 50, 100):
                states[f"fixA_{k}"] = A[k].ravel()
        if name == "default":
            for k in (1, 10, 30, 100):
                states[f"fixB_{k}"] = B[k].ravel()
        # U < 0 (forward differences + curvature ghost for CA, CC): low porosity with this presum
        lowphi = states["noise"].copy()
        lowphi[4 * N:] = 0.25 + 0.1 * rng.uniform(0, 1, N)
        states["U_negative"] = lowphi
        # wide Peclet spread: porosity ramp 0.05..0.97 makes |Pe| cross 1e-2 and, for Phi, be O(1)
        ramp = states["noise"].copy()
        ramp[4 * N:] = np.linspace(0.05, 0.97, N)
        ramp[2 * N:3 * N] *= np.linspace(0.2, 3.0, N)      # cCa*cCO3 on both sides of 1 and of 1/KRat
        states["ramp"] = ramp
        if name in ("default", "lattice_corner"):
            over1 = states["noise"].copy()
            over1[4 * N:] = np.linspace(0.7, 1.08, N)       # Phi crosses 1 (never exactly 1)
            states["Phi_above_1"] = over1
        for sname, y in states.items():
            y = np.ascontiguousarray(y, dtype=np.float64)
            eq.last_t = 0.0
            r_numba = eq.fun_numba(0.0, y.copy(), _NoBar(), 1e-5, 0.0)
            eq.last_t = 0.0
            r_numpy = eq.fun(0.0, y.copy(), _NoBar(), 1e-5, 0.0)
            ev = np.array([e(0.0, y, None, None, None) for e in
                           (eq.zeros, eq.zeros_CA, eq.zeros_CC, eq.ones_CA_plus_CC, eq.ones_Phi,
                            eq.zeros_U, eq.zeros_W)])
            key = f"{name}/{sname}"
            out[key + "/y"] = y
            out[key + "/rhs_numba"] = np.asarray(r_numba)
            out[key + "/rhs_numpy"] = np.asarray(r_numpy)
            out[key + "/events"] = ev
        meta[name] = {k: (float(v) if not isinstance(v, int) else v) for k, v in pde.items()}
        # the reference's own derived constants, as a cross-check of oracle.kernel_params
        out[f"{name}/derived"] = np.array([eq.presum, eq.rhorat, eq.Da, eq.lambda_, eq.dCa, eq.dCO3,
                                           eq.delta, eq.KRat, eq.nu1, eq.nu2, eq.dPhi_fixed, eq.delta_x,
                                           eq.auxcon, eq.rhorat0])
    out["__meta__"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "rhs_reference.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


def reference_model_with_time_varying_dPhi():
    """The reference's LHeureux_model module with its own commented-out dPhi formula switched on (text substitution
    of exactly two lines at import time; the file on disk is not modified)."""
    import types
    src = open(os.path.join(REF, "marlpde", "LHeureux_model.py")).read()
    a = "        # dPhi = self.auxcon * F * (Phi ** 3) / (1 - Phi)\n        dPhi = self.dPhi_fixed\n"
    b = "            # dPhi[i] = auxcon * F[i] * (Phi[i] ** 3) / one_minus_Phi[i]\n            dPhi[i] = dPhi_fixed\n"
    assert src.count(a) == 1 and src.count(b) == 1
    src = src.replace(a, "        dPhi = self.auxcon * F * (Phi ** 3) / (1 - Phi)\n")
    src = src.replace(b, "            dPhi[i] = auxcon * F[i] * (Phi[i] ** 3) / one_minus_Phi[i]\n")
    mod = types.ModuleType("LHeureux_model_vardphi")
    mod.__file__ = os.path.join(REF, "marlpde", "LHeureux_model.py")
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def main_vardphi():
    global ref_model
    data_dir = os.path.join(REF, "tests", "Regression_test", "data")
    A = np.asarray(hdf5lite.File(os.path.join(data_dir, "LMAHeureuxPorosityDiff_Phi0_0.6_PhiIni_0.5.hdf5"))["data"])
    B = np.asarray(hdf5lite.File(os.path.join(data_dir, "LMAHeureuxPorosityDiff_Phi0_PhiIni_0.8.hdf5"))["data"])
    base = asdict(ref_parameters.Map_Scenario())
    plain = ref_model
    ref_model = reference_model_with_time_varying_dPhi()
    rng = np.random.default_rng(20261019)
    out, meta = {}, {}
    for name, over in (("default", {}), ("scenario_A", {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6}),
                       ("fv_off", {"Phi0": 0.6, "PhiIni": 0.5, "PhiNR": 0.6, "FV_switch": 0})):
        pde = base | over
        eq = build_reference_model(pde)
        N = pde["N"]
        y0 = oracle.initial_state(pde)
        states = {"y0": y0, "noise": y0 * (1 + 0.05 * rng.uniform(-1, 1, y0.size))}
        fix = A if name != "default" else B
        for k in (1, 10, 100):
            states[f"fix_{k}"] = fix[k].ravel()
        ramp = states["noise"].copy()
        ramp[4 * N:] = np.linspace(0.05, 0.97, N)          # dPhi spans orders of magnitude: every Peclet branch
        states["ramp"] = ramp
        lowphi = states["noise"].copy()
        lowphi[4 * N:] = 0.25 + 0.1 * rng.uniform(0, 1, N)
        states["U_negative"] = lowphi
        for sname, y in states.items():
            y = np.ascontiguousarray(y, dtype=np.float64)
            eq.last_t = 0.0
            out[f"{name}/{sname}/y"] = y
            out[f"{name}/{sname}/rhs_numba"] = np.asarray(eq.fun_numba(0.0, y.copy(), _NoBar(), 1e-5, 0.0))
            eq.last_t = 0.0
            out[f"{name}/{sname}/rhs_numpy"] = np.asarray(eq.fun(0.0, y.copy(), _NoBar(), 1e-5, 0.0))
        # how much the variant differs from the shipped model on the same state (so the test cannot pass trivially)
        ref_model = plain
        eq0 = build_reference_model(pde)
        ref_model = sys.modules.get("LHeureux_model_vardphi") or reference_model_with_time_varying_dPhi()
        eq0.last_t = 0.0
        out[f"{name}/noise/rhs_numba_fixed_dPhi"] = np.asarray(eq0.fun_numba(0.0, np.ascontiguousarray(states["noise"]), _NoBar(), 1e-5, 0.0))
        meta[name] = {k: (float(v) if not isinstance(v, int) else v) for k, v in pde.items()}
    out["__meta__"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "rhs_reference_vardphi.npz"), **out)
    print("wrote rhs_reference_vardphi.npz", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vardphi":
        main_vardphi()
    else:
        main()
