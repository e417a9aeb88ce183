"""`LMAHeureuxPorosityDiff` — same constructor signature, attributes and callbacks as the
reference's marlpde/LHeureux_model.py, with the right-hand side evaluated by the sm_100a kernel
(csrc/rhs_batch.cu) through the C ABI instead of py-pde operators + numba.

`fun` and `fun_numba` keep the `(t, y, progress_proxy, progress_dt, t0) -> ndarray[5N]`
signature so SciPy's solve_ivp can drive them exactly as upstream does
(marlpde/Evolve_scenario.py:104-109); both call the same CUDA kernel (the reference's two
backends compute the same maths, LHeureux_model.py:162-288 vs :361-522).
The batched, fully on-device integration lives in Evolve_scenario.integrate_equations[_batch].
"""
from __future__ import annotations

import numpy as np

import marlpde_b200 as _mb
from marlpde_b200 import _cabi
from marlpde_b200.pde_standin import FieldCollection


class LMAHeureuxPorosityDiff:
    """Parameters, boundary conditions, RHS and event monitors of the diagenetic model."""

    def __init__(self, Depths, slices_all_fields, not_too_shallow, not_too_deep,
                 CA0, CC0, cCa0, cCO30, Phi0, sedimentationrate, Xstar, Tstar,
                 k1, k2, k3, k4, m1, m2, n1, n2, b, beta, rhos, rhow, rhos0,
                 KA, KC, muA, D0Ca, PhiNR, PhiInfty, PhiIni, DCa, DCO3,
                 FV_switch, time_varying_dPhi=False):
        # `time_varying_dPhi` is not an upstream argument: upstream switches between the fixed porosity diffusion
        # coefficient and dPhi = auxcon F Phi^3 / (1 - Phi) by (un)commenting a source line (:222-223, :430-431);
        # here it is a kernel template flag selected per column.  A `pde_parms` dictionary without the key behaves
        # exactly like upstream's shipped source.
        self.time_varying_dPhi = bool(time_varying_dPhi)
        self.no_fields = 5
        self.Depths = Depths
        # accepted for signature compatibility (reference :19-22); the stencils are in the kernel
        self.Depths.register_operator("grad_back", None)
        self.Depths.register_operator("grad_forw", None)
        x = self.Depths._axes_coords[0]
        self.delta_x = x[1] - x[0]
        self.slices_all_fields = slices_all_fields
        (self.CA_sl, self.CC_sl, self.cCa_sl, self.cCO3_sl, self.Phi_sl) = slices_all_fields
        # boundary conditions in py-pde's notation (reference :26-30); realised as ghost cells in
        # csrc/lheureux_device.cuh::load_triple
        self.bc_CA = [{"value": CA0}, {"curvature": 0}]
        self.bc_CC = [{"value": CC0}, {"curvature": 0}]
        self.bc_cCa = [{"value": cCa0}, {"derivative": 0}]
        self.bc_cCO3 = [{"value": cCO30}, {"derivative": 0}]
        self.bc_Phi = [{"value": Phi0}, {"derivative": 0}]
        for name, val in (("sedimentationrate", sedimentationrate), ("Xstar", Xstar), ("Tstar", Tstar),
                          ("k1", k1), ("k2", k2), ("k3", k3), ("k4", k4), ("m1", m1), ("m2", m2),
                          ("n1", n1), ("n2", n2), ("b", b), ("beta", beta), ("rhos", rhos), ("rhow", rhow),
                          ("rhos0", rhos0), ("KA", KA), ("KC", KC), ("muA", muA), ("D0Ca", D0Ca),
                          ("PhiNR", PhiNR), ("PhiInfty", PhiInfty), ("Phi0", Phi0), ("PhiIni", PhiIni),
                          ("DCa", DCa), ("DCO3", DCO3), ("FV_switch", FV_switch), ("CA0", CA0), ("CC0", CC0),
                          ("cCa0", cCa0), ("cCO30", cCO30)):
            setattr(self, name, val)
        self.not_too_shallow = np.asarray(not_too_shallow.data, dtype=np.float64)
        self.not_too_deep = np.asarray(not_too_deep.data, dtype=np.float64)
        self.Peclet_min = 1e-2
        self.Peclet_max = 1 / self.Peclet_min
        self.last_t = 0.0      # progress-bar helper, as upstream (:90)

        n = self.Depths.shape[0]
        lo_hi = self.Depths.axes_bounds[0]
        pde = dict(CA0=CA0, CC0=CC0, cCa0=cCa0, cCO30=cCO30, Phi0=Phi0, sedimentationrate=sedimentationrate,
                   Xstar=Xstar, Tstar=Tstar, k1=k1, k2=k2, k3=k3, k4=k4, m1=m1, m2=m2, n1=n1, n2=n2, b=b,
                   beta=beta, rhos=rhos, rhow=rhow, rhos0=rhos0, KA=KA, KC=KC, muA=muA, D0Ca=D0Ca,
                   PhiNR=PhiNR, PhiInfty=PhiInfty, PhiIni=PhiIni, DCa=DCa, DCO3=DCO3, FV_switch=FV_switch,
                   N=n, max_depth=(lo_hi[1] - lo_hi[0]) * Xstar, ShallowLimit=0.0, DeepLimit=0.0,
                   time_varying_dPhi=self.time_varying_dPhi)
        self.column_params = _mb.derive_column_params(pde)
        # the masks are given as arrays (Evolve_scenario.py:51-54): the kernel wants their support interval
        mask = self.not_too_shallow * self.not_too_deep
        nz = np.nonzero(mask)[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 0)
        if nz.size and (not np.all(mask[lo:hi] == 1.0) or np.any((mask != 0) & (mask != 1))):
            raise ValueError("not_too_shallow*not_too_deep must be 1 on one interval of cells and 0 elsewhere")
        self.column_params["mask_lo"], self.column_params["mask_hi"] = lo, hi
        self.column_params["delta_x"] = self.delta_x
        p = self.column_params[0]
        # derived constants under the reference's attribute names (:36-72, :131-133)
        self.nu1, self.nu2, self.KRat = p["nu1"], p["nu2"], p["KRat"]
        self.g = 100 * 9.81
        self.dCa, self.dCO3, self.delta = p["dCa"], p["dCO3"], p["delta"]
        self.Da, self.lambda_ = p["Da"], p["lambda_"]
        self.auxcon = beta / (D0Ca * b * self.g * rhow * (PhiNR - PhiInfty))
        self.rhorat0 = (rhos0 / rhow - 1) * beta / sedimentationrate
        self.rhorat, self.presum = p["rhorat"], p["presum"]
        self.F_fixed = 1 - np.exp(10 - 10 / PhiIni)
        self.dPhi_fixed = p["dPhi_fixed"]
        for ev in (self.zeros, self.zeros_CA, self.zeros_CC, self.ones_CA_plus_CC, self.ones_Phi,
                   self.zeros_U, self.zeros_W):
            setattr(ev.__func__, "terminal", False)     # monitors only (:113-122)
        self._dev = None

    # -- state ------------------------------------------------------------------------------
    def get_state(self, AragoniteSurface, CalciteSurface, CaSurface, CO3Surface, PorSurface):
        """Initial state as a FieldCollection (reference :135-145)."""
        for fld, label in zip((AragoniteSurface, CalciteSurface, CaSurface, CO3Surface, PorSurface),
                              ("ARA", "CAL", "Ca", "CO3", "Po")):
            fld.label = label
        return FieldCollection([AragoniteSurface, CalciteSurface, CaSurface, CO3Surface, PorSurface])

    # -- right-hand side ----------------------------------------------------------------------
    def _progress(self, t, progress_proxy, progress_dt, t0):
        if progress_proxy is None:
            return
        if self.last_t == 0.:
            self.last_t = t0
        n = int((t - self.last_t) / progress_dt)
        progress_proxy.update(n)
        self.last_t += n * progress_dt

    def _device_buffers(self):
        if self._dev is None:
            import torch
            if _cabi.lib().marlpde_device_count() < 1:
                raise _cabi.MarlpdeError("no CUDA device available; LMAHeureuxPorosityDiff has no CPU path")
            dev = torch.device("cuda", torch.cuda.current_device())
            n = self.Depths.shape[0]
            self._dev = (torch.empty((1, 5, n), dtype=torch.float64, device=dev),
                         torch.empty((1, 5, n), dtype=torch.float64, device=dev),
                         _mb.batch.params_to_device(self.column_params, dev))
        return self._dev

    def rhs(self, y):
        """dy/dt for one column on the GPU; returns a fresh array (SciPy keeps it in K[s])."""
        import torch
        d_y, d_out, d_par = self._device_buffers()
        y = np.ascontiguousarray(y, dtype=np.float64)
        d_y.copy_(torch.from_numpy(y).view(1, 5, -1))
        _mb.rhs_batch(d_y, d_par, out=d_out)
        return d_out.cpu().numpy().ravel()

    def fun(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        self._progress(t, progress_proxy, progress_dt, t0)
        return self.rhs(y)

    def fun_numba(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        self._progress(t, progress_proxy, progress_dt, t0)
        return self.rhs(y)

    # -- event monitors (reference :524-593); the extra arguments are unused, as upstream -----
    def _UW(self, y):
        Phi = y[self.Phi_sl]
        F = 1 - np.exp(10 - 10 / Phi)
        return (self.presum + self.rhorat * Phi ** 3 * F / (1 - Phi), self.presum - self.rhorat * Phi ** 2 * F)

    def zeros(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amin(y)

    def zeros_CA(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amin(y[self.CA_sl])

    def zeros_CC(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amin(y[self.CC_sl])

    def ones_CA_plus_CC(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amax(y[self.CA_sl] + y[self.CC_sl]) - 1

    def ones_Phi(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amax(y[self.Phi_sl]) - 1

    def zeros_U(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amin(self._UW(y)[0])

    def zeros_W(self, t, y, progress_proxy=None, progress_dt=None, t0=None):
        return np.amax(self._UW(y)[1])
