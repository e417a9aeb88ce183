"""A column the BDF kernel stalls on: integrate until it stops, then look at the state it stopped in — the device Jacobian
(marlpde_probe_jacobian) against the numpy restatement and central differences of the oracle RHS.
    python scripts/diag_bdf_stall.py col [t_end]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import marlpde_b200 as mb
from marlpde_b200 import _cabi
import lheureux_oracle as oracle, jacobian_blocks as jb
col = int(sys.argv[1]); t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lat = mb.sweep_lattice(oracle.default_scenario(), 16, 16, 16)
pde = {k: (float(v[col]) if np.ndim(v) else v) for k, v in lat.items()}
P, y0 = mb.derive_column_params(pde), mb.initial_state(pde)
p = oracle.kernel_params(pde)
r = mb.integrate_bdf_batch(y0, P, t_span=(0, t_end), first_step=1e-6)
print("status", r.status[0], "t", r.t[0], "h", r.h_abs[0], "acc", r.n_accepted[0], "newton fails", r.newton_failures[0])
y = r.y[0].reshape(-1).copy(); N = 200
print("state finite", np.all(np.isfinite(y)), "Phi range", y[4*N:].min(), y[4*N:].max(), "CA range", y[:N].min(), y[:N].max(), "cCa", y[2*N:3*N].min(), y[2*N:3*N].max())
f = oracle.rhs(y, p, np.empty(5 * N)); print("rhs finite", np.all(np.isfinite(f)), "max |f|", np.abs(f).max())
J = np.zeros((N, 3, 5, 5))
_cabi.check(_cabi.lib().marlpde_probe_jacobian(_cabi.ptr(y), _cabi.ptr(P), N, _cabi.ptr(J), 0))
print("device J finite", np.all(np.isfinite(J)), "non-finite cells", np.unique(np.nonzero(~np.isfinite(J))[0])[:20])
worst = 0
for i in range(N):
    want = jb.blocks(y, p, i, N)
    for b in range(3):
        got = J[i, b].T
        e = np.max(np.abs(got - want[b])) / max(1.0, np.abs(want[b]).max())
        if not (e <= worst): worst = e; wi = (i, b)
print("device vs numpy analytic J: worst rel", worst, wi)
# central differences
n = 5 * N; Jc = np.zeros((n, n)); fp, fm = np.empty(n), np.empty(n)
for j in range(n):
    h = 1e-7 * max(1e-3, abs(y[j])); yp, ym = y.copy(), y.copy(); yp[j] += h; ym[j] -= h
    oracle.rhs(yp, p, fp); oracle.rhs(ym, p, fm); Jc[:, j] = (fp - fm) / (2 * h)
worst = 0
for i in range(N):
    rows = [ff * N + i for ff in range(5)]
    for b, off in ((0, -1), (1, 0), (2, 1)):
        if 0 <= i + off < N:
            Jn = Jc[np.ix_(rows, [ff * N + i + off for ff in range(5)])]
            e = np.abs(J[i, b].T - Jn); rel = e / np.maximum(np.abs(Jn), 1e-6 * np.abs(Jc).max())
            if rel.max() > worst: worst = rel.max(); wi = (i, b, np.unravel_index(rel.argmax(), rel.shape), J[i, b].T.flat[rel.argmax()], Jn.flat[rel.argmax()])
print("device J vs central differences: worst rel", worst, wi)
np.save(os.path.join(ROOT, "gpurun_out", f"stall_state_{col}.npy"), y)
# where are the switches?  cells with U <= 0, Phi >= 1, |Pe| regimes
Phi = y[4*N:]; F = 1 - np.exp(10 - 10 / Phi); U = p[8] + p[9] * Phi**3 * F / (1 - Phi); W = p[8] - p[9] * Phi**2 * F
print("cells with U <= 0:", np.nonzero(U <= 0)[0][:20], " W > 0:", np.nonzero(W > 0)[0][:20], " Phi >= 1:", np.nonzero(Phi >= 1)[0][:20])
two = y[2*N:3*N] * y[3*N:4*N]; print("two range", two.min(), two.max(), "three range", (two * p[15]).min(), (two * p[15]).max())

# restart from the stall state (fresh order-1 BDF): does the kernel get going again?
t0 = float(r.t[0])
for fs in (1e-6, 1e-8, 1e-10):
    r2 = mb.integrate_bdf_batch(y.reshape(1, 5, N), P, t_span=(t0, t0 + 0.01), first_step=fs)
    print("restart first_step", fs, "status", r2.status[0], "t", r2.t[0], "acc", r2.n_accepted[0], "rej", r2.n_rejected[0], "newton", r2.newton_iterations[0], "fails", r2.newton_failures[0], "njev", r2.njev[0])
r3 = mb.integrate_radau_batch(y.reshape(1, 5, N), P, t_span=(t0, t0 + 0.01), first_step=1e-6)
print("radau restart: status", r3.status[0], "t", r3.t[0], "acc", r3.n_accepted[0], "fails", r3.newton_failures[0])
