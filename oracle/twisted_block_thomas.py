"""Two-ended ("twisted") block-Thomas for a block-tridiagonal system — the linear algebra of the batched Radau
kernel (csrc/radau_batch.cu: factorise(), solve()), restated in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/lheureux_oracle.py): imported by tests/, never by the product path.

The kernel solves (M I - J) x = b, J block tridiagonal in cell-major order with 5x5 blocks L_i (to cell i-1), D_i,
U_i (to cell i+1) — the structure of the reference's Jacobian (marlpde/parameters.py:150-199 hands SciPy a
27-diagonal superset of it) — for M = MU_REAL / h and M = MU_COMPLEX / h of scipy's Radau
(scipy/integrate/_ivp/radau.py: LU_real, LU_complex).  Cells 0 .. mid-1 are eliminated top-down, cells
N-1 .. mid+1 bottom-up, the chains meet in cell mid = N // 2; the two chains of a solve are independent, which is
what the kernel runs side by side in the two halves of a warp.
"""
import numpy as np


def factor(L, D, U, M):
    """Returns Sinv[N, b, b]: S_i^{-1} for i <= mid (top chain and meeting cell), T_i^{-1} for i > mid."""
    N, b = len(D), D.shape[1]
    mid = N // 2
    eye = np.eye(b)
    Sinv = np.zeros((N, b, b), dtype=np.result_type(D.dtype, type(M)))
    X = np.zeros((b, b), dtype=Sinv.dtype)              # X_{i-1} = S_{i-1}^{-1} U_{i-1}
    for i in range(mid):
        Sinv[i] = np.linalg.inv(M * eye - D[i] - L[i] @ X)
        X = Sinv[i] @ U[i]
    Y = np.zeros((b, b), dtype=Sinv.dtype)              # Y_{i+1} = T_{i+1}^{-1} L_{i+1}
    for i in range(N - 1, mid, -1):
        Sinv[i] = np.linalg.inv(M * eye - D[i] - U[i] @ Y)
        Y = Sinv[i] @ L[i]
    Sinv[mid] = np.linalg.inv(M * eye - D[mid] - L[mid] @ X - U[mid] @ Y)
    return Sinv


def solve(L, U, Sinv, rhs):
    """x with (M I - J) x = rhs, from the factors above: inward sweeps of both chains in lock-step, the meeting
    cell, outward sweeps in lock-step (the order of operations of the kernel's solve())."""
    N, b = Sinv.shape[0], Sinv.shape[1]
    mid = N // 2
    n_top, n_bot = mid, N - 1 - mid
    x = np.array(rhs, dtype=Sinv.dtype).reshape(N, b).copy()
    p = np.zeros(b, dtype=Sinv.dtype)
    q = np.zeros(b, dtype=Sinv.dtype)
    for j in range(max(n_top, n_bot)):
        if j < n_top:
            i = j
            p = Sinv[i] @ (x[i] + (L[i] @ p if j > 0 else 0))
            x[i] = p
        if j < n_bot:
            i = N - 1 - j
            q = Sinv[i] @ (x[i] + (U[i] @ q if j > 0 else 0))
            x[i] = q
    g = x[mid] + (L[mid] @ p if n_top > 0 else 0) + (U[mid] @ q if n_bot > 0 else 0)
    x[mid] = Sinv[mid] @ g
    xt = x[mid].copy()
    xb = x[mid].copy()
    for j in range(max(n_top, n_bot)):
        if j < n_top:
            i = mid - 1 - j
            xt = x[i] + Sinv[i] @ (U[i] @ xt)
            x[i] = xt
        if j < n_bot:
            i = mid + 1 + j
            xb = x[i] + Sinv[i] @ (L[i] @ xb)
            x[i] = xb
    return x.ravel()


def dense(L, D, U, M):
    N, b = len(D), D.shape[1]
    A = np.zeros((b * N, b * N), dtype=np.result_type(D.dtype, type(M)))
    for i in range(N):
        A[b * i:b * i + b, b * i:b * i + b] = M * np.eye(b) - D[i]
        if i > 0:
            A[b * i:b * i + b, b * i - b:b * i] = -L[i]
        if i < N - 1:
            A[b * i:b * i + b, b * i + b:b * i + 2 * b] = -U[i]
    return A
