"""Stand-in for pde.grids.operators.cartesian (py-pde 0.32.2) — test infrastructure.
Raw 1-D stencils acting on a ghost-padded array `arr[N+2]`, writing `out[N]`."""
from numba import njit


def _make_derivative(grid, axis=0, method="central"):
    dx = grid.discretization[axis]
    n = grid.shape[axis]
    if method == "forward":
        @njit
        def diff(arr, out):
            for i in range(1, n + 1):
                out[i - 1] = (arr[i + 1] - arr[i]) / dx
    elif method == "backward":
        @njit
        def diff(arr, out):
            for i in range(1, n + 1):
                out[i - 1] = (arr[i] - arr[i - 1]) / dx
    elif method == "central":
        @njit
        def diff(arr, out):
            for i in range(1, n + 1):
                out[i - 1] = (arr[i + 1] - arr[i - 1]) / (2 * dx)
    else:
        raise ValueError(method)
    return diff


def make_laplace(grid):
    scale = grid.discretization[0] ** -2
    n = grid.shape[0]

    @njit
    def laplace(arr, out):
        for i in range(1, n + 1):
            out[i - 1] = (arr[i - 1] - 2 * arr[i] + arr[i + 1]) * scale
    return laplace
