"""`from pde.grids.operators.cartesian import _make_derivative` (Evolve_scenario.py:12, :43-46).

The reference registers two factories built from this function on its grid; on the GPU path the
factories are stored and never called — forward / backward differences are formed inside
csrc/lheureux_device.cuh.  Calling the returned operator is therefore an error, not a fallback."""


def _make_derivative(grid, axis=0, method="central"):
    def stencil(arr, out):
        raise NotImplementedError(f"py-pde '{method}' derivative stencil: evaluated by the CUDA kernels, "
                                  "there is no CPU path on this side of the boundary")
    stencil.method = method
    return stencil
