"""`import matplotlib` no-op (Evolve_scenario.py:15-17: only `matplotlib.use("AGG")`)."""


def use(backend, **kwargs):
    return None
