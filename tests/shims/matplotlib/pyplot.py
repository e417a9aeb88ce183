"""`import matplotlib.pyplot as plt` no-op: subplots() returns objects that swallow every call
(Evolve_scenario.py:190-205 plot / set_xlabel / legend / suptitle / savefig)."""


class _Sink:
    def __getattr__(self, name):
        return lambda *a, **k: None


def subplots(*args, **kwargs):
    return _Sink(), _Sink()
