"""Drop-in mirror of the reference's `marlpde` package (same module names, class names,
signatures and dictionary keys) whose hot path runs on the B200 through marlpde_b200's C ABI.

Like the reference (marlpde/__init__.py:1-2) the package directory is put on sys.path so that the
bare-module imports used inside the package (`from parameters import ...`) and the prefixed ones
used by the tests (`from marlpde.parameters import ...`) both work."""
import pathlib
import sys

_here = pathlib.Path(__file__).parent
for _p in (str(_here), str(_here.parent)):
    if _p not in sys.path:
        sys.path.append(_p)
