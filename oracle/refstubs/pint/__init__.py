"""Stand-in for pint 0.22 (test infrastructure; see ../README.md).
Only what marlpde/parameters.py:2-48,:99 uses: UnitRegistry attributes, Quantity, * / **, .magnitude."""


class Quantity:
    def __init__(self, magnitude=1.0, units=""):
        self.magnitude = magnitude
        self.units = units

    def _m(self, o):
        return o.magnitude if isinstance(o, Quantity) else o

    def __mul__(self, o):
        return Quantity(self.magnitude * self._m(o))

    def __rmul__(self, o):
        return Quantity(self._m(o) * self.magnitude)

    def __truediv__(self, o):
        # units are symbolic upstream: dividing by a *unit* (magnitude 1) leaves the number intact
        return Quantity(self.magnitude / self._m(o))

    def __rtruediv__(self, o):
        return Quantity(self._m(o) / self.magnitude)

    def __pow__(self, e):
        return Quantity(self.magnitude ** e)


class UnitRegistry:
    Quantity = Quantity

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Quantity(1.0, name)
