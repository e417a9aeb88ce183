"""`from pint import UnitRegistry` (parameters.py:2, :6-7).  Units are cosmetic upstream: only
`.magnitude` is consumed (parameters.py:99), so a quantity here is a number that survives
multiplication / division by unit symbols."""


class Quantity:
    __slots__ = ("magnitude", "units")

    def __init__(self, magnitude=1.0, units=""):
        self.magnitude, self.units = magnitude, units

    @staticmethod
    def _num(other):
        return other.magnitude if isinstance(other, Quantity) else other

    def __mul__(self, other):
        return Quantity(self.magnitude * self._num(other))

    def __rmul__(self, other):
        return Quantity(self._num(other) * self.magnitude)

    def __truediv__(self, other):
        return Quantity(self.magnitude / self._num(other))

    def __rtruediv__(self, other):
        return Quantity(self._num(other) / self.magnitude)

    def __pow__(self, exponent):
        return Quantity(self.magnitude ** exponent)

    def __repr__(self):
        return f"<Quantity({self.magnitude})>"


class UnitRegistry:
    Quantity = Quantity

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Quantity(1.0, name)
