"""A small stand-in for `simtk.unit`, enough for the reference's call sites on this path
(`298.0 * unit.kelvin`, `1.0 / unit.picoseconds`, `dt.value_in_unit(unit.femtosecond)` ...).

Quantities carry a value and a Unit = (factor to the MD unit system, dimension exponents).
The MD unit system is OpenMM's: nm, ps, amu, kJ/mol (= amu nm^2/ps^2), K, e.  A bare number
handed to the engine is taken to be in those units already, as OpenMM does.
"""
import numpy as _np

_DIMS = ("length", "time", "mass", "temperature", "charge", "amount")


class Unit:
    __array_priority__ = 100  # ndarray * Unit -> one Quantity holding the array (not an object array)

    def __init__(self, factor, dims, name):
        self.factor = float(factor)
        self.dims = tuple(dims)
        self.name = name

    def is_compatible(self, other):
        return isinstance(other, Unit) and self.dims == other.dims

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor * other.factor, [a + b for a, b in zip(self.dims, other.dims)],
                        "%s*%s" % (self.name, other.name))
        if isinstance(other, Quantity):
            return Quantity(other._value, self * other.unit)
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.factor / other.factor, [a - b for a, b in zip(self.dims, other.dims)],
                        "%s/(%s)" % (self.name, other.name))
        if isinstance(other, Quantity):
            return Quantity(1.0 / other._value, self / other.unit)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.factor, [-a for a in self.dims], "1/(%s)" % self.name)
        if isinstance(other, Quantity):
            return Quantity(other._value, other.unit * inv)
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit(self.factor ** p, [a * p for a in self.dims], "(%s)**%s" % (self.name, p))

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and self.factor == other.factor

    def __hash__(self):
        return hash((self.factor, self.dims))

    def __repr__(self):
        return self.name


class Quantity:
    __array_priority__ = 100

    def __init__(self, value, unit):
        self._value = value
        self.unit = unit

    def value_in_unit(self, unit):
        if not self.unit.is_compatible(unit):
            raise TypeError("Unit %r is not compatible with %r" % (self.unit, unit))
        return self._value * (self.unit.factor / unit.factor)

    def in_units_of(self, unit):
        return Quantity(self.value_in_unit(unit), unit)

    def value_in_md_units(self):
        return self._value * self.unit.factor

    def _coerce(self, other):
        return other if isinstance(other, Quantity) else Quantity(other, dimensionless)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self._value, self.unit * other)
        o = self._coerce(other)
        return _strip(Quantity(self._value * o._value, self.unit * o.unit))

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return _strip(Quantity(self._value, self.unit / other))
        o = self._coerce(other)
        return _strip(Quantity(self._value / o._value, self.unit / o.unit))

    def __rtruediv__(self, other):
        o = self._coerce(other)
        return _strip(Quantity(o._value / self._value, o.unit / self.unit))

    def __add__(self, other):
        o = self._coerce(other)
        return Quantity(self._value + o.value_in_unit(self.unit), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        o = self._coerce(other)
        return Quantity(self._value - o.value_in_unit(self.unit), self.unit)

    def __rsub__(self, other):
        o = self._coerce(other)
        return Quantity(o.value_in_unit(self.unit) - self._value, self.unit)

    def __neg__(self):
        return Quantity(-self._value, self.unit)

    def __pow__(self, p):
        return Quantity(self._value ** p, self.unit ** p)

    def _cmp(self, other):
        return self._coerce(other).value_in_unit(self.unit)

    def __lt__(self, o): return self._value < self._cmp(o)
    def __le__(self, o): return self._value <= self._cmp(o)
    def __gt__(self, o): return self._value > self._cmp(o)
    def __ge__(self, o): return self._value >= self._cmp(o)

    def __eq__(self, o):
        try:
            return self._value == self._cmp(o)
        except TypeError:
            return False

    def __hash__(self):
        return hash((self.value_in_md_units(), self.unit.dims))

    def __len__(self):
        return len(self._value)

    def __getitem__(self, i):
        return Quantity(self._value[i], self.unit)

    def __float__(self):
        if any(self.unit.dims):
            raise TypeError("cannot convert a dimensioned Quantity to float")
        return float(self._value * self.unit.factor)

    def __repr__(self):
        return "Quantity(value=%r, unit=%r)" % (self._value, self.unit)

    def __str__(self):
        return "%s %s" % (self._value, self.unit)


def _strip(q):
    """dimensionless results collapse to plain numbers, as simtk.unit does"""
    if not any(q.unit.dims):
        return q._value * q.unit.factor
    return q


def is_quantity(x):
    return isinstance(x, Quantity)


def md_value(x, expected=None):
    """plain number (already MD units) or Quantity -> float/array in MD units"""
    if isinstance(x, Quantity):
        if expected is not None and not x.unit.is_compatible(expected):
            raise TypeError("expected a quantity compatible with %r, got %r" % (expected, x.unit))
        return x.value_in_md_units()
    return x


def _u(factor, name, **dims):
    return Unit(factor, [dims.get(d, 0) for d in _DIMS], name)


dimensionless = _u(1.0, "dimensionless")
nanometer = nanometers = _u(1.0, "nm", length=1)
angstrom = angstroms = _u(0.1, "A", length=1)
picosecond = picoseconds = _u(1.0, "ps", time=1)
femtosecond = femtoseconds = _u(1e-3, "fs", time=1)
nanosecond = nanoseconds = _u(1e3, "ns", time=1)
amu = dalton = daltons = _u(1.0, "amu", mass=1)
kelvin = _u(1.0, "K", temperature=1)
elementary_charge = elementary_charges = _u(1.0, "e", charge=1)
mole = moles = _u(1.0, "mol", amount=1)
kilojoule_per_mole = kilojoules_per_mole = _u(1.0, "kJ/mol", mass=1, length=2, time=-2)
kilocalorie_per_mole = kilocalories_per_mole = _u(4.184, "kcal/mol", mass=1, length=2, time=-2)
radian = radians = _u(1.0, "rad")
degree = degrees = _u(_np.pi / 180.0, "deg")
atmosphere = atmospheres = _u(1.0, "atm")  # only carried around; the engine has no barostat

# kB * N_A in kJ/mol/K (CODATA 2018, the value OpenMM >= 7.6 uses).  OpenMM 7.1-7.5, which the
# reference ran on, used 8.314472e-3; the relative difference (1.1e-6) is below the parity budget.
MOLAR_GAS_CONSTANT_R = Quantity(8.31446261815324e-3, kilojoule_per_mole / kelvin)
