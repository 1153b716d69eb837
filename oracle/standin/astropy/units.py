"""Unit scalars + a thin ndarray-subclass Quantity (stand-in)."""
import contextlib
import functools
import math
import numpy as np

_BASES = ('m', 's', 'rad', 'J')
_HC_J_M = 6.62607015e-34 * 299792458.0  # CODATA 2018 exact h*c


class UnitConversionError(Exception):
    pass


class Unit:
    __array_ufunc__ = None  # make ndarray defer to Unit.__rmul__

    def __init__(self, scale, dims, name=None):
        self.scale = float(scale)
        self.dims = tuple(dims)
        self.name = name

    # -- algebra -----------------------------------------------------
    def _comb(self, other, sign):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale ** sign,
                        [a + sign * b for a, b in zip(self.dims, other.dims)])
        return NotImplemented

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._comb(other, 1)
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._comb(other, -1)
        return Quantity(1. / np.asanyarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit(1. / self.scale, [-d for d in self.dims]))

    def __pow__(self, p):
        return Unit(self.scale ** p, [d * p for d in self.dims])

    def __eq__(self, other):
        return (isinstance(other, Unit) and self.dims == other.dims
                and math.isclose(self.scale, other.scale, rel_tol=1e-14))

    def __hash__(self):
        return hash(self.dims)

    def __repr__(self):
        return 'Unit({0})'.format(self.name or (self.scale, self.dims))

    def to(self, other, value=1.0, equivalencies=None):
        return _convert(value, self, other, equivalencies)

    def decompose(self):
        return self

    def is_equivalent(self, other):
        return self.dims == other.dims


def _convert(value, src, dst, equivalencies=None):
    value = np.asanyarray(value, dtype=float).view(np.ndarray)
    if src.dims == dst.dims:
        return value * (src.scale / dst.scale)
    # numpy's inverse trigonometric ufuncs drop the unit here (astropy returns radians): a bare
    # number converted to an angle is taken to be in radians
    if src.dims == (0, 0, 0, 0) and dst.dims == (0, 0, 1, 0):
        return value * (src.scale / dst.scale)
    eq = set(equivalencies or ()) | set(_ENABLED[-1])
    # angles are dimensionless
    strip = lambda d: (d[0], d[1], 0, d[3])
    if 'dimensionless_angles' in eq and strip(src.dims) == strip(dst.dims):
        return value * (src.scale / dst.scale)
    if 'spectral' in eq:
        E, L = (0, 0, 0, 1), (1, 0, 0, 0)
        if src.dims == E and dst.dims == L:
            return _HC_J_M / (value * src.scale) / dst.scale
        if src.dims == L and dst.dims == E:
            return _HC_J_M / (value * src.scale) / dst.scale
    raise UnitConversionError('{0} -> {1}'.format(src, dst))


class Quantity(np.ndarray):
    __array_priority__ = 10000

    def __new__(cls, value, unit=None, dtype=float, copy=True):
        if isinstance(value, Quantity) and unit is None:
            unit = value.unit
        elif isinstance(value, (list, tuple)) and len(value) and \
                isinstance(value[0], Quantity):
            u0 = value[0].unit if unit is None else unit
            value = [v.to(u0).value for v in value]
            unit = u0
        elif isinstance(value, Quantity) and unit is not None:
            value = value.to(unit).value
        obj = np.array(value, dtype=dtype).view(cls)
        obj.unit = unit if unit is not None else dimensionless_unscaled
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, 'unit', dimensionless_unscaled)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        """exp / log need a dimensionless argument and astropy first converts a SCALED dimensionless
        unit (mm / um = 1000) to a bare number (catgrating.py:192 relies on it); every other ufunc
        keeps the plain ndarray-subclass behaviour (result carries the unit of the first Quantity)."""
        unscale = ufunc in _UNSCALED_UFUNCS
        args, first_unit = [], None
        for x in inputs:
            if isinstance(x, Quantity):
                v = x.view(np.ndarray)
                if first_unit is None:
                    first_unit = x.unit
                if unscale:
                    if any(x.unit.dims):
                        raise UnitConversionError('{0} needs a dimensionless argument'.format(ufunc.__name__))
                    v = v * x.unit.scale
                args.append(v)
            else:
                args.append(x)
        if out is not None:
            kwargs['out'] = tuple(o.view(np.ndarray) if isinstance(o, Quantity) else o for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        unit = dimensionless_unscaled if unscale else first_unit

        def wrap(r):
            if isinstance(r, np.ndarray) and r.dtype.kind == 'f':
                q = r.view(Quantity)
                q.unit = unit
                return q
            if isinstance(r, (float, np.floating)):
                return Quantity(r, unit)
            return r
        return tuple(wrap(r) for r in res) if isinstance(res, tuple) else wrap(res)

    @property
    def value(self):
        v = self.view(np.ndarray)
        return v if v.ndim else float(v)

    @property
    def isscalar(self):
        return self.ndim == 0

    def to(self, unit, equivalencies=None):
        return Quantity(_convert(self.view(np.ndarray), self.unit, unit,
                                 equivalencies), unit)

    def to_value(self, unit, equivalencies=None):
        return self.to(unit, equivalencies).value

    def decompose(self):
        return Quantity(self.view(np.ndarray) * self.unit.scale,
                        Unit(1., self.unit.dims))

    @property
    def si(self):
        return self.decompose()

    def _binary_unit(self, other, sign):
        ou = other.unit if isinstance(other, Quantity) else (
            other if isinstance(other, Unit) else dimensionless_unscaled)
        return self.unit._comb(ou, sign)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.view(np.ndarray), self.unit * other)
        out = np.multiply(self.view(np.ndarray), np.asanyarray(other).view(np.ndarray))
        return Quantity(out, self._binary_unit(other, 1))

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.view(np.ndarray), self.unit / other)
        out = np.divide(self.view(np.ndarray), np.asanyarray(other).view(np.ndarray))
        return Quantity(out, self._binary_unit(other, -1))

    def __rtruediv__(self, other):
        out = np.divide(np.asanyarray(other).view(np.ndarray), self.view(np.ndarray))
        return Quantity(out, Unit(1. / self.unit.scale, [-d for d in self.unit.dims]))

    def __pow__(self, p):
        if np.ndim(p) > 0:
            # array exponents need a dimensionless base (astropy raises otherwise)
            if any(self.unit.dims):
                raise UnitConversionError('array exponent needs a dimensionless base')
            return Quantity((self.view(np.ndarray) * self.unit.scale) ** np.asanyarray(p).view(np.ndarray),
                            dimensionless_unscaled)
        return Quantity(self.view(np.ndarray) ** p, self.unit ** p)


_UNSCALED_UFUNCS = (np.exp, np.exp2, np.expm1, np.log, np.log2, np.log10, np.log1p)

dimensionless_unscaled = Unit(1., (0, 0, 0, 0), 'dimensionless')
one = dimensionless_unscaled
m = Unit(1., (1, 0, 0, 0), 'm')
mm = Unit(1e-3, (1, 0, 0, 0), 'mm')
cm = Unit(1e-2, (1, 0, 0, 0), 'cm')
um = Unit(1e-6, (1, 0, 0, 0), 'um')
micron = um
micrometer = um
nm = Unit(1e-9, (1, 0, 0, 0), 'nm')
Angstrom = Unit(1e-10, (1, 0, 0, 0), 'Angstrom')
AA = Angstrom
s = Unit(1., (0, 1, 0, 0), 's')
ks = Unit(1e3, (0, 1, 0, 0), 'ks')
rad = Unit(1., (0, 0, 1, 0), 'rad')
radian = rad
deg = Unit(math.pi / 180., (0, 0, 1, 0), 'deg')
degree = deg
arcmin = Unit(math.pi / 180. / 60., (0, 0, 1, 0), 'arcmin')
arcsec = Unit(math.pi / 180. / 3600., (0, 0, 1, 0), 'arcsec')
J = Unit(1., (0, 0, 0, 1), 'J')
eV = Unit(1.602176634e-19, (0, 0, 0, 1), 'eV')
keV = Unit(1.602176634e-16, (0, 0, 0, 1), 'keV')
photon = dimensionless_unscaled
ph = photon
ct = photon
pix = dimensionless_unscaled
pixel = pix

_ENABLED = [()]


def spectral():
    return ('spectral',)


def dimensionless_angles():
    return ('dimensionless_angles',)


@contextlib.contextmanager
def set_enabled_equivalencies(eq):
    _ENABLED.append(tuple(eq))
    try:
        yield
    finally:
        _ENABLED.pop()


def quantity_input(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(f):
        @functools.wraps(f)
        def wrapper(*a, **k):
            return f(*a, **k)
        return wrapper
    return deco
