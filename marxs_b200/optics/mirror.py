"""Ideal focusing optics (reference marxs/optics/mirror.py)."""
import numpy as np

from ..program import NotFusable, UnsupportedCallable
from .base import FlatOpticalElement

__all__ = ['PerfectLens', 'ReflectivityTable']


class ReflectivityTable:
    """Reflectivity R(energy [keV], grazing angle [rad]) tabulated on a rectilinear grid and interpolated
    bilinearly with the query clamped to the table: exactly what the reference's usual interpolator,
    ``scipy.interpolate.RectBivariateSpline(energy, angle, refl, kx=1, ky=1)``, computes
    (reference optics/mirror.py:31-36, optics/tests/test_mirror.py:62-80).  Callable like it on the host."""

    def __init__(self, energy, angle, refl):
        self.x = np.ascontiguousarray(energy, dtype=float)
        self.y = np.ascontiguousarray(angle, dtype=float)
        self.z = np.ascontiguousarray(refl, dtype=float)
        if self.x.ndim != 1 or self.y.ndim != 1 or self.z.shape != (len(self.x), len(self.y)):
            raise ValueError('refl must have shape (len(energy), len(angle))')
        if len(self.x) < 2 or len(self.y) < 2 or np.any(np.diff(self.x) <= 0) or np.any(np.diff(self.y) <= 0):
            raise ValueError('energy and angle must be strictly increasing with at least two points each')

    @classmethod
    def from_interpolator(cls, interp):
        """From a degree-(1, 1) ``RectBivariateSpline``: its interior knots are the grid and its values on
        the grid the table."""
        if isinstance(interp, cls):
            return interp
        if tuple(getattr(interp, 'degrees', ())) == (1, 1) and hasattr(interp, 'get_knots'):
            tx, ty = interp.get_knots()
            x, y = np.asarray(tx, dtype=float)[1:-1], np.asarray(ty, dtype=float)[1:-1]
            return cls(x, y, np.asarray(interp(x, y), dtype=float))
        raise UnsupportedCallable('PerfectLens.reflectivity_interpolator must be a ReflectivityTable or a '
                                  'scipy RectBivariateSpline with kx = ky = 1 to run on the device; got {0!r}'.format(interp))

    def __call__(self, energy, angle, grid=False):
        if grid:
            raise ValueError('only grid=False is supported')
        xq = np.clip(np.asarray(energy, dtype=float), self.x[0], self.x[-1])
        yq = np.clip(np.asarray(angle, dtype=float), self.y[0], self.y[-1])
        i = np.clip(np.searchsorted(self.x, xq, side='right') - 1, 0, len(self.x) - 2)
        j = np.clip(np.searchsorted(self.y, yq, side='right') - 1, 0, len(self.y) - 2)
        tx = (xq - self.x[i]) / (self.x[i + 1] - self.x[i])
        ty = (yq - self.y[j]) / (self.y[j + 1] - self.y[j])
        z = self.z
        f0 = z[i, j] + tx * (z[i + 1, j] - z[i, j])
        f1 = z[i, j + 1] + tx * (z[i + 1, j + 1] - z[i, j + 1])
        return f0 + ty * (f1 - f0)


class PerfectLens(FlatOpticalElement):
    """Infinitely large lens that focuses every ray exactly (reference :12-82).

    ``reflectivity_interpolator`` (default: reflectivity 1) is a `ReflectivityTable` or the reference's
    ``RectBivariateSpline(kx=1, ky=1)``; the probability is multiplied by R(energy, angle / 4)**2 with
    angle = arccos|d_new . d_old| (reference :68-81)."""

    display = {'color': (0., 0.5, 0.), 'opacity': 0.5, 'shape': 'box'}
    loc_coos_name = ['mirror_x', 'mirror_y']

    def __init__(self, **kwargs):
        self.focallength = kwargs.pop('focallength')
        self.d_center_optax = kwargs.pop('d_center_optical_axis', 0)
        self.reflectivity_interpolator = kwargs.pop('reflectivity_interpolator', None)
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        p_opt_axis = self.geometry['center'] - self.d_center_optax * self.geometry['e_z']
        base = [p_opt_axis[0], p_opt_axis[1], p_opt_axis[2], self.focallength]
        if self.reflectivity_interpolator is None:
            lw.op('LENS', pf=lw.eparams(base))
            return
        tab = ReflectivityTable.from_interpolator(self.reflectivity_interpolator)
        if lw.array is not None:
            raise NotFusable('a PerfectLens with a reflectivity table cannot be a facet of a Parallel')
        block = lw.params_with_table(base + [len(tab.x), len(tab.y), -1.], 6, np.concatenate([tab.x, tab.y, tab.z.ravel()]))
        lw.op('LENS', flags=1, pf=block)
