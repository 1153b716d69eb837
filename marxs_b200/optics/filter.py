"""Energy-dependent filters (reference marxs/optics/filter.py).

``filterfunc`` runs on the device: numbers, tabulated curves (``Tabulated1D`` or a
scipy ``interp1d``-like object with ``x``/``y``) are fused into the trace kernel;
any other callable is evaluated on the device energy tensor with torch semantics
(no CPU fallback) and must therefore accept a torch tensor."""
import numpy as np
import torch

from ..base import MarxsElement
from ..program import NotFusable
from ..simulator import run_fused
from .base import FlatOpticalElement

__all__ = ['Tabulated1D', 'EnergyFilter', 'GlobalEnergyFilter']


class Tabulated1D:
    """Linear interpolation table p(E) with the out-of-range modes of scipy.interpolate.interp1d:
    ``bounds_error=True`` (scipy's default) raises ValueError, otherwise ``fill_value`` applies - a number
    or a ``(below, above)`` pair (NaN by default, like scipy) or ``'extrapolate'``."""

    def __init__(self, x, y, bounds_error=True, fill_value=np.nan):
        self.x = np.asarray(x, dtype=float)
        self.y = np.asarray(y, dtype=float)
        self.bounds_error = bounds_error
        self.fill_value = fill_value
        if np.any(np.diff(self.x) <= 0):
            raise ValueError('x must be strictly increasing')


def _fill_mode(f):
    """(flags, extra params) of the out-of-range behaviour of a Tabulated1D / scipy interp1d object."""
    if getattr(f, 'bounds_error', True):
        return 1, []
    fill = getattr(f, 'fill_value', np.nan)
    if (isinstance(fill, str) and fill == 'extrapolate') or getattr(f, '_extrapolate', False):
        return 4, []
    fill = np.asarray(fill, dtype=float).ravel()      # scipy stores (below, above) or one broadcastable value
    if fill.size == 1:
        return 2, [float(fill[0]), float(fill[0])]
    if fill.size == 2:
        return 2, [float(fill[0]), float(fill[1])]
    raise NotFusable('unsupported fill_value of the filter table')


def lower_filter(f):
    """-> (params, flags) for MXB_OP_FILTER / GFILTER, or raises NotFusable."""
    if np.isscalar(f):
        if f < 0. or f > 1.:
            raise ValueError('Probabilities returned by filterfunc must be in interval [0, 1].')
        return np.array([0., float(f)]), 0
    if isinstance(f, Tabulated1D) or (hasattr(f, 'x') and hasattr(f, 'y') and
                                      getattr(f, '_kind', 'linear') == 'linear'):
        x, y = np.asarray(f.x, dtype=float), np.asarray(f.y, dtype=float)
        if x.ndim != 1 or y.shape != x.shape:
            raise NotFusable('filter table must be one-dimensional')
        if np.any(y < 0.) or np.any(y > 1.):
            raise ValueError('Probabilities returned by filterfunc must be in interval [0, 1].')
        flags, extra = _fill_mode(f)
        for v in extra:
            if v < 0. or v > 1.:       # NaN passes, as in the reference's range check
                raise ValueError('Probabilities returned by filterfunc must be in interval [0, 1].')
        return np.concatenate([[len(x)], x, y, extra]), flags
    raise NotFusable('filterfunc is an arbitrary callable')


def _torch_filter(f, energy):
    p = torch.as_tensor(f(energy), device=energy.device, dtype=torch.float64)
    if bool(torch.any(p < 0.)) or bool(torch.any(p > 1.)):
        raise ValueError('Probabilities returned by filterfunc must be in interval [0, 1].')
    return p


class GlobalEnergyFilter(MarxsElement):
    """Energy-dependent probability factor applied to ALL photons (reference :10-53)."""

    def __init__(self, **kwargs):
        self.filterfunc = kwargs.pop('filterfunc')
        super().__init__(**kwargs)

    def _can_lower(self):
        try:
            lower_filter(self.filterfunc)
            return True
        except NotFusable:
            return False

    def _lower(self, lw):
        params, flags = lower_filter(self.filterfunc)
        lw.op('GFILTER', flags=flags, pg=lw.params(params))

    def _call_unfused(self, photons):
        photons['probability'] *= _torch_filter(self.filterfunc, photons['energy'])
        return photons

    def __call__(self, photons):
        return run_fused([self], photons) if self._can_lower() else self._call_unfused(photons)


class EnergyFilter(FlatOpticalElement):
    """Energy-dependent filter with a position and size (reference :56-94)."""

    display = {'color': (1.0, 0., 0.), 'opacity': 0.5, 'shape': 'box'}

    def __init__(self, **kwargs):
        self.filterfunc = kwargs.pop('filterfunc')
        super().__init__(**kwargs)

    def _can_lower(self):
        if not super()._can_lower():
            return False
        try:
            lower_filter(self.filterfunc)
            return True
        except NotFusable:
            return False

    def _lower_specific(self, lw):
        params, flags = lower_filter(self.filterfunc)
        lw.op('FILTER', flags=flags, pg=lw.params(params))

    def specific_process_photons(self, photons, intersect, interpos, intercoos):
        return {'probability': _torch_filter(self.filterfunc, photons['energy'][intersect])}
