"""Ideal focusing optics (reference marxs/optics/mirror.py)."""
from ..program import UnsupportedCallable
from .base import FlatOpticalElement

__all__ = ['PerfectLens']


class PerfectLens(FlatOpticalElement):
    """Infinitely large lens that focuses every ray exactly (reference :12-82).

    ``reflectivity_interpolator`` is only supported at its default (probability 1)."""

    display = {'color': (0., 0.5, 0.), 'opacity': 0.5, 'shape': 'box'}
    loc_coos_name = ['mirror_x', 'mirror_y']

    def __init__(self, **kwargs):
        self.focallength = kwargs.pop('focallength')
        self.d_center_optax = kwargs.pop('d_center_optical_axis', 0)
        self.reflectivity_interpolator = kwargs.pop('reflectivity_interpolator', None)
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        if self.reflectivity_interpolator is not None:
            raise UnsupportedCallable('PerfectLens.reflectivity_interpolator is not supported on the device yet')
        p_opt_axis = self.geometry['center'] - self.d_center_optax * self.geometry['e_z']
        lw.op('LENS', pf=lw.eparams([p_opt_axis[0], p_opt_axis[1], p_opt_axis[2], self.focallength]))
