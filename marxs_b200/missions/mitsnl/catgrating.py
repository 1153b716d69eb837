"""CAT-grating efficiency table and non-parallel bars
(reference marxs/missions/mitsnl/catgrating.py:63-144, 280-311)."""
import numpy as np

from ...optics import CATGrating
from ...program import SEL_INTERPTABLE

__all__ = ['InterpolateEfficiencyTable', 'NonParallelCATGrating']


class InterpolateEfficiencyTable:
    """Order selector from a (wavelength, blaze angle, order) efficiency table with
    bilinear interpolation (the reference's ``RectBivariateSpline(kx=ky=1)``).

    Parameters
    ----------
    wave : (nw,) wavelengths as they stand in the table (nm)
    theta : (nt,) blaze angles in rad
    prob : (nw, nt, n_orders) efficiencies
    orders : (n_orders,) integers
    """

    def __init__(self, wave, theta, prob, orders):
        self.wave = np.asarray(wave, dtype=float)
        self.theta = np.asarray(theta, dtype=float)
        self.prob = np.ascontiguousarray(prob, dtype=float)
        self.orders = np.asarray(orders)
        if self.prob.shape != (len(self.wave), len(self.theta), len(self.orders)):
            raise ValueError('prob must have shape (len(wave), len(theta), len(orders))')

    @classmethod
    def from_table(cls, tab):
        """From a flattened table with columns wave, theta[deg], then one column per order
        (the reference's file layout, marxs/utils.py:56-100)."""
        names = list(tab.keys()) if isinstance(tab, dict) else list(tab.colnames)
        x, y = np.asarray(tab[names[0]], dtype=float), np.asarray(tab[names[1]], dtype=float)
        n_x, n_y = len(set(x)), len(set(y))
        if len(x) != n_x * n_y:
            raise ValueError('Data is not on regular grid.')
        prob = np.stack([np.asarray(tab[n], dtype=float).reshape(n_x, n_y) for n in names[2:]], axis=-1)
        return cls(x[::n_y], np.deg2rad(y[:n_y]), prob, [int(n) for n in names[2:]])

    def device_table(self):
        nw, nt, no = self.prob.shape
        block = np.concatenate([[SEL_INTERPTABLE, nw, nt, no, -1.], self.wave, self.theta,
                                self.orders.astype(float)])
        return block, self.prob


class NonParallelCATGrating(CATGrating):
    """CAT grating whose bar side-walls tilt linearly across the facet:
    blaze += blaze_center + d_blaze_mm * local_y (reference :280-311)."""

    def __init__(self, **kwargs):
        self.d_blaze_mm = kwargs.pop('d_blaze_mm', 0)
        self.blaze_center = kwargs.pop('blaze_center', 0)
        super().__init__(**kwargs)

    def _blaze_modifier(self):
        return (self.blaze_center, self.d_blaze_mm)
