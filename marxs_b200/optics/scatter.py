"""Statistical scatter added at the point of last interaction (reference marxs/optics/scatter.py)."""
from .base import FlatOpticalElement

__all__ = ['RadialMirrorScatter', 'RandomGaussianScatter']


def _rad(x):
    """Angle in radian from a float or an astropy-like Quantity."""
    if hasattr(x, 'to'):
        import astropy.units as u
        return float(x.to(u.rad).value)
    return float(x)


class RadialMirrorScatter(FlatOpticalElement):
    """Gaussian scatter in / perpendicular to the plane of reflection (reference :24-77).
    ``inplanescatter`` / ``perpplanescatter``: sigma in rad (float or Quantity)."""

    inplanescattercol = 'inplanescatter'
    perpplanescattercol = 'perpplanescatter'

    def __init__(self, **kwargs):
        self.inplanescatter = _rad(kwargs.pop('inplanescatter'))
        self.perpplanescatter = _rad(kwargs.pop('perpplanescatter', 0.))
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        c = self.pos4d[:-1, -1]
        lw.op('RSCATTER', pf=lw.eparams([c[0], c[1], c[2], self.inplanescatter, self.perpplanescatter]),
              cols=[lw.fcol(self.inplanescattercol), lw.fcol(self.perpplanescattercol)],
              s0=lw.slot('normal'), s1=lw.slot('normal'))


class RandomGaussianScatter(FlatOpticalElement):
    """Scatter by a Gaussian angle in a random direction (reference :80-145); constant sigma only."""

    scattername = 'scatter'

    def __init__(self, **kwargs):
        if 'scatter' in kwargs:
            self.scatter = kwargs.pop('scatter')
        elif not hasattr(self, 'scatter'):
            raise ValueError('Keyword "scatter" missing.')
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        if callable(self.scatter):
            from ..program import UnsupportedCallable
            raise UnsupportedCallable('callable scatter(photons, ...) is not supported on the device '
                                      '(L2Diffraction has its own fused form)')
        sigma = _rad(self.scatter)
        if sigma == 0:
            return                      # reference returns {} : nothing changes (:121-123)
        lw.op('GSCATTER', pf=lw.eparams([sigma]), cols=[lw.fcol(self.scattername)],
              s0=lw.slot('normal'), s1=lw.slot('uniform'))
