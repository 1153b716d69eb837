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

    _ANGLE_COLUMN = '_mxb_angle'      # per-photon angle of a callable scatter (exists only during the launch)

    def _can_lower(self):
        # a callable scatter(photons, intersect, interpos, intercoos) needs the intersect first: kernel
        # intersect, the user function on the device table, then the rotation kernel with the angle read per photon
        return super()._can_lower() and not callable(self.scatter)

    def process_photons(self, photons, intersect, interpos, intercoos):
        if not callable(self.scatter) or not FlatOpticalElement._can_lower(self):
            return super().process_photons(photons, intersect, interpos, intercoos)
        import numpy as np
        import torch
        hit = torch.as_tensor(intersect, device=photons.device)
        try:
            angle = self.scatter(photons, intersect, interpos, intercoos)
        except (TypeError, RuntimeError):
            # numpy-only user function: evaluate it on a host copy of the table (user code, not the trace path)
            def host(x):
                return torch.as_tensor(x).cpu().numpy()
            angle = self.scatter(photons.to_numpy(), host(intersect), host(interpos), host(intercoos))
        if hasattr(angle, 'to') and hasattr(angle, 'unit'):      # astropy Quantity (reference: .to(u.rad).value)
            angle = angle.to('rad').value
        if hasattr(angle, 'data') and not isinstance(angle, (torch.Tensor, np.ndarray)):
            angle = angle.data                                     # column facade -> tensor
        angle = torch.as_tensor(angle if isinstance(angle, torch.Tensor) else np.asarray(angle, dtype=float),
                                dtype=torch.float64, device=photons.device)
        n, n_hit = len(photons), int(hit.sum())
        if angle.ndim == 0:
            angle = angle.expand(n_hit)
        full = torch.full((n,), float('nan'), dtype=torch.float64, device=photons.device)
        if angle.shape[0] == n_hit:
            full[hit] = angle
        elif angle.shape[0] == n:
            full = torch.where(hit, angle, full)
        else:
            raise ValueError('scatter(photons, intersect, interpos, intercoos) must return one angle per '
                             'intersecting photon')
        photons[self._ANGLE_COLUMN] = full
        try:
            return self._process_photons_kernel(photons, intersect, interpos, intercoos)
        finally:
            if self._ANGLE_COLUMN in photons:
                photons.remove_column(self._ANGLE_COLUMN)

    def _lower_specific(self, lw):
        if callable(self.scatter):
            if self._ANGLE_COLUMN not in lw.existing:
                from ..program import UnsupportedCallable
                raise UnsupportedCallable('a callable scatter(photons, ...) cannot be fused: call the element on a '
                                          'photon table (kernel intersect, scatter(), rotation kernel)')
            lw.op('GSCATTER', flags=2, pf=lw.eparams([0.]), cols=[lw.fcol(self.scattername), lw.fcol(self._ANGLE_COLUMN)],
                  s0=-1, s1=lw.slot('uniform'))
            return
        sigma = _rad(self.scatter)
        if sigma == 0:
            return                      # reference returns {} : nothing changes (:121-123)
        lw.op('GSCATTER', pf=lw.eparams([sigma]), cols=[lw.fcol(self.scattername)],
              s0=lw.slot('normal'), s1=lw.slot('uniform'))
