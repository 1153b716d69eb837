"""Apertures: where parallel rays from a source start the simulation
(reference marxs/optics/aperture.py).  Photon start positions are born on the device."""
import numpy as np
import torch

from ..base import GeometryError
from ..geometry import RectangleHole, CircularHole
from ..program import NotFusable
from ..simulator import BaseContainer, run_fused
from .base import FlatOpticalElement

__all__ = ['RectangleAperture', 'CircleAperture', 'MultiAperture']


class BaseAperture:
    display = {'color': (0.0, 0.75, 0.75), 'opacity': 0.3, 'shape': 'triangulation'}

    @staticmethod
    def add_colpos(photons):
        """Add the ``pos`` column (w = 1) if missing (reference :19-27)."""
        if 'pos' not in photons.colnames:
            t = photons.new_column('pos', torch.float64, fill=0., vector=True)
            t[3] = 1.


class FlatAperture(BaseAperture, FlatOpticalElement):
    """Flat aperture: position = center + x v_y + y v_z; probability *= projected area
    (reference :39-82)."""
    _circle = False

    def _aperture_params(self, cum_lo=0., cum_hi=1.):
        g = self.geometry
        c, vy, vz, ex = g['center'], g['v_y'], g['v_z'], g['e_x']
        phi = getattr(g, 'phi', [0., 0.])
        r_in = g['r_inner'] / np.linalg.norm(g['v_y'])
        return np.concatenate([c[:3], vy[:3], vz[:3], -ex[:3],
                               [phi[0], phi[1] - phi[0], r_in ** 2, cum_lo, cum_hi]])

    def _can_lower(self):
        return '_lower_specific' not in type(self).__dict__ and super()._can_lower()

    def _lower_specific(self, lw):
        raise NotFusable('apertures are lowered as a whole')

    def _lower(self, lw, aper_slot=-1, aper_id=0, cum=(0., 1.), slots=None):
        if lw.array is not None:
            raise NotFusable('apertures cannot be facets of a Parallel')
        lw.needs_pos = True
        if slots is None:
            slots = (lw.slot('uniform'), lw.slot('uniform'))
        lw.op('APERTURE', flags=1 if self._circle else 0, pf=lw.eparams(self._aperture_params(*cum)),
              s0=slots[0], s1=slots[1], w14=aper_slot, w15=aper_id)
        lw.commit(self.loc_coos_name, self.id_col, self.id_num)
        return slots

    def __call__(self, photons):
        self.add_colpos(photons)
        return run_fused([self], photons)


class AreaMM2(float):
    """An aperture area in mm**2.  The reference returns ``... * u.mm**2`` (aperture.py:100,152,199) and
    sources convert it against a flux per cm**2 (basesources.py:174); astropy is not a dependency here, so
    the unit travels as the type: ``PointSource(geomarea=aperture.area)`` divides by 100, a bare number
    is taken as cm**2.  Scaling by plain numbers and sums of areas keep the unit."""
    unit = 'mm2'

    @property
    def value(self):
        return float(self)

    def to(self, unit):
        name = unit if isinstance(unit, str) else getattr(unit, 'to_string', lambda: str(unit))()
        name = name.replace(' ', '').replace('**', '').replace('^', '')
        if name == 'mm2':
            return float(self)
        if name == 'cm2':
            return float(self) / 100.
        raise ValueError('AreaMM2 converts to mm2 or cm2, not {0}'.format(unit))

    def __mul__(self, other):
        return AreaMM2(float(self) * other) if type(other) in (int, float) else float.__mul__(self, other)
    __rmul__ = __mul__

    def __truediv__(self, other):
        return AreaMM2(float(self) / other) if type(other) in (int, float) else float.__truediv__(self, other)

    def __add__(self, other):
        return AreaMM2(float(self) + float(other)) if isinstance(other, AreaMM2) else float.__add__(self, other)

    def __radd__(self, other):
        if isinstance(other, AreaMM2) or (type(other) is int and other == 0):      # sum() starts at 0
            return AreaMM2(float(self) + float(other))
        return float.__radd__(self, other)


class RectangleAperture(FlatAperture):
    """Rectangular opening (reference :85-100)."""
    default_geometry = RectangleHole

    @property
    def area(self):
        return AreaMM2(4 * np.linalg.norm(self.geometry['v_y']) * np.linalg.norm(self.geometry['v_z']))


class CircleAperture(FlatAperture):
    """Circular / ring / wedge opening (reference :103-152); ``phi``, ``r_inner`` keywords."""
    default_geometry = CircularHole
    _circle = True

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.geometry['r_inner'] > np.linalg.norm(self.geometry['v_y']):
            raise ValueError('r_inner must be less than size of full aperture.')
        if not np.isclose(np.linalg.norm(self.geometry['v_y']), np.linalg.norm(self.geometry['v_z'])):
            raise GeometryError('Aperture does not have the same size in y, z direction.')

    @property
    def area(self):
        A_circ = np.pi * (np.linalg.norm(self.geometry['v_y']) ** 2 - self.geometry['r_inner'] ** 2)
        return AreaMM2((self.geometry.phi[1] - self.geometry.phi[0]) / (2 * np.pi) * A_circ)


class MultiAperture(BaseAperture, BaseContainer):
    """Several non-overlapping apertures; a photon enters through one of them with
    probability proportional to its area (reference :155-218)."""
    display = {'shape': 'container'}

    def __init__(self, **kwargs):
        self.elements = kwargs.pop('elements')
        self.id_col = kwargs.pop('id_col', 'aperture')
        self.id_num_offset = kwargs.pop('id_num_offset', 0)
        super().__init__(**kwargs)
        for i, elem in enumerate(self.elements):
            elem.id_col = self.id_col
            elem.id_num = self.id_num_offset + i

    @property
    def area(self):
        return AreaMM2(sum(float(e.area) for e in self.elements))

    def _can_lower(self):
        return (not self.preprocess_steps and not self.postprocess_steps
                and all(isinstance(e, FlatAperture) and e._can_lower() for e in self.elements))

    def _lower(self, lw):
        areas = np.array([float(e.area) for e in self.elements])
        cum = np.concatenate([[0.], np.cumsum(areas) / areas.sum()])
        cum[-1] = 1.0
        aper_slot = lw.slot('aperid')
        slots = None
        for i, e in enumerate(self.elements):
            slots = e._lower(lw, aper_slot=aper_slot, aper_id=i, cum=(cum[i], cum[i + 1]), slots=slots)

    def __call__(self, photons):
        self.add_colpos(photons)
        if not self._can_lower():
            raise NotImplementedError('MultiAperture with pre/post steps or custom apertures is not supported')
        return run_fused([self], photons)
