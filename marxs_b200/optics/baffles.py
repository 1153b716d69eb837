"""Baffles: plates with a hole (reference marxs/optics/baffles.py)."""
from .. import geometry
from .base import FlatOpticalElement

__all__ = ['Baffle', 'CircularBaffle']


class Baffle(FlatOpticalElement):
    """Photons that miss the rectangular opening get probability 0; the others are
    moved to the plane (reference :8-28)."""

    default_geometry = geometry.RectangleHole
    display = {'color': (1.0, 0.5, 0.4)}

    def _lower_specific(self, lw):
        lw.op('BAFFLE')

    def _lower(self, lw):
        if lw.array is not None:
            from ..program import NotFusable
            raise NotFusable('arrays of baffles are not fused')
        lw.plane(self.pos4d, circular=getattr(self.geometry, 'circular', False))
        lw.op('BAFFLE')


class CircularBaffle(Baffle):
    """Circular opening; note the reference compares the ABSOLUTE radius with 1.0
    (math/geometry.py:376-380, pinned by optics/tests/test_baffle.py)."""

    default_geometry = geometry.CircularHole
