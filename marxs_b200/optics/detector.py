"""Pixelated detectors (reference marxs/optics/detector.py)."""
import warnings

import numpy as np

from ..affines import decompose44
from ..geometry import Cylinder
from .base import FlatOpticalElement, OpticalElement

__all__ = ['FlatDetector', 'CircularDetector']


class SimulationSetupWarning(Warning):
    pass


class FlatDetector(FlatOpticalElement):
    """Flat detector with square pixels: adds det_x/det_y [mm] and the fractional
    pixel coordinates detpix_x/detpix_y, (0, 0) at the centre of the corner pixel
    (reference :18-75)."""

    loc_coos_name = ['det_x', 'det_y']
    detpix_name = ['detpix_x', 'detpix_y']
    display = {'color': (1.0, 1.0, 0.), 'shape': 'box', 'box-half': '+x'}

    def __init__(self, pixsize=1, ignore_pixel_warning=False, **kwargs):
        self.pixsize = pixsize
        super().__init__(**kwargs)
        t, r, zoom, s = decompose44(self.pos4d)
        self.npix = [0, 0]
        self.centerpix = [0, 0]
        for i in (0, 1):
            z = zoom[i + 1]
            self.npix[i] = int(np.round(2. * z / self.pixsize))
            if (np.abs(2. * z / self.pixsize - self.npix[i]) > 1e-2) and not ignore_pixel_warning:
                warnings.warn('Detector size is not an integer multiple of pixel size in direction {0}. '
                              'It will be rounded.'.format('xy'[i]), SimulationSetupWarning)
            self.centerpix[i] = (self.npix[i] - 1) / 2

    def _lower_specific(self, lw):
        image = lw.image if lw.image is not None else (
            (self.image, self.id_num) if getattr(self, 'image', None) is not None else None)
        pg, s0 = -1, -1
        if image is not None:
            # detector image fused into the trace: image[id_num - sel_lo, round(detpix_y), round(detpix_x)] += probability
            t, sel_lo = image
            pg = lw.params([t.shape[2], t.shape[1], sel_lo, t.shape[0]])
            s0 = lw.aux_ptr(t)
        lw.op('DETPIX', flags=1 if lw.array is not None else 0, pg=pg,
              pf=lw.eparams([self.pixsize, self.centerpix[0], self.centerpix[1]]),
              cols=[lw.fcol(self.detpix_name[0]), lw.fcol(self.detpix_name[1])], s0=s0,
              w14=self.id_num if lw.array is None else 0)


class CircularDetector(OpticalElement):
    """Detector shaped like a ring or tube following the Rowland circle (reference :78-118): adds
    det_phi [rad], det_y [mm] and detpix_x = phi * R / pixsize, detpix_y = y / pixsize."""

    loc_coos_name = ['det_phi', 'det_y']
    detpix_name = ['detpix_x', 'detpix_y']
    display = {'color': (1.0, 1.0, 0.), 'opacity': 0.7}
    centerpix = [0, 0]
    default_geometry = Cylinder
    _lowerable_geometries = (Cylinder,)

    def __init__(self, pixsize=1, **kwargs):
        self.pixsize = pixsize
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        lw.op('DETPIX', flags=2, pf=lw.eparams([self.pixsize, self.centerpix[0], self.centerpix[1], self.geometry['R']]),
              cols=[lw.fcol(self.detpix_name[0]), lw.fcol(self.detpix_name[1])], w14=self.id_num)
