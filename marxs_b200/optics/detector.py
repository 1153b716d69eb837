"""Pixelated detectors (reference marxs/optics/detector.py)."""
import warnings

import numpy as np

from ..affines import decompose44
from .base import FlatOpticalElement

__all__ = ['FlatDetector']


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
        lw.op('DETPIX', pf=lw.eparams([self.pixsize, self.centerpix[0], self.centerpix[1]]),
              cols=[lw.fcol(self.detpix_name[0]), lw.fcol(self.detpix_name[1])])
