"""ACIS detector: 10 chips as a Parallel of flat detectors with Chandra chip / TDET /
DET / sky coordinates (reference marxs/missions/chandra/det_acis.py, data.py:169-190)."""
import numpy as np

from ...optics import FlatDetector
from ...simulator import Parallel
from .data import (NOMINAL_FOCALLENGTH, AIMPOINTS, TDET, ODET, PIXSIZE, ACIS_name, load_acis_corners)

__all__ = ['ACISChip', 'ACIS', 'ACIS_name']


class ACISChip(FlatDetector):
    """One ACIS chip; adds chipx/y, tdetx/y, detx/y and sky x/y (reference :19-58)."""

    def __init__(self, **kwargs):
        self.TDET = TDET['ACIS']
        self.ODET = ODET['ACIS']
        self.pixsize_in_rad = np.deg2rad(PIXSIZE['ACIS'])
        kwargs['ignore_pixel_warning'] = True
        super().__init__(**kwargs)

    @property
    def chip_name(self):
        return 'ACIS-{0}'.format(ACIS_name[self.id_num])

    def _lower_specific(self, lw):
        t = self.TDET
        theta = t['theta'][self.id_num]
        sh = t['scale'][self.id_num] * t['handedness'][self.id_num]
        ox, oy = t['origin'][self.id_num]
        # the pointing may be part of the same program (source.observe): its meta is pending in meta_updates
        roll = np.deg2rad(lw.meta_updates['ROLL_PNT'][0] if 'ROLL_PNT' in lw.meta_updates else lw.meta['ROLL_PNT'][0])
        glob = [NOMINAL_FOCALLENGTH, self.pixsize_in_rad, self.ODET[0], self.ODET[1], np.cos(roll), np.sin(roll)]
        s0 = -1
        if lw.image is not None:
            # fused detector image: image[CCD_ID - sel_lo, round(chipy) - 1, round(chipx) - 1] += probability
            img, sel_lo = lw.image
            glob += [img.shape[2], img.shape[1], sel_lo, img.shape[0]]
            s0 = lw.aux_ptr(img)
        pg = lw.params(glob)
        pf = lw.eparams([self.pixsize, self.centerpix[0], self.centerpix[1], sh, np.cos(theta),
                         np.sin(theta), ox + 0.5, oy + 0.5])
        lw.op('ACIS', pg=pg, pf=pf, s0=s0,
              cols=[lw.fcol(n) for n in ('chipx', 'chipy', 'tdetx', 'tdety', 'detx', 'dety', 'x', 'y')])
        lw.meta_updates['ACSYS1'] = ('CHIP:AXAF-ACIS-1.0', 'reference for chip coord system')
        lw.meta_updates['ACSYS2'] = ('TDET:{0}'.format(t['version']), 'reference for tiled detector coord system')
        lw.meta_updates['ACSYS3'] = ('DET:ASC-FP-1.1', 'reference for focal plane coord system')
        lw.meta_updates['ACSYS4'] = ('SKY:ASC-FP-1.1', 'reference for sky coord system')


class ACIS(Parallel):
    """The ACIS instrument (ideal detection only).  ``chips``: indices into I0..S5,
    ``aimpoint``: one of ``AIMPOINTS`` (reference :61-150)."""

    OLSI = np.array([0.684, 0.750, 236.552])
    id_col = 'CCD_ID'

    def __init__(self, chips, **kwargs):
        self.aimpoint = kwargs.pop('aimpoint')
        self.detoffset = np.array([kwargs.pop('DetOffsetX', 0), 0, kwargs.pop('DetOffsetY', 0)])
        kwargs['elem_pos'] = self.calculate_elempos()
        kwargs['elem_class'] = ACISChip
        kwargs['elem_args'] = {'pixsize': 0.024}
        super().__init__(**kwargs)
        self.chips = chips
        self.elements = [self.elements[i] for i in chips]

    def generate_elements(self):
        super().generate_elements()
        if hasattr(self, 'chips'):
            self.elements = [self.elements[i] for i in self.chips]

    def calculate_elempos(self):
        corners = load_acis_corners()
        pos4d = []
        for i in range(len(ACIS_name)):
            p0 = corners[i, 0]
            e_y = corners[i, 1] - p0
            e_z = corners[i, 3] - p0
            e_x = np.cross(e_y, e_z)
            center = p0 + 0.5 * e_y + 0.5 * e_z
            A = np.eye(4)
            A[:3, :3] = np.vstack([e_x / np.sqrt((e_x ** 2).sum()), e_y / 2, e_z / 2]).T
            A[:3, 3] = center
            A[:3, 3] += self.OLSI
            A[:3, 3] += self.aimpoint
            A[:3, 3] -= self.detoffset
            pos4d.append(A)
        return pos4d
