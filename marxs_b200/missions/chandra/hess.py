"""Chandra HETG: 336 flat grating facets on the HESS structure
(reference marxs/missions/chandra/hess.py:11-85)."""
import numpy as np

from ...affines import compose
from ...optics import FlatGrating, OrderSelector
from ...simulator import Parallel
from .data import load_hess

__all__ = ['HETG']


class HETG(Parallel):
    """High-Energy Transmission Grating: 192 MEG + 144 HEG facets, orders -3..3 with
    equal probability (the reference's placeholder efficiency)."""

    _fingerprint_skip = Parallel._fingerprint_skip + ('hess',)     # the design table the facets were built from

    id_col = 'facet'

    def __init__(self, **kwargs):
        self.hess = load_hess()
        kwargs['elem_pos'] = self.calculate_elempos()
        kwargs['elem_class'] = FlatGrating
        # gratings are listed MEG first, then HEG
        d = [4001.95 * 1e-7] * 192 + [2000.81 * 1e-7] * 144
        ul = np.vstack([self.hess[s + 'ul'] for s in 'xyz'])
        uyf = np.vstack([self.hess[s + 'uyf'] for s in 'xyz'])
        uxf = np.vstack([self.hess[s + 'uxf'] for s in 'xyz'])
        groove_ang = -np.arctan2(np.einsum('i...,i...->...', ul, uxf),
                                 np.einsum('i...,i...->...', ul, uyf))
        kwargs['elem_args'] = {'order_selector': OrderSelector(np.arange(-3, 4)),
                               'd': d, 'name': list(self.hess['hessloc']),
                               'groove_angle': groove_ang.tolist()}
        super().__init__(**kwargs)
        # centre of the HETG on its (approximate) hinge point
        move = np.eye(4)
        move[:, 3] = [8618., -650 * np.sin(np.deg2rad(60)), 650 * np.cos(np.deg2rad(60)), 1.]
        self.move_center(move)

    def calculate_elempos(self):
        hess = self.hess
        cen = np.vstack([hess[s + 'c'] for s in 'xyz']).astype(float)
        cen[0, :] += (339.909 - 1.7) * 25.4
        facnorm = np.vstack([hess[s + 'u'] for s in 'xyz'])
        uxf = np.vstack([hess[s + 'uxf'] for s in 'xyz'])
        uyf = np.vstack([hess[s + 'uyf'] for s in 'xyz'])
        pos4ds = []
        for i in range(facnorm.shape[1]):
            if int(str(hess['hessloc'][i])[0]) <= 3:
                zoom = np.array([1, 1.035 * 25.4 / 2, 1.035 * 25.4 / 2])     # MEG facet size
            else:
                zoom = np.array([1, 0.920 * 25.4 / 2, 0.920 * 25.4 / 2])     # HEG facet size
            rot = np.eye(3)
            rot[:, 0] = facnorm[:, i]
            rot[:, 1] = uxf[:, i]
            rot[:, 2] = uyf[:, i]
            pos4ds.append(compose(cen[:, i], rot, zoom))
        return pos4ds
