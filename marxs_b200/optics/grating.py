"""Gratings and order selectors (reference marxs/optics/grating.py).

Order selection and the diffraction-direction update are one fused op
(``MXB_OP_GRATING``); selectors are lowered to small device tables."""
import math

import numpy as np

from ..program import (NotFusable, UnsupportedCallable, SEL_ORDERSELECTOR, SEL_EFFFILE,
                       SEL_INTERPTABLE)
from .base import FlatOpticalElement

__all__ = ['OrderSelector', 'EfficiencyFile', 'FlatGrating', 'CATGrating', 'lower_selector']


class OrderSelector:
    """Select from a list of orders independent of energy (reference :12-57).

    The reference calls ``np.random.choice(orderlist, p=p/sum(p))``, i.e. an inverse
    CDF lookup ``orderlist[searchsorted(cdf, u, 'right')]``; the device does the same
    lookup with ``u`` from Philox (or injected)."""

    def __init__(self, orderlist, p=None):
        self.orderlist = np.asarray(orderlist)
        if p is None:
            self.p = np.ones(len(orderlist)) / len(orderlist)
        else:
            p = np.asanyarray(p, dtype=float)
            if len(p) != len(orderlist):
                raise ValueError('Number of elements in orderlist and probabilities does not match.')
            if p.sum() > 1.:
                raise ValueError('Sum of all probabilities must be <= 1.')
            if np.any(p < 0):
                raise ValueError('Probabilities cannot be negative')
            self.p = p

    def device_table(self):
        p = self.p / self.p.sum()
        cdf = p.cumsum()
        cdf /= cdf[-1]
        n = len(cdf)
        return np.concatenate([[SEL_ORDERSELECTOR, n, self.p.sum()], cdf,
                               np.asarray(self.orderlist, dtype=float)]), None


class EfficiencyFile:
    """Order probabilities per energy from a text table (reference :60-96):
    nearest tabulated energy, first order whose cumulative probability exceeds u."""

    def __init__(self, filename, orders):
        dat = np.loadtxt(filename) if isinstance(filename, str) else np.asarray(filename, dtype=float)
        self.energy = dat[:, 0]
        if len(orders) != (dat.shape[1] - 1):
            raise ValueError('orders has len={0}, but data files has {1} order columns.'.format(
                len(orders), dat.shape[1] - 1))
        self.orders = np.array(orders)
        self.prob = dat[:, 1:]
        self.totalprob = np.sum(self.prob, axis=1)
        self.cumprob = np.cumsum(self.prob, axis=1) / self.totalprob[:, None]

    def device_table(self):
        nE, nO = self.cumprob.shape
        return np.concatenate([[SEL_EFFFILE, nE, nO], self.energy, self.totalprob,
                               self.orders.astype(float), self.cumprob.ravel()]), None


def lower_selector(sel, lw):
    """Order selector object -> offset of its device block.

    Recognised: this package's selectors (``device_table``), and duck-typed reference
    objects (``orderlist``+``p``; ``energy``+``cumprob``+``totalprob``+``orders``)."""
    if hasattr(sel, 'device_table'):
        block, big = sel.device_table()
    elif hasattr(sel, 'orderlist') and hasattr(sel, 'p'):
        block, big = OrderSelector(sel.orderlist, sel.p).device_table()
    elif all(hasattr(sel, a) for a in ('energy', 'cumprob', 'totalprob', 'orders')):
        tab = np.hstack([np.asarray(sel.energy)[:, None], np.asarray(sel.prob)])
        block, big = EfficiencyFile(tab, sel.orders).device_table()
    else:
        raise UnsupportedCallable(
            'order_selector {0!r} cannot run on the device: use OrderSelector, EfficiencyFile or '
            'InterpolateEfficiencyTable (arbitrary Python callables have no CPU fallback)'.format(sel))
    if big is None:
        return lw.params(block)
    return lw.params_with_table(block, 4, big)


class FlatGrating(FlatOpticalElement):
    """Flat transmission / reflection grating (reference :99-277)."""

    loc_coos_name = ['grat_y', 'grat_z']
    order_name = 'order'
    blaze_name = 'blaze'
    _cat = False

    def __init__(self, **kwargs):
        self.order_selector = kwargs.pop('order_selector')
        self.transmission = kwargs.pop('transmission', True)
        if 'd' not in kwargs:
            raise ValueError('Input parameter "d" (Grating constant) is required.')
        self._d = kwargs.pop('d')
        groove_angle = kwargs.pop('groove_angle', 0.)
        super().__init__(**kwargs)
        self.geometry._geometry['groove_angle'] = groove_angle

    def e_groove_coos(self, intercoos):
        """(e_groove, e_perp_groove, n), each tiled to (N, 4) (reference :186-207)."""
        groove = self.geometry['groove_angle']
        ex, ey, en = self.geometry.get_local_euklid_bases(intercoos)
        e_groove = math.sin(-groove) * ex + math.cos(-groove) * ey
        e_perp_groove = math.cos(-groove) * ex - math.sin(-groove) * ey
        return e_groove, e_perp_groove, en

    def d(self, intercoos):
        return self._d if not callable(self._d) else self._d(intercoos)

    def _blaze_modifier(self):
        """(blaze_center, d_blaze_mm) or None (reference blaze_angle_modifier :222-231)."""
        return None

    _D_COLUMN = '_mxb_d'      # per-photon grating constant of a callable d (exists only during the launch)

    def _can_lower(self):
        # a callable d(intercoos) needs the intersect first: kernel intersect, d() on the device tensors, then
        # the diffraction kernel with d read from a column (process_photons below)
        return super()._can_lower() and not callable(self._d)

    def process_photons(self, photons, intersect, interpos, intercoos):
        if not callable(self._d) or not FlatOpticalElement._can_lower(self):
            return super().process_photons(photons, intersect, interpos, intercoos)
        import torch
        hit = torch.as_tensor(intersect, device=photons.device)
        loc = torch.as_tensor(intercoos, device=photons.device).as_subclass(torch.Tensor)[hit]
        try:
            d_hit = self._d(loc)
        except (TypeError, RuntimeError, ValueError, AttributeError):
            d_hit = self._d(loc.cpu().numpy())          # numpy-only user function: evaluate it on a host copy
        d_full = torch.full((len(photons),), float('nan'), dtype=torch.float64, device=photons.device)
        if isinstance(d_hit, torch.Tensor):
            d_full[hit] = d_hit.to(device=photons.device, dtype=torch.float64)
        else:
            d_full[hit] = torch.as_tensor(np.asarray(d_hit, dtype=float), device=photons.device)
        photons[self._D_COLUMN] = d_full
        try:
            return self._process_photons_kernel(photons, intersect, interpos, intercoos)
        finally:
            if self._D_COLUMN in photons:
                photons.remove_column(self._D_COLUMN)

    def _lower_specific(self, lw):
        d_callable = callable(self._d)
        if d_callable and self._D_COLUMN not in lw.existing:
            raise UnsupportedCallable('a callable grating constant d(intercoos) cannot be fused: call the grating '
                                      'on a photon table (kernel intersect, d(), diffraction kernel)')
        l, e_perp, n = self.e_groove_coos(np.zeros((1, 2)))
        dd = -e_perp[0]
        mod = self._blaze_modifier()
        flags = (1 if self._cat else 0) | (0 if self.transmission else 2) | (4 if mod is not None else 0)
        pf = lw.eparams(np.concatenate([l[0][:3], dd[:3], [0. if d_callable else self._d],
                                        list(mod) if mod is not None else [0., 0.]]))
        extra = self._lower_l1(lw)         # (flag, second draw slot, block offset) for the L1 support variant
        cols = [lw.fcol(self.order_name), lw.fcol(self.blaze_name)]
        s0 = lw.slot('uniform')
        if extra is not None:
            cols.append(extra[1])
            flags |= 8
        if d_callable:
            cols += [-1] * (3 - len(cols)) + [lw.fcol(self._D_COLUMN)]
            flags |= 16
        lw.op('GRATING', flags=flags, pg=lower_selector(self.order_selector, lw), pf=pf, cols=cols, s0=s0,
              s1=lw.slot('uniform') if extra is not None else -1)
        lw.last_order_col = self.order_name

    def _lower_l1(self, lw):
        return None


class CATGrating(FlatGrating):
    """Critical-angle transmission grating: blazing on the side of negative orders (reference :280-301)."""
    _cat = True
