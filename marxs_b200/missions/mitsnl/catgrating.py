"""CAT-grating efficiency table and non-parallel bars
(reference marxs/missions/mitsnl/catgrating.py:63-144, 280-311)."""
import os

import numpy as np

from ...optics import CATGrating, OrderSelector, FlatStack, FlatOpticalElement, Tabulated1D
from ...optics.scatter import RandomGaussianScatter
from ...program import SEL_INTERPTABLE, NotFusable

__all__ = ['l1transtab', 'l1_order_selector', 'l1_dims', 'l2_dims', 'qualityfactor', 'd',
           'InterpolateEfficiencyTable', 'QualityFactor', 'L1', 'L2Abs', 'L2Diffraction', 'CATL1L2Stack',
           'NonParallelCATGrating', 'catsupportbars']

d = 0.0002
'''Spacing of grating bars [mm]'''


def _mm(x):
    """Length in mm from a float (mm) or an astropy-like Quantity."""
    if hasattr(x, 'to'):
        import astropy.units as u
        return float(x.to(u.mm).value)
    return float(x)


def _load_si_transmission():
    dat = np.loadtxt(os.path.join(os.path.dirname(__file__), 'data', 'si_transmission_1um.csv'),
                     delimiter=',', skiprows=2)
    return {'energy': dat[:, 0] * 0.001, 'transmission': dat[:, 1]}      # eV -> keV


l1transtab = _load_si_transmission()
'''Transmission through 1 mu of Si: ``energy`` [keV], ``transmission``'''

l1_order_selector = OrderSelector(orderlist=np.array([-4, -3, -2, -1, 0, 1, 2, 3, 4]),
                                  p=np.array([0.006, 0.0135, 0.022, 0.028, 0.861, 0.028, 0.022, 0.0135, 0.006]))
'''Simple order selector for diffraction on L1 (reference :44-45).'''

l1_dims = {'bardepth': 0.004, 'period': 0.005, 'barwidth': 0.0009}
'''Dimensions of L1 support bars [mm] (floats or astropy Quantities are accepted)'''

l2_dims = {'bardepth': 0.5, 'period': 0.9622504, 'barwidth': 0.0962250}
'''Dimensions of hexagonal L2 support [mm]'''

qualityfactor = {'d': 200., 'sigma': 1.75}
'''Debye-Waller parameterization of the grating quality factor [um]'''


def check_lx_dims(lx_dims):
    if not (_mm(lx_dims['barwidth']) < _mm(lx_dims['period'])):
        raise ValueError('Period of grating must be larger than bar width.')


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
        # global table: the efficiencies, then their running sums over the order axis (the fast build
        # bisects the interpolated cumulative curve instead of summing all orders: mxb_ops.cuh select_order)
        return block, np.concatenate([self.prob.ravel(), np.cumsum(self.prob, axis=2).ravel()])


class NonParallelCATGrating(CATGrating):
    """CAT grating whose bar side-walls tilt linearly across the facet:
    blaze += blaze_center + d_blaze_mm * local_y (reference :280-311)."""

    def __init__(self, **kwargs):
        self.d_blaze_mm = kwargs.pop('d_blaze_mm', 0)
        self.blaze_center = kwargs.pop('blaze_center', 0)
        super().__init__(**kwargs)

    def _blaze_modifier(self):
        return (self.blaze_center, self.d_blaze_mm)


def log_double_double(x):
    """ln(x) as an unevaluated sum of two doubles (40-digit arithmetic): the fast kernel build forms
    ``x ** y = exp(y * ln x)`` from it without losing the digits a double logarithm would (csrc/mxb_ops.cuh
    pow_loghost).  NaN, NaN when x is not a positive finite number (the kernel then calls libm's pow)."""
    import decimal
    x = float(x)
    if not (x > 0.0 and np.isfinite(x)):
        return float('nan'), float('nan')
    with decimal.localcontext() as ctx:
        ctx.prec = 40
        ln = decimal.Decimal(x).ln()
        hi = float(ln)
        lo = float(ln - decimal.Decimal(hi))
    return hi, lo


class QualityFactor(FlatOpticalElement):
    """Scale probabilities of theoretical curves to measured values:
    ``probability *= factor ** order**2`` (reference :147-161)."""

    def __init__(self, qualityfactor=qualityfactor, **kwargs):
        sigma, dd = qualityfactor['sigma'], qualityfactor['d']
        if hasattr(sigma, 'to'):
            import astropy.units as u
            sigma, dd = sigma.to(u.um).value, dd.to(u.um).value
        self.factor = np.exp(- (2 * np.pi * sigma / dd) ** 2)
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        if lw.last_order_col != 'order':
            raise NotFusable('QualityFactor needs the order of a grating in the same stack')
        lw.op('QFACTOR', pf=lw.eparams([self.factor, *log_double_double(self.factor)]))


class L1(CATGrating):
    """The L1 support structure as a second (cross-dispersing) CAT grating; photons that hit a
    support bar instead go through 4 um of solid Si (reference :170-219)."""

    blaze_name = 'blaze_L1'
    order_name = 'order_L1'

    def __init__(self, l1_dims=l1_dims, transtab=None, **kwargs):
        check_lx_dims(l1_dims)
        self.openfraction = 1 - _mm(l1_dims['barwidth']) / _mm(l1_dims['period'])
        tab = l1transtab if transtab is None else transtab
        with np.errstate(divide='ignore'):
            logtranstab = np.log(np.asarray(tab['transmission'], dtype=float))
        trans = np.exp(logtranstab * (_mm(l1_dims['bardepth']) * 1e3))     # bardepth / (1 um)
        self.transfunc = Tabulated1D(np.asarray(tab['energy'], dtype=float), trans, bounds_error=True)
        kwargs['d'] = _mm(l1_dims['period'])
        super().__init__(**kwargs)

    def _lower_l1(self, lw):
        x, y = self.transfunc.x, self.transfunc.y
        block = lw.params_with_table([self.openfraction, -1.], 1, np.concatenate([[len(x)], x, y]))
        return (8, block)


class L2Abs(FlatOpticalElement):
    """L2 absorption and shadowing by the hexagonal support mesh (reference :222-259)."""

    def __init__(self, l2_dims=l2_dims, **kwargs):
        check_lx_dims(l2_dims)
        self.bardepth = _mm(l2_dims['bardepth'])
        self.period = _mm(l2_dims['period'])
        self.barwidth = _mm(l2_dims['barwidth'])
        super().__init__(**kwargs)
        self.innerfree = self.period - self.barwidth

    def _lower_specific(self, lw):
        openfraction = (self.innerfree / self.period) ** 2
        totalarea = self.period ** 2 / 2 * np.sqrt(3)
        lw.op('L2ABS', pf=lw.eparams([openfraction, self.bardepth * self.innerfree, totalarea]))


class L2Diffraction(RandomGaussianScatter):
    """Broadening by the single-slit function of the L2 mesh, approximated as Gaussian scatter with
    the Airy-disk radius as sigma (reference :262-285)."""

    scattername = 'L2Diffraction'
    scatter = 'airy'        # energy dependent: evaluated per photon inside the kernel

    def __init__(self, l2_dims=l2_dims, **kwargs):
        check_lx_dims(l2_dims)
        self.innerfree = _mm(l2_dims['period']) - _mm(l2_dims['barwidth'])
        super().__init__(**kwargs)

    def _lower_specific(self, lw):
        lw.op('GSCATTER', flags=1, pf=lw.eparams([self.innerfree]), cols=[lw.fcol(self.scattername)],
              s0=lw.slot('normal'), s1=lw.slot('uniform'))


def catsupportbars(photons):
    """The metal frame absorbs every photon that does not pass through a facet (reference :314-323)."""
    if 'facet' in photons.colnames:
        photons['probability'][photons['facet'] < 0] = 0.
    else:
        photons['probability'] = 0.
    return photons


class CATL1L2Stack(FlatStack):
    """SNL fabricated CAT grating: membrane + quality factor + L1 + L2 absorption + L2 diffraction on
    ONE intersect (reference :326-374); the whole stack is five ops of the fused kernel."""

    def __init__(self, l1_dims=l1_dims, l2_dims=l2_dims, l1_order_selector=l1_order_selector,
                 qualityfactor=qualityfactor, **kwargs):
        kwargs['elements'] = [NonParallelCATGrating, QualityFactor, L1, L2Abs, L2Diffraction]
        groove_angle = kwargs.pop('groove_angle', 0.)
        kwargs['keywords'] = [{'order_selector': kwargs.pop('order_selector'),
                               'd': kwargs.pop('d', d),
                               'groove_angle': groove_angle},
                              {'qualityfactor': qualityfactor},
                              {'l1_dims': l1_dims,
                               'order_selector': l1_order_selector,
                               'groove_angle': np.pi / 2. + groove_angle},
                              {'l2_dims': l2_dims},
                              {'l2_dims': l2_dims}]
        super().__init__(**kwargs)
