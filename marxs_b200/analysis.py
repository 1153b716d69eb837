"""Analysis helpers that drive the trace path (reference marxs/analysis/analysis.py:9-111,
analysis/gratings.py:26-132) - SURVEY 8(f) rank 4: the heaviest CALLERS of the hot path.

Everything stays on the device: sigma clipping is a kernel (csrc/mxb_stats.cu), the other reductions are torch
reductions over photon columns, and every
trial detector of ``find_best_detector_position`` is one out-of-place launch of the trace kernel on
the resident photon list (``simulator.trace_from`` into a reused work table; the kernel specialised
for "one flat detector" is compiled once because detector positions are parameters, not structure).
"""
import numpy as np
import torch

from . import optics
from .geometry import Cylinder
from .simulator import trace_from

__all__ = ['sigma_clipped_stats', 'sigma_clipped_std', 'mean_width_2d', 'find_best_detector_position',
           'detected_fraction', 'resolvingpower_per_order', 'AnalysisError',
           'resolvingpower_from_photonlist', 'resolvingpower_from_photonlist_robust', 'effectivearea_from_photonlist',
           'average_R_Aeff', 'weighted_per_order', 'identify_photon_in_subaperture',
           'CaptureResAeff', 'CaptureResAeff_CCDgaps']


class AnalysisError(Exception):
    pass


def _column(photons, colname):
    c = photons[colname]
    return c.as_subclass(torch.Tensor) if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c))


def sigma_clipped_stats(data, sigma=3.0, maxiters=5):
    """(mean, median, std) after iterative clipping at ``sigma`` standard deviations around the median
    (astropy.stats.sigma_clipped_stats defaults: cenfunc='median', stdfunc='std', non-finite values masked).

    One call of ``mxb_sigma_clip_stats`` (csrc/mxb_stats.cu): every clipping round is a few streaming passes over the
    resident column (count / sum, squared deviations, an exact radix-select median) - no sort, no compaction, no host
    round trip between the rounds.  Host data is copied to the device first; there is no CPU path."""
    from . import _lib
    x = torch.as_tensor(data)
    if x.device.type != 'cuda':
        x = x.to('cuda')
    x = x.flatten().to(torch.float64).contiguous()
    lib = _lib.load()
    work = torch.empty(int(lib.mxb_sigma_clip_workspace()), dtype=torch.uint8, device=x.device)
    out = torch.empty(4, dtype=torch.float64, device=x.device)
    iters = 100 if maxiters is None else int(maxiters)       # astropy: None = until nothing is clipped
    with torch.cuda.device(x.device):
        rc = lib.mxb_sigma_clip_stats(x.data_ptr(), x.numel(), float(sigma), iters, out.data_ptr(), work.data_ptr(),
                                      torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(lib, rc, 'mxb_sigma_clip_stats')
    mean, med, std, _ = out.tolist()
    return mean, med, std


def sigma_clipped_std(photons, colname='det_x', **kwargs):
    """Standard deviation of the sigma-clipped column (reference analysis.py:9-25)."""
    return sigma_clipped_stats(_column(photons, colname), **kwargs)[2]


def mean_width_2d(photons):
    """Average distance from the centre of the det_x, det_y distribution (reference :28-40)."""
    x, y = _column(photons, 'det_x'), _column(photons, 'det_y')
    r = torch.sqrt((x - x.mean()) ** 2 + (y - y.mean()) ** 2)
    return float(r.sum() / r.numel())


def find_best_detector_position(photons, objective_func=sigma_clipped_std, objective_func_args={'colname': 'det_x'},
                                orientation=np.eye(3), **kwargs):
    """Numerically find the position of best focus (reference :43-84): a flat detector is moved along its
    normal and the width of the photon distribution minimised with ``scipy.optimize.minimize_scalar``.
    Each trial is one launch on the resident photons; as in the reference, ``photons`` itself is not
    modified (the reference intersects a copy)."""
    import scipy.optimize
    work = [None]

    def width(x):
        mdet = optics.FlatDetector(position=np.dot(orientation, np.array([x, 0, 0])), orientation=orientation,
                                   zoom=1e5, pixsize=1.)
        work[0] = trace_from(mdet, photons, out=work[0])
        return objective_func(work[0], **objective_func_args)

    return scipy.optimize.minimize_scalar(width, **kwargs)


def detected_fraction(photons, labels, col='order'):
    """Fraction of the photons detected per integer label, e.g. effective area per order (reference :87-111)."""
    labels = np.asarray(labels)
    c, p = _column(photons, col), _column(photons, 'probability')
    prob = np.zeros(labels.shape, dtype=float)
    for i, o in enumerate(labels.ravel()):
        prob.ravel()[i] = float(p[c == float(o)].sum()) / len(photons)
    return prob


def resolvingpower_per_order(gratings, photons, orders, detector=None, colname='det_x'):
    """Resolving power per grating order (reference analysis/gratings.py:26-132): all photons are sent into
    one order at a time through ``gratings`` (its order selector is REPLACED, like in the reference) and
    projected onto ``detector`` - an element instance, or None for a flat detector whose x position is
    optimised per order.  Returns (res, fwhm, info)."""
    orders = np.asarray(orders)
    res = np.zeros(orders.shape, dtype=float)
    fwhm = np.zeros(orders.shape, dtype=float)
    info = {}
    if detector is None:
        info['method'] = 'Detector position numerically optimized'
        info['fit_results'] = []
        col, zeropos, det = 'det_x', 0., None
    else:
        if isinstance(detector, Cylinder):
            detector = optics.CircularDetector(geometry=detector)
            colname = 'detpix_x'
        det, col = detector, colname
        info['method'] = 'User defined detector'
        pg = det(photons.copy())
        pg = pg[_column(pg, 'probability') > 0.]
        zeropos = sigma_clipped_stats(_column(pg, col))[0]
    for i, order in enumerate(orders):
        sel = optics.OrderSelector([order])
        gratings.elem_args['order_selector'] = sel
        for elem in gratings.elements:
            elem.order_selector = sel
        pg = gratings(photons.copy())
        if 'order' not in pg.colnames:
            raise AnalysisError('no photon reaches a grating')
        pg = pg[(_column(pg, 'order') == float(order)) & (_column(pg, 'probability') > 0.)]
        if detector is None:
            xbest = find_best_detector_position(pg, objective_func=sigma_clipped_std)
            info['fit_results'].append(xbest)
            det = optics.FlatDetector(position=np.array([xbest.x, 0, 0]), zoom=1e5)
        pg = det(pg)
        meanpos, medianpos, stdpos = sigma_clipped_stats(_column(pg, col))
        fwhm[i] = 2.3548 * stdpos
        res[i] = np.abs((meanpos - zeropos) / fwhm[i])
    return res, fwhm, info


# ---------------------------------------------------------------------------------------------
# figures of merit read off a traced photon list (reference analysis/gratings.py:133-549): the
# ``analyzefunc`` of the tolerancing loops.  Selections are boolean device tensors; nothing but
# the per-order scalars leaves the GPU.
# ---------------------------------------------------------------------------------------------
def resolvingpower_from_photonlist(photons, orders, col='proj_x', zeropos=None, ordercol='order', ind=None):
    """Resolving power, mean position and width of column ``col`` for every order in ``orders`` (reference
    analysis/gratings.py:211-269).  Orders with 20 photons or fewer give NaN; ``zeropos=None`` measures the
    position of order 0 (AnalysisError below 20 photons).  ``ind`` (bool mask) restricts the list without
    materialising a sub-table."""
    o, x = _column(photons, ordercol), _column(photons, col)
    if ind is not None:
        o, x = o[ind], x[ind]
    if zeropos is None:
        zero = o == 0.
        if int(zero.sum()) < 20:
            raise AnalysisError('Too few photons in list to determine position of order 0 automatically.')
        zeropos = sigma_clipped_stats(x[zero])[0]
    orders = np.asarray(orders)
    pos = np.zeros(orders.shape, dtype=float)
    std = np.zeros(orders.shape, dtype=float)
    for i, order in enumerate(orders):
        sel = o == float(order)
        if int(sel.sum()) > 20:
            pos[i], _, std[i] = sigma_clipped_stats(x[sel])
        else:
            pos[i], std[i] = np.nan, np.nan
    with np.errstate(divide='ignore', invalid='ignore'):
        res = np.abs(pos - zeropos) / (std * 2.3548)
    return res, pos, std


def resolvingpower_from_photonlist_robust(lphotons, orders, cols, zeropositions, ordercol='order'):
    """Worst case over several (photon list, column, zero position) triples, per order (reference :272-330)."""
    if len(cols) != len(zeropositions):
        raise ValueError('Number of elements in cols and zeropositions is not the same.')
    if len(cols) != len(lphotons):
        raise ValueError('Number of elements in cols and photon lists is not the same.')
    allres = [resolvingpower_from_photonlist(p, orders, col=c, zeropos=z, ordercol=ordercol)
              for p, c, z in zip(lphotons, cols, zeropositions)]
    res, pos, std = (np.array([a[k] for a in allres]) for k in range(3))
    worst = np.argmin(res, axis=0)
    k = np.arange(res.shape[1])
    return res[worst, k], pos[worst, k], std[worst, k]


def effectivearea_from_photonlist(photons, orders, n_photons, A_geom=1., ordercol='order', ind=None):
    """Effective area per order: summed probability of the photons of that order / n_photons x A_geom
    (reference :333-362).  One pass: orders are binned with a scatter-add on the device."""
    o, p = _column(photons, ordercol), _column(photons, 'probability')
    if ind is not None:
        o, p = o[ind], p[ind]
    orders = np.asarray(orders)
    aeff = np.zeros(len(orders))
    for i, order in enumerate(orders):
        aeff[i] = float(p[o == float(order)].sum())
    return aeff / n_photons * A_geom


def identify_photon_in_subaperture(angle, max_ang, ang_0=np.pi / 2):
    """Photons in two mirrored sectors of half-width ``max_ang`` around ``ang_0`` and ``-ang_0`` (reference :365-386)."""
    a = torch.as_tensor(angle)
    a = torch.remainder(a, 2 * np.pi)
    tol = max_ang + 1e-5 * abs(ang_0)               # np.isclose(a, b, atol): |a - b| <= atol + rtol |b|, rtol = 1e-5
    tol2 = max_ang + 1e-5 * abs(2 * np.pi - ang_0)
    return ((a - ang_0).abs() <= tol) | ((a - (2 * np.pi - ang_0)).abs() <= tol2)


def average_R_Aeff(r, aeff, axis=None):
    """Sum of Aeff and the Aeff-weighted mean of R, ignoring non-finite R (reference :181-208).
    Returns (res_avg, aeff_sum)."""
    r, aeff = np.asarray(r, dtype=float), np.asarray(aeff, dtype=float)
    return np.ma.average(np.ma.masked_invalid(r), weights=aeff, axis=axis), aeff.sum(axis=axis)


def weighted_per_order(data, orders, energy, gratingeff):
    """Mean over orders of ``data`` (orders x energies) weighted with the tabulated probability of each
    order at each energy (reference :133-178)."""
    data = np.asarray(data)
    if len(orders) != data.shape[0]:
        raise ValueError('First dimension of "data" must match length of "orders".')
    if len(energy) != data.shape[1]:
        raise ValueError('Second dimension of "data" must match length of "energy".')
    weights = np.zeros_like(data, dtype=float)
    en_sort = np.argsort(gratingeff.energy)
    for i, o in enumerate(orders):
        k = np.nonzero(np.asarray(gratingeff.orders) == o)[0]
        if len(k) != 1:
            raise KeyError('No data for order {0} in gratingeff'.format(o))
        weights[i] = np.interp(energy, gratingeff.energy[en_sort], gratingeff.prob[:, k[0]][en_sort])
    return np.ma.average(data, axis=0, weights=weights)


class CaptureResAeff:
    """Resolving power and effective area per order of one traced photon list, plus the totals of the
    dispersed orders and of order 0 (reference :389-500); the usual ``analyzefunc`` of
    ``marxs_b200.design.tolerancing.run_tolerances``.  Subclass and override ``aeff_filter`` /
    ``res_filter`` (bool device masks) to change which photons count."""

    def __init__(self, A_geom=1, order_col='order', orders=np.arange(-10, 11), dispersion_coord='det_x', zeropos=None):
        self.A_geom = A_geom
        self.order_col = order_col
        self.orders = np.asanyarray(orders)
        self.dispersion_coord = dispersion_coord
        self.zeropos = zeropos

    def aeff_filter(self, photons):
        return None                                             # all photons

    def res_filter(self, photons):
        return torch.isfinite(_column(photons, self.dispersion_coord)) & (_column(photons, 'probability') > 0)

    def __call__(self, photons, n_photons=None):
        if n_photons is None:
            n_photons = len(photons)
        aeff = effectivearea_from_photonlist(photons, self.orders, n_photons, self.A_geom, self.order_col,
                                             ind=self.aeff_filter(photons))
        try:
            if self.dispersion_coord not in photons.colnames or self.order_col not in photons.colnames:
                raise AnalysisError('no photon reached the detector')
            res = resolvingpower_from_photonlist(photons, self.orders, col=self.dispersion_coord, zeropos=self.zeropos,
                                                 ordercol=self.order_col, ind=self.res_filter(photons))[0]
        except AnalysisError:
            res = np.nan * np.ones(len(self.orders))
        disp = self.orders != 0
        avggratres, aeffgrat = average_R_Aeff(res[disp], aeff[disp])
        return {'Aeff0': np.sum(aeff[~disp]), 'Aeffgrat': aeffgrat, 'Aeff': aeff, 'Rgrat': avggratres, 'R': res}


class CaptureResAeff_CCDgaps(CaptureResAeff):
    """As CaptureResAeff, but only photons with ``photons[aeff_filter_col] >= 0`` (e.g. a CCD_ID) count
    towards the effective area, while R ignores chip gaps (reference :503-549)."""

    def __init__(self, aeff_filter_col='CCD_ID', **kwargs):
        super().__init__(**kwargs)
        self.aeff_filter_col = aeff_filter_col

    def aeff_filter(self, photons):
        return _column(photons, self.aeff_filter_col) >= 0
