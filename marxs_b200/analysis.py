"""Analysis helpers that drive the trace path (reference marxs/analysis/analysis.py:9-111,
analysis/gratings.py:26-132) - SURVEY 8(f) rank 4: the heaviest CALLERS of the hot path.

Everything stays on the device: the reductions are torch reductions over photon columns, and every
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
           'detected_fraction', 'resolvingpower_per_order', 'AnalysisError']


class AnalysisError(Exception):
    pass


def _column(photons, colname):
    c = photons[colname]
    return c.as_subclass(torch.Tensor) if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c))


def sigma_clipped_stats(data, sigma=3.0, maxiters=5):
    """(mean, median, std) after iterative clipping at ``sigma`` standard deviations around the median
    (astropy.stats.sigma_clipped_stats defaults: cenfunc='median', stdfunc='std', NaN ignored)."""
    x = torch.as_tensor(data).flatten().to(torch.float64)
    x = x[torch.isfinite(x)]
    for _ in range(maxiters):
        if x.numel() == 0:
            break
        med = torch.median(x) if x.numel() % 2 else 0.5 * (torch.kthvalue(x, x.numel() // 2).values
                                                           + torch.kthvalue(x, x.numel() // 2 + 1).values)
        std = torch.std(x, unbiased=False)
        keep = (x >= med - sigma * std) & (x <= med + sigma * std)
        if bool(keep.all()):
            break
        x = x[keep]
    if x.numel() == 0:
        nan = float('nan')
        return nan, nan, nan
    n = x.numel()
    med = torch.median(x) if n % 2 else 0.5 * (torch.kthvalue(x, n // 2).values + torch.kthvalue(x, n // 2 + 1).values)
    return float(x.mean()), float(med), float(torch.std(x, unbiased=False))


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
