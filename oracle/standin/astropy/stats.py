"""Stand-in for the one astropy.stats routine the reference's analysis helpers call.

``sigma_clipped_stats`` restates the published algorithm of astropy.stats.SigmaClip with the
defaults the reference relies on (sigma=3, maxiters=5, cenfunc='median', stdfunc='std', NaNs
ignored): repeat { centre = median, spread = population std, drop values outside
centre +- sigma * spread } until nothing is dropped or maxiters is reached; then return
(mean, median, std) of what is left.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def sigma_clipped_stats(data, mask=None, mask_value=None, sigma=3.0, sigma_lower=None, sigma_upper=None,
                        maxiters=5, cenfunc='median', stdfunc='std', std_ddof=0, axis=None, grow=False):
    x = np.asarray(getattr(data, 'value', data), dtype=float).ravel()
    if mask is not None:
        x = x[~np.asarray(mask).ravel()]
    x = x[np.isfinite(x)]
    lo = sigma if sigma_lower is None else sigma_lower
    hi = sigma if sigma_upper is None else sigma_upper
    for _ in range(maxiters if maxiters is not None else 1000000):
        if x.size == 0:
            break
        c, s = np.median(x), np.std(x, ddof=0)
        keep = (x >= c - lo * s) & (x <= c + hi * s)
        if keep.all():
            break
        x = x[keep]
    if x.size == 0:
        return np.nan, np.nan, np.nan
    return np.mean(x), np.median(x), np.std(x, ddof=std_ddof)
