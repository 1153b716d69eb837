"""Sources (reference source/basesources.py, source/labSource.py) and pointing (source/pointing.py)
as ops of the fused trace kernel (GENERATE, LABCONE, FARLAB, POINTING in include/mxb.h).

Supported specifications (the rest raises, there is no CPU fallback): ``flux`` a number or
Quantity (constant flux -> evenly spaced times, basesources.py:171-174); ``energy`` a number /
Quantity [keV] or a table with ``energy`` (upper bin edges) and ``fluxdensity`` (basesources.py:191-197);
``polarization`` None (unpolarised: uniform angle), an angle, or a table with ``angle`` and
``probabilitydensity``.  Sky coordinates: (ra, dec) in degrees or an astropy SkyCoord."""
from collections import OrderedDict

import numpy as np
import torch

from ..base import SimulationSequenceElement, _parse_position_keywords
from ..photons import PhotonBatch
from ..program import Lowering, NotFusable
from .. import program as _program
from .. import rng as _rng

__all__ = ['Source', 'PointSource', 'LabPointSourceCone', 'FarLabPointSource', 'FixedPointing', 'JitterPointing',
           'RandomArbitraryPdfTable', 'observe', 'SourceSpecificationError']


class SourceSpecificationError(Exception):
    pass


def _val(x, unit=None):
    """Plain float from a number or an astropy-like Quantity (converted to ``unit`` first)."""
    if hasattr(x, 'to') and unit is not None:
        import astropy.units as u
        return float(x.to(getattr(u, unit) if isinstance(unit, str) else unit).value)
    if hasattr(x, 'value'):
        return float(x.value)
    return float(x)


def _area_cm2(x):
    """Geometric area in cm**2 (the unit of ``flux`` is 1 / s / cm**2, reference basesources.py:151-174 where
    ``(flux * geomarea * u.s).decompose()`` converts): ``aperture.area`` values carry mm**2
    (``marxs_b200.optics.aperture.AreaMM2``, the reference returns ``u.mm**2`` quantities), astropy
    quantities are converted, bare numbers are cm**2 like the reference default ``1 * u.cm**2``."""
    from ..optics.aperture import AreaMM2
    if isinstance(x, AreaMM2):
        return float(x) / 100.
    if hasattr(x, 'to') and hasattr(x, 'unit'):
        import astropy.units as u
        return float(x.to(u.cm ** 2).value)
    return float(x)


def _flux_per_s_cm2(x):
    if hasattr(x, 'to') and hasattr(x, 'unit'):
        import astropy.units as u
        return float(x.to(1 / u.s / u.cm ** 2).value)
    return float(x)


def _strip(x, unit):
    """Array (numpy or torch) from an array or an astropy-like Quantity array (converted to ``unit`` first)."""
    if hasattr(x, 'to') and hasattr(x, 'unit'):
        import astropy.units as u
        return x.to(getattr(u, unit)).value
    return x


def poisson_process(rate):
    """Poisson distributed photon arrival times with expectation ``rate`` [1 / s / cm**2] (reference
    basesources.py:15-62 ``poisson_process``), generated ON THE DEVICE: exponential waiting times (torch's
    Philox stream) and their running sum, cut at the exposure time.  Returns the function
    ``flux(exposuretime, geomarea)`` a source takes as its ``flux``."""
    rate = _flux_per_s_cm2(rate)

    def poisson_rate(exposuretime, geomarea, device='cuda'):
        full = rate * _area_cm2(geomarea)
        t_exp = _val(exposuretime, 's')
        # 10 % more numbers than expected (+ 5 sigma), more if that was not enough
        n = int(t_exp * full * 1.1 + 5. * np.sqrt(t_exp * full) + 16)
        times = torch.cumsum(torch.empty(n, dtype=torch.float64, device=device).exponential_(full), 0)
        while float(times[-1]) < t_exp:
            more = torch.cumsum(torch.empty(n, dtype=torch.float64, device=device).exponential_(full), 0) + times[-1]
            times = torch.cat([times, more])
        return times[times < t_exp]
    return poisson_rate


def _radec(coords):
    if hasattr(coords, 'ra') and hasattr(coords, 'dec'):
        c = coords.icrs if hasattr(coords, 'icrs') else coords
        return float(c.ra.deg), float(c.dec.deg)
    ra, dec = coords
    return float(ra), float(dec)


class RandomArbitraryPdfTable:
    """Device table of math/random.py:4-95 RandomArbitraryPdf (sort=True, randomize_in_bin=True):
    n, cdf[n], sortindex[n], x[n], bin_width[n]."""

    def __init__(self, x, pdf):
        x, pdf = np.asarray(x, dtype=float), np.asarray(pdf, dtype=float)
        if len(x) != len(pdf):
            raise ValueError('x and pdf must have same number of elements.')
        if not np.all(pdf >= 0):
            raise ValueError('pdf cannot have negative elements.')
        self.x = x
        self.bin_width = np.hstack(([0], np.diff(x)))
        if not np.all(self.bin_width >= 0):
            raise ValueError('x must be input in increasing order.')
        pdf = pdf * self.bin_width
        self.sortindex = np.argsort(pdf)
        self.cdf = np.cumsum(pdf[self.sortindex])

    def table(self):
        return np.concatenate([[len(self.x)], self.cdf, self.sortindex.astype(float), self.x, self.bin_width])


def _column(tab, name):
    col = tab[name]
    return np.asarray(col.value if hasattr(col, 'value') else col, dtype=float)


class Source(SimulationSequenceElement):
    """Base class of all photon sources (reference basesources.py:85-277)."""

    sky = False

    def __init__(self, energy=1., flux=1., polarization=None, geomarea=1., **kwargs):
        self.energy, self.flux, self.polarization = energy, flux, polarization
        self.geomarea = 1 if geomarea is None else geomarea
        super().__init__(**kwargs)

    def __call__(self, *args, **kwargs):
        return self.generate_photons(*args, **kwargs)

    # ---- specification -> numbers ---------------------------------------------------------
    def _rate(self):
        if callable(self.flux):
            return float('inf')      # dt = 0: the time column comes from the callable (GENERATE flag bit 2)
        return _flux_per_s_cm2(self.flux) * _area_cm2(self.geomarea)

    def n_photons(self, exposuretime):
        if callable(self.flux):
            raise SourceSpecificationError('the photon count of a callable flux is only known after calling it: '
                                           'use generate_photons() / observe()')
        # len(np.arange(0, T, dt)) without materialising it: numpy computes ceil((stop - start) / step)
        return max(int(np.ceil((_val(exposuretime, 's') - 0.) / (1. / self._rate()))), 0)

    def has_callables(self):
        return callable(self.flux) or callable(self.energy) or callable(self.polarization)

    def evaluate_callables(self, exposuretime, device):
        """Callable flux(exposuretime, geomarea), energy(times), polarization(times, energies) (reference
        basesources.py:167-214) are user code: they are called once per observation, on the host or - when
        they return torch tensors, like `poisson_process` here - on the device, and their results enter the
        kernel as input columns.  Returns {'time': t, 'energy': e, 'polangle': p} (device tensors, only the
        callable ones) and the photon count (None when the flux is constant)."""
        out = {}

        def as_dev(x, unit):
            x = _strip(x, unit)
            if isinstance(x, torch.Tensor):
                return x.to(device=device, dtype=torch.float64)
            return torch.as_tensor(np.asarray(x, dtype=float), device=device)

        def for_user(t):      # user functions get what they can digest: numpy unless they produced tensors
            return t if getattr(self, '_callables_take_tensors', False) else t.cpu().numpy()
        n = None
        if callable(self.flux):
            t = self.flux(exposuretime, self.geomarea)
            self._callables_take_tensors = isinstance(_strip(t, 's'), torch.Tensor)
            out['time'] = as_dev(t, 's')
            n = int(out['time'].shape[0])
        if callable(self.energy) or callable(self.polarization):
            if 'time' in out:
                times = out['time']
            else:
                n_const = self.n_photons(exposuretime)
                times = torch.arange(n_const, dtype=torch.float64, device=device) * (1. / self._rate())
            if callable(self.energy):
                e = self.energy(for_user(times))
                out['energy'] = as_dev(e, 'keV')
                if out['energy'].shape[0] != times.shape[0]:
                    raise SourceSpecificationError('`energy` has to return an array of same size as input time array.')
            if callable(self.polarization):
                if 'energy' not in out and not (hasattr(self.energy, 'columns') or isinstance(self.energy, dict)):
                    energies = torch.full_like(times, _val(self.energy, 'keV'))
                elif 'energy' in out:
                    energies = out['energy']
                else:
                    raise SourceSpecificationError('a callable polarization(times, energies) needs a constant or a '
                                                   'callable energy: tabulated energies are drawn inside the kernel')
                pol = self.polarization(for_user(times), for_user(energies))
                out['polangle'] = as_dev(pol, 'rad')
                if out['polangle'].shape[0] != times.shape[0]:
                    raise SourceSpecificationError('`polarization` has to return an array of same size as input time '
                                                   'and energy arrays.')
        return out, n

    def _energy_spec(self):
        e = self.energy
        if callable(e):
            return 0, 0., None       # read from the energy plane (GENERATE flag bit 0)
        if hasattr(e, 'columns') or isinstance(e, dict):
            x = _column(e, 'energy')
            y = np.hstack(([0], _column(e, 'fluxdensity')[1:]))
            return 1, 0., RandomArbitraryPdfTable(x, y).table()
        return 0, _val(e, 'keV'), None

    def _pol_spec(self):
        p = self.polarization
        if callable(p):
            return 0, 0., None       # read from the polangle column (GENERATE flag bit 1)
        if p is None:
            return 1, 0., None
        if hasattr(p, 'columns') or isinstance(p, dict):
            return 2, 0., RandomArbitraryPdfTable(_column(p, 'angle'), _column(p, 'probabilitydensity')).table()
        return 0, _val(p, 'rad'), None

    # ---- lowering ------------------------------------------------------------------------------
    def _lower(self, lw):
        """GENERATE op: time, energy, polangle, probability (+ ra, dec for sky sources)."""
        if lw.ops:
            raise NotFusable('a source must be the first element')
        lw.born = True
        e_mode, e_const, e_tab = self._energy_spec()
        p_mode, p_const, p_tab = self._pol_spec()
        flags = (1 if callable(self.energy) else 0) | (2 if callable(self.polarization) else 0) | (4 if callable(self.flux) else 0)
        vals = [1. / self._rate(), e_mode, e_const, -1., p_mode, p_const, -1., 1. if self.sky else 0.,
                getattr(self, 'ra', 0.), getattr(self, 'dec', 0.)]
        if e_tab is not None and p_tab is not None:
            # two global tables: store them back to back and point both fields into the block
            off = lw.params_with_table(vals, 3, np.concatenate([e_tab, p_tab]))
            lw.fixup_relative(off + 6, off + 3, len(e_tab))
        elif e_tab is not None:
            off = lw.params_with_table(vals, 3, e_tab)
        elif p_tab is not None:
            off = lw.params_with_table(vals, 6, p_tab)
        else:
            off = lw.params(vals)
        s = [-1, -1, -1, -1]
        if e_mode == 1:
            s[0], s[1] = lw.slot('uniform'), lw.slot('uniform')
        if p_mode >= 1:
            s[2] = lw.slot('uniform')
        if p_mode == 2:
            s[3] = lw.slot('uniform')
        cols = [lw.fcol('time'), lw.fcol('polangle')]
        if self.sky:
            cols += [lw.fcol('ra'), lw.fcol('dec')]
        lw.op('GENERATE', flags=flags, pg=off, cols=cols, s0=s[0], s1=s[1], w14=s[2], w15=s[3])

    def _drop_after_birth(self):
        return ('pos', 'dir', 'polarization')

    def generate_photons(self, exposuretime, device=None, id0=0):
        """Photon table born on the device (reference :234-277): columns time, energy, polangle,
        probability (+ ra, dec / pos, dir, polarization depending on the source)."""
        photons, given = _born_table(self, exposuretime, device, id0, None, None)
        _run_born([self], photons, given=given)
        for c in self._drop_after_birth():
            photons.remove_column(c)
        return photons


class AstroSource(Source):
    sky = True

    def __init__(self, **kwargs):
        self.ra, self.dec = _radec(kwargs.pop('coords'))
        super().__init__(**kwargs)


class PointSource(AstroSource):
    """Astrophysical point source (reference basesources.py:337-352)."""

    def generate_photons(self, exposuretime, device=None, id0=0):
        photons = super().generate_photons(exposuretime, device, id0)
        photons.meta['COORDSYS'] = ('ICRS', 'Type of coordinate system')
        return photons


class LabPointSourceCone(Source):
    """In-lab point source shining into a cone (reference labSource.py:62-137)."""

    def __init__(self, position=[0, 0, 0], half_opening=np.pi, direction=[1., 0., 0.], **kwargs):
        if (len(position) != 3) or (len(direction) != 3):
            raise ValueError('Direction and position are expected in Euclidean coordinates.')
        d = np.asanyarray(direction, dtype=float) / np.linalg.norm(direction)
        self.dir = np.array([d[0], d[1], d[2], 0.])
        self.position = np.array([position[0], position[1], position[2], 1.], dtype=float)
        self.half_opening = half_opening
        kwargs.setdefault('flux', 1.)
        super().__init__(geomarea=None, **kwargs)

    def _rotation(self):
        from ..affines import axangle2mat
        axis = np.cross(self.dir[:3], [0, 0, 1])
        return axangle2mat(axis, -np.arccos(self.dir[2]))

    def _lower(self, lw):
        super()._lower(lw)
        frac = 2 * np.pi * (1 - np.cos(self.half_opening)) / (4 * np.pi)
        lw.op('LABCONE', pg=lw.params(np.concatenate([self.position[:3], self._rotation().ravel(), [frac]])),
              cols=[lw.input_col('polangle')], s0=lw.slot('uniform'), s1=lw.slot('uniform'))

    def _drop_after_birth(self):
        return ()


class FarLabPointSource(Source):
    """Far in-lab point source seen through a rectangular aperture (reference labSource.py:13-59)."""

    def __init__(self, sourcePos, **kwargs):
        self.sourcePos = np.asarray(sourcePos, dtype=float)
        self.pos4d = _parse_position_keywords(kwargs)
        kwargs.setdefault('flux', 1.)
        super().__init__(geomarea=None, **kwargs)

    def _lower(self, lw):
        super()._lower(lw)
        lw.op('FARLAB', pg=lw.params(np.concatenate([self.pos4d[:3].ravel(), self.sourcePos[:3]])),
              cols=[lw.input_col('polangle')], s0=lw.slot('uniform'), s1=lw.slot('uniform'))

    def _drop_after_birth(self):
        return ()


# -----------------------------------------------------------------------------------------------
# pointing
# -----------------------------------------------------------------------------------------------
def skyoffset_matrix(ra0, dec0, roll):
    """ICRS cartesian -> SkyOffsetFrame(origin=(ra0, dec0), rotation=roll) cartesian (angles in rad):
    R_x(-roll) R_y(-dec0) R_z(ra0) with passive rotations, as astropy's skyoffset frame does."""
    def rot(angle, axis):
        c, s = np.cos(angle), np.sin(angle)
        i, j = {'x': (1, 2), 'y': (2, 0), 'z': (0, 1)}[axis]
        R = np.eye(3)
        R[i, i], R[j, j] = c, c
        R[i, j], R[j, i] = s, -s
        return R
    return rot(-roll, 'x') @ rot(-dec0, 'y') @ rot(ra0, 'z')


class PointingModel(SimulationSequenceElement):
    """Base class of the pointing models (reference pointing.py:14-42): a pointing turns ra / dec / polangle of
    a photon list into ``dir`` and ``polarization`` in the spacecraft system."""


class FixedPointing(PointingModel):
    """Photon directions and polarization vectors from ra, dec, polangle for a fixed pointing
    (reference pointing.py:45-177): x axis to the aimpoint, z axis North at roll 0."""

    jitter = 0.

    def __init__(self, **kwargs):
        self.ra0, self.dec0 = _radec(kwargs.pop('coords'))
        self.roll = _val(kwargs.pop('roll', 0.), 'rad')
        self.reference_transform = np.asarray(kwargs.pop('reference_transform', np.eye(4)), dtype=float)
        super().__init__(**kwargs)

    def _params(self):
        M = skyoffset_matrix(np.deg2rad(self.ra0), np.deg2rad(self.dec0), self.roll)
        T = self.reference_transform[:3, :3]
        ra, dec = np.deg2rad(0.), np.deg2rad(90.)
        v = np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])
        o = np.array([M[r, 0] * v[0] + M[r, 1] * v[1] + M[r, 2] * v[2] for r in range(3)])
        north = np.array([T[r, 0] * o[0] + T[r, 1] * o[1] + T[r, 2] * o[2] for r in range(3)])
        return np.concatenate([M.ravel(), T.ravel(), north, [self.jitter]])

    def _lower(self, lw):
        jit = isinstance(self, JitterPointing)
        lw.op('POINTING', flags=1 if jit else 0, pg=lw.params(self._params()),
              cols=[lw.input_col('ra'), lw.input_col('dec'), lw.input_col('polangle')],
              s0=lw.slot('uniform') if jit else -1, s1=lw.slot('normal') if jit else -1)
        lw.creates_core = True
        lw.meta_updates['RA_PNT'] = (self.ra0, '[deg] Pointing RA')
        lw.meta_updates['DEC_PNT'] = (self.dec0, '[deg] Pointing Dec')
        lw.meta_updates['ROLL_PNT'] = (np.rad2deg(self.roll), '[deg] Pointing Roll')
        lw.meta_updates['RA_NOM'] = (self.ra0, '[deg] Nominal Pointing RA')
        lw.meta_updates['DEC_NOM'] = (self.dec0, '[deg] Nominal Pointing Dec')
        lw.meta_updates['ROLL_NOM'] = (np.rad2deg(self.roll), '[deg] Nominal Pointing Roll')

    def __call__(self, photons):
        n = len(photons)
        created = []
        for name in ('pos', 'dir', 'polarization'):
            if name not in photons:
                t = photons.new_column(name, torch.float64, vector=True, n=n)
                t[3] = 1. if name == 'pos' else 0.
                t[:3] = float('nan') if name == 'pos' else 0.
                created.append(name)
        lw = Lowering(photons.colnames, meta=photons.meta)
        self._lower(lw)
        prog = lw.finish()
        draws = _rng.take_injected(len(prog.slot_kinds))
        prog.run(photons, draws=draws, seed=_rng.next_launch_seed(), id0=getattr(photons, 'id0', 0))
        if 'pos' in created:
            photons.remove_column('pos')        # the reference adds pos in the aperture (aperture.py:19-27)
        return photons


class JitterPointing(FixedPointing):
    """FixedPointing plus uncorrelated Gaussian pointing jitter (reference pointing.py:180-211)."""

    def __init__(self, **kwargs):
        self.jitter = abs(_val(kwargs.pop('jitter'), 'rad'))
        super().__init__(**kwargs)


# -----------------------------------------------------------------------------------------------
def _empty_batch(n, device, id0):
    photons = PhotonBatch(device=device)
    for name in ('pos', 'dir', 'polarization'):
        t = photons.new_column(name, torch.float64, vector=True, n=n)
        t[3] = 1. if name == 'pos' else 0.
    photons.new_column('energy', torch.float64, n=n)
    photons.new_column('probability', torch.float64, n=n)
    photons.id0 = id0
    return photons


def _born_table(source, exposuretime, device, id0, n, out):
    """Empty table for one observation; columns of callable source specifications already filled.
    Returns (photons, names of the filled columns)."""
    given = {}
    n_call = None
    if source.has_callables():
        dev = torch.device(device if device is not None else 'cuda')
        given, n_call = source.evaluate_callables(exposuretime, dev)
    if n is not None:
        n_tot = int(n)
        given = {k: v[id0:id0 + n_tot] for k, v in given.items()}      # this rank's shard of the host-made columns
    else:
        n_tot = n_call if n_call is not None else source.n_photons(exposuretime)
    if out is not None and len(out) == n_tot:
        photons = out                      # reuse the table (and its memory) of a previous observation
        photons.id0 = id0
    else:
        photons = _empty_batch(n_tot, device, id0)
    photons.meta['EXTNAME'] = 'EVENTS'
    photons.meta['EXPOSURE'] = (_val(exposuretime, 's'), 'total exposure time [s]')
    names = []
    for k, v in given.items():
        if v.shape[0] != n_tot:
            raise SourceSpecificationError('callable source specifications returned {0} values for {1} photons'.format(
                v.shape[0], n_tot))
        if k == 'energy':
            photons['energy'][...] = v
        else:
            photons[k] = v
            names.append(k)
    return photons, names


_born_cache = {}


def _run_born(elements, photons, check=True, given=()):
    """Lower (cached under the fingerprint of the element trees, like simulator._lower_run) and launch.
    ``given``: columns the host filled before the launch (results of callable source specifications)."""
    from ..simulator import fingerprint, Uncacheable
    pins = []
    try:
        key = fingerprint(elements, ('born', photons.meta, tuple(given), _program.EXHAUSTIVE_SEARCH), pins)
    except Uncacheable:
        key = None
    prog = _born_cache.get(key, (None,))[0] if key is not None else None
    if prog is None:
        lw = Lowering(list(given), meta=photons.meta)
        for e in elements:
            e._lower(lw)
        prog = lw.finish()
        if key is not None:
            if len(_born_cache) >= 32:
                _born_cache.pop(next(iter(_born_cache)))
            _born_cache[key] = (prog, pins)      # pins: identity-hashed objects stay alive with the entry
    draws = _rng.take_injected(len(prog.slot_kinds))
    prog.run(photons, draws=draws, seed=_rng.next_launch_seed(), id0=getattr(photons, 'id0', 0), check=check)
    return prog


def observe(source, pointing, elements, exposuretime, device=None, id0=0, n=None, check=True, out=None):
    """One observation as ONE kernel launch: source -> pointing -> elements (aperture, mirror, gratings,
    detectors ...) lowered into a single born-on-device program.  ``pointing`` may be None for lab
    sources.  ``n`` overrides the photon count (e.g. the size of this rank's shard; ``id0`` is then the
    global index of its first photon).  Falls back to the reference call sequence when an element
    cannot be fused."""
    from ..simulator import _lowerable, Sequence
    chain = [source] + ([pointing] if pointing is not None else []) + list(elements)
    photons, given = _born_table(source, exposuretime, device, id0, n, out)
    if all(_lowerable(e) or isinstance(e, (Source, FixedPointing)) for e in chain):
        try:
            _run_born(chain, photons, check=check, given=given)
            return photons
        except NotFusable:
            pass
    photons = source.generate_photons(exposuretime, device=device, id0=id0)
    if pointing is not None:
        photons = pointing(photons)
    return Sequence(elements=list(elements))(photons)
