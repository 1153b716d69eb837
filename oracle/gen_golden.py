"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE, build container only (needs /root/reference; the stand-ins
under oracle/standin supply astropy/transforms3d).  Re-run with

    python oracle/gen_golden.py

Every case stores its seeded inputs, the injected random draws and the
reference's outputs.  Random draws are injected by wrapping
``OpticalElement.process_photons`` (to learn the current intersect mask) and
replacing the ``np.random`` entry points the reference calls
 with per-photon lookups.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import tier_r  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
SEED = 20261017


# ---------------------------------------------------------------------------
# draw injection into the reference
# ---------------------------------------------------------------------------
class Injector:
    """Context manager: route the reference's np.random calls to ``table``."""

    def __init__(self, marxs, table):
        self.marxs = marxs
        self.table = table
        self.stack = []

    # -- current context ------------------------------------------------
    def _next(self, n=None):
        ctx = self.stack[-1]
        elem, mask = ctx['elem'], ctx['mask']
        slots = getattr(elem, '_slots', None)
        if not slots:
            raise RuntimeError('element {0} draws random numbers but has no slots'.format(elem))
        if n is None:   # scalar draw in a python loop (EfficiencyFile)
            vals = np.asarray(self.table[slots[ctx['i']]])[mask]
            v = vals[ctx['cursor']]
            ctx['cursor'] += 1
            if ctx['cursor'] == len(vals):
                ctx['cursor'] = 0
                ctx['i'] += 1
            return v
        vals = np.asarray(self.table[slots[ctx['i']]])[mask]
        ctx['i'] += 1
        assert len(vals) == n, (len(vals), n)
        return vals

    # -- patched numpy.random entry points ------------------------------
    def normal(self, loc=0., scale=1., size=None):
        return loc + scale * self._next(size)

    def random(self, size=None):
        return self._next(size)

    def rand(self, *shape):
        if len(shape) == 0:
            return self._next(None)
        if self.stack and self.stack[-1]['elem'] == 'multiaperture':
            return np.zeros(shape[0])
        return self._next(shape[0])

    def uniform(self, low=0., high=1., size=None):
        return low + (high - low) * self._next(size)

    def choice(self, a, size=None, p=None):
        a = np.asarray(a)
        u = self._next(size)
        cdf = np.asarray(p).cumsum()
        cdf /= cdf[-1]
        return a[cdf.searchsorted(u, side='right')]

    def shuffle(self, x):
        ctx = self.stack[-1]
        assert ctx['elem'] == 'multiaperture'
        x[:] = np.asarray(self.table[ctx['slot']]).astype(int)

    def __enter__(self):
        import marxs.optics.base as ob
        import marxs.optics.aperture as ap
        inj = self
        self._orig_pp = ob.OpticalElement.process_photons
        self._orig_ma = ap.MultiAperture.__call__
        self._orig = {k: getattr(np.random, k) for k in
                      ('normal', 'random', 'rand', 'uniform', 'choice', 'shuffle')}

        def process_photons(elem, photons, intersect, interpos, intercoos):
            inj.stack.append({'elem': elem, 'mask': np.asarray(intersect), 'i': 0, 'cursor': 0})
            try:
                return inj._orig_pp(elem, photons, intersect, interpos, intercoos)
            finally:
                inj.stack.pop()

        def multi_call(elem, photons):
            inj.stack.append({'elem': 'multiaperture', 'slot': elem._slots[0]})
            try:
                return inj._orig_ma(elem, photons)
            finally:
                inj.stack.pop()

        ob.OpticalElement.process_photons = process_photons
        ap.MultiAperture.__call__ = multi_call
        for k in self._orig:
            setattr(np.random, k, getattr(self, k))
        return self

    def __exit__(self, *exc):
        import marxs.optics.base as ob
        import marxs.optics.aperture as ap
        ob.OpticalElement.process_photons = self._orig_pp
        ap.MultiAperture.__call__ = self._orig_ma
        for k, v in self._orig.items():
            setattr(np.random, k, v)


def table_to_dict(photons, prefix='out_'):
    out = {}
    for c in photons.colnames:
        out[prefix + c] = np.asarray(photons[c].data)
    return out


def rand_pos4d(rng, zoom=(1., 8., 5.), shift=20.):
    from transforms3d.euler import euler2mat
    from transforms3d.affines import compose
    R = euler2mat(*rng.uniform(-0.4, 0.4, 3))
    return compose(rng.uniform(-shift, shift, 3), R, zoom)


def make_photons(rng, n, spread=0.3, x0=100., lateral=10., e_lo=0.3, e_hi=8.):
    """Photons flying roughly along -x from x~x0 with un-normalised dirs."""
    from astropy.table import Table
    pos = np.ones((n, 4))
    pos[:, 0] = x0 + rng.uniform(-5, 5, n)
    pos[:, 1:3] = rng.uniform(-lateral, lateral, (n, 2))
    dir = np.zeros((n, 4))
    dir[:, 0] = -1.
    dir[:, 1:3] = rng.normal(0, spread, (n, 2))
    dir[:, :3] *= rng.uniform(0.5, 2., n)[:, None]
    # random unit polarization perpendicular to dir
    v = rng.normal(size=(n, 3))
    d = dir[:, :3] / np.linalg.norm(dir[:, :3], axis=1)[:, None]
    v -= d * np.einsum('ij,ij->i', v, d)[:, None]
    v /= np.linalg.norm(v, axis=1)[:, None]
    pol = np.zeros((n, 4))
    pol[:, :3] = v
    t = Table({'pos': pos, 'dir': dir, 'energy': rng.uniform(e_lo, e_hi, n),
               'polarization': pol, 'probability': rng.uniform(0.2, 1., n)})
    return t


def inputs_dict(p):
    return {'in_' + c: np.asarray(p[c].data).copy() for c in p.colnames}


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, '{0:.1f} kB'.format(os.path.getsize(path) / 1e3))


# ---------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------
def case_intersect(marxs, rng):
    from marxs.math.geometry import FinitePlane, CircularHole
    n = 1500
    p = make_photons(rng, n, spread=0.12, lateral=10., x0=60.)
    # a few special rays: parallel to the plane, flying away
    pos4d = rand_pos4d(rng, shift=4.)
    g = FinitePlane({'pos4d': pos4d})
    ey = g['e_y']
    p['dir'][0] = ey            # exactly parallel -> k_den == 0 is not guaranteed; keep anyway
    p['dir'][1] = np.array([1., 0., 0., 0.])   # backwards
    hit, ipos, loc = g.intersect(p['dir'].data, p['pos'].data)
    pos4d_c = rand_pos4d(rng, zoom=(1., 6., 0.9), shift=.4)
    gc = CircularHole({'pos4d': pos4d_c})
    pc = make_photons(rng, n, spread=0.01, lateral=1.6, x0=30.)
    hit_c, ipos_c, loc_c = gc.intersect(pc['dir'].data, pc['pos'].data)
    save('intersect', pos4d=pos4d, pos4d_circ=pos4d_c, hit=hit, interpos=ipos, loc=loc,
         hit_circ=hit_c, interpos_circ=ipos_c, loc_circ=loc_c,
         circ_in_dir=np.asarray(pc['dir'].data), circ_in_pos=np.asarray(pc['pos'].data),
         **inputs_dict(p))


def case_parallel_transport(marxs, rng):
    from marxs.math.polarization import parallel_transport
    n = 1000
    d1 = np.zeros((n, 4))
    d2 = np.zeros((n, 4))
    pol = np.zeros((n, 4))
    d1[:, :3] = rng.normal(size=(n, 3))
    d2[:, :3] = d1[:, :3] + rng.normal(scale=0.2, size=(n, 3))
    d2[:50] = d1[:50]                      # unchanged directions -> identity
    d2[50:60, :3] = d1[50:60, :3] * 3.     # parallel, different length
    pol[:, :3] = rng.normal(size=(n, 3))
    out = parallel_transport(d1, d2, pol)
    save('parallel_transport', dir_old=d1, dir_new=d2, pol_old=pol, pol_new=out)


def case_gratings(marxs, rng):
    from marxs.optics import FlatGrating, CATGrating, OrderSelector
    n = 1500
    arrays = {}
    for tag, cls, kw in [
            ('flat', FlatGrating, dict(d=2e-4, groove_angle=0.2)),
            ('cat', CATGrating, dict(d=2e-4, groove_angle=-0.1)),
            ('flat_refl', FlatGrating, dict(d=4e-4, transmission=False))]:
        p = make_photons(rng, n, spread=0.05, lateral=6., e_lo=0.3, e_hi=3.)
        pos4d = rand_pos4d(rng, zoom=(1., 9., 7.), shift=3.)
        sel = OrderSelector(np.arange(-3, 4), p=np.array([.05, .1, .2, .25, .2, .1, .05]))
        g = cls(pos4d=pos4d, order_selector=sel, **kw)
        g._slots = [0]
        draws = [rng.random(n)]
        inp = inputs_dict(p)
        with Injector(marxs, draws):
            out = g(p)
        arrays.update({tag + '_' + k: v for k, v in inp.items()})
        arrays.update({tag + '_' + k: v for k, v in table_to_dict(out).items()})
        arrays[tag + '_pos4d'] = pos4d
        arrays[tag + '_u'] = draws[0]
    save('gratings', **arrays)


def case_order_selectors(marxs, rng):
    """np.random.choice == inverse cdf; EfficiencyFile; InterpolateEfficiencyTable."""
    from marxs.optics import OrderSelector, EfficiencyFile
    import tempfile
    n = 2000
    sel = OrderSelector(np.arange(-2, 3), p=np.array([.1, .2, .3, .2, .1]))
    np.random.seed(1234)
    m_ref, p_ref = sel(np.ones(n))
    np.random.seed(1234)
    u = np.random.random_sample(n)
    # efficiency file
    en = np.array([0.3, 0.5, 1.0, 2.0, 4.0, 8.0])
    tab = np.hstack([en[:, None], rng.uniform(0.01, 0.2, (6, 5))])
    with tempfile.NamedTemporaryFile('w', suffix='.dat', delete=False) as f:
        np.savetxt(f, tab)
    ef = EfficiencyFile(f.name, [-2, -1, 0, 1, 2])
    os.unlink(f.name)
    energies = rng.uniform(0.2, 9., n)
    ue = rng.random(n)

    class Ctx:
        _slots = [0]
    inj = Injector(marxs, [ue])
    inj.stack.append({'elem': Ctx, 'mask': np.ones(n, bool), 'i': 0, 'cursor': 0})
    orig = np.random.rand
    np.random.rand = inj.rand
    try:
        m_ef, p_ef = ef(energies)
    finally:
        np.random.rand = orig
    # interpolated table (k=1) with a synthetic regular table
    from marxs.missions.mitsnl.catgrating import InterpolateEfficiencyTable
    from astropy.table import Table
    import astropy.units as uu
    wave = np.array([0.1, 0.2, 0.35, 0.6, 1.0, 1.7, 2.5, 4.0])
    theta_deg = np.array([0.5, 1.0, 1.5, 2.0, 3.0])
    orders = [2, 1, 0, -1, -2, -3]
    W, T = np.meshgrid(wave, theta_deg, indexing='ij')
    data = {'lambda': W.ravel(), 'theta': T.ravel()}
    prob = rng.uniform(0.0, 0.15, (len(wave), len(theta_deg), len(orders)))
    for k, o in enumerate(orders):
        data[str(o)] = prob[:, :, k].ravel()
    t = Table(data)
    t['theta'].unit = uu.deg
    iet = InterpolateEfficiencyTable(t)
    en_i = rng.uniform(0.25, 14., n)     # some outside the table range (clamped)
    blaze = np.deg2rad(rng.uniform(0.3, 3.3, n))
    ui = rng.random(n)
    o_ref, pr_ref = iet.probabilities(en_i, None, blaze)
    inj2 = Injector(marxs, [ui])
    inj2.stack.append({'elem': Ctx, 'mask': np.ones(n, bool), 'i': 0, 'cursor': 0})
    np.random.rand = inj2.rand
    try:
        m_iet, p_iet = iet(en_i, None, blaze)
    finally:
        np.random.rand = orig
    save('order_selectors', sel_orders=sel.orderlist, sel_p=sel.p, sel_u=u, sel_m=m_ref,
         sel_prob=p_ref, ef_table=tab, ef_orders=np.array([-2, -1, 0, 1, 2]), ef_energy=energies,
         ef_u=ue, ef_m=m_ef, ef_prob=p_ef,
         iet_wave=wave, iet_theta=np.deg2rad(theta_deg), iet_prob=prob, iet_orders=np.array(orders),
         iet_energy=en_i, iet_blaze=blaze, iet_u=ui, iet_probs=pr_ref.T, iet_m=m_iet, iet_total=p_iet)


def case_lens_scatter(marxs, rng):
    from marxs.optics import PerfectLens, RadialMirrorScatter, RandomGaussianScatter, FlatStack, EnergyFilter
    import astropy.units as u
    n = 1500
    arrays = {}
    # lens alone with off-axis centre
    p = make_photons(rng, n, spread=0.01, x0=300., lateral=40.)
    pos4d = rand_pos4d(rng, zoom=(1., 60., 60.), shift=2.)
    lens = PerfectLens(focallength=250., d_center_optical_axis=7.5, pos4d=pos4d)
    inp = inputs_dict(p)
    out = lens(p)
    arrays.update({'lens_' + k: v for k, v in inp.items()})
    arrays.update({'lens_' + k: v for k, v in table_to_dict(out).items()})
    arrays['lens_pos4d'] = pos4d
    # standalone radial scatter (uses the OLD position for "radial")
    p = make_photons(rng, n, spread=0.02, x0=50., lateral=30.)
    pos4d = rand_pos4d(rng, zoom=(1., 50., 50.), shift=1.)
    sc = RadialMirrorScatter(inplanescatter=2e-3 * u.rad, perpplanescatter=5e-4 * u.rad, pos4d=pos4d)
    sc._slots = [0, 1]
    draws = [rng.standard_normal(n), rng.standard_normal(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = sc(p)
    arrays.update({'rms_' + k: v for k, v in inp.items()})
    arrays.update({'rms_' + k: v for k, v in table_to_dict(out).items()})
    arrays['rms_pos4d'] = pos4d
    arrays['rms_z0'], arrays['rms_z1'] = draws
    # random gaussian scatter
    p = make_photons(rng, n, spread=0.02, x0=50., lateral=30.)
    p['dir'][:5] = np.array([1., 1e-7, 0., 0.])   # |pdir_x| >= 0.99999 branch (and flying away: miss)
    p['dir'][5:10] = np.array([-1., 1e-7, 0., 0.])
    sg = RandomGaussianScatter(scatter=1e-3 * u.rad, pos4d=pos4d)
    sg._slots = [0, 1]
    draws = [rng.standard_normal(n), rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = sg(p)
    arrays.update({'rgs_' + k: v for k, v in inp.items()})
    arrays.update({'rgs_' + k: v for k, v in table_to_dict(out).items()})
    arrays['rgs_z0'], arrays['rgs_u1'] = draws
    # HRMA-like stack
    p = make_photons(rng, n, spread=0.003, x0=400., lateral=45.)
    pos4d = rand_pos4d(rng, zoom=(1., 65., 65.), shift=1.)
    st = FlatStack(pos4d=pos4d, elements=[PerfectLens, RadialMirrorScatter, EnergyFilter],
                   keywords=[{'focallength': 300.},
                             {'inplanescatter': 3e-4 * u.rad, 'perpplanescatter': 1e-4 * u.rad},
                             {'filterfunc': lambda x: np.ones_like(x) * 0.66}])
    st.elements[1]._slots = [0, 1]
    draws = [rng.standard_normal(n), rng.standard_normal(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = st(p)
    arrays.update({'stack_' + k: v for k, v in inp.items()})
    arrays.update({'stack_' + k: v for k, v in table_to_dict(out).items()})
    arrays['stack_pos4d'] = pos4d
    arrays['stack_z0'], arrays['stack_z1'] = draws
    save('lens_scatter', **arrays)


def case_detectors(marxs, rng):
    from marxs.optics import FlatDetector
    n = 1500
    arrays = {}
    p = make_photons(rng, n, spread=0.1, lateral=8.)
    pos4d = rand_pos4d(rng, zoom=(1., 12.288, 6.144), shift=2.)
    det = FlatDetector(pixsize=0.024, pos4d=pos4d)
    inp = inputs_dict(p)
    out = det(p)
    arrays.update({'det_' + k: v for k, v in inp.items()})
    arrays.update({'det_' + k: v for k, v in table_to_dict(out).items()})
    arrays['det_pos4d'] = pos4d
    arrays['det_npix'] = np.array(det.npix)
    arrays['det_centerpix'] = np.array(det.centerpix)
    save('detectors', **arrays)


def case_mlmirror(marxs, rng):
    from marxs.optics import FlatBrewsterMirror
    from marxs.optics.multiLayerMirror import MultiLayerMirror, MultiLayerEfficiency
    from astropy.io import ascii
    from transforms3d.euler import euler2mat
    from transforms3d.affines import compose
    n = 1500
    arrays = {}
    p = make_photons(rng, n, spread=0.02, x0=60., lateral=5., e_lo=0.25, e_hi=0.45)
    R = euler2mat(0.1, -0.75, 0.2)
    pos4d = compose([1., 0.5, -0.5], R, [1., 24.5, 12.])
    m = FlatBrewsterMirror(pos4d=pos4d)
    inp = inputs_dict(p)
    out = m(p)
    arrays.update({'brew_' + k: v for k, v in inp.items()})
    arrays.update({'brew_' + k: v for k, v in table_to_dict(out).items()})
    arrays['brew_pos4d'] = pos4d
    ddir = os.path.join(tier_r.REFERENCE_ROOT, 'marxs', 'optics', 'data')
    refl = ascii.read(os.path.join(ddir, 'A12113.txt'))
    polf = ascii.read(os.path.join(ddir, 'ALSpolarization2.txt'))
    arrays['ml_x_mm'] = np.asarray(refl['X(mm)'].data, float)
    arrays['ml_peak_lambda'] = np.asarray(refl['Peak lambda'].data, float)
    arrays['ml_peak'] = np.asarray(refl['Peak'].data, float)
    arrays['ml_fwhm'] = np.asarray(refl['FWHM(nm)'].data, float)
    arrays['ml_pol_energy_ev'] = np.asarray(polf['Photon energy'].data, float)
    arrays['ml_pol'] = np.asarray(polf['Polarization'].data, float)
    p = make_photons(rng, n, spread=0.02, x0=60., lateral=5., e_lo=0.25, e_hi=0.45)
    mm = MultiLayerMirror(reflFile=os.path.join(ddir, 'A12113.txt'),
                          testedPolarization=os.path.join(ddir, 'ALSpolarization2.txt'),
                          pos4d=pos4d)
    inp = inputs_dict(p)
    out = mm(p)
    arrays.update({'mlm_' + k: v for k, v in inp.items()})
    arrays.update({'mlm_' + k: v for k, v in table_to_dict(out).items()})
    save('mlmirror', **arrays)


def case_apertures_baffle(marxs, rng):
    from marxs.optics import RectangleAperture, CircleAperture, MultiAperture, Baffle, CircularBaffle
    from astropy.table import Table
    n = 1500
    arrays = {}

    def src(n):
        d = np.zeros((n, 4))
        d[:, 0] = -1.
        d[:, 1:3] = rng.normal(0, 0.01, (n, 2))
        pol = np.zeros((n, 4))
        pol[:, 1] = 1.
        return Table({'dir': d, 'energy': rng.uniform(0.5, 2, n), 'polarization': pol,
                      'probability': np.ones(n)})
    p = src(n)
    pos4d = rand_pos4d(rng, zoom=(1., 4., 2.), shift=5.)
    ap = RectangleAperture(pos4d=pos4d)
    ap._slots = [0, 1]
    draws = [rng.random(n), rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = ap(p)
    arrays.update({'rect_' + k: v for k, v in inp.items()})
    arrays.update({'rect_' + k: v for k, v in table_to_dict(out).items()})
    arrays['rect_pos4d'] = pos4d
    arrays['rect_u0'], arrays['rect_u1'] = draws
    # multi aperture of rings (Chandra-like)
    p = src(n)
    radii = np.array([[59.8, 61.0], [48.1, 49.1], [42.4, 43.3]])
    aps = [CircleAperture(position=[100., 0, 0], zoom=[1, r[1], r[1]], r_inner=r[0]) for r in radii]
    ma = MultiAperture(elements=aps, id_col='mirror_shell')
    for a in aps:
        a._slots = [1, 2]
    ma._slots = [0]
    areas = np.array([a.area.value for a in aps])
    aperid = rng.choice(3, size=n, p=areas / areas.sum()).astype(float)
    draws = [aperid, rng.random(n), rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = ma(p)
    arrays.update({'multi_' + k: v for k, v in inp.items()})
    arrays.update({'multi_' + k: v for k, v in table_to_dict(out).items()})
    arrays['multi_radii'] = radii
    arrays['multi_aperid'], arrays['multi_u0'], arrays['multi_u1'] = draws
    # baffles
    p = make_photons(rng, n, spread=0.1, lateral=3.)
    pos4d = rand_pos4d(rng, zoom=(1., 3., 2.), shift=1.)
    b = Baffle(pos4d=pos4d)
    inp = inputs_dict(p)
    out = b(p)
    arrays.update({'baffle_' + k: v for k, v in inp.items()})
    arrays.update({'baffle_' + k: v for k, v in table_to_dict(out).items()})
    arrays['baffle_pos4d'] = pos4d
    p = make_photons(rng, n, spread=0.05, lateral=1.5, x0=20.)
    pos4d = rand_pos4d(rng, zoom=(1., 1.5, 0.8), shift=.2)
    b = CircularBaffle(pos4d=pos4d)
    inp = inputs_dict(p)
    out = b(p)
    arrays.update({'cbaffle_' + k: v for k, v in inp.items()})
    arrays.update({'cbaffle_' + k: v for k, v in table_to_dict(out).items()})
    arrays['cbaffle_pos4d'] = pos4d
    save('apertures_baffle', **arrays)


def chandra_photons(rng, n):
    """Synthetic C2 input: on-axis photons on the 4 HRMA annuli."""
    from astropy.table import Table
    radii = np.array([[598., 610.], [481, 491], [424, 433], [315, 322]])
    area = radii[:, 1] ** 2 - radii[:, 0] ** 2
    shell = rng.choice(4, size=n, p=area / area.sum())
    r = np.sqrt(rng.uniform(radii[shell, 0] ** 2, radii[shell, 1] ** 2))
    phi = rng.uniform(0, 2 * np.pi, n)
    pos = np.ones((n, 4))
    pos[:, 0] = 10061.65 + 100.
    pos[:, 1] = r * np.cos(phi)
    pos[:, 2] = r * np.sin(phi)
    dir = np.zeros((n, 4))
    dir[:, 0] = -1.
    ang = rng.uniform(0, 2 * np.pi, n)
    pol = np.zeros((n, 4))
    pol[:, 1] = np.cos(ang)
    pol[:, 2] = np.sin(ang)
    t = Table({'pos': pos, 'dir': dir, 'energy': rng.uniform(0.5, 8., n),
               'polarization': pol, 'probability': np.ones(n)})
    t.meta['ROLL_PNT'] = (0., 'roll')
    return t


def case_chandra(marxs, rng):
    from marxs.missions import chandra
    n = 4000
    hrma = chandra.HRMA()
    hetg = chandra.HETG()
    acis = chandra.ACIS(chips=[4, 5, 6, 7, 8, 9], aimpoint=chandra.AIMPOINTS['ACIS-S'])
    hrma.elements[1]._slots = [0, 1]
    for e in hetg.elements:
        e._slots = [2]
    draws = [rng.standard_normal(n), rng.standard_normal(n), rng.random(n)]
    p = chandra_photons(rng, n)
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        p = hrma(p)
        after_hrma = {'hrma_' + c: np.asarray(p[c].data).copy() for c in ('pos', 'dir', 'polarization', 'probability')}
        p = hetg(p)
        after_hetg = {'hetg_' + c: np.asarray(p[c].data).copy() for c in ('pos', 'dir', 'polarization', 'probability')}
        p = acis(p)
    save('chandra_c2', z0=draws[0], z1=draws[1], u2=draws[2],
         hetg_pos4d=np.array([e.pos4d for e in hetg.elements]),
         hetg_groove=np.array([e.geometry['groove_angle'] for e in hetg.elements]),
         hetg_d=np.array([e._d for e in hetg.elements]),
         acis_pos4d=np.array([e.pos4d for e in acis.elements]),
         acis_id=np.array([e.id_num for e in acis.elements]),
         hrma_pos4d=hrma.pos4d,
         **inp, **after_hrma, **after_hetg, **table_to_dict(p))


def case_parallel_overlap(marxs, rng):
    """Overlapping facets: pins the sequential 'last hit wins' semantics."""
    from marxs.simulator import Parallel
    from marxs.optics import FlatGrating, OrderSelector, FlatDetector
    n = 1500
    p = make_photons(rng, n, spread=0.02, lateral=12., x0=50., e_lo=0.5, e_hi=2.)
    pos = [[0., -4., 0.], [-3., 2., 1.], [2., 4., -3.], [-6., -1., 5.]]
    par = Parallel(elem_class=FlatGrating,
                   elem_args={'d': [2e-4, 3e-4, 2.5e-4, 4e-4], 'zoom': [1, 5., 6.],
                              'order_selector': OrderSelector([-1, 0, 1]),
                              'groove_angle': [0., 0.1, -0.2, 0.05]},
                   elem_pos={'position': pos}, id_col='facet')
    for e in par.elements:
        e._slots = [0]
    draws = [rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = par(p)
    save('parallel_overlap', u0=draws[0], pos4d=np.array([e.pos4d for e in par.elements]),
         **inp, **table_to_dict(out))


def case_cat_stack(marxs, rng):
    """SNL CAT grating stack (membrane + quality factor + L1 + L2 absorption + L2 diffraction) and a
    Parallel of such stacks: catgrating.py:147-374."""
    from marxs.missions.mitsnl import catgrating as cg
    from marxs.optics import OrderSelector
    from marxs.simulator import Parallel
    n = 1200
    arrays = {}
    arrays['trans_1um'] = np.asarray(cg.l1transtab['transmission'].data, float)
    sel = OrderSelector(np.arange(-2, 9), p=np.array([.01, .02, .2, .05, .05, .05, .1, .15, .15, .1, .02]))
    pos4d = rand_pos4d(rng, zoom=(1., 14., 15.), shift=2.)
    p = make_photons(rng, n, spread=0.03, lateral=20., x0=70., e_lo=0.25, e_hi=2.0)
    st = cg.CATL1L2Stack(pos4d=pos4d, order_selector=sel, groove_angle=0.05)
    st.elements[0]._slots = [0]
    st.elements[2]._slots = [1, 2]
    st.elements[4]._slots = [3, 4]
    arrays['trans_energy'] = np.asarray(st.elements[2].transfunc.x, float)     # keV, as the reference uses it
    draws = [rng.random(n), rng.random(n), rng.random(n), rng.standard_normal(n), rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = st(p)
    arrays.update({'stack_' + k: v for k, v in inp.items()})
    arrays.update({'stack_' + k: v for k, v in table_to_dict(out).items()})
    arrays['stack_pos4d'] = pos4d
    arrays['stack_sel_orders'], arrays['stack_sel_p'] = sel.orderlist, sel.p
    for k, d in enumerate(draws):
        arrays['stack_draw{0}'.format(k)] = d
    # a small array of stacks + the support-bar rule
    pos = [[0., y, z] for y in (-16., 0., 16.) for z in (-17., 0., 17.)]
    par = Parallel(elem_class=cg.CATL1L2Stack, elem_pos={'position': pos}, id_col='facet',
                   elem_args={'zoom': [1, 7.5, 8.], 'order_selector': sel,
                              'orientation': rand_pos4d(rng)[:3, :3] * 0 + np.array(
                                  [[np.cos(0.03), -np.sin(0.03), 0], [np.sin(0.03), np.cos(0.03), 0], [0, 0, 1.]])})
    for e in par.elements:
        e.elements[0]._slots = [0]
        e.elements[2]._slots = [1, 2]
        e.elements[4]._slots = [3, 4]
    p = make_photons(rng, n, spread=0.02, lateral=26., x0=60., e_lo=0.3, e_hi=1.5)
    draws = [rng.random(n), rng.random(n), rng.random(n), rng.standard_normal(n), rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = par(p)
    out = cg.catsupportbars(out)
    arrays.update({'par_' + k: v for k, v in inp.items()})
    arrays.update({'par_' + k: v for k, v in table_to_dict(out).items()})
    arrays['par_pos4d'] = np.array([e.pos4d for e in par.elements])
    for k, d in enumerate(draws):
        arrays['par_draw{0}'.format(k)] = d
    save('cat_stack', **arrays)


def case_cylinder(marxs, rng):
    """Cylinder.intersect (geometry.py:470-564) and CircularDetector (detector.py:78-118)."""
    from marxs.math.geometry import Cylinder
    from marxs.optics import CircularDetector
    from transforms3d.euler import euler2mat
    from transforms3d.affines import compose
    n = 1000
    arrays = {}
    for tag, zoom, phi_lim, inside in (('full', [40., 40., 6.], None, False),
                                       ('half', [30., 30., 10.], [-0.15, 2.2], False),
                                       ('wrap', [25., 35., 8.], [2.5, -2.5], False)):
        R = euler2mat(*rng.uniform(-0.3, 0.3, 3))
        pos4d = compose(rng.uniform(-3, 3, 3), R, zoom)
        kw = {'pos4d': pos4d}
        if phi_lim is not None:
            kw['phi_lim'] = phi_lim
        p = make_photons(rng, n, spread=0.25, lateral=12., x0=(5. if inside else 90.))
        p['dir'][:20, :3] = np.array([0., 0., 1.])           # along the axis: a == 0
        g = Cylinder(dict(kw))
        hit, ipos, loc = g.intersect(p['dir'].data, p['pos'].data)
        arrays[tag + '_pos4d'] = pos4d
        arrays[tag + '_phi_lim'] = np.array(phi_lim if phi_lim is not None else [-np.pi, np.pi])
        arrays[tag + '_hit'], arrays[tag + '_interpos'], arrays[tag + '_loc'] = hit, ipos, loc
        arrays.update({tag + '_' + k: v for k, v in inputs_dict(p).items()})
        det = CircularDetector(pixsize=0.05, **kw)
        out = det(p)
        arrays.update({tag + '_' + k: v for k, v in table_to_dict(out).items()})
    save('cylinder', **arrays)


def case_euklid_bases(marxs, rng):
    """Geometry.get_local_euklid_bases of FinitePlane (geometry.py:263-281) and Cylinder (:602-626)."""
    from marxs.math.geometry import Cylinder, FinitePlane
    from transforms3d.euler import euler2mat
    from transforms3d.affines import compose
    arrays = {}
    for tag, cls, zoom in (('plane', FinitePlane, [1., 7., 3.]), ('cyl', Cylinder, [40., 25., 6.])):
        pos4d = compose(rng.uniform(-3, 3, 3), euler2mat(*rng.uniform(-0.5, 0.5, 3)), zoom)
        g = cls({'pos4d': pos4d})
        loc = np.column_stack([rng.uniform(-3, 3, 50), rng.uniform(-1, 1, 50)])
        e1, e2, n = g.get_local_euklid_bases(loc)
        arrays[tag + '_pos4d'], arrays[tag + '_loc'] = pos4d, loc
        arrays[tag + '_e1'], arrays[tag + '_e2'], arrays[tag + '_n'] = np.asarray(e1), np.asarray(e2), np.asarray(n)
    save('euklid_bases', **arrays)


def case_pt_matrix(marxs, rng):
    """paralleltransport_matrix (math/polarization.py:90-149) with the identity and a general Jones matrix."""
    from marxs.math.polarization import paralleltransport_matrix
    n = 40
    d1 = rng.normal(size=(n, 3))
    d2 = d1 + rng.normal(scale=0.3, size=(n, 3))
    d2[:5] = d1[:5] * rng.uniform(0.5, 2., 5)[:, None]          # parallel: the undefined case
    jones = np.array([[0.8, 0.1], [-0.2, 0.5]])
    save('pt_matrix', d1=d1, d2=d2, jones=jones,
         ident=paralleltransport_matrix(d1, d2), ident_nan=paralleltransport_matrix(d1, d2, replace_nans=False),
         gen=paralleltransport_matrix(d1, d2, jones=jones), gen_nan=paralleltransport_matrix(d1, d2, jones=jones, replace_nans=False))


class SeqFeeder:
    """Replace np.random.uniform / rand for code that draws outside any optical element (sources):
    the k-th call returns low + (high - low) * table[k]."""

    def __init__(self, table):
        self.table, self.k = table, 0

    def _next(self, n):
        v = np.asarray(self.table[self.k])
        self.k += 1
        assert len(v) == n, (len(v), n)
        return v

    def uniform(self, low=0., high=1., size=None):
        return low + (high - low) * self._next(size)

    def rand(self, n):
        return self._next(n)

    def __enter__(self):
        self._orig = (np.random.uniform, np.random.rand)
        np.random.uniform, np.random.rand = self.uniform, self.rand
        return self

    def __exit__(self, *exc):
        np.random.uniform, np.random.rand = self._orig


def case_sources(marxs, rng):
    """Photon birth without sky frames: RandomArbitraryPdf, polarization_vectors, LabPointSourceCone,
    FarLabPointSource (source/labSource.py, math/random.py, math/polarization.py:12-62)."""
    import astropy.units as u
    from astropy.table import QTable
    from marxs.math.random import RandomArbitraryPdf
    from marxs.math.polarization import polarization_vectors
    from marxs.source import LabPointSourceCone, FarLabPointSource
    arrays = {}
    n = 1000
    # arbitrary pdf (spectrum): upper bin edges + flux densities
    x = np.array([0.3, 0.5, 0.9, 1.4, 2.2, 3.5, 6.0, 10.])
    pdf = np.array([0., 3., 1., 5., 0.2, 4., 0.05, 1.])
    u0, u1 = rng.random(n), rng.random(n)
    with SeqFeeder([u0, u1]):
        arrays['pdf_out'] = RandomArbitraryPdf(x, pdf)(n)
    arrays['pdf_x'], arrays['pdf_pdf'], arrays['pdf_u0'], arrays['pdf_u1'] = x, pdf, u0, u1
    # polarization vectors incl. rays along +-y
    d = np.zeros((n, 4))
    d[:, :3] = rng.normal(size=(n, 3))
    d[:10, :3] = [0., 1., 0.]
    d[10:20, :3] = [0., -2.5, 0.]
    ang = rng.uniform(0, 2 * np.pi, n)
    arrays['polvec_dir'], arrays['polvec_angle'] = d, ang
    arrays['polvec_out'] = polarization_vectors(d, ang)
    # cone source with a tabulated spectrum and random polarization
    spec = QTable({'energy': x * u.keV, 'fluxdensity': pdf / u.s / u.cm**2 / u.keV})
    cone = LabPointSourceCone(position=[200., 3., -2.], direction=[-1., 0.2, 0.1], half_opening=0.02,
                              flux=100. / u.s, energy=spec)
    draws = [rng.random(n) for _ in range(5)]      # energy bin, energy in bin, polangle, theta, v
    with SeqFeeder(draws):
        p = cone.generate_photons(10. * u.s)
    assert len(p) == n
    for c in ('time', 'energy', 'polangle', 'probability', 'pos', 'dir', 'polarization'):
        arrays['cone_' + c] = np.asarray(p[c].data if hasattr(p[c], 'data') else p[c], dtype=float)
    for k, dr in enumerate(draws):
        arrays['cone_draw{0}'.format(k)] = dr
    # far source behind a rectangular aperture, fixed energy and polarization angle
    far = FarLabPointSource([500., 20., -30.], position=[50., 1., 2.], zoom=[1., 4., 7.],
                            flux=100. / u.s, energy=1.5 * u.keV, polarization=0.7 * u.rad)
    draws = [rng.random(n) for _ in range(2)]
    with SeqFeeder(draws):
        p = far.generate_photons(10. * u.s)
    for c in ('time', 'energy', 'polangle', 'probability', 'pos', 'dir', 'polarization'):
        arrays['far_' + c] = np.asarray(p[c].data if hasattr(p[c], 'data') else p[c], dtype=float)
    arrays['far_pos4d'] = far.pos4d
    for k, dr in enumerate(draws):
        arrays['far_draw{0}'.format(k)] = dr
    save('sources', **arrays)


def case_rowland(marxs, rng):
    """Facet placement on a Rowland torus (design/rowland.py): the Config-3 grating array (561 facets) and
    CCD strip (16), a tilted torus with a partial ring of facets, torus normals and parametrisation."""
    from marxs.design.rowland import RowlandTorus, GratingArrayStructure, RectangularGrid, design_tilted_torus
    from marxs.optics import CATGrating, FlatDetector, OrderSelector
    from transforms3d.axangles import axangle2mat
    arrays = {}
    rt = RowlandTorus(6000., 6000.)
    gas = GratingArrayStructure(rowland=rt, d_element=[30., 30.], radius=[300., 500.], elem_class=CATGrating,
                                elem_args={'d': 2e-4, 'zoom': [1, 13.5, 13.5], 'order_selector': OrderSelector([0]),
                                           'orientation': axangle2mat([0, 0, 1], np.deg2rad(1.91))})
    arrays['gas_pos4d'] = np.array([e.pos4d for e in gas.elements])
    arrays['gas_elem_pos'] = np.array(gas.elem_pos)
    det = RectangularGrid(rowland=rt, d_element=[49.652, 49.652], y_range=[-50, 700], elem_class=FlatDetector,
                          elem_args={'zoom': [1, 24.576, 12.288], 'pixsize': 0.024}, id_col='CCD_ID', guess_distance=25.)
    arrays['det_pos4d'] = np.array([e.pos4d for e in det.elements])
    R, r, pos4d = design_tilted_torus(9e3, np.deg2rad(3.8), np.deg2rad(7.6))
    arrays['tilt_Rr'] = np.array([R, r])
    arrays['tilt_pos4d'] = pos4d
    rt2 = RowlandTorus(R, r, pos4d=pos4d)
    gas2 = GratingArrayStructure(rowland=rt2, d_element=[25., 40.], radius=[200., 420.], phi=[-0.5 + np.pi / 2, 0.5 + np.pi / 2],
                                 elem_class=CATGrating, normal_spec=np.array([0, 0, 0, 1.]),
                                 elem_args={'d': 2e-4, 'zoom': [1, 10., 18.], 'order_selector': OrderSelector([0])})
    arrays['gas2_pos4d'] = np.array([e.pos4d for e in gas2.elements])
    th, ph = rng.uniform(-3, 3, 50), rng.uniform(-3, 3, 50)
    pts = rt2.parametric(th, ph)
    arrays['par_theta'], arrays['par_phi'], arrays['par_xyzw'] = th, ph, pts
    arrays['par_normal'] = rt2.normal(pts)
    t2, p2 = rt2.xyzw2parametric(pts)
    arrays['par_theta_back'], arrays['par_phi_back'] = t2, p2
    arrays['quartic_offsurface'] = rt2.quartic(pts[:, :3] + 1.)
    save('rowland', **arrays)


def case_tolerancing(marxs, rng):
    """Tolerancing drivers (design/tolerancing.py, design/uncertainties.py) and the figure-of-merit class they
    are used with (analysis/gratings.py CaptureResAeff): facet pos4d after each mutator (np.random seeded for
    the Gaussian ones), the 6-dof parameter lists, CaptureResAeff on a synthetic event list, and a full
    run_tolerances loop (moveglobal of a grating array in front of a detector, single-order selector so the
    trace is deterministic)."""
    import astropy.units as u
    from astropy.table import Table
    from marxs.design import RowlandTorus, GratingArrayStructure
    from marxs.design import tolerancing as tol
    from marxs.optics import FlatGrating, OrderSelector, FlatDetector
    from marxs.simulator import Sequence
    from marxs.analysis.gratings import CaptureResAeff, CaptureResAeff_CCDgaps, resolvingpower_from_photonlist
    arrays = {}
    torus = RowlandTorus(0.5, 0.5, position=[1.5, 0, -3])        # design/tests/test_tolerancing.py:36-48

    def gsa():
        return GratingArrayStructure(rowland=torus, d_element=[0.1, 0.1], radius=[0.1, .2], elem_class=FlatGrating,
                                     elem_args={'zoom': 0.2, 'd': 0.002, 'order_selector': OrderSelector([1])})

    def stack(g):
        return np.array([e.pos4d for e in g.elements])
    arrays['gsa_pos4d'] = stack(gsa())
    g = gsa()
    np.random.seed(1234)
    tol.wiggle(g, dx=0.01, dy=0.02, dz=0.03, rx=0.01, ry=0.02, rz=0.03)
    arrays['wiggle_pos4d'] = stack(g)
    g = gsa()
    tol.moveglobal(g, dx=0.1, dy=-0.2, dz=0.3, rx=0.1, ry=-0.2, rz=0.3)
    arrays['moveglobal_pos4d'] = stack(g)
    g = gsa()
    tol.moveindividual(g, dx=0.1, dy=-0.2, dz=0.3, rx=0.1, ry=-0.2, rz=0.3)
    arrays['moveindividual_pos4d'] = stack(g)
    det = FlatDetector(zoom=[1, 100, 50], position=[3., 2., 1.])
    tol.moveelem(det, dx=1., dz=5., ry=0.3)
    tol.moveelem(det, dy=2., rx=-0.2, rz=0.1)                       # relative to pos4d_orig, not cumulative
    arrays['moveelem_pos4d'] = det.geometry.pos4d
    np.random.seed(99)
    g = gsa()
    tol.varyperiod(g.elements, 2e-3, 1e-4)
    arrays['varyperiod_d'] = np.array([e._d for e in g.elements])
    cglob, cind = tol.generate_6d_wigglelist([0., .1, .2, .4] * u.cm, [0., 2., 5., 10.] * u.arcmin)
    names = ['dx', 'dy', 'dz', 'rx', 'ry', 'rz']
    arrays['wigglelist_global'] = np.array([[d[k] for k in names] for d in cglob])
    arrays['wigglelist_individual'] = np.array([[d[k] for k in names] for d in cind])
    # figure of merit on a synthetic event list
    n = 6000
    order = rng.integers(-3, 4, n).astype(float)
    order[rng.random(n) < 0.1] = np.nan
    det_x = order * 12.5 + rng.normal(0, 0.05, n) + (rng.random(n) < 0.02) * rng.normal(0, 3., n)
    det_x[rng.random(n) < 0.05] = np.nan
    prob = rng.random(n) * (rng.random(n) > 0.1)
    ccd = rng.integers(-1, 4, n)
    ev = Table({'order': order, 'det_x': det_x, 'probability': prob, 'CCD_ID': ccd})
    arrays.update(ev_order=order, ev_det_x=det_x, ev_probability=prob, ev_ccd=ccd)
    orders = np.array([-3, -2, -1, 0, 1, 2, 3, 7])
    arrays['ev_orders'] = orders
    for tag, cap in (('cap', CaptureResAeff(A_geom=5., orders=orders)),
                     ('capz', CaptureResAeff(A_geom=5., orders=orders, zeropos=0.3)),
                     ('gaps', CaptureResAeff_CCDgaps(A_geom=5., orders=orders))):
        out = cap(ev, n_photons=8000)
        for k, v in out.items():
            arrays[tag + '_' + k] = np.ma.filled(np.ma.asarray(v, dtype=float), np.nan)
    ok = np.isfinite(det_x)
    res, pos, std = resolvingpower_from_photonlist(ev[ok], orders, col='det_x')
    arrays.update(rp_res=res, rp_pos=pos, rp_std=std)
    # a tolerancing loop: 40 gratings 12 m from a detector, moved as a whole
    n = 20000
    p = make_photons(rng, n, spread=0.0, lateral=1., x0=13000., e_lo=0.5, e_hi=0.5)
    p['pos'][:, 1:3] = rng.uniform(-250., 250., (n, 2))
    focus = np.array([0., 0., 0.])
    d = focus - p['pos'][:, :3]
    p['dir'][:, :3] = d / np.linalg.norm(d, axis=1)[:, None]
    p['probability'] = 1.
    pos = [[12000., y, z] for y in np.arange(-200, 201, 100.) for z in np.arange(-200, 201, 100.)]
    from marxs.simulator import Parallel
    gratings = Parallel(elem_class=FlatGrating, elem_pos={'position': pos}, id_col='facet',
                        elem_args={'d': 2e-3, 'zoom': [1, 40., 40.], 'order_selector': OrderSelector([1])})
    detector = FlatDetector(zoom=[1, 500, 500], pixsize=0.024)
    instrum = Sequence(elements=[gratings, detector])
    pars = [{'dx': 0., 'rz': 0.}, {'dx': 30., 'rz': 0.}, {'dx': 0., 'rz': 0.002}, {'dx': -50., 'rz': -0.004, 'ry': 0.01}]
    cap = CaptureResAeff(A_geom=2., orders=np.array([0, 1]), dispersion_coord='det_x', zeropos=0.)
    out = tol.run_tolerances(p, instrum, tol.moveglobal, gratings, pars, cap)
    arrays.update({'loop_' + k: np.asarray(p[k]) for k in ('pos', 'dir', 'energy', 'polarization', 'probability')})
    arrays['loop_grating_pos'] = np.array(pos)
    arrays['loop_pars'] = np.array([[q.get(k, 0.) for k in names] for q in pars])
    for k in ('Aeff0', 'Aeffgrat', 'Aeff', 'Rgrat', 'R'):
        arrays['loop_' + k] = np.array([np.ma.filled(np.ma.asarray(o[k], dtype=float), np.nan) for o in out])
    save('tolerancing', **arrays)


def case_lens_reflectivity(marxs, rng):
    """PerfectLens with a reflectivity_interpolator (mirror.py:31-36,68-81): the reference's own
    RectBivariateSpline(kx=ky=1), queries inside the table, clamped below / above it, and strongly
    off-axis rays (angle / 4 up to ~0.1 rad)."""
    from marxs.optics import PerfectLens
    from scipy.interpolate import RectBivariateSpline
    n = 2000
    egrid = np.concatenate([np.linspace(0.4, 2., 9), np.geomspace(2.5, 7., 6)])     # non-uniform on purpose
    agrid = np.linspace(0.002, 0.08, 23)
    eg, ag = np.meshgrid(egrid, agrid, indexing='ij')
    refl = np.exp(-ag * eg * 9.) * (0.6 + 0.4 * np.cos(eg))
    refl = np.clip(refl, 0., 1.)
    interp = RectBivariateSpline(egrid, agrid, refl, kx=1, ky=1)
    p = make_photons(rng, n, spread=0.25, x0=300., lateral=40., e_lo=0.2, e_hi=9.)
    pos4d = rand_pos4d(rng, zoom=(1., 80., 80.), shift=2.)
    lens = PerfectLens(focallength=150., d_center_optical_axis=-3., pos4d=pos4d, reflectivity_interpolator=interp)
    inp = inputs_dict(p)
    out = lens(p)
    arrays = {'refl_' + k: v for k, v in inp.items()}
    arrays.update({'refl_' + k: v for k, v in table_to_dict(out).items()})
    arrays.update(refl_pos4d=pos4d, refl_egrid=egrid, refl_agrid=agrid, refl_table=refl)
    save('lens_reflectivity', **arrays)


def chirp_d(intercoos):
    """Grating constant varying over the facet (grating.py:209-220); works on numpy arrays and torch tensors."""
    return 2e-4 * (1. + 0.02 * intercoos[:, 0] + 0.001 * intercoos[:, 1] ** 2)


def case_grating_callable_d(marxs, rng):
    """FlatGrating / CATGrating with a callable grating constant d(intercoos) (grating.py:209-220, the
    reference's test_grating_d_callable), orders -2..2."""
    from marxs.optics import FlatGrating, CATGrating, OrderSelector
    n = 2000
    arrays = {}
    for tag, cls in (('flat', FlatGrating), ('cat', CATGrating)):
        p = make_photons(rng, n, spread=0.2, x0=60., lateral=8.)
        pos4d = rand_pos4d(rng, zoom=(1., 9., 6.), shift=1.)
        g = cls(d=chirp_d, order_selector=OrderSelector(np.arange(-2, 3)), pos4d=pos4d, groove_angle=0.3)
        g._slots = [0]
        u = rng.random(n)
        inp = inputs_dict(p)
        with Injector(marxs, [u]):
            out = g(p)
        arrays.update({'%s_%s' % (tag, k): v for k, v in inp.items()})
        arrays.update({'%s_%s' % (tag, k): v for k, v in table_to_dict(out).items()})
        arrays[tag + '_pos4d'] = pos4d
        arrays[tag + '_u'] = u
    save('grating_callable_d', **arrays)


def case_scatter_callable(marxs, rng):
    """RandomGaussianScatter with a callable scatter (scatter.py:127-129; the reference's
    test_scatteredfunction): the angle depends on energy and on where the photon hits."""
    from marxs.optics import RandomGaussianScatter
    import astropy.units as u
    n = 2000

    def anglefunc(photons, intersect, interpos, intercoos):
        return (2e-3 * np.asarray(photons['energy'])[intersect] * np.sin(3. * intercoos[intersect, 0])) * u.rad

    p = make_photons(rng, n, spread=0.1, x0=50., lateral=20.)
    pos4d = rand_pos4d(rng, zoom=(1., 30., 30.), shift=1.)
    sg = RandomGaussianScatter(scatter=anglefunc, pos4d=pos4d)
    sg._slots = [0]
    draws = [rng.random(n)]
    inp = inputs_dict(p)
    with Injector(marxs, draws):
        out = sg(p)
    arrays = {'cs_' + k: v for k, v in inp.items()}
    arrays.update({'cs_' + k: v for k, v in table_to_dict(out).items()})
    arrays['cs_pos4d'] = pos4d
    arrays['cs_u'] = draws[0]
    save('scatter_callable', **arrays)


def main():
    marxs = tier_r.load_reference()
    import marxs.missions.chandra  # noqa: F401
    import marxs.missions.mitsnl.catgrating  # noqa: F401
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])      # python oracle/gen_golden.py [case_name ...]: regenerate only those
    for i, case in enumerate([case_intersect, case_parallel_transport, case_gratings,
                              case_order_selectors, case_lens_scatter, case_detectors,
                              case_mlmirror, case_apertures_baffle, case_chandra,
                              case_parallel_overlap, case_cat_stack, case_cylinder, case_sources, case_rowland, case_tolerancing,
                              case_lens_reflectivity, case_grating_callable_d, case_scatter_callable, case_euklid_bases, case_pt_matrix]):
        rng = np.random.Generator(np.random.PCG64(SEED + i))
        if only and case.__name__ not in only:
            continue
        case(marxs, rng)


if __name__ == '__main__':
    main()
