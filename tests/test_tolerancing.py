"""marxs_b200.design.tolerancing / uncertainties and the figure-of-merit classes of marxs_b200.analysis against
outputs of the unmodified reference (tests/golden/tolerancing.npz, oracle/gen_golden.py case_tolerancing) and
the known answers of the reference's own tests (design/tests/test_tolerancing.py)."""
import os

import numpy as np
import pytest
import torch

from marxs_b200 import optics, simulator, analysis, affines
from marxs_b200.design import RowlandTorus, GratingArrayStructure
from marxs_b200.design import tolerancing as tol
from marxs_b200.design.uncertainties import generate_facet_uncertainty

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'tolerancing.npz'))
NAMES = ['dx', 'dy', 'dz', 'rx', 'ry', 'rz']
torus = RowlandTorus(0.5, 0.5, position=[1.5, 0, -3])


def gsa(elem_class=optics.FlatGrating):
    return GratingArrayStructure(rowland=torus, d_element=[0.1, 0.1], radius=[0.1, .2], elem_class=elem_class,
                                 elem_args={'zoom': 0.2, 'd': 0.002, 'order_selector': optics.OrderSelector([1])})


def stack(g):
    return np.array([e.pos4d for e in g.elements])


def test_euler_convention():
    """'sxyz' = static x, then y, then z: R = Rz Ry Rx."""
    a, b, c = 0.3, -0.7, 1.9
    Rx = affines.axangle2mat([1, 0, 0], a)
    Ry = affines.axangle2mat([0, 1, 0], b)
    Rz = affines.axangle2mat([0, 0, 1], c)
    np.testing.assert_allclose(affines.euler2mat(a, b, c, 'sxyz'), Rz @ Ry @ Rx, atol=1e-15)
    with pytest.raises(NotImplementedError):
        affines.euler2mat(a, b, c, 'rzyx')


def test_mutators_match_the_reference():
    np.testing.assert_allclose(stack(gsa()), GOLD['gsa_pos4d'], rtol=0, atol=1e-10)
    g = gsa()
    np.random.seed(1234)                 # same global generator, same draw order as the reference
    tol.wiggle(g, dx=0.01, dy=0.02, dz=0.03, rx=0.01, ry=0.02, rz=0.03)
    np.testing.assert_allclose(stack(g), GOLD['wiggle_pos4d'], rtol=0, atol=1e-10)
    g = gsa()
    tol.moveglobal(g, dx=0.1, dy=-0.2, dz=0.3, rx=0.1, ry=-0.2, rz=0.3)
    np.testing.assert_allclose(stack(g), GOLD['moveglobal_pos4d'], rtol=0, atol=1e-10)
    g = gsa()
    tol.moveindividual(g, dx=0.1, dy=-0.2, dz=0.3, rx=0.1, ry=-0.2, rz=0.3)
    np.testing.assert_allclose(stack(g), GOLD['moveindividual_pos4d'], rtol=0, atol=1e-10)
    det = optics.FlatDetector(zoom=[1, 100, 50], position=[3., 2., 1.])
    tol.moveelem(det, dx=1., dz=5., ry=0.3)
    tol.moveelem(det, dy=2., rx=-0.2, rz=0.1)
    np.testing.assert_allclose(det.geometry.pos4d, GOLD['moveelem_pos4d'], rtol=0, atol=1e-12)
    np.random.seed(99)
    g = gsa()
    tol.varyperiod(g.elements, 2e-3, 1e-4)
    np.testing.assert_allclose([e._d for e in g.elements], GOLD['varyperiod_d'], rtol=1e-15)
    assert len(generate_facet_uncertainty(5, [1, 2, 3], [0.1, 0.2, 0.3])) == 5


def test_reference_known_answers():
    """design/tests/test_tolerancing.py:50-217."""
    @tol.oneormoreelements
    def func(a, b, c):
        a.value += 1

    class Hold:
        def __init__(self, value):
            self.value = value
    one, two = Hold(2), [Hold(4), Hold(6)]
    func(one, 2, c=4)
    func(two, 'a', None)
    assert (one.value, two[0].value, two[1].value) == (3, 5, 7)
    ref = stack(gsa())
    for function in (tol.wiggle, tol.moveglobal, tol.moveindividual):
        g = gsa()
        function(g, 0., 0., 0.)
        assert np.all(stack(g) == ref)
        for key in NAMES:
            function(g, **{key: 1.23})
            assert not np.all(stack(g) == ref)
    g1, g2 = gsa(), gsa()
    tol.moveglobal(g1, dy=-20)
    tol.moveindividual(g2, dy=-20)
    assert np.allclose(stack(g1), stack(g2)) and not np.allclose(ref, stack(g2))
    g1, g2 = gsa(), gsa()
    tol.moveglobal(g1, rz=-1, ry=.2)
    tol.moveindividual(g2, rz=-1, ry=.2)
    assert not np.allclose(stack(g1), stack(g2))
    det = optics.FlatDetector(zoom=[1, 100, 100])
    tol.moveelem(det, dz=5)
    assert det.geometry['center'][2] == 5 and np.all(det.geometry.pos4d[:3, :3] == np.eye(3) * [1, 100, 100])
    g = gsa()
    tol.wiggle(g, dx=10, dy=.1)
    diff = ref - stack(g)
    assert np.std(diff[:, 0, 3]) > np.std(diff[:, 1, 3])
    for function in (tol.varyperiod, tol.varyorderselector):
        with pytest.raises(ValueError, match='does not have'):
            function(gsa, 1., 2.)
    with pytest.raises(ValueError, match='does not have'):
        tol.varyattribute(gsa, attributenotpresent=1., notpresenteither=2.)
    g = gsa()
    tol.varyperiod(g.elements, 1., .1)
    periods = [e._d for e in g.elements]
    assert 0.01 < np.std(periods) < 5. and np.mean(periods) > .5
    scat = optics.RadialMirrorScatter(inplanescatter=1e-4, perpplanescatter=1e-5)
    tol.varyattribute(scat, inplanescatter=2e-5, perpplanescatter=3e-3)
    assert scat.inplanescatter == 2e-5 and scat.perpplanescatter == 3e-3


def test_6d_lists_and_selection():
    cglob, cind = tol.generate_6d_wigglelist(np.array([0., .1, .2, .4]) * 10., np.deg2rad(np.array([0., 2., 5., 10.]) / 60.))
    np.testing.assert_allclose(np.array([[d[k] for k in NAMES] for d in cglob]), GOLD['wigglelist_global'], rtol=1e-14)
    np.testing.assert_allclose(np.array([[d[k] for k in NAMES] for d in cind]), GOLD['wigglelist_individual'], rtol=1e-14)
    cglob, cind = tol.generate_6d_wigglelist([0, 10.], np.deg2rad([0., 1.]), names=['x', 'y', 'z', 'rx', 'ry', 'rz'])
    assert len(cind) == 7 and len(cglob) == 13 and set(cind[5].keys()) == {'x', 'y', 'z', 'rx', 'ry', 'rz'}
    tab = tol.ResultTable.from_rows(cglob)
    for col in tab.colnames:
        assert -np.min(tab[col]) == np.max(tab[col])
    with pytest.warns(UserWarning):
        tol.generate_6d_wigglelist([1.], [0., 1.])
    tab = tol.ResultTable(par1=np.array([-1, -1, 0, 0, 0, 1]), par2=np.array([-1, 0, 0, 3, 0, 0]), id=np.arange(6))
    t = tol.select_1dof_changed(tab, 'par1', parlist=['par1', 'par2'])
    assert set(t['id']) == {1, 2, 4, 5} and len(t) == 4
    assert set(tol.reset_6d) == set(NAMES) and not any(tol.reset_6d.values())


def _events():
    return {'order': GOLD['ev_order'], 'det_x': GOLD['ev_det_x'], 'probability': GOLD['ev_probability'],
            'CCD_ID': GOLD['ev_ccd']}


class _Events(dict):
    colnames = property(lambda self: list(self.keys()))

    def __len__(self):
        return len(self['order'])


@pytest.mark.gpu
def test_capture_res_aeff_matches_the_reference():
    """(GPU: the statistics behind it are a kernel, mxb_sigma_clip_stats - there is no CPU path)"""
    ev = _Events(_events())
    orders = GOLD['ev_orders']
    for tag, cap in (('cap', analysis.CaptureResAeff(A_geom=5., orders=orders)),
                     ('capz', analysis.CaptureResAeff(A_geom=5., orders=orders, zeropos=0.3)),
                     ('gaps', analysis.CaptureResAeff_CCDgaps(A_geom=5., orders=orders))):
        out = cap(ev, n_photons=8000)
        for k in ('Aeff0', 'Aeffgrat', 'Aeff', 'Rgrat', 'R'):
            np.testing.assert_allclose(np.ma.filled(np.ma.asarray(out[k], dtype=float), np.nan), GOLD[tag + '_' + k],
                                       rtol=1e-10, equal_nan=True, err_msg=tag + ' ' + k)
    ok = torch.isfinite(torch.as_tensor(ev['det_x']))
    res, pos, std = analysis.resolvingpower_from_photonlist(ev, orders, col='det_x', ind=ok)
    np.testing.assert_allclose(res, GOLD['rp_res'], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(pos, GOLD['rp_pos'], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(std, GOLD['rp_std'], rtol=1e-10, equal_nan=True)
    # too few zero-order photons
    few = _Events({k: v[:15] for k, v in _events().items()})
    with pytest.raises(analysis.AnalysisError):
        analysis.resolvingpower_from_photonlist(few, orders, col='det_x')
    assert np.isnan(analysis.CaptureResAeff(orders=orders)(few)['R']).all()
    # robust variant = worst case over the lists
    r2, p2, s2 = analysis.resolvingpower_from_photonlist_robust([ev, ev], orders, ['det_x', 'det_x'], [None, 0.3])
    both = np.array([analysis.resolvingpower_from_photonlist(ev, orders, col='det_x', zeropos=z)[0] for z in (None, 0.3)])
    np.testing.assert_allclose(both[0], GOLD['rp_res'], rtol=1e-10, equal_nan=True)      # NaN det_x never enters the statistics
    np.testing.assert_allclose(r2[:7], np.min(both, axis=0)[:7], rtol=1e-10)
    with pytest.raises(ValueError):
        analysis.resolvingpower_from_photonlist_robust([ev], orders, ['det_x', 'det_x'], [None, 0.3])


def test_analysis_host_helpers():
    res_avg, aeff_sum = analysis.average_R_Aeff(np.array([1., np.nan, 3.]), np.array([1., 5., 3.]))
    assert aeff_sum == 9. and np.isclose(res_avg, (1. + 9.) / 4.)
    ang = np.deg2rad(np.array([80., 100., 120., 260., 290., 10.]))
    got = analysis.identify_photon_in_subaperture(ang, np.deg2rad(20.))
    assert got.tolist() == [True, True, False, True, True, False]


def loop_instrument():
    gratings = simulator.Parallel(elem_class=optics.FlatGrating, elem_pos={'position': [list(p) for p in GOLD['loop_grating_pos']]},
                                  id_col='facet', elem_args={'d': 2e-3, 'zoom': [1, 40., 40.],
                                                             'order_selector': optics.OrderSelector([1])})
    with pytest.warns(Warning):
        detector = optics.FlatDetector(zoom=[1, 500, 500], pixsize=0.024)
    return gratings, detector


@pytest.mark.gpu
def test_run_tolerances_matches_the_reference_loop():
    """tolerancing.run_tolerances (moveglobal of a grating array, CaptureResAeff) on the device against the same
    loop run by the unmodified reference; every step is one out-of-place launch and re-uses one kernel."""
    import marxs_b200 as mb
    from marxs_b200 import _lib
    gratings, detector = loop_instrument()
    instrum = simulator.Sequence(elements=[gratings, detector])
    table = {k: GOLD['loop_' + k] for k in ('pos', 'dir', 'energy', 'polarization', 'probability')}
    photons = mb.PhotonBatch(table, device='cuda')
    before = photons.to_numpy()
    pars = [dict((k, v) for k, v in zip(NAMES, row) if v != 0 or k in ('dx', 'rz')) for row in GOLD['loop_pars']]
    cap = analysis.CaptureResAeff(A_geom=2., orders=np.array([0, 1]), dispersion_coord='det_x', zeropos=0.)
    out = tol.run_tolerances(photons, instrum, tol.moveglobal, gratings, pars, cap)
    assert len(out) == 4 and out[3]['ry'] == 0.01 and 'Aeff' not in pars[0]
    for k in ('Aeff0', 'Aeffgrat', 'Aeff', 'Rgrat', 'R'):
        got = np.array([np.ma.filled(np.ma.asarray(o[k], dtype=float), np.nan) for o in out])
        np.testing.assert_allclose(got, GOLD['loop_' + k], rtol=1e-7, equal_nan=True, err_msg=k)
    after = photons.to_numpy()
    for k in before:                                               # the input list is left untouched
        np.testing.assert_array_equal(before[k], after[k])
    tab = tol.ResultTable.from_rows(out)
    assert len(tab) == 4 and tab['R'].shape == (4, 2)


@pytest.mark.gpu
def test_run_tolerances_for_energies():
    """design/tests/test_tolerancing.py:222-297 with the elements this engine lowers: source -> pointing -> aperture
    -> lens -> grating -> detector, the grating period varied, two energies."""
    from marxs_b200 import source as msource
    coords = (12., -45.)
    src = msource.PointSource(coords=coords)
    pnt = msource.FixedPointing(coords=coords)
    aper = optics.RectangleAperture(position=[5000, 0, 0], zoom=[1, 10, 10])
    lens = optics.PerfectLens(position=[4900, 0, 0], zoom=[1, 10, 10], focallength=4900)
    grat = optics.FlatGrating(d=.002, order_selector=optics.OrderSelector([0, 1]), position=[4800, 0, 0], zoom=[1, 10, 10])
    det = optics.FlatDetector(zoom=[1, 100, 100])
    parameters = [{'period_mean': 0.003, 'period_sigma': 0.}, {'period_mean': 0.004, 'period_sigma': 0.}]
    for variant in (1, 2):
        if variant == 1:
            res = tol.run_tolerances_for_energies(src, [.1, 1], simulator.Sequence(elements=[pnt, aper, lens]),
                                                  simulator.Sequence(elements=[grat, det]), tol.varyperiod, grat, parameters,
                                                  analysis.CaptureResAeff(orders=[0, 1, 2]),
                                                  reset={'period_mean': 0.005, 'period_sigma': 0.}, t_source=1000.)
        else:
            instrum = simulator.Sequence(elements=[pnt, aper, lens, grat, det])
            res = tol.run_tolerances_for_energies2(src, [.1, 1], instrum, optics.FlatGrating, tol.varyperiod, parameters,
                                                   analysis.CaptureResAeff(orders=[0, 1, 2]),
                                                   reset={'period_mean': 0.005, 'period_sigma': 0.}, t_source=1000.)
        assert grat._d == 0.005
        assert 1 in res['energy'] and .1 in res['energy'] and len(res) == 4
        assert np.allclose(res['wave'], 12.398419843320026 / res['energy'])
        R = np.asarray(res['R'])
        assert np.all(R[:, 0] == 0) and not np.any(np.isfinite(R[:, 2])) and R[2, 1] > R[0, 1]
    with pytest.raises(Exception, match='not part of'):
        tol.run_tolerances_for_energies2(src, [1.], instrum, optics.CATGrating, tol.varyperiod, parameters,
                                         analysis.CaptureResAeff(orders=[0, 1]))
