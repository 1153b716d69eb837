"""CPU tests of the host side: table facade, element API / error behaviour,
lowering (program encoding), culling-grid conservativeness, slot order."""
import os

import numpy as np
import pytest
import torch

import marxs_b200 as mb
from marxs_b200 import optics, simulator, program, affines
from marxs_b200.base import _parse_position_keywords
from marxs_b200.missions import chandra
from marxs_b200.missions.mitsnl import InterpolateEfficiencyTable, NonParallelCATGrating
from marxs_b200.program import Lowering, OP, COL_INIT
from oracle import marxs_oracle as mo

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CORE = ['pos', 'dir', 'polarization', 'energy', 'probability']


# ---- photon table facade -------------------------------------------------------
def test_photonbatch_table_protocol():
    p = mb.generate_test_photons(5, device='cpu')
    assert p.colnames == ['pos', 'dir', 'energy', 'polarization', 'probability'] and len(p) == 5
    assert p['pos'].shape == (5, 4) and p.storage('pos').shape == (4, 5)
    p['pos'][:, 3] = 2.                       # writes through the (N, 4) view
    assert torch.all(p.storage('pos')[3] == 2.)
    mask = torch.tensor([True, False, True, False, False])
    p['probability'][mask] = 0.               # reference idiom (catgrating.py:320-323)
    assert p['probability'].tolist() == [0., 1., 0., 1., 1.]
    p['pos'] = p['pos'] + 3. * p['dir']       # propagate idiom (simulator.py:640)
    assert p['pos'][0, 0] == -2.
    p['energy'] = 2                           # scalar broadcast (tests/test_simulator.py:12-18)
    assert torch.all(p['energy'] == 2.)
    q = p.copy()
    q['energy'][0] = 5.
    assert p['energy'][0] == 2.               # copy is deep
    sub = p[mask]
    assert len(sub) == 2 and sub['pos'].shape == (2, 4)
    p.meta['X'] = 1
    assert p.copy().meta['X'] == 1
    assert float(p['energy'].mean()) == 2. and p['energy'].copy() is not p['energy']
    out = p.to_numpy()
    assert out['dir'].shape == (5, 4) and out['dir'].flags['C_CONTIGUOUS']
    with pytest.raises(ValueError):
        p['bad'] = np.zeros(3)


def test_add_output_cols_defaults():
    p = mb.generate_test_photons(3, device='cpu')
    el = optics.FlatDetector(pixsize=0.1, id_col='CCD', id_num=2)
    el.add_output_cols(p, ['a'])
    assert torch.isnan(p['a']).all() and p['CCD'].dtype == torch.int64 and torch.all(p['CCD'] == -1)


# ---- constructor / error behaviour of the reference -----------------------------
def test_position_keywords():
    with pytest.raises(ValueError):
        _parse_position_keywords({'zoom': [1., 0., 1.]})
    with pytest.raises(ValueError):
        _parse_position_keywords({'pos4d': np.eye(4), 'position': [1, 2, 3]})
    with pytest.raises(ValueError):
        _parse_position_keywords({'pos4d': np.zeros((4, 4))})
    p = _parse_position_keywords({'position': [1, 2, 3], 'zoom': 2.})
    assert np.allclose(p, mo.compose([1, 2, 3], np.eye(3), [2, 2, 2]))
    with pytest.raises(ValueError, match='not understood'):
        optics.FlatDetector(pixsize=1., nonsense=3)
    with pytest.raises(ValueError, match='Grating constant'):
        optics.FlatGrating(order_selector=optics.OrderSelector([0]))
    with pytest.raises(ValueError):
        optics.OrderSelector([0, 1], p=[0.7, 0.7])
    with pytest.raises(simulator.SimulationSetupError):
        simulator.Sequence(elements=[5])


def test_affines_roundtrip_and_conventions():
    """math/tests/test_utils.py:7-28, simulator/tests/test_parallel.py:87-99 conventions."""
    rng = np.random.default_rng(0)
    R = affines.axangle2mat([1, 2, 3], 0.7)
    A = affines.compose([1., -2., 3.], R, [2., 3., 4.])
    T, R2, Z, S = affines.decompose44(A)
    assert np.allclose(T, [1, -2, 3]) and np.allclose(R, R2) and np.allclose(Z, [2, 3, 4]) and np.allclose(S, 0)
    T, R2, Z, S = mo.decompose44(A)
    assert np.allclose(Z, [2, 3, 4])
    assert np.allclose(affines.ex2vec_fix([0, 0, 1.], [0, 1., 0])[:, 0], [0, 0, 1])


def test_detector_known_answers():
    """optics/tests/test_detector.py:16-31"""
    det = optics.FlatDetector(zoom=100., pixsize=0.5)
    assert det.npix == [400, 400] and det.centerpix == [199.5, 199.5]


def test_parallel_matches_oracle_and_golden():
    pos = [[0, -10.1, -10.1], [0, .1, -10.1], [0, -10.1, .1], [0, .1, .1]]
    det = simulator.Parallel(elem_class=optics.FlatDetector, elem_args={'pixsize': 0.01, 'zoom': 5},
                             elem_pos={'position': pos}, id_col='CCD_ID')
    assert [e.id_num for e in det.elements] == [0, 1, 2, 3]          # test_parallel.py:11-23
    odet = mo.Parallel(mo.FlatDetector, {'position': pos}, {'pixsize': 0.01, 'zoom': 5}, id_col='CCD_ID')
    for a, b in zip(det.elements, odet.elements):
        assert np.array_equal(a.pos4d, b.pos4d)
    g = dict(np.load(os.path.join(GOLD, 'chandra_c2.npz')))
    hetg = chandra.HETG()
    assert len(hetg.elements) == 336
    np.testing.assert_allclose(np.array([e.pos4d for e in hetg.elements]), g['hetg_pos4d'], rtol=1e-14, atol=1e-12)
    np.testing.assert_allclose([e.geometry['groove_angle'] for e in hetg.elements], g['hetg_groove'], rtol=1e-14)
    np.testing.assert_allclose([e._d for e in hetg.elements], g['hetg_d'])
    acis = chandra.ACIS(chips=[4, 5, 6, 7, 8, 9], aimpoint=chandra.AIMPOINTS['ACIS-S'])
    np.testing.assert_allclose(np.array([e.pos4d for e in acis.elements]), g['acis_pos4d'], rtol=1e-14, atol=1e-12)
    assert [e.id_num for e in acis.elements] == list(g['acis_id'])
    # geometry is mutable: uncertainty + generate_elements moves the facets (tolerancing use case)
    before = hetg.elements[5].pos4d.copy()
    hetg.uncertainty = affines.translation2aff([0., 1., 0.])
    hetg.generate_elements()
    assert np.allclose(hetg.elements[5].pos4d[:3, 3] - before[:3, 3], [0., 1., 0.])


def test_hetg_orientation_known_answers():
    """missions/chandra/tests/test_hetg.py:6-38"""
    hetg = chandra.HETG()
    h = hetg.hess
    for i, e in enumerate(hetg.elements):
        ex, ey, ez = (e.geometry[k][:3] for k in ('e_x', 'e_y', 'e_z'))
        assert abs(np.dot([h['xu'][i], h['yu'][i], h['zu'][i]], ex)) > 0.99
        assert abs(np.dot([h['xuxf'][i], h['yuxf'][i], h['zuxf'][i]], ey)) > 0.99
        assert abs(np.dot([h['xuyf'][i], h['yuyf'][i], h['zuyf'][i]], ez)) > 0.99
        a, c, n = e.e_groove_coos(np.zeros((1, 2)))
        assert abs(np.dot(a[0, :3], [h['xul'][i], h['yul'][i], h['zul'][i]])) > 0.99
        assert abs(np.dot(c[0, :3], [h['xud'][i], h['yud'][i], h['zud'][i]])) > 0.99


# ---- lowering -----------------------------------------------------------------------
def lower(elems, cols=CORE, meta=None):
    lw = Lowering(cols, meta=meta or {'ROLL_PNT': (0., '')})
    for e in elems:
        e._lower(lw)
    return lw, lw.finish()


def test_program_encoding_c2():
    lw, prog = lower([chandra.HRMA(), chandra.HETG(),
                      chandra.ACIS(chips=[4, 5, 6, 7, 8, 9], aimpoint=chandra.AIMPOINTS['ACIS-S'])])
    b = prog.blob
    assert int(b[0]) == program.MXB_MAGIC and int(b[2]) == prog.n_ops == 11 and int(b[3]) == b.size
    assert prog.stage_words * 8 < program.MAX_STAGE_BYTES and prog.stage_words % 2 == 0
    types = [o['type'] for o in prog.ops]
    # the COMMIT of each element is folded into its own op (flag 256) except after ACIS (8 columns used)
    assert types == [OP['PLANE'], OP['LENS'], OP['RSCATTER'], OP['FILTER'], OP['ARRAY_BEGIN'], OP['GRATING'],
                     OP['ARRAY_END'], OP['ARRAY_BEGIN'], OP['ACIS'], OP['COMMIT'], OP['ARRAY_END']]
    assert all(prog.ops[k]['flags'] & 256 for k in (0, 1, 2, 3, 5))
    # the 21 diagnostic columns of the reference
    assert len(prog.out_f64) + len(prog.out_i64) == 21
    assert prog.slot_kinds == ['normal', 'normal', 'uniform']
    # same slot order as the oracle
    orac = mo.Sequence([mo.chandra_hrma()])
    assert mo.assign_slots(orac) == ['normal', 'normal']
    begin = prog.ops[4]
    F, stride, rows_off = begin['cols'][:3]
    assert F == 336 and (stride // 2) % 2 == 1 and begin['w14'] == 6
    rows = b[rows_off:rows_off + F * stride].reshape(F, stride)
    np.testing.assert_array_equal(rows[:, :14], np.array([program.geom14(e.pos4d) for e in chandra.HETG().elements]))
    assert list(rows[:, prog.ops[5]['w15']].astype(int)) == list(range(336))       # facet id_num per row
    # the array header describes the SAME grid the cell lists were built for (the certificate makes the lowering
    # rebuild the grid with a smaller margin: frame, origin, cell size and lists must come from one build), and the
    # blob's lists are that grid's
    for pc, n_fac in ((4, 336), (7, 6)):
        o = prog.ops[pc]
        H = b[o['pg']:o['pg'] + program.ARRAY_HEADER_WORDS]
        G = b[o['cols'][2]:o['cols'][2] + n_fac * o['cols'][1]].reshape(n_fac, o['cols'][1])[:, :14]
        ref_grid = program.build_cull_grid(G, hops=0)           # both arrays are certified beyond the culling cone
        assert H[19] == 1. and H[17] > H[15] == ref_grid['T2']
        assert (o['cols'][4], o['cols'][5]) == (ref_grid['nu'], ref_grid['nv'])
        np.testing.assert_array_equal(H[12:15], [ref_grid['u0'], ref_grid['v0'], ref_grid['inv_cell']])
        start = b[o['cols'][6]:o['cols'][6] + (ref_grid['nu'] * ref_grid['nv'] + 2) // 2].view(np.int32)[:ref_grid['nu'] * ref_grid['nv'] + 1]
        np.testing.assert_array_equal(start, ref_grid['start'])
        cand = b[o['cols'][7]:o['cols'][7] + (len(ref_grid['cand']) + 1) // 2].view(np.int32)[:len(ref_grid['cand'])]
        np.testing.assert_array_equal(cand, ref_grid['cand'])
    # 'y' is first written (and initialised) by the HRMA stack, later overwritten by ACIS sky y
    ycol = prog.out_f64.index('y') + program.FIRST_OUT
    assert prog.ops[0]['cols'][5] == ycol + COL_INIT and prog.ops[8]['cols'][7] == ycol
    # a second run on a table that already has the columns initialises nothing
    lw2, prog2 = lower([chandra.HRMA()], cols=CORE + prog.out_f64)
    assert all(c < COL_INIT for o in prog2.ops for c in o['cols'])


def test_parallel_with_per_facet_args_and_heterogeneous_fallback():
    pos = {'position': [[0., -4., 0.], [-3., 2., 1.], [2., 4., -3.]]}
    sel = optics.OrderSelector([-1, 0, 1])
    par = simulator.Parallel(elem_class=optics.FlatGrating, elem_pos=pos, id_col='facet',
                             elem_args={'d': [2e-4, 3e-4, 2.5e-4], 'zoom': [1, 5., 6.], 'order_selector': sel})
    lw, prog = lower([par])
    assert [o['type'] for o in prog.ops] == [OP['ARRAY_BEGIN'], OP['GRATING'], OP['ARRAY_END']]
    assert len(prog.slot_kinds) == 1            # one slot for the whole array, not per facet
    # different selectors per facet cannot share one body: NotFusable -> caller falls back
    par2 = simulator.Parallel(elem_class=optics.FlatGrating, elem_pos=pos,
                              elem_args={'d': 2e-4, 'order_selector': [optics.OrderSelector([0]),
                                                                       optics.OrderSelector([1]),
                                                                       optics.OrderSelector([2])]})
    with pytest.raises(program.NotFusable):
        lower([par2])


def test_user_plugin_is_not_lowered():
    """A subclass that overrides the Python hook must not silently run the parent's kernel."""
    class MyGrating(optics.FlatGrating):
        def specific_process_photons(self, photons, intersect, interpos, intercoos):
            return {'probability': torch.ones(int(intersect.sum()), dtype=torch.float64) * 0.5}

    g = MyGrating(d=1e-3, order_selector=optics.OrderSelector([0]))
    assert not g._can_lower()
    assert optics.FlatGrating(d=1e-3, order_selector=optics.OrderSelector([0]))._can_lower()
    with pytest.raises(program.UnsupportedCallable):
        lower([optics.FlatGrating(d=1e-3, order_selector=lambda e, p, b: (0, 1))])
    with pytest.raises(program.UnsupportedCallable):
        lower([optics.FlatGrating(d=lambda ic: 1e-3, order_selector=optics.OrderSelector([0]))])
    f = optics.EnergyFilter(filterfunc=lambda en: en * 0 + 0.5)
    assert not f._can_lower()                   # evaluated with torch on the device instead
    seq = simulator.Sequence(elements=[optics.FlatDetector(pixsize=1.)], postprocess_steps=[simulator.KeepCol('pos')])
    assert not seq._can_lower()                 # KeepCol must see the table after every element


def test_selector_tables():
    sel = optics.OrderSelector(np.arange(-2, 3), p=[.1, .2, .3, .2, .1])
    blk, big = sel.device_table()
    assert big is None and int(blk[0]) == program.SEL_ORDERSELECTOR and int(blk[1]) == 5
    np.testing.assert_allclose(blk[3:8], mo.OrderSelector(np.arange(-2, 3), [.1, .2, .3, .2, .1]).cdf())
    wave, theta = np.array([1., 2., 3.]), np.array([0.01, 0.02])
    prob = np.random.default_rng(0).uniform(0, .1, (3, 2, 4))
    iet = InterpolateEfficiencyTable(wave, theta, prob, [1, 0, -1, -2])
    g = optics.CATGrating(d=2e-4, order_selector=iet)
    lw, prog = lower([g])
    sel_off = [o for o in prog.ops if o['type'] == OP['GRATING']][0]['pg']
    tab_off = int(prog.blob[sel_off + 4])
    assert tab_off >= prog.stage_words                     # the big table stays in global memory
    np.testing.assert_array_equal(prog.blob[tab_off:tab_off + prob.size], prob.ravel())
    tab = {'lambda': np.repeat(wave, 2), 'theta': np.tile(np.rad2deg(theta), 3)}
    for k, o in enumerate([1, 0, -1, -2]):
        tab[str(o)] = prob[:, :, k].ravel()
    iet2 = InterpolateEfficiencyTable.from_table(tab)
    np.testing.assert_allclose(iet2.prob, prob)
    np.testing.assert_allclose(iet2.theta, theta)
    assert NonParallelCATGrating(d=2e-4, order_selector=iet, d_blaze_mm=1e-3)._blaze_modifier() == (0, 1e-3)


# ---- culling grid ---------------------------------------------------------------------
@pytest.mark.parametrize('case', ['hetg', 'random', 'hetg-incoming-only'])
def test_cull_grid_is_conservative(case):
    """Every facet a ray inside the validity cone hits (oracle arithmetic) is listed in its cell (also for the tighter
    grid that only has to serve the incoming ray, hops=0)."""
    rng = np.random.default_rng(5)
    hops = 0 if case == 'hetg-incoming-only' else 2
    if case.startswith('hetg'):
        elems = chandra.HETG().elements
    else:
        F = 60
        elems = [optics.FlatDetector(position=[rng.uniform(-5, 5), rng.uniform(-60, 60), rng.uniform(-60, 60)],
                                     orientation=affines.axangle2mat(rng.normal(size=3), rng.uniform(0, 0.08)),
                                     zoom=[1, rng.uniform(3, 9), rng.uniform(3, 9)]) for _ in range(F)]
    G = np.array([program.geom14(e.pos4d) for e in elems])
    g = program.build_cull_grid(G, hops=hops)
    assert g is not None
    if hops == 0:
        assert g['margin'] < 0.4 * program.build_cull_grid(G)['margin'] and np.isclose(g['T2'], 0.12 ** 2)
    n = 60000
    # rays through random points of the array, directions inside the cone around nbar
    target = G[rng.integers(0, len(G), n), :3] + rng.uniform(-20, 20, (n, 3))
    tmax = np.sqrt(g['T2'])
    t = rng.uniform(0, tmax, n) * 0.999
    phi = rng.uniform(0, 2 * np.pi, n)
    d = -(g['nbar'] + t[:, None] * (np.cos(phi)[:, None] * g['u'] + np.sin(phi)[:, None] * g['v']))
    pos = np.ones((n, 4))
    pos[:, :3] = target - d * rng.uniform(500, 900, n)[:, None]
    dirs = np.zeros((n, 4))
    dirs[:, :3] = d * rng.uniform(0.5, 2, n)[:, None]
    dn = dirs[:, :3] @ g['nbar']
    tt = ((g['O'] - pos[:, :3]) @ g['nbar']) / dn
    q = pos[:, :3] + tt[:, None] * dirs[:, :3] - g['O']
    fu = (q @ g['u'] - g['u0']) * g['inv_cell']
    fv = (q @ g['v'] - g['v0']) * g['inv_cell']
    inside = (fu >= 0) & (fv >= 0) & (fu < g['nu']) & (fv < g['nv'])
    cell = np.where(inside, fv.astype(int) * g['nu'] + fu.astype(int), 0)
    nhit = 0
    for j in range(len(G)):
        hit, _, _ = mo.plane_intersect(mo.PlaneConsts(elems[j].pos4d), dirs, pos)
        for i in np.nonzero(hit)[0]:
            assert inside[i]
            lst = g['cand'][g['start'][cell[i]]:g['start'][cell[i] + 1]]
            assert j in lst
            nhit += 1
    assert nhit > 5000
    assert g['mean_candidates'] < 6


@pytest.mark.parametrize('case', ['hetg', 'random', 'staggered', 'overlapping', 'rowland'])
def test_single_hit_certificate_is_conservative(case):
    """program.single_hit_successors: inside the certified cone a ray that starts ON facet A hits no LATER facet that
    is not in A's successor list (oracle arithmetic, rays in both directions; earlier facets are never revisited by
    the reference's loop).  Tiled arrays get empty lists and a wide cone, overlapping neighbours become successors."""
    rng = np.random.default_rng(7)
    if case == 'hetg':
        elems = chandra.HETG().elements
    elif case == 'random':
        elems = [optics.FlatDetector(position=[rng.uniform(-5, 5), 25. * (k % 8) + rng.uniform(-2, 2), 25. * (k // 8) + rng.uniform(-2, 2)],
                                     orientation=affines.axangle2mat(rng.normal(size=3), rng.uniform(0, 0.08)),
                                     zoom=[1, rng.uniform(3, 9), rng.uniform(3, 9)]) for k in range(48)]
    elif case == 'staggered':      # two layers 30 mm apart, laterally interleaved: small gaps against large height steps
        elems = [optics.FlatDetector(position=[30. * (k % 2), 11. * k, 0.], zoom=[1, 5., 20.]) for k in range(12)]
    elif case == 'overlapping':
        elems = [optics.FlatDetector(position=[2. * k, 8. * k, 0.], zoom=[1, 5., 20.]) for k in range(6)]
    else:                          # ring-placed facets of a Rowland array: diagonal neighbours overlap in projection
        from marxs_b200.design import RowlandTorus, GratingArrayStructure
        gas = GratingArrayStructure(rowland=RowlandTorus(6000., 6000.), d_element=[30., 30.], radius=[300., 420.],
                                    elem_class=optics.FlatGrating,
                                    elem_args={'zoom': [1, 13.5, 13.5], 'd': 2e-4, 'order_selector': optics.OrderSelector([0])})
        elems = gas.elements
    G = np.array([program.geom14(e.pos4d) for e in elems])
    g = program.build_cull_grid(G)
    assert g is not None
    t2, start, succ = program.single_hit_successors(G, g)
    F = len(G)
    assert t2 > 0.
    if case in ('hetg', 'random'):
        assert succ is None and t2 > 0.09                  # tiled: every facet is the last one a photon can hit
    else:
        assert succ is not None and len(start) == F + 1 and start[-1] == len(succ)
        for a in range(F):
            lst = succ[start[a]:start[a + 1]]
            assert np.all(lst > a) and np.all(np.diff(lst) > 0)
        if case == 'overlapping':
            assert [list(succ[start[a]:start[a + 1]]) for a in range(F)][:2] == [[1], [2]]
        if case == 'rowland':
            assert 0 < (np.diff(start) > 0).sum() < F and np.diff(start).max() <= 4
    lists = [set() if succ is None else set(int(x) for x in succ[start[a]:start[a + 1]]) for a in range(F)]
    per = 400 if F < 100 else 120
    a = np.repeat(np.arange(F), per)
    n = len(a)
    ly, lz = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    start_pt = G[a, 0:3] + (ly * G[a, 12])[:, None] * G[a, 6:9] + (lz * G[a, 13])[:, None] * G[a, 9:12]
    t = np.sqrt(t2) * rng.uniform(0, 1, n) ** 0.25 * 0.999999             # biased towards the edge of the cone
    phi = rng.uniform(0, 2 * np.pi, n)
    d = g['nbar'] + t[:, None] * (np.cos(phi)[:, None] * g['u'] + np.sin(phi)[:, None] * g['v'])
    d *= rng.choice([-1., 1.], n)[:, None] * rng.uniform(0.5, 2, n)[:, None]
    pos = np.ones((n, 4))
    pos[:, :3] = start_pt
    dirs = np.zeros((n, 4))
    dirs[:, :3] = d
    later_hits = 0
    for j in range(F):
        hit, _, _ = mo.plane_intersect(mo.PlaneConsts(elems[j].pos4d), dirs, pos)
        for i in np.nonzero(hit & (a < j))[0]:
            assert j in lists[a[i]], (case, int(a[i]), j)
            later_hits += 1
    if case in ('overlapping', 'rowland', 'staggered'):
        assert later_hits > 0          # the lists are exercised, not vacuous


# ---------------------------------------------------------------------------
# later additions: row hoisting, plan-cache digest, sources / pointing host logic
# ---------------------------------------------------------------------------
def test_facet_invariant_blocks_leave_the_rows():
    """Blocks that are identical for every facet (quality factor, L2 dimensions, L2 scatter) are lowered
    once as shared parameters; per-facet blocks (grating vectors) stay in the rows."""
    from marxs_b200 import optics, simulator
    from marxs_b200.missions import mitsnl
    from marxs_b200.program import Lowering, OP
    sel = optics.OrderSelector([0, 1])
    pos = [[0., y, z] for y in (-16., 0., 16.) for z in (-17., 0., 17.)]
    par = simulator.Parallel(elem_class=mitsnl.CATL1L2Stack, elem_pos={'position': pos}, id_col='facet',
                             elem_args=dict(zoom=[1, 7.5, 8.], order_selector=sel))
    lw = Lowering(['pos', 'dir', 'polarization', 'energy', 'probability'])
    par._lower(lw)
    prog = lw.finish()
    by_type = {}
    for o in prog.ops:
        by_type.setdefault(o['type'], []).append(o)
    for t in ('QFACTOR', 'L2ABS', 'GSCATTER'):
        (o,) = by_type[OP[t]]
        assert o['pf'] == -1 and o['pg'] > 0, t
    g0, g1 = by_type[OP['GRATING']]
    assert (g0['pf'], g1['pf']) == (14, 23) and g1['flags'] & 8           # membrane, then L1 support
    F, stride = prog.ops[0]['cols'][0], prog.ops[0]['cols'][1]
    assert F == 9 and stride == 34                                         # 14 geometry + 2 x 9 grating + id, padded
    # a facet with a different quality factor keeps the block in the rows
    par.elements[3].elements[1].factor = 0.5
    lw = Lowering(['pos', 'dir', 'polarization', 'energy', 'probability'])
    par._lower(lw)
    q = [o for o in lw.finish().ops if o['type'] == OP['QFACTOR']][0]
    assert q['pf'] >= 14 and q['pg'] == -1


def test_fingerprint_tracks_what_lowering_reads():
    from marxs_b200 import optics, simulator
    det = optics.FlatDetector(pixsize=0.1, zoom=[1, 20, 20])
    flt = optics.EnergyFilter(filterfunc=lambda e: e * 0 + 0.5, zoom=[1, 20, 20])
    seq = simulator.Sequence(elements=[flt, det])
    f0 = simulator.fingerprint([seq], ('a',))
    assert f0 == simulator.fingerprint([seq], ('a',)) and f0 != simulator.fingerprint([seq], ('b',))
    det.geometry.pos4d[2, 3] += 1e-12
    f1 = simulator.fingerprint([seq], ('a',))
    assert f1 != f0
    det.pixsize = 0.2
    f2 = simulator.fingerprint([seq], ('a',))
    assert f2 != f1
    flt.filterfunc = lambda e: e * 0 + 0.5          # a different function object, same text
    assert simulator.fingerprint([seq], ('a',)) != f2
    det.name = 'renamed'                             # labels do not matter
    flt.filterfunc = seq.elements[0].filterfunc
    assert simulator.fingerprint([seq], ('a',)) == simulator.fingerprint([seq], ('a',))


def test_source_and_pointing_host_logic():
    from marxs_b200 import source
    from marxs_b200.source.source import skyoffset_matrix, RandomArbitraryPdfTable
    for T, rate in ((10., 100.), (10., 2000.), (1., 1e6), (5., 0.7), (3.3, 7.7)):
        s = source.PointSource(coords=(1., 2.), flux=rate)
        assert s.n_photons(T) == len(np.arange(0, T, 1. / rate))
    x = np.array([0.3, 0.5, 0.9, 1.4, 2.2, 3.5, 6.0, 10.])
    pdf = np.array([0., 3., 1., 5., 0.2, 4., 0.05, 1.])
    t = RandomArbitraryPdfTable(x, pdf).table()
    o = mo.RandomArbitraryPdf(x, pdf)
    n = len(x)
    assert t[0] == n and np.array_equal(t[1:1 + n], o.cdf) and np.array_equal(t[1 + n:1 + 2 * n], o.sortindex)
    assert np.array_equal(t[1 + 2 * n:1 + 3 * n], x) and np.array_equal(t[1 + 3 * n:], o.bin_width)
    np.testing.assert_array_equal(skyoffset_matrix(0.3, -0.2, 0.1), mo.skyoffset_matrix(0.3, -0.2, 0.1))
    p = source.JitterPointing(coords=(25., -10.), roll=0.3, jitter=1e-5)._params()
    op = mo.FixedPointing((25., -10.), roll=0.3)
    np.testing.assert_array_equal(p[:9].reshape(3, 3), op.matrix())
    north = op._apply_ref(op.sky_to_offset(np.array([0.]), np.array([90.])))[0, :3]
    np.testing.assert_array_equal(p[18:21], north)
    assert p[21] == 1e-5
    with pytest.raises(source.SourceSpecificationError):
        source.PointSource(coords=(1., 2.), flux=lambda t, a: t).n_photons(1.)


def test_tagversion_writes_meta():
    """base/base.py:100-139: TagVersion adds its keywords, the date and the code version to photons.meta."""
    from marxs_b200.base import TagVersion

    class Table:
        meta = {}
    t = TagVersion(OBSERVER=('me', 'who ran this'))(Table(), OBS_ID=5)
    assert t.meta['OBSERVER'][0] == 'me' and t.meta['OBS_ID'] == 5
    assert t.meta['CREATOR'][0] == 'MARXS' and 'MARXSVER' in t.meta and len(t.meta['DATE'][0]) == 10


def test_check_meta_and_energy_consistent():
    """base/base.py:142-191 semantics."""
    import numpy as np
    from marxs_b200.base import check_meta_consistent, check_energy_consistent
    check_meta_consistent({'ORIGIN': 'a'}, {'ORIGIN': 'a'})
    with pytest.raises(AssertionError):
        check_meta_consistent({'ORIGIN': 'a'}, {'ORIGIN': 'b'})
    with pytest.raises(KeyError):
        check_meta_consistent({'ORIGIN': 'a'}, {})
    with pytest.raises(KeyError):
        check_meta_consistent({}, {}, allow_missing=False)

    class T(dict):
        colnames = ['energy']
    check_energy_consistent(T(energy=np.ones(5)))
    with pytest.raises(AssertionError):
        check_energy_consistent(T(energy=np.array([1., 1.1])))
    T2 = type('T2', (dict,), {'colnames': []})
    check_energy_consistent(T2())


def test_mathutils_numpy_and_torch_agree():
    """marxs/math/utils.py helpers: same numbers for numpy arrays and torch tensors, reference error cases."""
    import numpy as np
    import torch
    from marxs_b200 import mathutils as mu
    rng = np.random.default_rng(0)
    e = rng.normal(size=(7, 3))
    for w in (0, 1):
        h = mu.e2h(e, w)
        assert h.shape == (7, 4) and np.all(h[:, 3] == w)
        assert torch.equal(mu.e2h(torch.tensor(e), w), torch.tensor(h))
        assert np.array_equal(mu.h2e(h), e) and torch.equal(mu.h2e(torch.tensor(h)), torch.tensor(e))
    with pytest.raises(ValueError):
        mu.e2h(e, 2)
    h = mu.e2h(e, 1) * 2.
    assert np.allclose(mu.h2e(h), e) and torch.allclose(mu.h2e(torch.tensor(h)), torch.tensor(e))
    mixed = mu.e2h(e, 1)
    mixed[0, 3] = 0
    with pytest.raises(ValueError):
        mu.h2e(mixed)
    with pytest.raises(ValueError):
        mu.h2e(torch.tensor(mixed))
    p1, p2 = mu.e2h(e, 1), mu.e2h(e[::-1].copy(), 1)
    d = mu.distance_point_point(p1, p2)
    assert np.allclose(d, np.linalg.norm(e - e[::-1], axis=1))
    assert torch.allclose(mu.distance_point_point(torch.tensor(p1), torch.tensor(p2)), torch.tensor(d))
    assert np.isclose(mu.distance_point_point(p1[0], p2[0]), d[0])
    n = mu.norm_vector(e)
    assert np.allclose(np.linalg.norm(n, axis=1), 1) and torch.allclose(mu.norm_vector(torch.tensor(e)), torch.tensor(n))


def test_chip2tdet_matches_reference_known_answers():
    """missions/chandra/data.py:169-190; spot values of the reference's test_chandra.py coordinate table
    (chip centre of ACIS-S3 -> TDET) and numpy / torch agreement."""
    import numpy as np
    import torch
    from marxs_b200.missions.chandra.data import chip2tdet, TDET
    chip = np.array([[512.5, 512.5], [1., 1.], [1024., 1024.]])
    for idn in (0, 1, 7):
        a = chip2tdet(chip, TDET['ACIS'], idn)
        b = chip2tdet(torch.tensor(chip), TDET['ACIS'], idn).numpy()
        assert np.allclose(a, b, rtol=0, atol=1e-9)
    # S3 (id 7): theta = 0, origin (3917, 1702): tdet = chip - 0.5 + origin + 0.5 = chip + origin
    assert np.allclose(chip2tdet(chip, TDET['ACIS'], 7), chip + [3917, 1702])
    # I0 (id 0): theta = 90 deg: (x, y) -> (y, -x) about the chip origin
    assert np.allclose(chip2tdet(np.array([[1., 1.]]), TDET['ACIS'], 0), [[0.5 + 3061.5, -0.5 + 5131.5]])


# ---- round-2 regressions ---------------------------------------------
def test_plan_cache_stores_run_length_not_end_index():
    """The cache key describes the tail elements[i:]; the same tail reached at another start offset used
    to return the first call's ABSOLUTE end index (elements skipped or applied twice, or an endless loop)."""
    p = mb.generate_test_photons(4, device='cpu')
    a = optics.FlatDetector(pixsize=0.1, zoom=[1, 20, 20])
    b = optics.EnergyFilter(filterfunc=0.5, zoom=[1, 20, 20], position=[-1, 0, 0])
    c = optics.FlatDetector(pixsize=0.2, zoom=[1, 20, 20], position=[-2, 0, 0])

    class Opaque:          # a user element that cannot be lowered
        def __call__(self, photons):
            return photons
    simulator._plan_cache.clear()
    j0, prog0, _ = simulator._lower_run([a, b, c], 0, p)
    assert j0 == 3 and prog0 is not None
    hits = simulator.plan_cache_stats['hit']
    j1, prog1, _ = simulator._lower_run([Opaque(), a, b, c], 1, p)
    assert simulator.plan_cache_stats['hit'] == hits + 1 and prog1 is prog0
    assert j1 == 4                                   # was 3: run_fused would have run `c` twice
    j2, _, _ = simulator._lower_run([Opaque(), Opaque(), a, b, c], 2, p)
    assert j2 == 5
    # a run that stops at a non-lowerable element: length 2 at any offset
    simulator._plan_cache.clear()
    tail = [a, b, Opaque(), c]
    assert simulator._lower_run(tail, 0, p)[0] == 2
    assert simulator._lower_run([Opaque()] + tail, 1, p)[0] == 3


def test_fingerprint_tensor_content_pins_and_depth():
    det = optics.FlatDetector(pixsize=0.1, zoom=[1, 20, 20])
    det.table = torch.arange(6, dtype=torch.float64)
    f0 = simulator.fingerprint([det])
    det.table[2] = 7.                                # in-place change of a small tensor: content is hashed
    f1 = simulator.fingerprint([det])
    assert f1 != f0
    big = torch.zeros(simulator._FP_TENSOR_BY_CONTENT + 1, dtype=torch.float64)
    det.table = big
    pins = []
    f2 = simulator.fingerprint([det], pins=pins)
    assert any(x is big for x in pins)               # identity-hashed: kept alive by the cache entry
    big[5] = 1.                                      # torch bumps _version
    assert simulator.fingerprint([det]) != f2
    # output buffers (the fused detector image) are identified by address, not content
    det.table = None
    det.image = torch.zeros((1, 8, 8), dtype=torch.float64)
    f3 = simulator.fingerprint([det])
    det.image.zero_()
    det.image += 1.
    assert simulator.fingerprint([det]) == f3
    # functions are pinned as well
    flt = optics.EnergyFilter(filterfunc=lambda e: e * 0 + 0.5, zoom=[1, 20, 20])
    pins = []
    simulator.fingerprint([flt], pins=pins)
    assert any(x is flt.filterfunc for x in pins)
    # too deep: fail closed instead of hashing by id()
    class Node:
        pass
    root = cur = Node()
    for _ in range(simulator._FP_MAX_DEPTH + 2):
        cur.child = Node()
        cur = cur.child
    with pytest.raises(simulator.Uncacheable):
        simulator.fingerprint([root])
    p = mb.generate_test_photons(4, device='cpu')
    det2 = optics.FlatDetector(pixsize=0.1, zoom=[1, 20, 20])
    det2.deep = root
    n_cached = len(simulator._plan_cache)
    assert simulator._lower_run([det2], 0, p)[0] == 1 and len(simulator._plan_cache) == n_cached   # lowered, not cached


def test_source_rate_converts_area_units():
    """flux is per cm**2; aperture.area is mm**2 (reference aperture.py:100, basesources.py:174)."""
    from marxs_b200 import source
    from marxs_b200.optics.aperture import AreaMM2
    aper = optics.RectangleAperture(zoom=[1, 10, 10])
    assert isinstance(aper.area, AreaMM2) and float(aper.area) == 400. and aper.area.to('cm2') == 4.
    s = source.PointSource(coords=(0., 0.), flux=100., geomarea=aper.area)
    # reference: len(np.arange(0, T, 1 / (flux * geomarea * u.s).decompose())) with 400 mm2 = 4 cm2
    assert s.n_photons(10.) == len(np.arange(0, 10., 1. / (100. * 4.))) == 4000
    assert source.PointSource(coords=(0., 0.), flux=100., geomarea=4.).n_photons(10.) == 4000      # bare number: cm2
    assert source.PointSource(coords=(0., 0.), flux=100.).n_photons(10.) == 1000                    # default 1 cm2
    circ = optics.CircleAperture(zoom=[1, 5, 5], r_inner=3.)
    assert np.isclose(float(circ.area), np.pi * (25. - 9.)) and isinstance(circ.area, AreaMM2)
    multi = optics.MultiAperture(elements=[aper, circ])
    assert isinstance(multi.area, AreaMM2) and np.isclose(float(multi.area), 400. + np.pi * 16.)
    assert isinstance(2 * aper.area, AreaMM2) and isinstance(aper.area + circ.area, AreaMM2)
    assert isinstance(sum([aper.area, circ.area]), AreaMM2)
    # Quantity-like objects (astropy or the test stand-in) are converted, not stripped
    class Q:
        def __init__(self, v, unit):
            self.value, self.unit = v, unit
        def to(self, unit):
            assert self.unit == 'mm2'
            return Q(self.value / 100., 'cm2')
    import sys
    import types
    fake = types.ModuleType('astropy.units')
    class U:
        def __init__(self, name): self.name = name
        def __pow__(self, k): return U('{0}{1}'.format(self.name, k))
        def __rtruediv__(self, other): return U('1/' + self.name)
        def __truediv__(self, other): return U(self.name + '/' + other.name)
    fake.cm, fake.s = U('cm'), U('s')
    pkg = types.ModuleType('astropy')
    pkg.units = fake
    saved = {k: sys.modules.get(k) for k in ('astropy', 'astropy.units')}
    sys.modules['astropy'], sys.modules['astropy.units'] = pkg, fake
    try:
        s = source.PointSource(coords=(0., 0.), flux=100., geomarea=Q(300., 'mm2'))
        assert s.n_photons(10.) == 3000
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_filter_fill_modes_follow_scipy():
    """interp1d(bounds_error=False) returns fill_value (NaN by default) or extrapolates outside the table;
    the lowering used to treat every such object as clamped."""
    from scipy.interpolate import interp1d
    from marxs_b200.optics.filter import lower_filter
    x, y = np.array([1., 2., 4.]), np.array([0.2, 0.6, 0.4])
    q = np.array([0.5, 1., 1.5, 3., 4., 5.])
    cases = [dict(bounds_error=False), dict(bounds_error=False, fill_value=0.1),
             dict(bounds_error=False, fill_value=(0.1, 0.3)), dict(bounds_error=False, fill_value='extrapolate')]
    for kw in cases:
        want = interp1d(x, y, **kw)(q)
        assert np.allclose(mo.Tabulated1D(x, y, **kw)(q), want, rtol=1e-15, atol=1e-16, equal_nan=True), kw
    params, flags = lower_filter(interp1d(x, y, bounds_error=False))
    assert flags == 2 and np.isnan(params[-2:]).all()
    params, flags = lower_filter(interp1d(x, y, bounds_error=False, fill_value=(0.1, 0.3)))
    assert flags == 2 and params[-2:].tolist() == [0.1, 0.3]
    assert lower_filter(interp1d(x, y, fill_value='extrapolate'))[1] == 4
    assert lower_filter(interp1d(x, y))[1] == 1 and lower_filter(optics.Tabulated1D(x, y))[1] == 1
    with pytest.raises(ValueError):
        lower_filter(optics.Tabulated1D(x, y, bounds_error=False, fill_value=1.5))
    with pytest.raises(ValueError):
        mo.Tabulated1D(x, y)(q)


def test_get_local_euklid_bases_golden():
    """Geometry.get_local_euklid_bases (reference math/geometry.py:263-281 plane, :602-626 cylinder) against a
    run of the unmodified reference (tests/golden/euklid_bases.npz, oracle/gen_golden.py case_euklid_bases),
    plus the reference's own known answer (math/tests/test_geometry.py:240-252)."""
    import os
    from marxs_b200 import geometry
    g = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'euklid_bases.npz')))
    for tag, cls in (('plane', geometry.FinitePlane), ('cyl', geometry.Cylinder)):
        geo = cls({'pos4d': g[tag + '_pos4d']})
        e1, e2, n = geo.get_local_euklid_bases(g[tag + '_loc'])
        for got, name in ((e1, 'e1'), (e2, 'e2'), (n, 'n')):
            assert np.asarray(got).shape == (50, 4)
            np.testing.assert_allclose(np.asarray(got), g[tag + '_' + name], rtol=1e-14, atol=1e-15, err_msg=tag + name)
    # a plane's base is its own axes, whatever the position
    geo = geometry.FinitePlane({'position': [5., 1., 2.], 'zoom': [1., 3., 7.]})
    x, y, z = geo.get_local_euklid_bases(np.random.rand(5, 2))
    assert np.all(x == np.array([0., 1., 0., 0.])) and np.all(y == np.array([0., 0., 1., 0.]))
    assert np.all(z == np.array([1., 0., 0., 0.]))
    # cylinder: tangent, axis and normal are mutually perpendicular for an un-sheared, isotropic-radius tube
    geo = geometry.Cylinder({'zoom': [10., 10., 4.], 'position': [1., 2., 3.]})
    e_phi, e_z, e_n = geo.get_local_euklid_bases(np.column_stack([np.linspace(-3, 3, 7), np.zeros(7)]))
    assert np.allclose(np.einsum('ij,ij->i', e_phi, e_n), 0.) and np.allclose(np.einsum('ij,ij->i', e_phi, e_z), 0.)
    assert np.allclose(np.linalg.norm(e_n, axis=1), 1.)


def test_paralleltransport_matrix_golden():
    """paralleltransport_matrix incl. a general Jones matrix and replace_nans=False against the unmodified
    reference (tests/golden/pt_matrix.npz; reference math/polarization.py:90-149, tests test_polarization.py:40-72)."""
    import os
    import torch
    from marxs_b200 import polarization
    g = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'pt_matrix.npz')))
    d1, d2 = torch.as_tensor(g['d1']), torch.as_tensor(g['d2'])
    for key, kw in (('ident', {}), ('ident_nan', {'replace_nans': False}), ('gen', {'jones': g['jones']}),
                    ('gen_nan', {'jones': g['jones'], 'replace_nans': False})):
        got = polarization.paralleltransport_matrix(d1, d2, **kw).numpy()
        assert got.shape == (40, 3, 3)
        np.testing.assert_allclose(got, g[key], rtol=1e-12, atol=1e-13, equal_nan=True, err_msg=key)
    assert np.isnan(g['ident_nan'][:5]).all() and np.array_equal(g['ident'][:5], np.tile(np.eye(3), (5, 1, 1)))
