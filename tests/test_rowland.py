"""CPU tests of marxs_b200.design.rowland (facet placement on a Rowland torus) against outputs of the
unmodified reference (tests/golden/rowland.npz, made by oracle/gen_golden.py case_rowland) and the
known answers of the reference's own tests (design/tests/test_rowland.py)."""
import os

import numpy as np
import pytest

from marxs_b200 import optics, simulator, program
from marxs_b200.affines import axangle2mat
from marxs_b200.design import (RowlandTorus, GratingArrayStructure, RectangularGrid, CircularMeshGrid,
                               design_tilted_torus)
from marxs_b200.design.rowland import anglediff

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rowland.npz'))
ORD0 = optics.OrderSelector([0])


def c3_gas():
    rt = RowlandTorus(6000., 6000.)
    return rt, GratingArrayStructure(rowland=rt, d_element=[30., 30.], radius=[300., 500.], elem_class=optics.CATGrating,
                                     elem_args={'d': 2e-4, 'zoom': [1, 13.5, 13.5], 'order_selector': ORD0,
                                                'orientation': axangle2mat([0, 0, 1], np.deg2rad(1.91))})


def test_config3_arrays_match_the_reference():
    rt, gas = c3_gas()
    assert len(gas.elements) == 561                                        # SURVEY 8(d) probe count
    np.testing.assert_allclose(np.array([e.pos4d for e in gas.elements]), GOLD['gas_pos4d'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(np.array(gas.elem_pos), GOLD['gas_elem_pos'], rtol=0, atol=1e-8)
    det = RectangularGrid(rowland=rt, d_element=[49.652, 49.652], y_range=[-50, 700], elem_class=optics.FlatDetector,
                          elem_args={'zoom': [1, 24.576, 12.288], 'pixsize': 0.024}, id_col='CCD_ID', guess_distance=25.)
    assert len(det.elements) == 16 and det.id_col == 'CCD_ID'
    # the reference's root finder stops at xtol ~1e-8 relative; our closed form is the exact root
    np.testing.assert_allclose(np.array([e.pos4d for e in det.elements]), GOLD['det_pos4d'], rtol=0, atol=1e-7)
    # every facet centre is on the torus, normals point at the torus normal
    c = np.array([e.geometry['center'] for e in gas.elements])
    assert np.abs(rt.quartic(c[:, :3])).max() / rt.R ** 4 < 1e-14
    assert [e.id_num for e in gas.elements] == list(range(561))


def test_tilted_torus_and_partial_ring():
    R, r, pos4d = design_tilted_torus(9e3, np.deg2rad(3.8), np.deg2rad(7.6))
    np.testing.assert_allclose([R, r], GOLD['tilt_Rr'], rtol=1e-15)
    np.testing.assert_allclose(pos4d, GOLD['tilt_pos4d'], rtol=0, atol=1e-12)
    rt = RowlandTorus(R, r, pos4d=pos4d)
    gas = GratingArrayStructure(rowland=rt, d_element=[25., 40.], radius=[200., 420.],
                                phi=[-0.5 + np.pi / 2, 0.5 + np.pi / 2], elem_class=optics.CATGrating,
                                normal_spec=np.array([0, 0, 0, 1.]),
                                elem_args={'d': 2e-4, 'zoom': [1, 10., 18.], 'order_selector': ORD0})
    assert len(gas.elements) == GOLD['gas2_pos4d'].shape[0]
    np.testing.assert_allclose(np.array([e.pos4d for e in gas.elements]), GOLD['gas2_pos4d'], rtol=0, atol=1e-8)
    pts = rt.parametric(GOLD['par_theta'], GOLD['par_phi'])
    np.testing.assert_allclose(pts, GOLD['par_xyzw'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rt.normal(pts), GOLD['par_normal'], rtol=0, atol=1e-12)
    t, p = rt.xyzw2parametric(pts)
    np.testing.assert_allclose(t, GOLD['par_theta_back'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(p, GOLD['par_phi_back'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(rt.quartic(pts[:, :3] + 1.), GOLD['quartic_offsurface'], rtol=1e-9)


def test_known_answers_of_the_reference_tests():
    """design/tests/test_rowland.py: radii (:19-33), arc (:35-60), torus quartic for a sphere-like torus
    (:99-114), tilted-torus geometry (:198-214), anglediff (math/tests/test_utils.py)."""
    rt = RowlandTorus(1000., 1000.)
    gas = GratingArrayStructure(rowland=rt, d_element=[20., 20.], radius=[300., 600.], phi=[-0.2 * np.pi, 0.2 * np.pi],
                                elem_class=optics.FlatGrating, elem_args={'d': 1e-3, 'zoom': [1, 10, 10], 'order_selector': ORD0})
    assert gas.max_elements_on_radius([300., 600.]) == 15
    radii = gas.distribute_elements_on_radius()
    assert len(radii) == 15 and np.allclose(np.diff(radii), 20.) and radii[0] == 310. and radii[-1] == 590.
    ang = gas.distribute_elements_on_arc(300.)
    inphi = (ang < 0.2 * np.pi) | (ang > 1.8 * np.pi)
    assert inphi.all() and len(ang) == gas.max_elements_on_arc(290.)
    assert np.isclose(anglediff([-0.2 * np.pi, 0.2 * np.pi]), 0.4 * np.pi)
    assert np.isclose(anglediff([0.1, 3.]), 2.9) and np.isclose(anglediff([3., 0.1]), 2 * np.pi - 2.9)
    # points on the torus: quartic vanishes; off: it does not
    th, ph = np.mgrid[0:2 * np.pi:7j, 0:2 * np.pi:9j]
    pts = rt.parametric(th.ravel(), ph.ravel())
    assert np.abs(rt.quartic(pts[:, :3])).max() / 1000. ** 4 < 1e-12
    assert abs(rt.quartic(np.array([5000., 0, 0]))) > 1.
    # a line through the torus: the solution is on the line, on the torus, and nearest to the origin point
    hit = rt.solve_quartic(np.array([2100., 30., 40., 1.]), np.array([1., 0, 0, 0]))
    assert abs(rt.quartic(hit[:3])) / 1000. ** 4 < 1e-14 and np.allclose(hit[1:], [30., 40., 1.]) and abs(hit[0] - 2000.) < 2.
    with pytest.raises(Exception, match='not found'):
        rt.solve_quartic(np.array([0., 5000., 0., 1.]), np.array([1., 0, 0, 0]))
    # Heilmann et al. 2010: the focus (origin) and the point f along the axis are on the tilted torus
    R, r, pos4d = design_tilted_torus(10., 0.1, 0.2)
    tt = RowlandTorus(R, r, pos4d=pos4d)
    assert abs(tt.quartic(np.array([0., 0, 0]))) < 1e-9 and abs(tt.quartic(np.array([10., 0, 0]))) < 1e-9
    # circular mesh: centres inside the annulus only
    cm = CircularMeshGrid(rowland=RowlandTorus(5000., 5000.), d_element=[40., 40.], radius=[100., 300.],
                          elem_class=optics.FlatGrating, elem_args={'d': 1e-3, 'zoom': [1, 15, 15], 'order_selector': ORD0})
    c = np.array([e.geometry['center'] for e in cm.elements])
    rad = np.hypot(c[:, 1], c[:, 2])
    assert len(c) > 100 and (rad >= 100.).all() and (rad <= 300.).all()


def test_rowland_arrays_lower_to_one_facet_array():
    """The arrays are ordinary ParallelCalculated containers: they lower to one culled facet array."""
    rt, gas = c3_gas()
    assert isinstance(gas, simulator.ParallelCalculated) and gas._can_lower()
    G = np.array([program.geom14(e.pos4d) for e in gas.elements])
    g = program.build_cull_grid(G)
    assert g is not None and g['mean_candidates'] < 6
