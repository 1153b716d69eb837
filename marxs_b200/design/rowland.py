"""Rowland-torus geometry and the placement of facets / CCDs on it (reference marxs/design/rowland.py).

Host-side SETUP code (numpy): it produces the ``pos4d`` of every facet once; the facets then become
the rows of a fused facet array (``marxs_b200.simulator.Parallel``).  One deliberate difference:
the reference finds the line - torus intersection with ``scipy.optimize.root`` from a starting
guess; here the quartic in the line parameter is solved in closed form (``numpy.roots`` + two Newton
steps) and the real root nearest to the starting point is taken - the same point the reference's
solver converges to for facet placement (checked against reference output in the tests, 1e-7 mm),
without its iteration noise or its "Intersection with torus not found" failures.
"""
import numpy as np

from ..affines import axangle2mat, compose
from ..base import MarxsElement
from ..geometry import Geometry
from ..simulator import ParallelCalculated

__all__ = ['RowlandTorus', 'ElementsOnTorus', 'GratingArrayStructure', 'RectangularGrid', 'CircularMeshGrid',
           'design_tilted_torus', 'anglediff']


def _h2e(v):
    v = np.asarray(v, dtype=float)
    w = v[..., 3:4]
    return v[..., :3] / np.where(w == 0, 1., w)


def _e2h(v, w):
    v = np.asarray(v, dtype=float)
    return np.concatenate([v, np.full(v.shape[:-1] + (1,), float(w))], axis=-1)


def anglediff(phi):
    """Angle range covered by ``phi = [phi0, phi1]``, accounting for 2 pi (reference math/utils.py:166-178)."""
    d = phi[1] - phi[0]
    if (d < 0.) or (d > (2. * np.pi)):
        d = d % (2. * np.pi)
    return d


class RowlandTorus(MarxsElement, Geometry):
    """Torus with the local y axis as symmetry axis; the origin is the focal point.

    R : radius with which the Rowland circle is rotated around the symmetry axis,
    r : radius of the Rowland circle (reference :78-105)."""

    display = {'color': (1., 0.3, 0.3), 'opacity': 0.2, 'shape': 'torus; surface'}

    def __init__(self, R, r, **kwargs):
        self.R = R
        self.r = r
        Geometry.__init__(self, kwargs)
        MarxsElement.__init__(self, **kwargs)

    # ---- implicit form ---------------------------------------------------------------
    def _to_local(self, xyz):
        inv = np.linalg.inv(self.pos4d)
        return _h2e(np.einsum('...ij,...j', inv, _e2h(xyz, 1)))

    def quartic(self, xyz, transform=True):
        """Torus equation; its roots are the points on the torus (reference :107-134)."""
        xyz = np.asarray(xyz, dtype=float)
        if xyz.shape[-1] != 3:
            raise ValueError('Input coordinates must be defined in Euclidean space.')
        if transform:
            xyz = self._to_local(xyz)
        return (((xyz ** 2).sum(axis=-1) + self.R ** 2. - self.r ** 2.) ** 2.
                - 4. * self.R ** 2. * (xyz[..., [0, 2]] ** 2).sum(axis=-1))

    def solve_quartic(self, origin, v, transform=True):
        """Intersect the line ``origin + k v`` (homogeneous coordinates) with the torus and return the
        intersection nearest to ``origin`` as a homogeneous point (reference :136-183, see module docstring)."""
        o, d = _h2e(origin), _h2e(v)
        if transform:
            inv = np.linalg.inv(self.pos4d)
            ol = _h2e(inv @ _e2h(o, 1))
            dl = (inv @ _e2h(d, 0))[:3]
        else:
            ol, dl = o, d
        R, r = self.R, self.r
        a, b, c = dl @ dl, 2. * (ol @ dl), ol @ ol + R ** 2 - r ** 2
        A, B, C = dl[0] ** 2 + dl[2] ** 2, 2. * (ol[0] * dl[0] + ol[2] * dl[2]), ol[0] ** 2 + ol[2] ** 2
        coeff = [a * a, 2. * a * b, b * b + 2. * a * c - 4. * R ** 2 * A, 2. * b * c - 4. * R ** 2 * B,
                 c * c - 4. * R ** 2 * C]
        roots = np.roots(coeff)
        scale = max(abs(R), abs(r), 1.)
        real = roots[np.abs(roots.imag) <= 1e-6 * scale].real
        if len(real) == 0:
            raise Exception('Intersection with torus not found.')
        k = real[np.argmin(np.abs(real))]

        def f(k):
            return (a * k * k + b * k + c) ** 2 - 4. * R ** 2 * (A * k * k + B * k + C)

        def df(k):
            return 2. * (a * k * k + b * k + c) * (2. * a * k + b) - 4. * R ** 2 * (2. * A * k + B)
        for _ in range(3):                      # polish the closed-form root
            g = df(k)
            if g == 0:
                break
            k = k - f(k) / g
        return _e2h(o + k * d, 1)

    # ---- parametric form -----------------------------------------------------------------
    def parametric(self, theta, phi):
        """Points on the torus: theta along the Rowland circle (0 = on the optical axis opposite the
        focus), phi around the symmetry axis (reference :209-233)."""
        theta, phi = np.asarray(theta, dtype=float), np.asarray(phi, dtype=float)
        rho = self.R + self.r * np.cos(theta)
        local = np.array([rho * np.cos(phi), self.r * np.sin(theta), rho * np.sin(phi), np.ones_like(rho * phi)]).T
        return np.einsum('...ij,...j', self.pos4d, local)

    def parametric_surface(self, theta, phi, display=None):
        theta, phi = np.asarray(theta), np.asarray(phi)
        if (phi.ndim != 1) or (theta.ndim != 1):
            raise ValueError('input parameters have 1-dim shape.')
        theta, phi = np.meshgrid(theta, phi)
        return self.parametric(theta, phi)

    def xyzw2parametric(self, xyzw, transform=True, intersectvalid=True):
        """(theta, phi) of points on the torus surface (reference :235-280)."""
        xyzw = np.asarray(xyzw, dtype=float)
        if transform:
            xyzw = np.einsum('...ij,...j', np.linalg.inv(self.pos4d), xyzw)
        xyz = _h2e(xyzw)
        if not np.allclose(self.quartic(xyz, transform=False) / self.R ** 4., 0.):
            raise ValueError('Parametric representation is only defined for points on torus surface.')
        s = np.sign(np.sqrt(xyz[:, 0] ** 2 + xyz[:, 2] ** 2) - self.R)      # the two branches of the circle
        theta = np.arcsin(s * xyz[:, 1] / self.r) + (s < 0) * np.pi
        factor = self.R + self.r * np.cos(theta)
        phi = np.zeros_like(factor)
        ok = factor != 0
        phi[ok] = np.arctan2(xyz[ok, 2] / factor[ok], xyz[ok, 0] / factor[ok])
        if not intersectvalid:
            phi[~ok] = np.nan
        return theta, phi

    def normal_parametric(self, theta, phi):
        """Outward normal for (theta, phi) (reference :282-307)."""
        theta, phi = np.asarray(theta), np.asarray(phi)
        if not ((theta.ndim == 1) and (phi.ndim == 1)):
            raise ValueError('theta and phi must be 1-d arrays.')
        n = np.stack([np.cos(theta) * np.cos(phi), np.sin(theta), np.cos(theta) * np.sin(phi), np.zeros(len(theta))], axis=1)
        return np.einsum('...ij,...j', self.pos4d, n)

    def normal(self, xyzw):
        """Outward normal at points on the surface, homogeneous (reference :309-331)."""
        xyzw = np.asarray(xyzw, dtype=float)
        if xyzw.ndim != 2:
            raise ValueError('Shape of input array must be (N, 4).')
        return self.normal_parametric(*self.xyzw2parametric(xyzw, True))

    def xyz_from_radiusangle(self, radius, angle, start):
        """Point of the torus above the polar coordinates (radius, angle) of the plane perpendicular to the
        optical axis (reference :333-366)."""
        y, z = radius * np.cos(angle), radius * np.sin(angle)
        hit = self.solve_quartic(np.array([np.mean(start), y, z, 1.]), np.array([1., 0, 0, 0]), transform=False)
        return _h2e(np.einsum('...ij,...j', self.pos4d, hit))


def design_tilted_torus(f, alpha, beta):
    """Rowland torus from the focal distance and two angles, after Heilmann et al. 2010 (reference :369-422).
    Returns (R, r, pos4d)."""
    r = f / (2. * np.cos(alpha))
    R = r * np.sin(np.pi / 2 - beta)
    orientation = axangle2mat([0, 0, -1], beta - alpha)
    x_ct = f / 2 - R * np.cos(beta - alpha)
    y_ct = f / 2 * np.tan(alpha) + R * np.sin(beta - alpha)
    return R, r, compose([x_ct, y_ct, 0], orientation, np.ones(3))


class ElementsOnTorus(ParallelCalculated):
    """Elements placed on a Rowland torus: subclasses say where in the (y, z) plane of the torus, the
    x coordinate follows from the torus (reference :547-644).

    rowland : RowlandTorus;  d_element : [dy, dz] edge lengths reserved per element;
    guess_distance : where along ``optimize_axis`` the search for the torus starts (default R + r);
    normal_spec default: pointing into the torus; parallel_spec default: the torus symmetry axis."""

    def __init__(self, **kwargs):
        self.rowland = kwargs.pop('rowland')
        self.guess_distance = kwargs.pop('guess_distance', self.rowland.r + self.rowland.R)
        self.d_element = kwargs.pop('d_element')
        self.optimize_axis = np.asanyarray(kwargs.pop('optimize_axis', self.rowland.pos4d @ np.array([1, 0, 0, 0])))
        kwargs.setdefault('id_col', 'facet')
        self.id_col = kwargs['id_col']
        if 'normal_spec' not in kwargs:
            kwargs['normal_spec'] = lambda xyzw: -self.rowland.normal(xyzw)
        if 'parallel_spec' not in kwargs:
            kwargs['parallel_spec'] = self.rowland.pos4d @ np.array([0, 1, 0, 0])
        kwargs['pos_spec'] = self.elempos
        super().__init__(**kwargs)

    def elemposyz(self):
        raise NotImplementedError

    def elempos(self):
        ypos, zpos = self.elemposyz()
        start = np.stack([np.zeros_like(ypos), ypos, zpos, np.ones_like(ypos)], axis=1)
        origin = np.einsum('...ij,...j', self.rowland.pos4d, start) + self.guess_distance * self.optimize_axis
        return np.vstack([self.rowland.solve_quartic(o, self.optimize_axis) for o in origin])


class RectangularGrid(ElementsOnTorus):
    """Elements filling a rectangle ``y_range`` x ``z_range`` (torus coordinates; reference :647-700)."""

    def __init__(self, **kwargs):
        self.y_range = kwargs.pop('y_range')
        self.z_range = kwargs.pop('z_range', [-1e-10, 1e-10])
        super().__init__(**kwargs)

    def _axis(self, rng, d):
        n = max(1, int(np.ceil((rng[1] - rng[0]) / d)))
        return np.arange(0.5 * (rng[0] - n * d + rng[1] + d), rng[1], d)

    def elemposyz(self):
        ypos, zpos = np.meshgrid(self._axis(self.y_range, self.d_element[0]), self._axis(self.z_range, self.d_element[1]))
        return ypos.flatten(), zpos.flatten()


class CircularMeshGrid(ElementsOnTorus):
    """Elements on a rectangular mesh, kept where their centre lies within ``radius`` = [inner, outer]
    (reference :703-755)."""

    def __init__(self, **kwargs):
        self.radius = kwargs.pop('radius')
        super().__init__(**kwargs)

    def elemposyz(self):
        rmax = self.radius[1]
        n_y = int(np.ceil(2 * rmax / self.d_element[0]))
        n_z = int(np.ceil(2 * rmax / self.d_element[1]))
        ypos = (np.arange(n_y) - (n_y - 1) / 2.) * self.d_element[0]
        zpos = (np.arange(n_z) - (n_z - 1) / 2.) * self.d_element[1]
        ypos, zpos = np.meshgrid(ypos, zpos)
        rad = np.sqrt(ypos ** 2 + zpos ** 2)
        keep = (rad >= self.radius[0]) & (rad <= self.radius[1])
        return ypos[keep], zpos[keep]


class GratingArrayStructure(ElementsOnTorus):
    """Facets on concentric rings between ``radius`` = [inner, outer] (pairs of them for several rings of
    rings), optionally only the segment ``phi`` = [phi0, phi1] (reference :758-887)."""

    def __init__(self, **kwargs):
        self.phi = kwargs.pop('phi', [0., 2 * np.pi])
        self.radius = kwargs.pop('radius')
        super().__init__(**kwargs)

    def max_elements_on_radius(self, radius):
        return int(np.ceil((radius[1] - radius[0]) / self.d_element[0]))

    def distribute_elements_on_radius(self):
        """Radii of the ring centres; rings may reach beyond the limits when the width is not a multiple
        of the element size."""
        out = []
        for i in range(len(self.radius) // 2):
            lo, hi = self.radius[2 * i: 2 * i + 2]
            n = self.max_elements_on_radius([lo, hi])
            out.append(np.mean([lo, hi]) + np.arange(-n / 2 + 0.5, n / 2 + 0.5) * self.d_element[0])
        return np.hstack(out)

    def max_elements_on_arc(self, radius):
        return radius * anglediff(self.phi) // self.d_element[1]

    def distribute_elements_on_arc(self, radius):
        """Centre angles of the elements on the ring of this radius, evenly spread, never beyond ``phi``."""
        if len(self.phi) == 1:
            return self.phi
        n = self.max_elements_on_arc(radius - self.d_element[0] / 2)        # the inner edge is the crowded one
        element_angle = self.d_element[1] / (2. * np.pi * radius)
        gap = (anglediff(self.phi) - n * element_angle) / (n + 1)
        centers = gap + 0.5 * element_angle + np.arange(n) * (gap + element_angle)
        return (self.phi[0] + centers) % (2. * np.pi)

    def elemposyz(self):
        radii = self.distribute_elements_on_radius()
        angles = [self.distribute_elements_on_arc(r) for r in radii]
        radii = np.concatenate([[radii[i]] * len(a) for i, a in enumerate(angles)])
        angles = np.concatenate(angles)
        return radii * np.sin(angles), radii * np.cos(angles)
