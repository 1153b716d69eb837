"""Flat geometries: ray x plane intersection on the device.

API of the reference's marxs/math/geometry.py (Geometry :98-198, FinitePlane
:201-281, RectangleHole/CircularHole :335-380); ``intersect`` runs the
``mxb_plane_intersect`` kernel instead of numpy."""
import ctypes
from copy import copy

import numpy as np
import torch

from . import _lib
from .base import _parse_position_keywords
from .program import geom14

__all__ = ['NoGeometry', 'Geometry', 'FinitePlane', 'PlaneWithHole', 'RectangleHole', 'CircularHole', 'Cylinder']


class NoGeometry:
    _geometry = {}
    shape = 'none'

    def __init__(self, kwargs={}):
        self.geometry = self


class Geometry(NoGeometry):

    def __init__(self, kwargs={}):
        self.pos4d = _parse_position_keywords(kwargs)
        self._geometry = copy(self._geometry)
        super().__init__(kwargs=kwargs)

    def __getitem__(self, key):
        """Named access to the pos4d columns: center, v_x/y/z, e_x/y/z (unit)."""
        if key == 'center':
            return self.pos4d[:, 3]
        elif key in ['v_x', 'e_x']:
            val = self.pos4d[:, 0]
        elif key in ['v_y', 'e_y']:
            val = self.pos4d[:, 1]
        elif key in ['v_z', 'e_z']:
            val = self.pos4d[:, 2]
        elif key == 'plane':
            c, n = self['center'], self['e_x']
            return np.array([n[0], n[1], n[2], -np.dot(n[:3], c[:3])])
        else:
            val = self._geometry[key]
            if isinstance(val, np.ndarray) and (val.shape[-1] == 4):
                val = np.dot(self.pos4d, val)
        if key[:2] == 'e_':
            return val / np.linalg.norm(val)
        return val

    def intersect(self, dir, pos):
        raise NotImplementedError

    def get_local_euklid_bases(self, interpos_local):
        raise NotImplementedError


def _planes(t):
    """(N, 4) tensor (any strides) -> (3, N) contiguous component planes."""
    t = t.as_subclass(torch.Tensor)
    if t.dim() != 2 or t.shape[1] < 3:
        raise ValueError('expected an (N, 4) array of homogeneous vectors')
    return t[:, :3].T.contiguous().to(torch.float64)


class FinitePlane(Geometry):
    """Rectangle spanned by +-v_y, +-v_z around center, normal e_x."""

    shape = 'box'
    loc_coos_name = ['y', 'z']
    circular = False

    def intersect(self, dir, pos):
        """-> (intersect bool (N,), interpos (N, 4), interpos_local (N, 2)), NaN on miss."""
        if not isinstance(dir, torch.Tensor):
            dev = 'cuda' if torch.cuda.is_available() else None
            if dev is None:
                raise _lib.MxbError('Geometry.intersect needs a CUDA device (no CPU fallback)')
            dir = torch.as_tensor(np.asarray(dir), device=dev)
            pos = torch.as_tensor(np.asarray(pos), device=dev)
        if dir.device.type != 'cuda':
            raise _lib.MxbError('Geometry.intersect needs CUDA tensors (no CPU fallback)')
        lib = _lib.load()
        n = dir.shape[0]
        d, p = _planes(dir), _planes(pos)
        hit = torch.empty(n, dtype=torch.uint8, device=dir.device)
        ipos = torch.empty((4, n), dtype=torch.float64, device=dir.device)
        ipos[3] = pos.as_subclass(torch.Tensor)[:, 3]
        loc = torch.empty((2, n), dtype=torch.float64, device=dir.device)
        g = np.ascontiguousarray(geom14(self.pos4d))
        vp3 = ctypes.c_void_p * 3
        vp2 = ctypes.c_void_p * 2
        dd = vp3(*[d.data_ptr() + k * n * 8 for k in range(3)])
        pp = vp3(*[p.data_ptr() + k * n * 8 for k in range(3)])
        ii = vp3(*[ipos.data_ptr() + k * n * 8 for k in range(3)])
        ll = vp2(*[loc.data_ptr() + k * n * 8 for k in range(2)])
        with torch.cuda.device(dir.device):
            rc = lib.mxb_plane_intersect(g.ctypes.data, 1 if self.circular else 0, dd, pp, hit.data_ptr(),
                                         ii, ll, n, torch.cuda.current_stream(dir.device).cuda_stream)
        _lib.check(lib, rc, 'mxb_plane_intersect')
        return hit.bool(), ipos.T, loc.T

    def get_local_euklid_bases(self, interpos_local):
        n = interpos_local.shape[0]
        return (np.tile(self['e_y'], (n, 1)), np.tile(self['e_z'], (n, 1)), np.tile(self['e_x'], (n, 1)))


class PlaneWithHole(FinitePlane):
    shape = 'triangulation'

    def __init__(self, kwargs):
        self._geometry['r_inner'] = kwargs.pop('r_inner', 0.)
        super().__init__(kwargs)


class RectangleHole(PlaneWithHole):
    pass


class CircularHole(PlaneWithHole):
    circular = True

    def __init__(self, kwargs):
        phi = kwargs.pop('phi', [0, 2. * np.pi])
        if np.max(np.abs(phi)) > 10:
            raise ValueError('Input angles >> 2 pi. Did you use degrees (radian expected)?')
        if phi[0] > phi[1]:
            raise ValueError('phi[1] must be greater than phi[0].')
        if (phi[1] - phi[0]) > (2 * np.pi + 1e-6):
            raise ValueError('phi[1] - phi[0] must be less than 2 pi.')
        self.phi = phi
        super().__init__(kwargs)


class Cylinder(Geometry):
    """Ring / tube: circle (ellipse) in the local xy plane, flat along z (reference
    marxs/math/geometry.py:383-564).  ``zoom[0], zoom[1]`` are the radii, ``zoom[2]`` the half
    width; ``phi_lim`` restricts the covered angle range."""

    loc_coos_name = ['phi', 'z']
    shape = 'surface'

    def __init__(self, kwargs={}):
        self.coos_limits = [np.asanyarray(kwargs.pop('phi_lim', [-np.pi, np.pi]), dtype=float), np.array([-1., 1.])]
        super().__init__(kwargs)

    def __getitem__(self, value):
        if value == 'R':
            from .affines import decompose44
            return decompose44(self.pos4d)[2][0]
        return super().__getitem__(value)

    @property
    def phi_lim(self):
        return self.coos_limits[0]

    def get_local_euklid_bases(self, interpos_local):
        """Local base (e_phi, e_z, e_n) at the positions ``interpos_local`` = (phi, z) (reference :602-626):
        the tangent and the normal are the images of the unit-circle vectors under pos4d, normalised;
        e_z is the (un-normalised) global-frame axis vector as the reference returns it.  Host arrays."""
        loc = np.asarray(interpos_local.detach().cpu() if isinstance(interpos_local, torch.Tensor) else interpos_local,
                         dtype=float)
        phi = loc[:, 0]
        zeros = np.zeros_like(phi)
        e_phi = np.array([-np.sin(phi), np.cos(phi), zeros, zeros]).T
        e_n = np.array([np.cos(phi), np.sin(phi), zeros, zeros]).T
        e_z = np.tile(self['e_z'], (loc.shape[0], 1))
        e_phi = np.einsum('ij,nj->ni', self.pos4d, e_phi)
        e_n = np.einsum('ij,nj->ni', self.pos4d, e_n)

        def unit(v):
            return v / np.linalg.norm(v, axis=-1)[:, None]
        return unit(e_phi), e_z, unit(e_n)

    def intersect(self, dir, pos):
        """-> (intersect bool (N,), interpos (N, 4), interpos_local (N, 2) = (phi, z)), NaN on miss.
        Runs the CYLINDER op of the trace kernel on a scratch table."""
        from .photons import PhotonBatch
        from .program import Lowering
        if not isinstance(dir, torch.Tensor):
            if not torch.cuda.is_available():
                raise _lib.MxbError('Geometry.intersect needs a CUDA device (no CPU fallback)')
            dir = torch.as_tensor(np.asarray(dir), device='cuda')
            pos = torch.as_tensor(np.asarray(pos), device='cuda')
        if dir.device.type != 'cuda':
            raise _lib.MxbError('Geometry.intersect needs CUDA tensors (no CPU fallback)')
        if bool(torch.any(dir.as_subclass(torch.Tensor)[:, 3] != 0)):
            raise ValueError('First input must be direction vectors.')
        n = dir.shape[0]
        b = PhotonBatch(device=dir.device)
        b['pos'] = pos.as_subclass(torch.Tensor).clone()
        b['dir'] = dir.as_subclass(torch.Tensor).clone()
        b['polarization'] = torch.zeros((n, 4), dtype=torch.float64, device=dir.device)
        b['energy'] = torch.ones(n, dtype=torch.float64, device=dir.device)
        b['probability'] = torch.ones(n, dtype=torch.float64, device=dir.device)
        lw = Lowering(b.colnames)
        lw.cylinder(self.pos4d, self.phi_lim)
        lw.commit(['_phi', '_z'], None, 0)
        prog = lw.finish()
        prog.run(b, check=True)
        if '_phi' not in b:     # nothing hit: the columns were pruned like the reference does
            nan = torch.full((n,), float('nan'), dtype=torch.float64, device=dir.device)
            return torch.zeros(n, dtype=torch.bool, device=dir.device), torch.stack([nan] * 4, 1), torch.stack([nan] * 2, 1)
        loc = torch.stack([b['_phi'].as_subclass(torch.Tensor), b['_z'].as_subclass(torch.Tensor)], 1)
        hit = torch.isfinite(loc[:, 0])
        ipos = b['pos'].as_subclass(torch.Tensor).clone()
        ipos[~hit] = float('nan')
        return hit, ipos, loc
