"""Polarization vectors and parallel transport on the device (reference math/polarization.py)."""
import ctypes

import numpy as np

import torch

from . import _lib
from .geometry import _planes


def parallel_transport(dir_old, dir_new, pol_old):
    """Transport ``pol_old`` from ``dir_old`` to ``dir_new``; all (N, 4) CUDA tensors -> (N, 4)."""
    if not isinstance(dir_old, torch.Tensor) or dir_old.device.type != 'cuda':
        raise _lib.MxbError('parallel_transport needs CUDA tensors (no CPU fallback)')
    lib = _lib.load()
    n = dir_old.shape[0]
    a, b, p = _planes(dir_old), _planes(dir_new), _planes(pol_old)
    out = torch.zeros((4, n), dtype=torch.float64, device=dir_old.device)
    vp3 = ctypes.c_void_p * 3

    def ptrs(t):
        return vp3(*[t.data_ptr() + k * n * 8 for k in range(3)])
    with torch.cuda.device(dir_old.device):
        rc = lib.mxb_parallel_transport(ptrs(a), ptrs(b), ptrs(p), ptrs(out), n,
                                        torch.cuda.current_stream(dir_old.device).cuda_stream)
    _lib.check(lib, rc, 'mxb_parallel_transport')
    return out.T


def polarization_vectors(dir_array, angles):
    """Polarization angles -> unit vectors perpendicular to the photon direction (reference
    math/polarization.py:12-62): angle 0 is the direction closest to +y (closest to +x for photons flying
    along y).  ``dir_array`` (N, 4) and ``angles`` (N,) [rad] CUDA tensors -> (N, 4)."""
    if not isinstance(dir_array, torch.Tensor) or dir_array.device.type != 'cuda':
        raise _lib.MxbError('polarization_vectors needs CUDA tensors (no CPU fallback)')
    lib = _lib.load()
    n = dir_array.shape[0]
    if hasattr(angles, 'to') and hasattr(angles, 'unit'):
        angles = angles.to('rad').value
    ang = torch.as_tensor(angles, dtype=torch.float64, device=dir_array.device).contiguous()
    if ang.shape != (n,):
        raise ValueError('angles must have one entry per direction')
    d = _planes(dir_array)
    out = torch.zeros((4, n), dtype=torch.float64, device=dir_array.device)
    vp3 = ctypes.c_void_p * 3

    def ptrs(t):
        return vp3(*[t.data_ptr() + k * n * 8 for k in range(3)])
    with torch.cuda.device(dir_array.device):
        rc = lib.mxb_polarization_vectors(ptrs(d), ang.data_ptr(), ptrs(out), n,
                                          torch.cuda.current_stream(dir_array.device).cuda_stream)
    _lib.check(lib, rc, 'mxb_polarization_vectors')
    return out.T


def paralleltransport_matrix(dir1, dir2, jones=None, replace_nans=True):
    """(n, 3, 3) parallel-transport ray matrices (reference math/polarization.py:90-149).

    Identity Jones matrix (the default, and what every element of the hot path uses): column k is the transported
    unit vector e_k - three launches of the transport kernel, consistent with ``parallel_transport`` bit for bit.
    Any other 2 x 2 ``jones`` matrix (local s, p system of the element): P = O_out . diag(jones, 1) . O_in^-1 with
    O_in rows (s, p_in, dir1) and O_out columns (s, p_out, dir2), s = dir1 x dir2 normalised, evaluated with tensor
    operations on the device of the inputs.  Rays with dir1 parallel to dir2 (|dir1 x dir2| <= 1e-8, np.isclose)
    get the identity, or NaN with ``replace_nans=False``."""
    n = dir1.shape[0]
    identity = jones is None or torch.equal(torch.as_tensor(jones, dtype=torch.float64).cpu(), torch.eye(2, dtype=torch.float64))
    if not isinstance(dir1, torch.Tensor):
        dev = 'cuda' if torch.cuda.is_available() else 'cpu'
        dir1 = torch.as_tensor(np.asarray(dir1, dtype=float), device=dev)
        dir2 = torch.as_tensor(np.asarray(dir2, dtype=float), device=dev)
    d1 = dir1.as_subclass(torch.Tensor)[:, :3].to(torch.float64)
    d2 = dir2.as_subclass(torch.Tensor)[:, :3].to(torch.float64)
    d1 = d1 / d1.norm(dim=1, keepdim=True)
    d2 = d2 / d2.norm(dim=1, keepdim=True)
    s = torch.linalg.cross(d1, d2)
    s_norm = s.norm(dim=1)
    same = s_norm.abs() <= 1e-8                      # np.isclose(s_norm, 0)
    if identity and d1.device.type == 'cuda':
        def h(v):
            return torch.cat([v, torch.zeros((n, 1), dtype=v.dtype, device=v.device)], dim=1)
        cols = []
        for k in range(3):
            e = torch.zeros((n, 4), dtype=torch.float64, device=d1.device)
            e[:, k] = 1.
            cols.append(parallel_transport(h(d1), h(d2), e)[:, :3])
        pmat = torch.stack(cols, dim=2)
    else:
        j3 = torch.eye(3, dtype=torch.float64, device=d1.device)
        if jones is not None:
            j3[:2, :2] = torch.as_tensor(np.asarray(jones, dtype=float) if not isinstance(jones, torch.Tensor) else jones,
                                         dtype=torch.float64, device=d1.device)
        su = s / torch.where(same, torch.ones_like(s_norm), s_norm)[:, None]
        p_in = torch.linalg.cross(d1, su)
        p_out = torch.linalg.cross(d2, su)
        o_in = torch.stack([su, p_in, d1], dim=1)          # rows
        o_out = torch.stack([su, p_out, d2], dim=2)        # columns
        pmat = o_out @ (j3 @ o_in)
    fill = torch.eye(3, dtype=torch.float64, device=d1.device) * (1. if replace_nans else float('nan'))
    pmat = torch.where(same[:, None, None], fill[None], pmat)
    return pmat
