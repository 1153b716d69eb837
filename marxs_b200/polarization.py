"""Polarization vectors and parallel transport on the device (reference math/polarization.py)."""
import ctypes

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
    """(n, 3, 3) parallel-transport ray matrices (reference math/polarization.py:90-149) for the identity
    Jones matrix: column k is the transported unit vector e_k, three launches of the transport kernel.
    ``replace_nans=True`` (identity where dir1 is parallel to dir2) is the only mode; other Jones matrices are
    not implemented (no element of the reference's hot path uses them)."""
    if jones is not None and not torch.equal(torch.as_tensor(jones, dtype=torch.float64), torch.eye(2, dtype=torch.float64)):
        raise NotImplementedError('paralleltransport_matrix: only the identity Jones matrix')
    if not replace_nans:
        raise NotImplementedError('paralleltransport_matrix: replace_nans=False')
    n = dir1.shape[0]

    def h(v):      # (n, 3) or (n, 4) -> (n, 4) with w = 0
        if v.shape[1] == 4:
            return v
        return torch.cat([v, torch.zeros((n, 1), dtype=v.dtype, device=v.device)], dim=1)
    d1, d2 = h(dir1), h(dir2)
    cols = []
    for k in range(3):
        e = torch.zeros((n, 4), dtype=torch.float64, device=d1.device)
        e[:, k] = 1.
        cols.append(parallel_transport(d1, d2, e)[:, :3])
    return torch.stack(cols, dim=2)
