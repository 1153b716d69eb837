"""Polarization parallel transport on the device (reference math/polarization.py:151-170)."""
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
