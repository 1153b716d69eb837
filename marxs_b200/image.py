"""Detector images on the device (input of the multi-GPU reduction epilogue)."""
import torch

from . import _lib


def detector_image(photons, xcol, ycol, nx, ny, weight='probability', sel=None, sel_lo=0, n_sel=1,
                   x0=0., y0=0., image=None, counts=None, want_counts=False):
    """Accumulate photons into image[(plane, iy, ix)] with ix = round(x - x0), iy = round(y - y0).

    Returns (image fp64 (n_sel, ny, nx), counts int64 or None).  fp64 atomics: the
    weighted image is reproducible up to summation order, the count image bit for bit."""
    lib = _lib.load()
    dev = photons.device
    if dev.type != 'cuda':
        raise _lib.MxbError('detector_image needs CUDA tensors (no CPU fallback)')
    if image is None:
        image = torch.zeros((n_sel, ny, nx), dtype=torch.float64, device=dev)
    if counts is None and want_counts:
        counts = torch.zeros((n_sel, ny, nx), dtype=torch.int64, device=dev)
    x, y = photons.storage(xcol), photons.storage(ycol)
    w = photons.storage(weight) if weight is not None else None
    s = photons.storage(sel) if sel is not None else None
    with torch.cuda.device(dev):
        rc = lib.mxb_hist2d(x.data_ptr(), y.data_ptr(), w.data_ptr() if w is not None else None,
                            s.data_ptr() if s is not None else None, int(sel_lo), int(n_sel),
                            float(x0), float(y0), len(photons), int(nx), int(ny), image.data_ptr(),
                            counts.data_ptr() if counts is not None else None,
                            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib, rc, 'mxb_hist2d')
    return image, counts
