"""Detector event lists: compaction of the photons that reached a detector.

After a trace most analyses only need the detected photons (``CCD_ID >= 0``, ``probability > 0``);
for the multi-GPU epilogue each rank compacts its events on the device
(``mxb_compact_events``: count / scan / scatter, order preserved) and the ragged lists are gathered
with NCCL (``marxs_b200.dist.gather_events``)."""
import ctypes
from collections import OrderedDict

import torch

from . import _lib
from .photons import PhotonBatch, VECTOR_COLUMNS


def compact(photons, columns=None, sel='CCD_ID', sel_min=0, weight='probability'):
    """Rows of ``photons`` with ``photons[sel] >= sel_min`` and ``photons[weight] > 0`` as a new
    PhotonBatch holding ``columns`` (default: all).  ``sel`` / ``weight`` may be None."""
    if photons.device.type != 'cuda':
        raise _lib.MxbError('marxs_b200 runs on CUDA devices only (no CPU fallback)')
    lib = _lib.load()
    n = len(photons)
    names = list(columns) if columns is not None else photons.colnames
    src, shapes = [], []
    for name in names:
        st = photons.storage(name)
        if st.element_size() != 8 or not st.is_contiguous():
            raise ValueError('column {0} must be a contiguous 8-byte column'.format(name))
        if st.dim() == 2:
            for k in range(st.shape[0]):
                src.append(st[k])
        else:
            src.append(st)
        shapes.append(st.shape)
    if len(src) > 64:
        raise ValueError('at most 64 planes per call')
    with torch.cuda.device(photons.device):
        dst = [torch.empty(n, dtype=t.dtype, device=photons.device) for t in src]
        n_out = torch.zeros(1, dtype=torch.int64, device=photons.device)
        ws_bytes = lib.mxb_compact_workspace(n)
        ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device=photons.device)
        sp = (ctypes.c_void_p * len(src))(*[t.data_ptr() for t in src])
        dp = (ctypes.c_void_p * len(dst))(*[t.data_ptr() for t in dst])
        selp = photons.storage(sel).data_ptr() if sel is not None else None
        if sel is not None and photons.storage(sel).dtype != torch.int64:
            raise ValueError('the selector column must be int64')
        wp = photons.storage(weight).data_ptr() if weight is not None else None
        rc = lib.mxb_compact_events(sp, dp, len(src), selp, int(sel_min), wp, n, n_out.data_ptr(), ws.data_ptr(),
                                    ws_bytes, torch.cuda.current_stream(photons.device).cuda_stream)
        _lib.check(lib, rc, 'mxb_compact_events')
        m = int(n_out.item())
    out = PhotonBatch(device=photons.device, meta=OrderedDict(photons.meta))
    k = 0
    for name, shape in zip(names, shapes):
        if len(shape) == 2:
            out._store[name] = torch.stack([dst[k + j][:m] for j in range(shape[0])]).contiguous()
            k += shape[0]
        else:
            out._store[name] = dst[k][:m].clone()
            k += 1
    return out


class EventStore:
    """Preallocated event list that batches of a long observation are appended to ON THE DEVICE
    (``mxb_compact_append``: the write cursor lives in device memory, so consecutive batches need no host
    round trip).  ``columns``: names of scalar 8-byte columns kept per event; ``capacity``: rows.

    ``planes``: optional dict name -> 1-D tensor-like destination (``data_ptr()``, e.g. a window of another
    GPU's store mapped over NVLink); by default the store allocates its own."""

    def __init__(self, columns, dtypes, capacity, device, planes=None):
        self.names = list(columns)
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.cols = OrderedDict()
        for name, dt in zip(self.names, dtypes):
            if planes is not None and name in planes:
                self.cols[name] = planes[name]
            else:
                self.cols[name] = torch.empty(self.capacity, dtype=dt, device=self.device)
        self.cursor = torch.zeros(2, dtype=torch.int64, device=self.device)    # rows written, rows dropped
        self._ws = None
        self._dp = (ctypes.c_void_p * len(self.names))(*[self.cols[c].data_ptr() for c in self.names])

    def reset(self):
        self.cursor.zero_()

    def append(self, photons, sel='CCD_ID', sel_min=0, weight='probability', n=None):
        """Append the rows of ``photons`` (first ``n``) with photons[sel] >= sel_min and photons[weight] > 0."""
        lib = _lib.load()
        n = len(photons) if n is None else int(n)
        src = []
        for name in self.names:
            st = photons.storage(name)
            if st.dim() != 1 or st.element_size() != 8:
                raise ValueError('column {0} must be a scalar 8-byte column'.format(name))
            src.append(st)
        ws_bytes = lib.mxb_compact_workspace(n)
        if self._ws is None or self._ws.numel() * 8 < ws_bytes:
            self._ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device=self.device)
        sp = (ctypes.c_void_p * len(src))(*[t.data_ptr() for t in src])
        selp = photons.storage(sel).data_ptr() if sel is not None else None
        wp = photons.storage(weight).data_ptr() if weight is not None else None
        rc = lib.mxb_compact_append(sp, self._dp, len(src), selp, int(sel_min), wp, n, self.cursor.data_ptr(),
                                    self.capacity, self._ws.data_ptr(), self._ws.numel() * 8,
                                    torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(lib, rc, 'mxb_compact_append')

    def count(self):
        """(rows stored, rows dropped for lack of capacity): one host read."""
        c = self.cursor.tolist()
        return int(c[0]), int(c[1])

    def columns(self):
        m, dropped = self.count()
        if dropped:
            raise _lib.MxbError('event store overflow: {0} events dropped (capacity {1})'.format(dropped, self.capacity))
        return OrderedDict((c, self.cols[c][:m]) for c in self.names)
