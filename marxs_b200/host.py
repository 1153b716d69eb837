"""Host-resident photon tables traced through ``mxb_trace_host``.

``HostPhotonTable`` keeps every column in pinned host memory in the engine's
struct-of-arrays layout ((4, N) component planes for pos/dir/polarization) and
exposes the reference layout as zero-copy views (``table['pos']`` is the (N, 4)
transposed view).  ``trace_host`` is the end-to-end call: H2D of the photon
record, the fused kernel, D2H of the full result, chunked and overlapped inside
the C library."""
import ctypes
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, rng
from .program import Lowering, FIRST_OUT

VECTORS = ('pos', 'dir', 'polarization')


def _pinned(shape, dtype):
    pin = torch.cuda.is_available()
    t = torch.empty(shape, dtype=dtype, pin_memory=pin)
    return t, t.numpy()


class HostPhotonTable:
    def __init__(self, n=0, meta=None):
        self.n = int(n)
        self._keep = OrderedDict()      # name -> (torch tensor, numpy view) in SoA layout
        self.meta = OrderedDict() if meta is None else meta
        self.id0 = 0

    @classmethod
    def from_columns(cls, cols, meta=None):
        names = cols.colnames if hasattr(cols, 'colnames') else list(cols.keys())
        n = len(np.asarray(cols[names[0]]))
        out = cls(n, meta=OrderedDict(getattr(cols, 'meta', None) or (meta or {})))
        for k in names:
            out[k] = np.asarray(cols[k])
        return out

    @property
    def colnames(self):
        return list(self._keep.keys())

    def __len__(self):
        return self.n

    def __contains__(self, name):
        return name in self._keep

    def planes(self, name):
        """SoA storage: (4, N) for vector columns, (N,) otherwise."""
        return self._keep[name][1]

    def __getitem__(self, name):
        a = self._keep[name][1]
        return a.T if a.ndim == 2 else a

    def new_column(self, name, dtype=np.float64, vector=False):
        dt = np.dtype(dtype)
        tdt = torch.float64 if dt.kind == 'f' else (torch.int32 if dt.itemsize == 4 else torch.int64)
        self._keep[name] = _pinned((4, self.n) if vector else (self.n,), tdt)
        return self._keep[name][1]

    def __setitem__(self, name, value):
        v = np.asarray(value)
        if v.ndim == 2:
            if v.shape != (self.n, 4):
                raise ValueError('vector columns must have shape (N, 4)')
            if name not in self._keep:
                self.new_column(name, np.float64, vector=True)
            self._keep[name][1][...] = v.T
        else:
            if name not in self._keep:
                self.new_column(name, np.int64 if v.dtype.kind in 'iu' else np.float64)
            self._keep[name][1][...] = v

    def to_numpy(self):
        return OrderedDict((k, np.array(self[k])) for k in self.colnames)


def lower(elements, colnames, meta):
    """Lower ``elements`` (must be fully fusable) into one Program."""
    lw = Lowering(colnames, meta=meta)
    for e in elements:
        e._lower(lw)
    return lw.finish()


def _lean_struct(prog, table, lean):
    """Output pointers for a LEAN result set: only the columns named in ``lean`` come back; id columns
    travel as int32 (MxbHostOptions.i32_out) and are stored as int32 in ``table``."""
    cols = _lib.MxbColumns()
    opts = _lib.MxbHostOptions()
    unknown = [c for c in lean if c not in prog.out_f64 and c not in prog.out_i64 and c not in ('probability', 'energy')]
    if unknown:
        raise KeyError('lean columns {0} are not produced by this program'.format(unknown))
    if 'probability' in lean:
        if 'probability' not in table:
            table.new_column('probability', np.float64)
        cols.f64[10] = table.planes('probability').ctypes.data
    for k, name in enumerate(prog.out_f64):
        if name in lean:
            if name not in table:
                table.new_column(name, np.float64)
            cols.f64[FIRST_OUT + k] = table.planes(name).ctypes.data
    for k, name in enumerate(prog.out_i64):
        if name in lean:
            if name not in table or table.planes(name).dtype != np.int32:
                table.new_column(name, np.int32)
            opts.i32_out[k] = table.planes(name).ctypes.data
    return cols, opts


def _struct(prog, table, which, draws=None):
    cols = _lib.MxbColumns()
    n = len(table)
    if which in ('in', 'both'):
        for vi, name in enumerate(VECTORS):
            base = table.planes(name).ctypes.data
            for k in range(3):
                cols.f64[3 * vi + k] = base + k * n * 8
        cols.f64[9] = table.planes('energy').ctypes.data
        cols.f64[10] = table.planes('probability').ctypes.data
    if which in ('out', 'both'):
        for k, name in enumerate(prog.out_f64):
            if name not in table:
                table.new_column(name, np.float64)
            cols.f64[FIRST_OUT + k] = table.planes(name).ctypes.data
        for k, name in enumerate(prog.out_i64):
            if name not in table:
                table.new_column(name, np.int64)
            cols.i64[k] = table.planes(name).ctypes.data
    if draws is not None:
        for k, d in enumerate(draws):
            if d is not None:
                cols.draws[k] = d.ctypes.data
    return cols


def trace_host(instrument, table, out=None, draws=None, chunk=None, check=True, program=None, lean=None):
    """Trace a HOST photon table through ``instrument`` (an element or list of elements).

    In place by default (reference semantics); with ``out`` (another HostPhotonTable of
    the same length) the inputs stay untouched.  Returns (table_or_out, Program).

    ``lean``: names of the only columns to bring back (SURVEY 8d lean mode, e.g. ``['order', 'facet',
    'CCD_ID', 'chipx', 'chipy', 'probability']``); needs ``out``.  The kernel is specialised without the
    stores of the other diagnostics, pos / dir / polarization stay on the device, and id columns come
    back as int32 (``out[name].astype(np.int64)`` widens them; -1 = no hit).  The energy column is not
    modified by any element: read it from the input table."""
    lib = _lib.load()
    elements = instrument if isinstance(instrument, (list, tuple)) else [instrument]
    prog = program if program is not None else lower(elements, table.colnames, table.meta)
    if prog.aux:
        raise ValueError('fused device images (element.image) are not available on the host-buffer path: '
                         'the image lives on the device; unset it or use the device API')
    dst = table if out is None else out
    if lean is not None:
        if out is None:
            raise ValueError('lean output needs a separate `out` table (the input stays as it is)')
        return _trace_host_lean(lib, prog, table, out, draws, chunk, check, list(lean))
    if out is not None:
        for name in VECTORS:
            if name not in out:
                out.new_column(name, np.float64, vector=True)
                out.planes(name)[3] = table.planes(name)[3]
        for name in ('energy', 'probability'):
            if name not in out:
                out.new_column(name, np.float64)
        # (the energy plane of ``out`` is filled by the library: it comes back from the device with the
        #  other results, which is cheaper than an 8 B/photon host memcpy here)
    keep = None
    if draws is not None:
        keep = [np.ascontiguousarray(d, dtype=np.float64) if d is not None else None
                for d in (draws.table if hasattr(draws, 'table') else draws)]
    cin = _struct(prog, table, 'in', keep)
    cout = _struct(prog, dst, 'both')
    status = np.zeros(_lib.MXB_STATUS_WORDS, dtype=np.uint64)
    rc = lib.mxb_trace_host(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cin), ctypes.byref(cout),
                            len(table), int(chunk if chunk else os.environ.get('MXB_HOST_CHUNK', 0)), int(table.id0), rng.next_launch_seed(),
                            status.ctypes.data)
    _lib.check(lib, rc, 'mxb_trace_host')
    dst.meta.update(prog.meta_updates)
    prog.last_status = status
    if check:
        if status[_lib.MXB_ST_PROB_RANGE]:
            raise ValueError('Found probability outside of the 0..1 arange.')
        if status[_lib.MXB_ST_FILTER_BOUNDS]:
            raise ValueError('A value in x_new is outside the interpolation range.')
        if status[_lib.MXB_ST_INTENSITY]:
            raise ValueError('Intensity cannot be > 1')
    return dst, prog


def _trace_host_lean(lib, prog, table, out, draws, chunk, check, lean):
    keep = None
    if draws is not None:
        keep = [np.ascontiguousarray(d, dtype=np.float64) if d is not None else None
                for d in (draws.table if hasattr(draws, 'table') else draws)]
    cin = _struct(prog, table, 'in', keep)
    cout, opts = _lean_struct(prog, out, lean)
    status = np.zeros(_lib.MXB_STATUS_WORDS, dtype=np.uint64)
    rc = lib.mxb_trace_host_opts(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cin), ctypes.byref(cout),
                                 len(table), int(chunk if chunk else os.environ.get('MXB_HOST_CHUNK', 0)), int(table.id0),
                                 rng.next_launch_seed(), status.ctypes.data, ctypes.byref(opts))
    _lib.check(lib, rc, 'mxb_trace_host_opts')
    out.meta.update(prog.meta_updates)
    prog.last_status = status
    if check:
        _raise_status(status)
    return out, prog


def _raise_status(status):
    if status[_lib.MXB_ST_PROB_RANGE]:
        raise ValueError('Found probability outside of the 0..1 arange.')
    if status[_lib.MXB_ST_FILTER_BOUNDS]:
        raise ValueError('A value in x_new is outside the interpolation range.')
    if status[_lib.MXB_ST_INTENSITY]:
        raise ValueError('Intensity cannot be > 1')


def lean_bytes(prog, n, lean):
    """(h2d, d2h) bytes per call of the lean mode, counted from the planes copied."""
    d2h = 0
    for name in lean:
        d2h += 4 * n if name in prog.out_i64 else 8 * n
    return 11 * 8 * n, d2h


def h2d_d2h_bytes(prog, n, in_place=True):
    """Bytes moved per call by mxb_trace_host for n photons (counted from the planes copied).
    Out of place (``out=`` another table) the unchanged energy plane is brought back too."""
    h2d = 11 * 8 * n
    d2h = (10 + (0 if in_place else 1) + len(prog.out_f64) + len(prog.out_i64)) * 8 * n
    return h2d, d2h
