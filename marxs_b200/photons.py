"""Device-resident photon table with the column protocol MARXS elements use.

The reference keeps photons in an ``astropy.table.Table`` whose vector columns
``pos``/``dir``/``polarization`` are (N, 4) float64 and whose other columns are
(N,) (reference marxs/base/base.py:208-224, marxs/utils.py:16-51).  Here every
column is a torch tensor on the GPU, stored as a struct of arrays:

* vector columns: one contiguous (4, N) fp64 block (component planes), exposed
  as the transposed (N, 4) *view* so reference-style code such as
  ``photons['pos'][mask] = ...`` or ``photons['pos'][:, 3] = 1`` writes through;
* scalar columns: contiguous (N,) fp64, id columns (N,) int64.

Kernels receive raw plane pointers (include/mxb.h ``MxbColumns``).
"""
from collections import OrderedDict

import numpy as np
import torch

VECTOR_COLUMNS = ('pos', 'dir', 'polarization')


class Column(torch.Tensor):
    """torch.Tensor with the two astropy Column conveniences MARXS code uses."""

    @staticmethod
    def __new__(cls, data):
        return torch.as_tensor(data).as_subclass(cls)

    def copy(self):
        return self.clone()

    @property
    def value(self):
        return self.as_subclass(torch.Tensor)

    def numpy_array(self):
        return self.detach().cpu().as_subclass(torch.Tensor).numpy()


def _as_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().as_subclass(torch.Tensor).numpy()
    if hasattr(x, 'value') and not isinstance(x, np.ndarray):
        x = x.value
    data = getattr(x, 'data', x)
    if isinstance(data, memoryview):
        data = x
    return np.asarray(data)


class PhotonBatch:
    """Table-like container of photon columns living on one device."""

    def __init__(self, data=None, device=None, meta=None):
        self._store = OrderedDict()     # name -> storage tensor ((4,N) for vector columns)
        self.meta = OrderedDict() if meta is None else meta
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.device = torch.device(device)
        self.id0 = 0    # global index of photon 0 (Philox counter base; set by the sharding helpers)
        if data is None:
            return
        if isinstance(data, PhotonBatch):
            for n in data.colnames:
                self._store[n] = data._store[n].to(self.device).clone()
            self.meta = OrderedDict(data.meta)
            self.id0 = data.id0
            return
        names = data.colnames if hasattr(data, 'colnames') else list(data.keys())
        for n in names:
            self[n] = data[n]
        src_meta = getattr(data, 'meta', None)
        if src_meta:
            self.meta.update(src_meta)

    # ---- table protocol ---------------------------------------------------
    @property
    def colnames(self):
        return list(self._store.keys())

    def keys(self):
        return self.colnames

    def __contains__(self, name):
        return name in self._store

    def __len__(self):
        for n, t in self._store.items():
            return t.shape[-1]
        return 0

    def _view(self, name):
        t = self._store[name]
        if t.dim() == 2:
            return t.T.as_subclass(Column)
        return t.as_subclass(Column)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._view(key)
        if isinstance(key, (list, tuple)) and len(key) and isinstance(key[0], str):
            out = PhotonBatch(device=self.device, meta=self.meta)
            for n in key:
                out._store[n] = self._store[n]
            return out
        # row selection -> new table (copy), like astropy
        if isinstance(key, np.ndarray):
            key = torch.as_tensor(key, device=self.device)
        out = PhotonBatch(device=self.device, meta=OrderedDict(self.meta))
        for n, t in self._store.items():
            out._store[n] = (t[:, key] if t.dim() == 2 else t[key]).contiguous()
        return out

    def __setitem__(self, name, value):
        if not isinstance(name, str):
            raise TypeError('only column assignment is supported: photons[name] = value')
        n = len(self) if len(self._store) else None
        if isinstance(value, torch.Tensor):
            v = value.as_subclass(torch.Tensor).to(self.device)
        elif np.isscalar(value):
            if name in self._store:
                self._store[name].fill_(value)
                return
            if n is None:
                raise ValueError('cannot broadcast a scalar into an empty table')
            v = torch.full((n,), value, device=self.device,
                           dtype=torch.int64 if isinstance(value, (int, np.integer)) else torch.float64)
        else:
            arr = _as_numpy(value)
            if arr.dtype.kind == 'f':
                arr = arr.astype(np.float64, copy=False)
            elif arr.dtype.kind in 'iub':
                arr = arr.astype(np.int64 if arr.dtype.kind != 'b' else np.bool_, copy=False)
            v = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        if v.dim() == 2:
            if v.shape[1] != 4:
                raise ValueError('vector columns must have shape (N, 4)')
            v = v.to(torch.float64).T.contiguous()       # -> (4, N) component planes
        else:
            v = v.contiguous()
            if v.dtype in (torch.float32, torch.float16):
                v = v.to(torch.float64)
        if n is not None and v.shape[-1] != n:
            raise ValueError('column {0} has {1} rows, table has {2}'.format(name, v.shape[-1], n))
        if name in self._store and self._store[name].shape == v.shape and self._store[name].dtype == v.dtype:
            self._store[name].copy_(v)
        else:
            self._store[name] = v.clone() if isinstance(value, torch.Tensor) else v

    def add_column(self, col, name=None, index=None):
        if name is None:
            name = getattr(col, 'name', None)
        if name is None:
            raise ValueError('add_column needs a name')
        self[name] = col

    def new_column(self, name, dtype=torch.float64, fill=None, vector=False, n=None):
        """Allocate an output column (uninitialised unless ``fill`` is given); ``n`` is
        only needed for the first column of an empty table."""
        n = len(self) if n is None else int(n)
        shape = (4, n) if vector else (n,)
        t = torch.empty(shape, dtype=dtype, device=self.device)
        if fill is not None:
            t.fill_(fill)
        self._store[name] = t
        return t

    def remove_column(self, name):
        del self._store[name]

    def rename_column(self, old, new):
        self._store = OrderedDict((new if k == old else k, v) for k, v in self._store.items())

    def copy(self):
        return PhotonBatch(self, device=self.device)

    def to(self, device):
        out = PhotonBatch(device=device, meta=OrderedDict(self.meta))
        for n, t in self._store.items():
            out._store[n] = t.to(out.device)
        return out

    def storage(self, name):
        """The raw storage tensor ((4, N) for vector columns)."""
        return self._store[name]

    # ---- export -------------------------------------------------------------
    def to_numpy(self):
        """dict of numpy arrays in the reference layout ((N,4) vectors)."""
        out = OrderedDict()
        for n, t in self._store.items():
            a = t.detach().cpu().numpy()
            out[n] = np.ascontiguousarray(a.T) if a.ndim == 2 else a
        return out

    def to_table(self):
        """astropy Table if astropy is importable, else the numpy dict."""
        cols = self.to_numpy()
        try:
            from astropy.table import Table
        except ImportError:
            return cols
        return Table(cols, meta=dict(self.meta))

    def __repr__(self):
        return '<PhotonBatch n={0} device={1} columns={2}>'.format(len(self), self.device, self.colnames)


def generate_test_photons(n=1, device=None):
    """n identical photons at x=1 flying along -x, 1 keV, polarised along +y
    (reference marxs/utils.py:16-51)."""
    return PhotonBatch({'pos': np.tile(np.array([1., 0., 0., 1.]), (n, 1)),
                        'dir': np.tile(np.array([-1., 0., 0., 0.]), (n, 1)),
                        'energy': np.ones(n),
                        'polarization': np.tile(np.array([0., 1., 0., 0.]), (n, 1)),
                        'probability': np.ones(n)}, device=device)
