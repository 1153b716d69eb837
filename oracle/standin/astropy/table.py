"""Thin numpy holders standing in for astropy.table (TEST INFRASTRUCTURE).

Only what the reference's hot path touches: column get/set (in place),
boolean-mask row selection, add_column, meta, copy, and ``Table.read`` for the
handful of small text formats the reference ships (rdb, whitespace/tab tables,
ECSV headers, the in-source PIX_CORNER csv string)."""
import copy as _copy
import csv
import io
import os
import re
from collections import OrderedDict

import numpy as np

from . import units as u


class Column(np.ndarray):
    def __new__(cls, data=None, name=None, dtype=None, shape=(), length=0,
                unit=None, **kwargs):
        if data is None:
            arr = np.zeros((length,) + tuple(shape), dtype=dtype or float)
        else:
            arr = np.array(data, dtype=dtype)
        obj = arr.view(cls)
        obj.name = name
        obj.unit = unit if unit is not None else getattr(data, 'unit', None)
        if isinstance(obj.unit, u.Unit) and obj.unit == u.dimensionless_unscaled \
                and not isinstance(data, u.Quantity):
            obj.unit = None
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, 'name', None)
        self.unit = getattr(obj, 'unit', None)

    @property
    def data(self):
        return self.view(np.ndarray)

    @property
    def value(self):
        return self.view(np.ndarray)

    @property
    def quantity(self):
        return u.Quantity(self.view(np.ndarray), self.unit)

    def to(self, unit, equivalencies=None):
        return self.quantity.to(unit, equivalencies)

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        # ufunc results are plain arrays/scalars (astropy returns Column but
        # the reference never relies on that for arithmetic results)
        arr = np.asarray(arr)
        if arr.ndim == 0:
            return arr[()]
        return arr


class TableColumns(OrderedDict):
    """OrderedDict that also takes int / slice keys like astropy's."""

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            return list(self.values())[item]
        if isinstance(item, slice):
            return TableColumns(list(self.items())[item])
        return OrderedDict.__getitem__(self, item)


class MaskedColumn(Column):
    pass


class Row:
    def __init__(self, table, index):
        self._table = table
        self._index = index

    def __getitem__(self, key):
        return self._table[key][self._index]

    def __setitem__(self, key, val):
        self._table[key][self._index] = val

    @property
    def colnames(self):
        return self._table.colnames

    @property
    def meta(self):
        return self._table.meta


class Table:
    def __init__(self, data=None, names=None, meta=None, rows=None, copy=True,
                 dtype=None):
        self.columns = TableColumns()
        self.meta = OrderedDict() if meta is None else meta
        if isinstance(data, Row):
            t = data._table
            for n in t.colnames:
                c = t[n][data._index:data._index + 1]
                self.columns[n] = Column(c, name=n, unit=t[n].unit)
            self.meta = t.meta
        elif isinstance(data, Table):
            for n in data.colnames:
                self.columns[n] = Column(data[n].data.copy() if copy else data[n],
                                         name=n, unit=data[n].unit)
            self.meta = _copy.deepcopy(data.meta) if copy else data.meta
        elif isinstance(data, dict):
            for n, v in data.items():
                self.columns[n] = Column(v, name=n)
        elif isinstance(data, (list, tuple)) and names is not None:
            for n, v in zip(names, data):
                self.columns[n] = Column(v, name=n)
        elif rows is not None and names is not None:
            cols = list(zip(*rows))
            for n, v in zip(names, cols):
                self.columns[n] = Column(list(v), name=n)
        elif data is not None:
            raise TypeError('stand-in Table: unsupported input')

    # ---- container protocol --------------------------------------------
    @property
    def colnames(self):
        return list(self.columns.keys())

    def keys(self):
        return self.colnames

    def __len__(self):
        for c in self.columns.values():
            return len(c)
        return 0

    def __contains__(self, name):
        return name in self.columns

    def __iter__(self):
        for i in range(len(self)):
            yield Row(self, i)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.columns[key]
        if isinstance(key, (int, np.integer)):
            return Row(self, key)
        if isinstance(key, (list, tuple)) and len(key) and isinstance(key[0], str):
            out = Table()
            for n in key:
                out.columns[n] = self.columns[n]
            out.meta = self.meta
            return out
        out = Table()
        for n, c in self.columns.items():
            out.columns[n] = Column(c.data[key], name=n, unit=c.unit)
        out.meta = OrderedDict(self.meta)
        return out

    def __setitem__(self, key, value):
        if isinstance(key, str):
            if key in self.columns and np.ndim(value) == 0:
                self.columns[key][...] = value
                return
            if np.ndim(value) == 0:
                value = np.full(len(self), value)
            unit = getattr(value, 'unit', None)
            if unit is None and key in self.columns:
                unit = self.columns[key].unit
            self.columns[key] = Column(np.asarray(value), name=key, unit=unit)
        else:
            raise TypeError('stand-in Table: row assignment not supported')

    def add_column(self, col, index=None, name=None):
        name = name or col.name
        if not isinstance(col, Column):
            col = Column(col, name=name)
        col.name = name
        self.columns[name] = col

    def remove_column(self, name):
        del self.columns[name]

    def rename_column(self, old, new):
        self.columns = TableColumns((new if k == old else k, v)
                                   for k, v in self.columns.items())
        self.columns[new].name = new

    def copy(self, copy_data=True):
        return Table(self, copy=copy_data)

    def __repr__(self):
        return '<standin Table n={0} cols={1}>'.format(len(self), self.colnames)

    # ---- readers ----------------------------------------------------------
    @classmethod
    def read(cls, source, format=None, **kwargs):
        if isinstance(source, str) and ('\n' in source or not os.path.exists(source)):
            text = source
            fname = ''
        else:
            with open(source) as f:
                text = f.read()
            fname = str(source)
        if fname.endswith('.rdb') or format == 'ascii.rdb':
            return cls._read_rdb(text)
        if format == 'ascii.ecsv' or text.lstrip().startswith('# %ECSV'):
            return cls._read_ecsv(text)
        lines = [l for l in text.splitlines() if l.strip() and not l.lstrip().startswith('#')]
        if ',' in lines[0] and '"' in text:
            return cls._read_csv_noheader(lines)
        return cls._read_whitespace(lines)

    @staticmethod
    def _convert(strs):
        try:
            return np.array([int(s) for s in strs])
        except ValueError:
            pass
        try:
            return np.array([float(s) for s in strs])
        except ValueError:
            return np.array(strs)

    @classmethod
    def _read_rdb(cls, text):
        lines = [l for l in text.splitlines() if l.strip() and not l.startswith('#')]
        names = lines[0].split('\t')
        types = lines[1].split('\t')
        rows = [l.split('\t') for l in lines[2:]]
        out = cls()
        for i, (n, t) in enumerate(zip(names, types)):
            col = [r[i] for r in rows]
            if t.strip().upper().endswith('N'):
                out.columns[n] = Column(np.array([float(c) for c in col]), name=n)
            else:
                out.columns[n] = Column(np.array(col), name=n)
        return out

    @classmethod
    def _read_ecsv(cls, text):
        units = {}
        for mt in re.finditer(r"name: '?([^,'}]+)'?(?:, unit: ([^,}]+))?", text):
            if mt.group(2):
                units[mt.group(1)] = mt.group(2).strip()
        lines = [l for l in text.splitlines() if l.strip() and not l.startswith('#')]
        delim = ',' if ',' in lines[0] else None
        names = [n.strip() for n in lines[0].split(delim)]
        rows = [l.split(delim) for l in lines[1:]]
        out = cls()
        for i, n in enumerate(names):
            unit = getattr(u, units[n], None) if n in units else None
            out.columns[n] = Column(cls._convert([r[i] for r in rows]), name=n,
                                    unit=unit)
        return out

    @classmethod
    def _read_csv_noheader(cls, lines):
        rows = list(csv.reader(io.StringIO('\n'.join(lines))))
        ncol = max(len(r) for r in rows)
        out = cls()
        for i in range(ncol):
            out.columns['col{0}'.format(i + 1)] = Column(
                cls._convert([r[i] if i < len(r) else '' for r in rows]),
                name='col{0}'.format(i + 1))
        return out

    @classmethod
    def _read_whitespace(cls, lines):
        delim = '\t' if '\t' in lines[0] else None
        names = [n.strip() for n in lines[0].split(delim)]
        rows = [[c.strip() for c in l.split(delim)] for l in lines[1:]]
        out = cls()
        for i, n in enumerate(names):
            out.columns[n] = Column(cls._convert([r[i] for r in rows]), name=n)
        return out


class QTable(Table):
    pass


def join(*args, **kwargs):
    raise NotImplementedError('stand-in: table join is out of scope')
