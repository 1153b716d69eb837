"""Element base classes and ``pos4d`` parsing (host side).

Mirrors the public surface of the reference's marxs/base/base.py (MarxsElement
:66-92, SimulationSequenceElement :194-289, _parse_position_keywords :292-336)
so element constructors take the same keywords and raise the same errors."""
from collections import OrderedDict
from copy import deepcopy

import numpy as np
import torch

from . import affines

__all__ = ['GeometryError', 'MarxsElement', 'SimulationSequenceElement', 'TagVersion', 'check_meta_consistent',
           'check_energy_consistent', '_parse_position_keywords']


class GeometryError(Exception):
    pass


class MarxsElement:
    """Base class for all elements in a simulation."""

    display = {'shape': 'None'}

    def __init__(self, **kwargs):
        if 'name' in kwargs:
            self.name = kwargs.pop('name')
        else:
            self.name = self.__class__
        if len(kwargs) > 0:
            raise ValueError('Initialization arguments {0} not understood'.format(', '.join(kwargs.keys())))
        self.display = deepcopy(self.display)

    def describe(self):
        return OrderedDict(element=self.name)


class TagVersion(MarxsElement):
    """Tag a photon list with diagnostic information (reference base/base.py:100-139): every keyword given at
    construction or at call time goes into ``photons.meta`` together with the date and the engine version.
    Host-side bookkeeping only; the table stays on the device."""

    def __init__(self, ORIGIN=('unkwown', 'Institution where file was created'),
                 CREATOR=('MARXS', 'Person or program creating file'), **kwargs):
        super().__init__(name=kwargs.pop('name', self.__class__))
        from . import _lib
        kwargs['ORIGIN'], kwargs['CREATOR'] = ORIGIN, CREATOR
        kwargs['MARXSVER'] = ('marxs_b200 ABI {0}'.format(_lib.MXB_ABI_VERSION), 'MARXS code version')
        self.tags = kwargs

    def __call__(self, photons, *args, **kwargs):
        from datetime import datetime
        photons.meta['DATE'] = (datetime.now().isoformat()[:10], 'Date/time of computation')
        for k, v in self.tags.items():
            photons.meta[k] = v
        for k, v in kwargs.items():
            photons.meta[k] = v
        return photons


def check_meta_consistent(meta1, meta2, keywords=['ORIGIN', 'CREATOR', 'MARXSVER', 'MARXSGIT', 'MARXSTIM'],
                          allow_missing=True):
    """Raise AssertionError / KeyError unless the two headers agree on ``keywords`` (reference
    base/base.py:142-182): a keyword may be missing from both (``allow_missing``), never from only one."""
    for k in keywords:
        if (k in meta1) and (k in meta2):
            assert meta1[k] == meta2[k]
        else:
            if not allow_missing:
                raise KeyError('{0} not found in both dicts.'.format(k))
            if (k in meta1) or (k in meta2):
                raise KeyError('{0} found in one, but not both dicts.'.format(k))


def check_energy_consistent(photons):
    """Assert that all photons have the same energy (reference base/base.py:185-191); passes without an
    energy column.  One reduction on the device."""
    if 'energy' in photons.colnames:
        e = photons['energy']
        e = e.data if hasattr(e, 'data') and isinstance(getattr(e, 'data'), torch.Tensor) else torch.as_tensor(np.asarray(e))
        if e.numel():
            assert bool(torch.isclose(e, e[0], rtol=1e-05, atol=1e-08).all())


class SimulationSequenceElement(MarxsElement):
    """Base class for all elements that process photons."""

    output_columns = []
    id_col = None

    def __init__(self, **kwargs):
        self.id_num = kwargs.pop('id_num', -9)
        if 'id_col' in kwargs:
            self.id_col = kwargs.pop('id_col')
        super().__init__(**kwargs)

    def add_output_cols(self, photons, colnames=[]):
        """Add float columns (NaN) and the id column (int, -1) if missing
        (reference base/base.py:249-286)."""
        for n in list(self.output_columns) + list(colnames):
            if (n is not None) and (n not in photons.colnames):
                if isinstance(n, dict):
                    spec = dict(n)
                    name = spec['name']
                    dtype = torch.int64 if spec.get('dtype', float) in (int, np.int64, 'int') else torch.float64
                    photons.new_column(name, dtype, fill=spec.get('value', 0))
                else:
                    photons.new_column(n, torch.float64, fill=float('nan'))
        if (self.id_col is not None) and (self.id_col not in photons.colnames):
            photons.new_column(self.id_col, torch.int64, fill=-1)

    def __call__(self, photons, *args, **kwargs):
        return self.process_photons(photons, *args, **kwargs)


def _parse_position_keywords(kwargs):
    """``pos4d`` xor ``position`` / ``orientation`` / ``zoom`` -> 4x4 affine."""
    pos4d = kwargs.pop('pos4d', None)
    if pos4d is None:
        position = kwargs.pop('position', np.zeros(3))
        orientation = kwargs.pop('orientation', np.eye(3))
        zoom = kwargs.pop('zoom', 1.)
        if np.isscalar(zoom):
            zoom = np.ones(3) * zoom
        if not len(zoom) == 3:
            raise ValueError('zoom must have three elements for x,y,z or be a scalar (global zoom).')
        if np.any(np.array(zoom) <= 0):
            raise ValueError('All values in zoom must be positive-definite to keep pos4d matrix valid. '
                             'Specify zero-thickness elements with display properties.')
        pos4d = affines.compose(position, orientation, zoom)
    else:
        if ('position' in kwargs) or ('orientation' in kwargs) or ('zoom' in kwargs):
            raise ValueError('If pos4d is specificed, the following keywords cannot be given at the '
                             'same time: position, orientation, zoom.')
        pos4d = np.array(pos4d, dtype=float)
    if np.linalg.det(pos4d) == 0:
        raise ValueError('pos4d matrix is invalid (determinant is 0).')
    return pos4d
