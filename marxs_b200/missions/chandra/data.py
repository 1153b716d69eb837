"""Chandra constants and design tables (numbers, not code).

Values follow the Chandra coordinate memo I (http://cxc.harvard.edu/contrib/jcm/ncoords.ps)
as used by the reference in ``marxs/missions/chandra/data.py:4-70``; the facet and
chip-corner tables were converted by ``tools/make_chandra_data.py``."""
import os

import numpy as np

NOMINAL_FOCALLENGTH = 10061.65

AIMPOINTS = {'ACIS-I': [-.782, 0, -233.592],
             'ACIS-S': [-.684, 0, -190.133],
             'HRC-I': [-1.040, 0, 126.985],
             'HRC-S': [-1.43, 0, 250.456]}

TDET = {'ACIS': {'version': 'ACIS-2.2',
                 'theta': np.deg2rad(np.array([90., 270., 90., 270., 0, 0, 0, 0, 0, 0])),
                 'scale': np.ones(10),
                 'handedness': np.ones(10),
                 'origin': np.array([[3061, 5131], [5131, 4107], [3061, 4085], [5131, 3061],
                                     [791, 1702], [1833, 1702], [2875, 1702], [3917, 1702],
                                     [4959, 1702], [6001, 1702]])}}
ODET = {'ACIS': [4096.5, 4096.5]}
PIXSIZE = {'ACIS': 0.492 / 3600.}
ACIS_name = ['I0', 'I1', 'I2', 'I3', 'S0', 'S1', 'S2', 'S3', 'S4', 'S5']

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def load_hess():
    """dict of columns of the HESS facet table: 'hessloc' (str) + 18 float columns."""
    path = os.path.join(_DATA, 'hess_facets.csv')
    lines = [l.strip() for l in open(path) if l.strip() and not l.startswith('#')]
    names = lines[0].split(',')
    rows = [l.split(',') for l in lines[1:]]
    out = {'hessloc': np.array([r[0] for r in rows])}
    for i, n in enumerate(names[1:], start=1):
        out[n] = np.array([float(r[i]) for r in rows])
    return out


def load_acis_corners():
    """(10, 4, 3) LSI corner coordinates, chips I0..S5, corners LL, LR, UR, UL."""
    path = os.path.join(_DATA, 'acis_corners_lsi.csv')
    lines = [l.strip() for l in open(path) if l.strip() and not l.startswith('#')][1:]
    out = np.zeros((10, 4, 3))
    corner_index = {'LL': 0, 'LR': 1, 'UR': 2, 'UL': 3}
    for l in lines:
        chip, corner, x, y, z = l.split(',')
        out[ACIS_name.index(chip), corner_index[corner]] = [float(x), float(y), float(z)]
    return out


def chip2tdet(chip, tdet, id_num):
    """CHIP -> TDET coordinates, eqn (5) of the Chandra coordinate memo (reference missions/chandra/data.py:169-190).
    The trace kernel applies the same transformation fused into the ACIS op; this standalone form takes an
    (N, 2) numpy array or torch tensor (any device) and returns the same kind."""
    import torch
    scale = tdet['scale'][id_num]
    handedness = tdet['handedness'][id_num]
    origin_tdet = tdet['origin'][id_num]
    theta = tdet['theta'][id_num]
    rotation = np.array([[np.cos(theta), np.sin(theta)],
                         [-np.sin(theta), np.cos(theta)]])
    if isinstance(chip, torch.Tensor):
        c = chip.as_subclass(torch.Tensor)
        rot = torch.as_tensor(rotation, dtype=c.dtype, device=c.device)
        org = torch.as_tensor(np.asarray(origin_tdet, dtype=float) + 0.5, dtype=c.dtype, device=c.device)
        return float(scale * handedness) * ((c - 0.5) @ rot.T) + org
    return scale * handedness * np.dot(rotation, (np.asarray(chip) - 0.5).T).T + (origin_tdet + 0.5)
