"""Brewster-angle and multilayer mirrors (reference marxs/optics/multiLayerMirror.py)."""
import numpy as np

from .base import FlatOpticalElement, FlatStack

__all__ = ['FlatBrewsterMirror', 'MultiLayerEfficiency', 'MultiLayerMirror', 'read_table']


def read_table(filename):
    """Tab/whitespace separated text table with a header line -> dict of columns."""
    with open(filename) as f:
        lines = [l.rstrip('\n') for l in f if l.strip() and not l.startswith('#')]
    delim = '\t' if '\t' in lines[0] else None
    names = [n.strip() for n in lines[0].split(delim)]
    rows = [[c.strip() for c in l.split(delim)] for l in lines[1:]]
    out = {}
    for i, n in enumerate(names):
        col = [r[i] for r in rows]
        try:
            out[n] = np.array([float(c) for c in col])
        except ValueError:
            out[n] = np.array(col)
    return out


class FlatBrewsterMirror(FlatOpticalElement):
    """Flat mirror at the Brewster angle: only s-polarisation is reflected (reference :9-91)."""

    display = {'color': (0., 1., 0.), 'shape': 'box', 'box-half': '+x'}

    def _lower_specific(self, lw):
        P = self.pos4d
        Pinv = np.linalg.inv(P)
        ex = self.geometry['e_x']
        lw.op('BREWSTER', pf=lw.eparams(np.concatenate([Pinv[:3, :3].ravel(), P[:3, :3].ravel(), ex[:3]])))


class MultiLayerEfficiency(FlatOpticalElement):
    """Position-dependent multilayer reflectivity (reference :94-170).  The two text
    tables are read ONCE at construction (the reference re-reads them per call)."""

    def __init__(self, **kwargs):
        self.fileName = kwargs.pop('reflFile')
        self.polFile = kwargs.pop('testedPolarization')
        super().__init__(**kwargs)
        refl = self.fileName if isinstance(self.fileName, dict) else read_table(self.fileName)
        pol = self.polFile if isinstance(self.polFile, dict) else read_table(self.polFile)
        self._refl = {k: np.asarray(refl[k], dtype=float) for k in ('X(mm)', 'Peak lambda', 'Peak', 'FWHM(nm)')}
        self._pol = {k: np.asarray(pol[k], dtype=float) for k in ('Photon energy', 'Polarization')}

    def _lower_specific(self, lw):
        Ly = np.linalg.norm(self.geometry['v_y'])
        xs = self._refl['X(mm)'] / Ly - 1
        pe = self._pol['Photon energy'] / 1000
        lw.op('MLEFF', pf=lw.eparams(np.concatenate([
            [Ly, len(xs), len(pe)], xs, self._refl['Peak lambda'], self._refl['Peak'],
            self._refl['FWHM(nm)'], pe, self._pol['Polarization']])))


class MultiLayerMirror(FlatStack):
    """Brewster reflection + multilayer efficiency at one position (reference :173-179)."""

    def __init__(self, **kwargs):
        super().__init__(elements=[FlatBrewsterMirror, MultiLayerEfficiency],
                         keywords=[{}, {'reflFile': kwargs.pop('reflFile'),
                                        'testedPolarization': kwargs.pop('testedPolarization')}],
                         **kwargs)
