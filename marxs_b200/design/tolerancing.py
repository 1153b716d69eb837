"""Tolerancing drivers: change one aspect of an instrument, re-trace, take a figure of merit (reference
marxs/design/tolerancing.py:26-551) - SURVEY 8(f) rank 4, the heaviest callers of the trace path.

What is different on the GPU.  The reference copies the photon table and walks every element in
Python for each parameter set; here the photon list stays resident on the device and every parameter
set is ONE out-of-place launch (``simulator.trace_from``: the kernel reads the pristine list and writes
a work table, no copy).  Moving, tilting or re-ruling elements changes only NUMBERS of the lowered
program, never its structure, so the specialised kernel is compiled once per study and re-used for
every step (``libmxb`` keys its kernels on the op table); only the host-side
lowering (facet rows + culling grid) is redone.  Figures of merit (``marxs_b200.analysis.CaptureResAeff``)
are device reductions; per step only a few dozen scalars reach the host.

The element mutators keep the reference's names, signatures and error messages.  Plot helpers
(``WigglePlotter``) are not part of the trace path and are not provided.
"""
import collections.abc
import warnings
from collections import OrderedDict
from copy import copy
from functools import wraps

import numpy as np

from ..affines import compose, euler2mat
from ..simulator import Sequence, trace_from
from .uncertainties import generate_facet_uncertainty as genfacun

__all__ = ['oneormoreelements', 'wiggle', 'moveglobal', 'moveindividual', 'moveelem',
           'varyperiod', 'varyorderselector', 'varyattribute',
           'run_tolerances', 'run_tolerances_for_energies', 'run_tolerances_for_energies2',
           'generate_6d_wigglelist', 'reset_6d', 'select_1dof_changed', 'ResultTable']

HC_KEV_ANGSTROM = 12.398419843320026      # h c in keV Angstrom (CODATA 2018, what astropy's u.spectral() uses)


def oneormoreelements(func):
    """Let a function written for ONE element also take an iterable of elements (reference :26-42)."""
    @wraps(func)
    def func_wrapper(elements, *args, **kwargs):
        if isinstance(elements, collections.abc.Iterable):
            for e in elements:
                func(e, *args, **kwargs)
        else:
            func(elements, *args, **kwargs)
    return func_wrapper


def _shift(dx, dy, dz, rx, ry, rz):
    return compose([dx, dy, dz], euler2mat(rx, ry, rz, 'sxyz'), np.ones(3))


@oneormoreelements
def wiggle(e, dx=0, dy=0, dz=0, rx=0., ry=0., rz=0.):
    """Misalign every facet of a Parallel independently: Gaussian sigmas in mm / rad (reference :45-60)."""
    e.elem_uncertainty = genfacun(len(e.elements), [dx, dy, dz], [rx, ry, rz])
    e.generate_elements()


@oneormoreelements
def moveglobal(e, dx=0, dy=0, dz=0, rx=0., ry=0., rz=0.):
    """Move / rotate the origin of a whole Parallel (reference :63-79)."""
    e.uncertainty = _shift(dx, dy, dz, rx, ry, rz)
    e.generate_elements()


@oneormoreelements
def moveindividual(e, dx=0, dy=0, dz=0, rx=0, ry=0, rz=0):
    """Move / rotate every facet of a Parallel by the SAME amount about its own centre (reference :82-104)."""
    e.elem_uncertainty = [_shift(dx, dy, dz, rx, ry, rz)] * len(e.elements)
    e.generate_elements()


@oneormoreelements
def moveelem(e, dx=0, dy=0, dz=0, rx=0., ry=0., rz=0.):
    """Move / rotate one element (not a container) relative to where it was FIRST (kept as
    ``geometry.pos4d_orig``), so repeated calls do not accumulate (reference :107-133)."""
    if not hasattr(e.geometry, 'pos4d_orig'):
        e.geometry.pos4d_orig = e.geometry.pos4d.copy()
    e.geometry.pos4d = _shift(dx, dy, dz, rx, ry, rz) @ e.geometry.pos4d_orig


@oneormoreelements
def varyattribute(element, **kwargs):
    """Set existing attributes of an element; unknown names raise (reference :136-173)."""
    for key, val in kwargs.items():
        if not hasattr(element, key):
            raise ValueError('Object {0} does not have {1} attribute.'.format(element, key))
        setattr(element, key, val)


@oneormoreelements
def varyperiod(element, period_mean, period_sigma):
    """Draw a new grating constant from N(period_mean, period_sigma) [mm] for each grating (reference :176-201)."""
    if not hasattr(element, '_d'):
        raise ValueError('Object {0} does not have grating period `_d` attribute.'.format(element))
    element._d = np.random.normal(period_mean, period_sigma)


@oneormoreelements
def varyorderselector(element, order_selector, *args, **kwargs):
    """Replace the order selector of a grating by ``order_selector(*args, **kwargs)`` (reference :204-221)."""
    if not hasattr(element, 'order_selector'):
        raise ValueError('Object {0} does not have an order_selector attribute.'.format(element))
    element.order_selector = order_selector(*args, **kwargs)


def run_tolerances(photons_in, instrum, wigglefunc, wiggleparts, parameters, analyzefunc, verbose=False):
    """For every dict in ``parameters``: ``wigglefunc(wiggleparts, **pars)``, trace ``photons_in`` (left
    untouched) through ``instrum``, and merge ``analyzefunc(photons)`` into a COPY of the dict
    (reference :224-283).  Returns the list of result dicts."""
    out = []
    for i, pars in enumerate(parameters):
        if verbose:
            print('Working on simulation {0}/{1}'.format(i, len(parameters)))
        wigglefunc(wiggleparts, **pars)
        # one launch, no copy of the input; a fresh table per step because a step in which nothing hits an
        # element must not inherit that element's columns from the step before
        photons = trace_from(instrum, photons_in)
        cpars = copy(pars)
        cpars.update(analyzefunc(photons))
        out.append(cpars)
        del photons
    return out


class ResultTable(OrderedDict):
    """Column store for tolerancing results: name -> numpy array with one entry (or one row) per run.
    ``len()`` is the number of runs; a boolean / integer array index selects runs (the subset of
    astropy.table.Table the tolerancing helpers rely on)."""

    @classmethod
    def from_rows(cls, rows):
        out = cls()
        if rows:
            for k in rows[0]:
                vals = [r[k] for r in rows]
                try:
                    out[k] = np.ma.stack(vals) if any(isinstance(v, np.ma.MaskedArray) for v in vals) else np.array(vals)
                except ValueError:
                    col = np.empty(len(vals), dtype=object)
                    col[:] = vals
                    out[k] = col
        return out

    @classmethod
    def vstack(cls, tables):
        out = cls()
        for k in (tables[0] if tables else ()):
            out[k] = np.ma.concatenate([t[k] for t in tables]) if any(isinstance(t[k], np.ma.MaskedArray) for t in tables) \
                else np.concatenate([t[k] for t in tables])
        return out

    @property
    def colnames(self):
        return list(self.keys())

    def __len__(self):
        for v in self.values():
            return len(v)
        return 0

    def __getitem__(self, key):
        if isinstance(key, str):
            return OrderedDict.__getitem__(self, key)
        sub = type(self)()
        for k, v in self.items():
            sub[k] = v[key]
        return sub


def _energies_kev(energies):
    if hasattr(energies, 'to'):                     # an astropy Quantity, if the caller has astropy
        energies = energies.to('keV').value
    return np.atleast_1d(np.asarray(energies, dtype=float))


def run_tolerances_for_energies(source, energies, instrum_before, instrum_remaining, wigglefunc, wiggleparts,
                                parameters, analyzefunc, reset=None, t_source=1.):
    """Loop ``run_tolerances`` over monoenergetic photon lists (reference :286-367).

    For every energy [keV] the source (flux 1, so ``t_source`` seconds = number of photons) is set to that
    energy, photons are generated and run ONCE through ``instrum_before``; the list then stays on the device
    for all parameter sets.  ``reset`` (a parameter dict) restores the wiggled parts at the end.  Returns a
    ResultTable with one row per (energy, parameter set): the parameters, the results, 'energy' [keV] and
    'wave' [Angstrom]."""
    energies = _energies_kev(energies)
    tabs = []
    for e in energies:
        source.energy = float(e)
        photons_in = source.generate_photons(t_source)
        photons_in = instrum_before(photons_in)
        data = run_tolerances(photons_in, instrum_remaining, wigglefunc, wiggleparts, parameters, analyzefunc)
        tab = ResultTable.from_rows(data)
        tab['energy'] = np.full(len(data), float(e))
        tab['wave'] = np.full(len(data), HC_KEV_ANGSTROM / float(e))
        tabs.append(tab)
    if reset is not None:
        wigglefunc(wiggleparts, **reset)
    return ResultTable.vstack(tabs)


def run_tolerances_for_energies2(source, energies, instrum, cls, wigglefunc, parameters, analyzefunc, reset=None,
                                 subclass_ok=False, t_source=1.):
    """As ``run_tolerances_for_energies``, splitting ``instrum`` at its first top-level element of class
    ``cls``; all elements of that class are wiggled (reference :379-451)."""
    ind = instrum.first_of_class_top_level(cls, subclass_ok)
    if ind is None:
        raise Exception('{0} not part of {1}'.format(cls, instrum))
    return run_tolerances_for_energies(source, energies, Sequence(elements=instrum.elements[:ind]),
                                       Sequence(elements=instrum.elements[ind:]), wigglefunc,
                                       instrum.elements_of_class(cls), parameters, analyzefunc, reset, t_source=t_source)


def generate_6d_wigglelist(trans, rot, names=['dx', 'dy', 'dz', 'rx', 'ry', 'rz']):
    """Parameter dicts that step ONE of six degrees of freedom at a time (reference :454-516).

    trans [mm] and rot [rad] are the step lists (first entry 0; astropy Quantities are converted).  Returns
    (both signs - for ``moveglobal`` / ``moveindividual``, positive side only - for ``wiggle``), each a list of
    dicts sorted by value with the all-zero entry once."""
    trans = np.asarray(trans.to('mm').value if hasattr(trans, 'to') else trans, dtype=float)
    rot = np.asarray(rot.to('rad').value if hasattr(rot, 'to') else rot, dtype=float)
    if trans[0] != 0 or rot[0] != 0:
        warnings.warn('First element of trans and rot should be 0.')
    steps = []
    for axis in range(6):
        for v in (trans if axis < 3 else rot):
            row = np.zeros(6)
            row[axis] = v
            steps.append(row)
    one_sided = np.array(steps)
    two_sided = np.unique(np.vstack([one_sided, -one_sided]), axis=0)
    one_sided = np.unique(one_sided, axis=0)
    return [dict(zip(names, row)) for row in two_sided], [dict(zip(names, row)) for row in one_sided]


reset_6d = {'dx': 0., 'dy': 0., 'dz': 0., 'rx': 0., 'ry': 0., 'rz': 0.}
"""The neutral element of a 6d wigglelist: resets all uncertainties."""


def select_1dof_changed(table, par, parlist=['dx', 'dy', 'dz', 'rx', 'ry', 'rz']):
    """Rows of a result table in which every parameter of ``parlist`` other than ``par`` is zero (reference :523-550)."""
    ind = np.ones(len(table), dtype=bool)
    for p in set(parlist) - {par}:
        ind &= np.asarray(table[p]) == 0
    return table[ind]
