"""OpticalElement protocol on device tensors.

Public surface of the reference's marxs/optics/base.py: ``elem(photons)`` =
``geometry.intersect`` + ``process_photons`` (:144-211), the plug-in hook
``specific_process_photons`` returning ``{column: values for the hit photons}``,
``FlatStack`` (:218-281).

Built-in elements implement ``_lower_specific`` and execute inside the fused
kernel.  User subclasses that only provide the Python hook keep working: the
intersect runs as a kernel, the hook sees torch tensors on the device, and the
masked write-back follows the reference rules (probability multiplies and is
range-checked, ``None`` keys are ignored)."""
import numpy as np
import torch

from ..base import SimulationSequenceElement
from ..geometry import Geometry, FinitePlane, Cylinder
from ..program import NotFusable
from ..simulator import BaseContainer, run_fused

__all__ = ['OpticalElement', 'FlatOpticalElement', 'FlatStack', 'photonlocalcoords']

_PY_HOOKS = ('specific_process_photons', 'process_photons', 'process_photon', '__call__')


def _assign_col_value(photons, col, index, value):
    """Masked write-back with the probability rules of optics/base.py:12-51."""
    if col == 'probability':
        value = torch.as_tensor(value, device=photons.device, dtype=torch.float64)
        if bool(torch.any(value < 0)) or bool(torch.any(value > 1)):
            raise ValueError('Found probability outside of the 0..1 arange.')
        photons[col][index] *= value
    elif col is None:
        pass
    else:
        v = torch.as_tensor(value, device=photons.device)
        photons[col][index] = v.to(photons[col].dtype)


class OpticalElement(SimulationSequenceElement):
    """Base class for all optical elements."""

    default_geometry = FinitePlane
    display = {}

    def __init__(self, **kwargs):
        geometry = kwargs.pop('geometry', self.default_geometry)
        if isinstance(geometry, Geometry):
            self.geometry = geometry
        elif issubclass(geometry, Geometry):
            self.geometry = geometry(kwargs)
        super().__init__(**kwargs)
        if not hasattr(self, 'loc_coos_name'):
            self.loc_coos_name = self.geometry.loc_coos_name

    @property
    def pos4d(self):
        return self.geometry.pos4d

    # ---- fused path ---------------------------------------------------------
    def _can_lower(self):
        """True when no Python hook was overridden below the class that owns the lowering."""
        mro = type(self).__mro__
        owner = next((k for k, c in enumerate(mro) if '_lower_specific' in c.__dict__), None)
        if owner is None:
            return False
        for c in mro[:owner]:
            if any(h in c.__dict__ for h in _PY_HOOKS):
                return False
        return isinstance(self.geometry, self._lowerable_geometries)

    _lowerable_geometries = (FinitePlane,)

    def _lower_intersect(self, lw):
        if isinstance(self.geometry, Cylinder):
            lw.cylinder(self.pos4d, self.geometry.phi_lim)
        else:
            lw.plane(self.pos4d, circular=getattr(self.geometry, 'circular', False))

    def _lower(self, lw):
        self._lower_intersect(lw)
        self._lower_specific(lw)
        lw.commit(self.loc_coos_name, self.id_col, self.id_num)

    def __call__(self, photons):
        if self._can_lower():
            return run_fused([self], photons)
        return self._call_unfused(photons)

    # ---- generic (plug-in) path: kernel intersect + torch hook ------------------
    def _call_unfused(self, photons):
        intersect_out = self.geometry.intersect(photons['dir'].data, photons['pos'].data)
        return self.process_photons(photons, *intersect_out)

    def _process_photons_kernel(self, photons, intersect, interpos, intercoos):
        """``process_photons`` of a built-in element for an externally supplied intersect:
        the mask / interpos / intercoos are loaded by a LOADHIT op instead of a PLANE op."""
        names = ['_mxb_hit', '_mxb_ipx', '_mxb_ipy', '_mxb_ipz', '_mxb_l0', '_mxb_l1']
        ip = torch.as_tensor(interpos, device=photons.device).as_subclass(torch.Tensor)
        lc = torch.as_tensor(intercoos, device=photons.device).as_subclass(torch.Tensor)
        photons['_mxb_hit'] = torch.as_tensor(intersect, device=photons.device).to(torch.float64)
        for k in range(3):
            photons[names[1 + k]] = ip[:, k].contiguous()
        for k in range(2):
            photons[names[4 + k]] = lc[:, k].contiguous()
        try:
            run_fused([_LoadHitWrapper(self, names)], photons)
        finally:
            for n in names:
                if n in photons:
                    photons.remove_column(n)
        return photons

    def _run_scalar_hook(self, photons, intersect):
        """The scalar plug-in hook ``process_photon(dir, pos, energy, polarization)`` (reference
        optics/base.py:98-142, loop :186-197).  It is user Python written against numpy scalars and
        4-vectors, so it runs as in the reference: the intersect is a kernel, then the hook is called once
        per HIT photon on host copies of that photon's values, and its return values
        ``(dir, pos, energy, polarization, probability, *output_columns)`` are written back to the device
        columns with the reference's rules (probability multiplies and is range-checked,
        optics/base.py:12-51).  One D2H of the four input columns of the hit photons and one H2D per
        returned column - not one transfer per photon."""
        idx = torch.nonzero(torch.as_tensor(intersect, device=photons.device)).ravel()
        n_hit = int(idx.numel())
        names = _scalar_hook_names(self)
        self.add_output_cols(photons)
        host = {c: photons[c].data.as_subclass(torch.Tensor)[idx].cpu().numpy()
                for c in ('dir', 'pos', 'energy', 'polarization')}
        outs = None
        for k in range(n_hit):
            res = self.process_photon(host['dir'][k], host['pos'][k], host['energy'][k], host['polarization'][k])
            if len(res) != len(names):
                raise ValueError('process_photon returned {0} values, expected {1}: {2}'.format(
                    len(res), len(names), names))
            if outs is None:
                outs = [np.empty((n_hit,) + np.shape(r), dtype=float) for r in res]
            for o, r in zip(outs, res):
                o[k] = r
        for col, o in zip(names, outs or []):
            _assign_col_value(photons, col, idx, o)

    def process_photons(self, photons, intersect, interpos, intercoos):
        """Reference semantics of optics/base.py:149-211 on device tensors."""
        if self._can_lower():
            return self._process_photons_kernel(photons, intersect, interpos, intercoos)
        if int(intersect.sum()) == 0:
            return photons
        if hasattr(self, 'specific_process_photons'):
            outcols = self.specific_process_photons(photons, intersect, interpos, intercoos)
            self.add_output_cols(photons, list(outcols.keys()))
            for col in outcols:
                _assign_col_value(photons, col, intersect, outcols[col])
        elif hasattr(self, 'process_photon'):
            self._run_scalar_hook(photons, intersect)
        else:
            raise AttributeError('Optical element must have one of three: specific_process_photons, '
                                 'process_photon, or override process_photons.')
        if self.loc_coos_name is not None:
            self.add_output_cols(photons, self.loc_coos_name)
            photons[self.loc_coos_name[0]][intersect] = intercoos[intersect, 0]
            photons[self.loc_coos_name[1]][intersect] = intercoos[intersect, 1]
        if self.id_col is not None:
            photons[self.id_col][intersect] = self.id_num
        photons['pos'][intersect] = interpos[intersect]
        return photons


def _scalar_hook_names(elem):
    return ['dir', 'pos', 'energy', 'polarization', 'probability'] + [
        (c['name'] if isinstance(c, dict) else c) for c in elem.output_columns]


def photonlocalcoords(f, colnames=['pos', 'dir']):
    """Decorator for ``process_photons``-like methods of user elements that want to calculate in the LOCAL
    coordinate system of the element (reference optics/base.py:284-316): the homogeneous columns ``colnames``
    are multiplied by ``inv(self.pos4d)`` before the call and by ``self.pos4d`` after it, on the device."""
    from functools import wraps

    @wraps(f)
    def wrapper(self, photons, *args, **kwargs):
        inv = torch.as_tensor(np.linalg.inv(self.pos4d), dtype=torch.float64, device=photons.device)
        fwd = torch.as_tensor(np.asarray(self.pos4d, dtype=float), dtype=torch.float64, device=photons.device)
        for n in colnames:       # einsum('...ij,...j', M, v) = v @ M.T
            photons[n] = photons[n].data.as_subclass(torch.Tensor) @ inv.T
        photons = f(self, photons, *args, **kwargs)
        for n in colnames:
            photons[n] = photons[n].data.as_subclass(torch.Tensor) @ fwd.T
        return photons
    return wrapper


class _LoadHitWrapper:
    """Lowers ``elem`` behind a LOADHIT op (external intersect)."""

    def __init__(self, elem, names):
        self.elem, self.names = elem, names

    def _can_lower(self):
        return True

    def _lower(self, lw):
        lw.op('LOADHIT', cols=[lw.fcol(n) for n in self.names], pg=lw.params(
            __import__('marxs_b200.program', fromlist=['geom14']).geom14(self.elem.pos4d)))
        self.elem._lower_specific(lw)
        lw.commit(self.elem.loc_coos_name, self.elem.id_col, self.elem.id_num)


class FlatOpticalElement(OpticalElement):
    pass


class FlatStack(FlatOpticalElement, BaseContainer):
    """Several flat layers at the same position sharing ONE intersect (reference :218-281)."""

    def __init__(self, **kwargs):
        elements = kwargs.pop('elements')
        keywords = kwargs.pop('keywords')
        super().__init__(**kwargs)
        self.elements = []
        for elem, k in zip(elements, keywords):
            self.elements.append(elem(pos4d=self.pos4d, **k))

    def _lower_specific(self, lw):
        pass

    def _can_lower(self):
        return (OpticalElement._can_lower(self) and not self.preprocess_steps and not self.postprocess_steps
                and all(getattr(e, '_can_lower', lambda: False)() for e in self.elements))

    def _lower(self, lw):
        lw.plane(self.pos4d, circular=getattr(self.geometry, 'circular', False))
        lw.commit(self.loc_coos_name, self.id_col, self.id_num)
        for e in self.elements:
            # layers were built with the stack's pos4d (reference :254-256)
            e._lower_specific(lw)
            lw.commit(e.loc_coos_name, e.id_col, e.id_num)

    def __call__(self, photons):
        if self._can_lower():
            return run_fused([self], photons)
        return self._call_unfused(photons)

    def specific_process_photons(self, *args, **kwargs):
        return {}

    def process_photons(self, photons, intersect=None, interpos=None, intercoos=None):
        if int(intersect.sum()) > 0:
            photons = OpticalElement.process_photons(self, photons, intersect, interpos, intercoos)
            for e in self.elements:
                photons = e.process_photons(photons, intersect, interpos, intercoos)
        return photons
