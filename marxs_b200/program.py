"""Lowering of MARXS element trees to the libmxb program blob, and its launcher.

An element tree (Sequence / FlatStack / Parallel / single elements) is *lowered*
into a flat list of ops plus parameter blocks (include/mxb.h documents the
encoding).  One program = one kernel launch = one pass of the photon batch over
HBM.  The host never touches photon data: lowering only packs per-element
constants (geometry from ``pos4d``, grating vectors, selector tables, culling
grids), which is what the reference recomputes on every ``geometry[...]``
access (marxs/math/geometry.py:108-160).

Geometry is mutable between calls in MARXS (optics/tests/test_grating.py:313-337)
so programs are re-lowered per call and cached by the hash of their blob.
"""
import ctypes
from collections import OrderedDict

import os

import numpy as np
import torch

from . import _lib

# ---- constants mirrored from include/mxb.h (checked by tests/test_abi.py) ----
MXB_MAGIC = 0x4D5842
HEADER_WORDS = 16
OP_WORDS = 16
MAX_OPS = 48
COL_INIT = 1024
FIRST_OUT = 11
CORE_F64 = ['pos_x', 'pos_y', 'pos_z', 'dir_x', 'dir_y', 'dir_z', 'pol_x', 'pol_y', 'pol_z',
            'energy', 'probability']

OP = dict(END=0, PLANE=1, COMMIT=2, ARRAY_BEGIN=3, ARRAY_END=4, BAFFLE=5, LENS=6, RSCATTER=7,
          GSCATTER=8, FILTER=9, GRATING=10, DETPIX=11, ACIS=12, BREWSTER=13, MLEFF=14, APERTURE=15,
          PROPAGATE=16, GFILTER=17, LOADHIT=18, QFACTOR=19, L2ABS=20, CYLINDER=21, GENERATE=22, POINTING=23,
          LABCONE=24, FARLAB=25)
SEL_ORDERSELECTOR, SEL_EFFFILE, SEL_INTERPTABLE = 1, 2, 3
ARRAY_HEADER_WORDS = 24
MAX_STAGE_BYTES = 200 * 1024
HIT_OPS = (OP['PLANE'], OP['APERTURE'], OP['ARRAY_BEGIN'], OP['LOADHIT'], OP['CYLINDER'])


class NotFusable(Exception):
    """The element cannot be expressed as program ops (caller runs it unfused)."""


class UnsupportedCallable(TypeError):
    """A user callable sits on the device path and cannot be lowered to a table."""


def geom14(pos4d):
    """[c, e_x, e_y, e_z, |v_y|, |v_z|] from pos4d exactly like Geometry.__getitem__
    (reference math/geometry.py:143-160: e = v / np.linalg.norm(v))."""
    p = np.asarray(pos4d, dtype=float)
    if p.shape != (4, 4) or not np.all(p[3] == [0., 0., 0., 1.]):
        raise ValueError('pos4d must be an affine 4x4 matrix with last row (0, 0, 0, 1)')
    out = np.empty(14)
    out[0:3] = p[:3, 3]
    for k in range(3):
        v = p[:, k]
        out[3 + 3 * k: 6 + 3 * k] = (v / np.linalg.norm(v))[:3]
    out[12] = np.linalg.norm(p[:, 1])
    out[13] = np.linalg.norm(p[:, 2])
    return out


# Verification switch (tests): lower every facet array WITHOUT culling grid and disjointness certificate, i.e. every
# photon is tested against every facet in order like the reference's loop.  Same arithmetic per facet, so the results
# must equal the culled search bit for bit.  Part of the plan-cache key (simulator._lower_run, source._run_born).
EXHAUSTIVE_SEARCH = False


def build_cull_grid(G, tan_max=0.12, max_cells=4096, hops=2):
    """Conservative 2-D culling grid for an array of rectangles.

    G : (F, 14) facet geometry rows.  Returns None when a grid is not worthwhile
    (few facets) or not valid (normals do not share a hemisphere).

    A photon is projected along its ray onto the reference plane (O, nbar); a
    facet is listed in every cell its footprint, inflated by ``margin``, touches.
    With H = max height of any facet point above the plane, a ray inside the cone
    tan(theta) <= margin / (3 H) cannot hit a facet that is not listed in its
    cell, even after ONE redirection inside the same cone (H t + 2 H t' <= margin).
    The kernel checks the cone per photon and falls back to the footprint scan outside.

    ``hops`` = 2 covers that redirection (margin = 3 H tan_max); ``hops`` = 0 (margin = H tan_max, a third of the
    inflation) is valid for the INCOMING ray only and is used when the disjointness certificate guarantees that a photon
    never continues through its cell list after a hit (``single_hit_successors`` with a cone at least this wide).
    """
    F = G.shape[0]
    if F < 3:
        return None
    c, ex, ey, ez = G[:, 0:3], G[:, 3:6].copy(), G[:, 6:9], G[:, 9:12]
    Ly, Lz = G[:, 12], G[:, 13]
    flip = (ex @ ex[0]) < 0
    ex[flip] *= -1
    nbar = ex.mean(axis=0)
    if np.linalg.norm(nbar) < 0.7:
        return None
    nbar /= np.linalg.norm(nbar)
    O = c.mean(axis=0)
    a = np.eye(3)[np.argmin(np.abs(nbar))]
    u = a - np.dot(a, nbar) * nbar
    u /= np.linalg.norm(u)
    v = np.cross(nbar, u)
    corners = np.stack([c + sy * Ly[:, None] * ey + sz * Lz[:, None] * ez
                        for sy in (-1, 1) for sz in (-1, 1)], axis=1)      # (F, 4, 3)
    rel = corners - O
    h = rel @ nbar
    pu, pv = rel @ u, rel @ v
    H = float(np.abs(h).max())
    extent = float(max(np.ptp(pu), np.ptp(pv), 1e-12))
    slack = 1e-6 + 1e-9 * extent
    margin = (1.0 + hops) * H * tan_max + slack
    T2 = min(((margin - slack) / ((1.0 + hops) * H)) ** 2, 1e12) if H > 0 else 1e12
    lo_u, hi_u = pu.min(axis=1) - margin, pu.max(axis=1) + margin
    lo_v, hi_v = pv.min(axis=1) - margin, pv.max(axis=1) + margin
    u0, v0 = float(lo_u.min()), float(lo_v.min())
    wu, wv = float(hi_u.max()) - u0, float(hi_v.max()) - v0
    size = np.median(np.minimum(np.ptp(pu, axis=1), np.ptp(pv, axis=1)))
    cell = max(0.5 * float(size), np.sqrt(wu * wv / max_cells), 1e-9)
    nu, nv = int(np.floor(wu / cell)) + 1, int(np.floor(wv / cell)) + 1
    inv_cell = 1.0 / cell
    lists = [[] for _ in range(nu * nv)]
    # (the margin carries 1e-6 mm of slack, far above the rounding of this arithmetic, which the
    #  kernel repeats in the same order: no extra ring of cells is needed)
    iu0 = np.clip(np.floor((lo_u - u0) * inv_cell).astype(int), 0, nu - 1)
    iu1 = np.clip(np.floor((hi_u - u0) * inv_cell).astype(int), 0, nu - 1)
    iv0 = np.clip(np.floor((lo_v - v0) * inv_cell).astype(int), 0, nv - 1)
    iv1 = np.clip(np.floor((hi_v - v0) * inv_cell).astype(int), 0, nv - 1)
    for j in range(F):
        for iv in range(iv0[j], iv1[j] + 1):
            base = iv * nu
            for iu in range(iu0[j], iu1[j] + 1):
                lists[base + iu].append(j)
    start = np.zeros(nu * nv + 1, dtype=np.int32)
    start[1:] = np.cumsum([len(l) for l in lists])
    cand = np.array([j for l in lists for j in l], dtype=np.int32)
    return dict(O=O, nbar=nbar, u=u, v=v, u0=u0, v0=v0, inv_cell=inv_cell, nu=nu, nv=nv,
                start=start, cand=cand, T2=T2, margin=margin, H=H,
                mean_candidates=float(np.mean([len(l) for l in lists if l])) if len(cand) else 0.)


def facet_reach_limits(G, grid, t_cap=1.0):
    """Sparse reachability of a facet array: arrays (A, B, L) over the pairs A < B with L <= t_cap, where L[A, B] is a
    lower bound of the tangent of the angle to the array's mean normal a ray needs to get from ANY point of facet A to
    facet B.  (Only later facets matter: the reference loops the facets in order and never returns to an earlier one,
    simulator.py:42-49.)

    The footprints of A and B on the mean plane are separated by gap_AB (the best of the four edge-normal axes of A's
    quadrilateral: a lower bound of their distance), their heights above the plane differ by at most h_AB, and a ray
    moves h tan(theta) sideways per height h: L = gap_AB / h_AB, shrunk by 1e-9 for rounding; 0 where the footprints
    overlap.  Pairs whose centres are further apart than the two half-diagonals plus t_cap times the array's height
    range cannot have L <= t_cap and are never looked at (blocks of 512 facets against all centres).
    Kernel side: csrc/mxb_ops.cuh array_revalidate (successor lists / single-hit stop)."""
    F = G.shape[0]
    c, ey, ez = G[:, 0:3], G[:, 6:9], G[:, 9:12]
    Ly, Lz = G[:, 12], G[:, 13]
    corners = np.stack([c + sy * Ly[:, None] * ey + sz * Lz[:, None] * ez
                        for sy, sz in ((-1, -1), (1, -1), (1, 1), (-1, 1))], axis=1)          # cyclic order
    rel = corners - grid['O']
    h = rel @ grid['nbar']
    q = np.stack([rel @ grid['u'], rel @ grid['v']], axis=2)                                  # (F, 4, 2)
    edges = np.roll(q, -1, axis=1) - q
    nrm = np.stack([edges[..., 1], -edges[..., 0]], axis=2)
    length = np.linalg.norm(nrm, axis=2, keepdims=True)
    ctr = q.mean(axis=1)
    rad = np.linalg.norm(q - ctr[:, None, :], axis=2).max(axis=1)
    reach = 2. * rad.max() + t_cap * (h.max() - h.min())
    idx = np.arange(F)
    pa, pb = [], []
    for a0 in range(0, F, 512):
        a1 = min(F, a0 + 512)
        d2 = ((ctr[a0:a1, None, :] - ctr[None, :, :]) ** 2).sum(axis=2)
        ia, ib = np.nonzero((d2 <= reach * reach) & (idx[None, :] > idx[a0:a1, None]))
        pa.append(ia + a0)
        pb.append(ib)
    pa, pb = np.concatenate(pa), np.concatenate(pb)
    if pa.size == 0:
        return pa, pb, np.zeros(0)
    if not np.all(length > 0):
        return pa, pb, np.zeros(pa.size)                      # degenerate footprint (facet edge-on): nothing is known
    nrm = nrm / length
    hlo, hhi = h.min(axis=1), h.max(axis=1)
    L = np.empty(pa.size)
    for p0 in range(0, pa.size, 65536):
        A, B = pa[p0:p0 + 65536], pb[p0:p0 + 65536]
        other = np.einsum('pkx,pmx->pkm', nrm[A], q[B])       # B's corners on A's four axes
        own = np.einsum('pkx,pmx->pkm', nrm[A], q[A])
        sep = np.maximum(other.min(axis=2) - own.max(axis=2), own.min(axis=2) - other.max(axis=2)).max(axis=1)
        gap = np.maximum(sep, 0.)
        hd = np.maximum(hhi[A] - hlo[B], hhi[B] - hlo[A])
        with np.errstate(divide='ignore', invalid='ignore'):
            L[p0:p0 + 65536] = np.where(hd > 0, gap / hd, np.inf)
    L = np.where(np.isfinite(L), L * (1. - 1e-9), L)
    keep = L <= t_cap
    return pa[keep], pb[keep], L[keep]


def single_hit_successors(G, grid, t_big=1.0, max_succ=8):
    """Disjointness certificate of a facet array.  Returns (t2, start, succ):

    * within the cone tan^2 <= t2 around the mean normal, a photon that leaves facet A can only reach the LATER facets
      succ[start[A]:start[A + 1]] (ascending) - none for most facets of a tiled array, the overlapping diagonal
      neighbours of a ring-placed one.  The kernel tests exactly those instead of the rest of the candidate list or the
      footprint scan; results equal the reference's loop over all facets by construction.
    * succ is None when no facet has a successor inside the cone (t2 is then the smallest pair limit, at most t_big^2).
    * (0., None, None): nothing can be certified."""
    F = G.shape[0]
    A, B, L = facet_reach_limits(G, grid, t_cap=t_big)
    t_min = float(L.min()) if L.size else t_big
    if t_min >= 0.3:                        # no pair closer than 17 degrees: one number for the array, no lists
        return min(t_min, t_big) ** 2, None, None
    for t in (t_big, 0.3):
        sel = L <= t
        count = np.bincount(A[sel], minlength=F)
        if count.max() <= max_succ:
            order = np.lexsort((B[sel], A[sel]))              # by facet, ascending successors
            start = np.zeros(F + 1, dtype=np.int64)
            start[1:] = np.cumsum(count)
            return t * t, start, B[sel][order]
    return 0., None, None


def _pack_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if len(a) % 2:
        a = np.concatenate([a, np.zeros(1, np.int32)])
    return a.view(np.float64)


class Lowering:
    """Accumulates ops and parameter blocks for one program."""

    def __init__(self, existing_columns=(), meta=None):
        self.existing = set(existing_columns)
        self.meta = meta if meta is not None else {}
        self.ops = []                     # dicts
        self.small = [np.zeros(HEADER_WORDS + MAX_OPS * OP_WORDS)]
        self.small_words = HEADER_WORDS + MAX_OPS * OP_WORDS
        self.big = []                     # global-only tables (offset fix-ups at finish)
        self.fixups = []                  # (small word index, big block index)
        self._dedupe = {}
        self.f64_cols = list(CORE_F64)
        self.i64_cols = []
        self.slot_kinds = []
        self.new_cols = OrderedDict()     # name -> dict(kind, creator)
        self.initialised = set()
        self.creator = -1                 # op index of the latest hit-defining op
        self.array = None                 # active array context
        self.meta_updates = OrderedDict()
        self.needs_pos = False            # an aperture creates the pos column
        self.aux = OrderedDict()          # name -> device tensor reached through an f64 pointer slot
        self._no_fold = None
        self.last_order_col = None        # column written by the latest GRATING op of the current element stack
        self.born = False                 # a source op creates the photons: header flag, no core planes are read
        self.rel_fixups = []              # (word, base word, delta): word = value(base word) + delta at finish
        self.image = None                 # (tensor, sel_lo) of the detector array being lowered

    # ---- transactions (an element that turns out not to be fusable is rolled back) ----
    def checkpoint(self):
        if self.array is not None:
            raise RuntimeError('checkpoint inside an array')
        return dict(n_ops=len(self.ops), n_small=len(self.small), small_words=self.small_words,
                    n_big=len(self.big), n_fix=len(self.fixups), dedupe=dict(self._dedupe),
                    f64=list(self.f64_cols), i64=list(self.i64_cols), slots=list(self.slot_kinds),
                    new_cols=OrderedDict(self.new_cols), init=set(self.initialised), creator=self.creator,
                    meta_updates=OrderedDict(self.meta_updates), aux=OrderedDict(self.aux))

    def rollback(self, cp):
        del self.ops[cp['n_ops']:]
        del self.small[cp['n_small']:]
        del self.big[cp['n_big']:]
        del self.fixups[cp['n_fix']:]
        self.small_words = cp['small_words']
        self._dedupe = cp['dedupe']
        self.f64_cols, self.i64_cols, self.slot_kinds = cp['f64'], cp['i64'], cp['slots']
        self.new_cols, self.initialised, self.creator = cp['new_cols'], cp['init'], cp['creator']
        self.meta_updates = cp['meta_updates']
        self.aux = cp['aux']
        self.array = None

    # ---- parameters -------------------------------------------------------
    def params(self, values):
        """Store a block of doubles (2-word aligned, content-deduplicated) -> word offset."""
        arr = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
        key = arr.tobytes()
        if key in self._dedupe:
            return self._dedupe[key]
        if len(arr) % 2:
            arr = np.concatenate([arr, [0.]])
        off = self.small_words
        self.small.append(arr)
        self.small_words += len(arr)
        self._dedupe[key] = off
        return off

    def params_with_table(self, values, index, table):
        """Like ``params`` but ``values[index]`` receives the absolute offset of
        ``table``, a large array kept in global memory only (patched at finish)."""
        tab = np.ascontiguousarray(table, dtype=np.float64).ravel()
        vals = np.array(values, dtype=np.float64).ravel()
        vals[index] = -1.0
        key = (vals.tobytes(), hash(tab.tobytes()), tab.size)
        if key in self._dedupe:
            return self._dedupe[key]
        self.big.append(tab)
        vals[index] = -1000.0 - (len(self.big) - 1)      # unique placeholder keeps de-duplication honest
        off = self.params(vals)
        self.fixups.append((off + index, len(self.big) - 1))
        self._dedupe[key] = off
        return off

    def fixup_relative(self, word, base_word, delta):
        """At finish: blob[word] = blob[base_word] + delta (second table inside one global block)."""
        self.rel_fixups.append((int(word), int(base_word), int(delta)))

    def input_col(self, name):
        """f64 slot of a column an op READS (plain index, like LOADHIT's): it must exist already or be
        created earlier in this program."""
        if name not in self.f64_cols:
            if name not in self.existing:
                raise NotFusable('input column {0} does not exist'.format(name))
            if len(self.f64_cols) >= _lib.MXB_MAX_F64_COLS:
                raise NotFusable('too many output columns in one program')
            self.f64_cols.append(name)
        return self.f64_cols.index(name)

    # ---- draw slots -----------------------------------------------------------
    def slot(self, kind):
        a = self.array
        if a is not None and a['facet'] > 0:
            s = a['slots'][a['slot_cursor']]
            a['slot_cursor'] += 1
            return s
        if len(self.slot_kinds) >= _lib.MXB_MAX_SLOTS:
            raise NotFusable('too many random-draw slots in one program')
        self.slot_kinds.append(kind)
        s = len(self.slot_kinds) - 1
        if a is not None:
            a['slots'].append(s)
        return s

    # ---- columns ----------------------------------------------------------------
    def _col(self, name, kind):
        a = self.array
        if a is not None and a['facet'] > 0:
            # facets after the first replay the column references of the template facet
            ref = a['colrefs'][a['col_cursor']]
            a['col_cursor'] += 1
            return ref
        ref = self._col_new(name, kind)
        if a is not None:
            a['colrefs'].append(ref)
        return ref

    def _col_new(self, name, kind):
        if name is None:
            return -1
        table = self.f64_cols if kind == 'f' else self.i64_cols
        if name in table:
            idx = table.index(name)
        else:
            limit = _lib.MXB_MAX_F64_COLS if kind == 'f' else _lib.MXB_MAX_I64_COLS
            if len(table) >= limit:
                raise NotFusable('too many output columns in one program')
            table.append(name)
            idx = len(table) - 1
        if name not in self.existing and name not in self.new_cols:
            self.new_cols[name] = dict(kind=kind, creator=self.creator)
        if name not in self.existing and name not in self.initialised:
            self.initialised.add(name)
            if self.array is not None:
                # the body's first pass initialises the column; the list serves warps whose
                # lanes all miss the array (the body is skipped for them)
                self.array['init'].append(idx if kind == 'f' else -idx - 2)
            return idx + COL_INIT
        return idx

    def fcol(self, name):
        return self._col(name, 'f')

    def aux_ptr(self, tensor):
        """Pointer slot (in the f64 table) for a device buffer that is not a photon column,
        e.g. a detector image accumulated by the kernel."""
        for name, t in self.aux.items():
            if t is tensor:
                return self.f64_cols.index(name)
        name = '_aux{0}'.format(len(self.aux))
        if len(self.f64_cols) >= _lib.MXB_MAX_F64_COLS:
            raise NotFusable('too many output columns in one program')
        self.f64_cols.append(name)
        self.aux[name] = tensor
        return len(self.f64_cols) - 1

    def icol(self, name):
        return self._col(name, 'i')

    # ---- ops ----------------------------------------------------------------------
    def op(self, type_, flags=0, pg=-1, pf=-1, cols=(), s0=-1, s1=-1, w14=0, w15=0):
        cols = list(cols) + [-1] * (8 - len(cols))
        rec = dict(type=OP[type_] if isinstance(type_, str) else type_, flags=int(flags), pg=int(pg),
                   pf=int(pf), cols=[int(c) for c in cols], s0=int(s0), s1=int(s1), w14=int(w14),
                   w15=int(w15))
        a = self.array
        if a is not None and a['facet'] > 0:
            tmpl = a['body'][a['op_cursor']] if a['op_cursor'] < len(a['body']) else None
            a['op_cursor'] += 1
            mask = ~(self.COMMIT_BIT | self.COMMIT_ROW_ID)
            same = (tmpl is not None and rec['type'] == tmpl['type'] and rec['flags'] == (tmpl['flags'] & mask)
                    and all(rec[k] == tmpl[k] for k in ('pg', 'pf', 's0', 's1'))
                    and (rec['cols'][:5] == tmpl['cols'][:5]))
            folded = tmpl is not None and bool(tmpl['flags'] & self.COMMIT_BIT)
            if same and not folded:
                same = rec['cols'] == tmpl['cols'] and rec['w14'] == tmpl['w14']
            if not same:
                raise NotFusable('facets of a Parallel differ in more than per-facet parameters')
            return -1
        if len(self.ops) >= MAX_OPS:
            raise NotFusable('program too long')
        self.ops.append(rec)
        if a is not None:
            a['body'].append(rec)
        if rec['type'] in HIT_OPS:
            self.creator = len(self.ops) - 1
        return len(self.ops) - 1

    def plane(self, pos4d, circular=False):
        """Start a single-plane element: intersect (math/geometry.py:211-261)."""
        if self.array is not None:
            # inside a facet the geometry comes from the facet row
            a = self.array
            if a['geom_taken']:
                raise NotFusable('only one intersect per facet')
            a['geom_taken'] = True
            a['geom'] = geom14(pos4d)          # always the first 14 words of the facet row
            if circular:
                raise NotFusable('arrays of circular elements are not supported')
            return
        self.op('PLANE', flags=1 if circular else 0, pg=self.params(geom14(pos4d)))
        self.last_order_col = None

    def cylinder(self, pos4d, phi_lim):
        """Start an element on a Cylinder geometry (math/geometry.py:470-564)."""
        if self.array is not None:
            raise NotFusable('arrays of cylinders are not supported')
        p = np.asarray(pos4d, dtype=float)
        twopi = 2 * np.pi
        b1 = (twopi + (float(phi_lim[0]) % twopi)) % twopi      # angle_between, math/utils.py:208-210
        b2 = (twopi + (float(phi_lim[1]) % twopi)) % twopi
        # zoom_z as transforms3d.affines.decompose44 gives it: norm of the third column of the 3x3 part
        from .affines import decompose44
        zoom = decompose44(p)[2]
        self.op('CYLINDER', pg=self.params(np.concatenate([np.linalg.inv(p).ravel(), p.ravel(), [b1, b2, zoom[2]]])))
        self.last_order_col = None

    HOISTABLE = tuple(OP[t] for t in ('QFACTOR', 'L2ABS', 'GSCATTER', 'LENS', 'RSCATTER', 'BREWSTER', 'PROPAGATE'))

    FOLDABLE = ('PLANE', 'LENS', 'RSCATTER', 'GSCATTER', 'FILTER', 'GRATING', 'DETPIX', 'BREWSTER', 'MLEFF')
    COMMIT_BIT, COMMIT_ROW_ID = 256, 512

    def commit(self, loc_names, id_col, id_num):
        """loc-coos columns, id column, pos = interpos (optics/base.py:201-209).  Folded into the
        element's own op when that op has its last three column slots free (one dispatch less)."""
        c0 = self.fcol(loc_names[0]) if loc_names is not None else -1
        c1 = self.fcol(loc_names[1]) if loc_names is not None else -1
        c2 = self.icol(id_col) if id_col is not None else -1
        a = self.array
        if a is not None:
            if id_col is not None:
                a['id_num'] = id_num                 # layers of a FlatStack carry no id column of their own
            if a['facet'] > 0:                       # replay the template facet's decision
                folded = a['folds'][a['fold_cursor']]
                a['fold_cursor'] += 1
                if not folded:
                    self.op('COMMIT', flags=1, cols=[c0, c1, c2], w15=0)
                return
        prev = self.ops[-1] if self.ops else None
        foldable = (prev is not None and prev is not self._no_fold
                    and prev['type'] in [OP[t] for t in self.FOLDABLE]
                    and prev['cols'][5:8] == [-1, -1, -1] and not (prev['flags'] & self.COMMIT_BIT)
                    and (a is None or (a['body'] and a['body'][-1] is prev)))
        if a is not None:
            a['folds'].append(bool(foldable))
        if foldable:
            prev['cols'][5:8] = [c0, c1, c2]
            prev['flags'] |= self.COMMIT_BIT | (self.COMMIT_ROW_ID if a is not None else 0)
            if a is None:
                prev['w14'] = id_num
            return
        if a is not None:
            self.op('COMMIT', flags=1, cols=[c0, c1, c2], w15=0)   # w15 patched to the row's id offset
        else:
            self.op('COMMIT', cols=[c0, c1, c2], w14=id_num)

    # ---- per-facet parameters -------------------------------------------------------
    def eparams(self, values):
        """Element parameters -> ``pf``: an absolute offset outside arrays (the kernel's
        facet-row base is the blob itself there), an offset inside the facet row otherwise."""
        if self.array is None:
            return self.params(values)
        a = self.array
        pf = 14 + len(a['row'])
        vals = [float(x) for x in np.asarray(values, dtype=float).ravel()]
        a['row'].extend(vals)
        a['block_len'][pf] = len(vals)
        return pf

    # ---- arrays -------------------------------------------------------------------------
    def begin_array(self):
        if self.array is not None:
            raise NotFusable('nested Parallel containers are not fused')
        self.array = dict(facet=-1, body=[], rows=[], geoms=[], ids=[], slots=[], init=[], colrefs=[],
                          folds=[], block_len={}, begin=self.op('ARRAY_BEGIN'))
        # the ARRAY_BEGIN op itself is not part of the per-facet body
        self.array['body'] = []

    def begin_facet(self):
        a = self.array
        a['facet'] += 1
        a['row'] = []
        a['geom'] = None
        a['geom_taken'] = False
        a['op_cursor'] = 0
        a['slot_cursor'] = 0
        a['col_cursor'] = 0
        a['fold_cursor'] = 0
        a['id_num'] = -9

    def end_facet(self):
        a = self.array
        if a['geom'] is None:
            raise NotFusable('facet without a plane geometry')
        if a['facet'] > 0 and a['op_cursor'] != len(a['body']):
            raise NotFusable('facets of a Parallel lower to different op lists')
        a['geoms'].append(a['geom'])
        a['rows'].append(list(a['row']))
        a['ids'].append(a['id_num'])

    def end_array(self, cull=True):
        cull = cull and not EXHAUSTIVE_SEARCH
        a = self.array
        F = len(a['rows'])
        nper = len(a['rows'][0]) if F else 0
        if any(len(r) != nper for r in a['rows']):
            raise NotFusable('facets of a Parallel carry different parameter counts')
        # Parameter blocks that are identical for every facet (quality factor, L2 dimensions, scatter
        # widths ...) move out of the facet rows into one shared block: smaller rows = more facets
        # per staged program.  Only for ops whose `pg` field is otherwise unused.
        if F > 1 and nper:
            rows_arr = np.array(a['rows'])
            keep = np.ones(nper, dtype=bool)
            for rec in a['body']:
                pf = rec['pf']
                if (rec['type'] in self.HOISTABLE and rec['pg'] < 0 and pf >= 14 and pf in a['block_len']):
                    cols = slice(pf - 14, pf - 14 + a['block_len'][pf])
                    if np.all(rows_arr[:, cols] == rows_arr[0, cols]):
                        rec['pg'] = self.params(rows_arr[0, cols])
                        rec['pf'] = -1
                        keep[cols] = False
            if not keep.all():
                new_index = np.cumsum(keep) - 1
                for rec in a['body']:
                    if rec['pf'] >= 14:
                        rec['pf'] = 14 + int(new_index[rec['pf'] - 14])
                a['rows'] = [list(r) for r in rows_arr[:, keep]]
                nper = int(keep.sum())
        width = 14 + nper + 1
        stride = width + (width % 2)
        if (stride // 2) % 2 == 0:
            stride += 2                      # stride/2 odd: spreads 16-byte smem accesses over banks
        rows = np.zeros((max(F, 1), stride))
        for j in range(F):
            rows[j, :14] = a['geoms'][j]
            rows[j, 14:14 + nper] = a['rows'][j]
            rows[j, 14 + nper] = a['ids'][j]
        G = rows[:F, :14]
        grid = build_cull_grid(G) if (cull and F) else None
        head = np.zeros(ARRAY_HEADER_WORDS)
        rows_off = self.params(rows)
        ints = [F, stride, rows_off, 0, 0, 0, 0, 0]
        if grid is not None:
            # disjointness certificate: head[17] = tan^2 of its cone, head[19] mode (1: every facet is the last one a
            # photon can hit; 2: successor lists), head[18] = offset of the per-facet list starts (int32, F + 1), which
            # index into the candidate array (the lists are appended to the grid's)
            t2, sstart, succ = single_hit_successors(G, grid)
            if t2 >= grid['T2'] and os.environ.get('MXB_GRID_HOPS', '0') == '0':      # (env: A/B switch)
                # after a hit the photon is either inside the certified cone (done / successors) or outside the culling
                # cone (footprint scan): it never walks on through its cell list, so the list only has to be complete
                # for the incoming ray (same frame O, nbar, u, v; finer inflation, hence other cells)
                grid = build_cull_grid(G, hops=0)
            head[0:3], head[3:6], head[6:9], head[9:12] = grid['O'], grid['nbar'], grid['u'], grid['v']
            head[12], head[13], head[14], head[15] = grid['u0'], grid['v0'], grid['inv_cell'], grid['T2']
            head[16] = grid['H'] * (1. + 1e-9) + 1e-6      # half thickness of the slab that holds every facet point
            ints[3] = 1
            ints[4], ints[5] = grid['nu'], grid['nv']
            cand_all = grid['cand']
            if t2 > 0. and succ is None:
                head[17], head[19] = t2, 1.
            elif t2 > 0.:
                head[17], head[19] = t2, 2.
                head[18] = float(self.params(_pack_i32(sstart + len(cand_all))))
                cand_all = np.concatenate([cand_all, succ])
            ints[6] = self.params(_pack_i32(grid['start']))
            ints[7] = self.params(_pack_i32(cand_all))
            grid['single_hit'] = (t2, sstart, succ)
        init = a['init']
        init_off = self.params(_pack_i32(init)) if init else 0
        begin = self.ops[a['begin']]
        begin['pg'] = self.params(head)
        begin['cols'] = ints
        begin['s0'], begin['s1'] = len(init), init_off
        for rec in a['body']:
            # ops that need the facet's id_num read it from the last word of the row
            if ((rec['type'] in (OP['COMMIT'], OP['DETPIX']) and rec['flags'] & 1) or rec['type'] == OP['ACIS']
                    or rec['flags'] & self.COMMIT_ROW_ID):
                rec['w15'] = 14 + nper
        self.array = None
        begin['w14'] = self.op('ARRAY_END')      # where to continue when no photon of a warp hits
        self.last_grid = grid
        return grid

    # ---- finish -----------------------------------------------------------------------------
    def finish(self):
        small = np.concatenate(self.small)
        if len(small) % 2:
            small = np.concatenate([small, [0.]])
        stage_words = len(small)
        blob = small
        # append global-only tables and patch their offsets
        offs = []
        for t in self.big:
            offs.append(len(blob))
            blob = np.concatenate([blob, t, np.zeros(len(t) % 2)])
        for word, idx in self.fixups:
            blob[word] = offs[idx]
        for word, base, delta in self.rel_fixups:
            blob[word] = blob[base] + delta
        blob[5] = 1 if self.born else 0
        blob[0], blob[1], blob[2], blob[3], blob[4] = MXB_MAGIC, _lib.MXB_ABI_VERSION, len(self.ops), len(blob), stage_words
        for i, rec in enumerate(self.ops):
            w = HEADER_WORDS + OP_WORDS * i
            blob[w:w + 4] = [rec['type'], rec['flags'], rec['pg'], rec['pf']]
            blob[w + 4:w + 12] = rec['cols']
            blob[w + 12:w + 16] = [rec['s0'], rec['s1'], rec['w14'], rec['w15']]
        return Program(blob, stage_words, list(self.ops), self.f64_cols[FIRST_OUT:], list(self.i64_cols),
                       list(self.slot_kinds), OrderedDict(self.new_cols), self.meta_updates,
                       OrderedDict(self.aux))


_status_names = {_lib.MXB_ST_PROB_RANGE: 'probability factor outside [0, 1]'}


class Program:
    """A lowered element tree: device blob + column/slot layout + launcher."""

    def __init__(self, blob, stage_words, ops, out_f64, out_i64, slot_kinds, new_cols, meta_updates,
                 aux=None):
        self.aux = aux or OrderedDict()
        self.blob = np.ascontiguousarray(blob, dtype=np.float64)
        self.stage_words = stage_words
        self.ops = ops
        self.out_f64 = out_f64
        self.out_i64 = out_i64
        self.slot_kinds = slot_kinds
        self.new_cols = new_cols
        self.meta_updates = meta_updates
        self._dev = {}
        self.last_status = None

    @property
    def n_ops(self):
        return len(self.ops)

    def device_blob(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.blob).to(device)
        return self._dev[key]

    def columns_struct(self, photons, draws=None):
        """Fill an MxbColumns struct with the plane pointers of ``photons``
        (allocating this program's new output columns).  Returns (struct, keepalive)."""
        cols = _lib.MxbColumns()
        keep = []
        n = len(photons)
        for vi, name in enumerate(('pos', 'dir', 'polarization')):
            st = photons.storage(name)
            if st.dtype != torch.float64 or not st.is_contiguous() or st.shape != (4, n):
                raise ValueError('column {0} must be a contiguous (4, N) float64 block'.format(name))
            for k in range(3):
                cols.f64[3 * vi + k] = st.data_ptr() + k * n * 8
        for idx, name in ((9, 'energy'), (10, 'probability')):
            st = photons.storage(name)
            if st.dtype != torch.float64 or not st.is_contiguous():
                raise ValueError('column {0} must be contiguous float64'.format(name))
            cols.f64[idx] = st.data_ptr()
        for k, name in enumerate(self.out_f64):
            if name in self.aux:
                t = self.aux[name]
                if t.device.type != photons.device.type or t.dtype != torch.float64 or not t.is_contiguous():
                    raise ValueError('auxiliary buffer must be a contiguous float64 tensor on the photon device')
                cols.f64[FIRST_OUT + k] = t.data_ptr()
                continue
            if name not in photons:
                photons.new_column(name, torch.float64)
            st = photons.storage(name)
            if st.dtype != torch.float64 or st.dim() != 1 or not st.is_contiguous():
                raise ValueError('output column {0} must be contiguous (N,) float64'.format(name))
            cols.f64[FIRST_OUT + k] = st.data_ptr()
        for k, name in enumerate(self.out_i64):
            if name not in photons:
                photons.new_column(name, torch.int64)
            st = photons.storage(name)
            if st.dtype != torch.int64 or st.dim() != 1 or not st.is_contiguous():
                raise ValueError('id column {0} must be contiguous (N,) int64'.format(name))
            cols.i64[k] = st.data_ptr()
        if draws is not None:
            table = draws.table if hasattr(draws, 'table') else draws
            if len(table) != len(self.slot_kinds):
                raise ValueError('expected {0} draw slots {1}, got {2}'.format(
                    len(self.slot_kinds), self.slot_kinds, len(table)))
            for k, d in enumerate(table):
                if d is None:
                    continue
                t = d if isinstance(d, torch.Tensor) else torch.from_numpy(
                    np.ascontiguousarray(d, dtype=np.float64))
                t = t.to(photons.device, torch.float64).contiguous()
                if t.shape != (n,):
                    raise ValueError('draw slot {0} must have shape (N,)'.format(k))
                keep.append(t)
                cols.draws[k] = t.data_ptr()
        return cols, keep

    def source_planes(self, source, n):
        """ctypes array of the 11 core plane pointers of ``source`` (mxb_trace_from)."""
        ptrs = (ctypes.c_void_p * 11)()
        for vi, name in enumerate(('pos', 'dir', 'polarization')):
            st = source.storage(name)
            if st.dtype != torch.float64 or not st.is_contiguous() or st.shape != (4, n):
                raise ValueError('source column {0} must be a contiguous (4, N) float64 block'.format(name))
            for k in range(3):
                ptrs[3 * vi + k] = st.data_ptr() + k * n * 8
        for idx, name in ((9, 'energy'), (10, 'probability')):
            st = source.storage(name)
            if st.dtype != torch.float64 or not st.is_contiguous() or st.shape != (n,):
                raise ValueError('source column {0} must be contiguous (N,) float64'.format(name))
            ptrs[idx] = st.data_ptr()
        return ptrs

    def run(self, photons, draws=None, seed=0, id0=0, check=True, strict=None, source=None):
        """Launch on ``photons`` (in place).  With ``check`` the status block is read
        back (one sync) and reference errors are raised (optics/base.py:45-46).
        ``source``: another table on the same device whose core columns are READ instead of
        those of ``photons`` (which then only receives results: mxb_trace_from)."""
        if photons.device.type != 'cuda':
            raise _lib.MxbError('marxs_b200 runs on CUDA devices only (no CPU fallback); '
                                'photons are on {0}'.format(photons.device))
        lib = _lib.load(strict)
        n = len(photons)
        if n == 0:
            return photons
        for name in ('pos', 'dir', 'polarization', 'energy', 'probability'):
            if name not in photons:
                raise KeyError('photon table has no column {0}'.format(name))
        with torch.cuda.device(photons.device):
            cols, keep = self.columns_struct(photons, draws)
            blob = self.device_blob(photons.device)
            status = torch.zeros(_lib.MXB_STATUS_WORDS, dtype=torch.int64, device=photons.device)
            stream = torch.cuda.current_stream(photons.device).cuda_stream
            src = None
            if source is not None:
                if source.device != photons.device or len(source) != n:
                    raise ValueError('source and destination tables must have the same length and device')
                src = self.source_planes(source, n)
            rc = lib.mxb_trace_from(blob.data_ptr(), self.blob.size, self.blob.ctypes.data, src, ctypes.byref(cols),
                                    n, int(id0), int(seed) & 0xFFFFFFFFFFFFFFFF, status.data_ptr(), stream)
            _lib.check(lib, rc, 'mxb_trace')
            for t in keep:
                t.record_stream(torch.cuda.current_stream(photons.device))
            photons.meta.update(self.meta_updates)
            if check:
                self.finalize(photons, status.cpu().numpy())
            else:
                self.last_status = status
        return photons

    def finalize(self, photons, st):
        """Host-side epilogue of a launch: error semantics + column pruning."""
        self.last_status = st
        # reference: process_photons adds no columns when nothing intersects (optics/base.py:176-177)
        for name, info in self.new_cols.items():
            if info['creator'] >= 0 and st[_lib.MXB_ST_OPHITS + info['creator']] == 0 and name in photons:
                photons.remove_column(name)
        if st[_lib.MXB_ST_PROB_RANGE]:
            raise ValueError('Found probability outside of the 0..1 arange.')
        if st[_lib.MXB_ST_FILTER_BOUNDS]:
            raise ValueError('A value in x_new is outside the interpolation range.')
        if st[_lib.MXB_ST_INTENSITY]:
            raise ValueError('Intensity cannot be > 1')

    def host_columns(self, table):
        """MxbColumns over HOST numpy planes for mxb_trace_host.  ``table`` maps the
        column names (pos/dir/polarization as (4,N) component planes) to arrays."""
        raise NotImplementedError
