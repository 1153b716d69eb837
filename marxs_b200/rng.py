"""Random-draw policy of the engine.

Production runs draw on the device: Philox4x32-10 keyed by ``seed`` and counted
by (global photon id, draw slot) — results do not depend on batch size, sharding
or GPU count.  Parity tests *inject* per-photon draws instead: one (N,) array per
draw slot, in the slot order of the lowered element tree (the same depth-first
order as ``oracle.marxs_oracle.assign_slots``; all facets of a Parallel share
the slots of facet 0 — SURVEY.md Appendix C)."""
import contextlib

_state = {'seed': 20261017, 'launch': 0, 'injected': None, 'cursor': 0}


def set_seed(seed):
    _state['seed'] = int(seed)
    _state['launch'] = 0


def next_launch_seed():
    """Distinct Philox key per launch so re-running an element draws fresh numbers."""
    s = (_state['seed'] + 0x9E3779B97F4A7C15 * _state['launch']) & 0xFFFFFFFFFFFFFFFF
    _state['launch'] += 1
    return s


@contextlib.contextmanager
def inject_draws(table):
    """``table``: sequence of (N,) arrays/tensors (or an object with ``.table``),
    consumed slot by slot by the programs launched inside the context."""
    old = (_state['injected'], _state['cursor'])
    _state['injected'] = list(table.table if hasattr(table, 'table') else table)
    _state['cursor'] = 0
    try:
        yield
    finally:
        _state['injected'], _state['cursor'] = old


def take_injected(n_slots):
    """Draw arrays for the next ``n_slots`` slots, or None when nothing is injected."""
    if _state['injected'] is None:
        return None
    c = _state['cursor']
    if c + n_slots > len(_state['injected']):
        raise ValueError('injected draw table has {0} slots, program needs slots {1}..{2}'.format(
            len(_state['injected']), c, c + n_slots - 1))
    _state['cursor'] = c + n_slots
    return _state['injected'][c:c + n_slots]


# ---------------------------------------------------------------------------
# verification: the draws a device launch makes, exported (mxb_debug_draws)
# ---------------------------------------------------------------------------
def device_draw_plan(prog):
    """Which device routine produces each draw slot of ``prog``: list of (kind, slot, partner) with kind
    0 uniform, 1 normal, 2 normal pair (RSCATTER with both widths non-zero takes both deviates of ONE Philox
    call keyed by its first slot: mxb_ops.cuh rscatter_draws).  Mirrors the kernels' own choices."""
    from .program import OP
    plan = {}

    def add(kind, slot, partner=-1):
        if slot is not None and slot >= 0 and slot not in plan:
            plan[slot] = (kind, slot, partner)
    in_array = False
    for o in prog.ops:
        t = o['type']
        if t == OP['ARRAY_BEGIN']:
            in_array = True
        elif t == OP['ARRAY_END']:
            in_array = False
        elif t == OP['RSCATTER']:
            if in_array or o['pf'] < 0:
                raise NotImplementedError('draw export for RSCATTER inside an array')
            sig_in, sig_perp = prog.blob[o['pf'] + 3], prog.blob[o['pf'] + 4]
            if sig_in != 0. and sig_perp != 0.:
                add(2, o['s0'], o['s1'])
                plan[o['s1']] = (None, o['s1'], o['s0'])       # filled by its partner
            else:
                add(1, o['s0'])
                add(1, o['s1'])
        elif t == OP['GSCATTER']:
            if not (o['flags'] & 2):
                add(1, o['s0'])
            add(0, o['s1'])
        elif t == OP['GRATING']:
            add(0, o['s0'])
            if o['flags'] & 8:
                add(0, o['s1'])
        elif t == OP['APERTURE']:
            add(0, o['s0'])
            add(0, o['s1'])
            add(0, o['w14'])
        elif t == OP['GENERATE']:
            for s in (o['s0'], o['s1'], o['w14'], o['w15']):
                add(0, s)
        elif t == OP['POINTING']:
            add(0, o['s0'])
            add(1, o['s1'])
        elif t in (OP['LABCONE'], OP['FARLAB']):
            add(0, o['s0'])
            add(0, o['s1'])
    return [plan[s] for s in sorted(plan)]


def export_device_draws(prog, seed, id0, n, device, strict=None):
    """Per-slot (n,) float64 device tensors with the draws a launch of ``prog`` with ``seed`` and first
    global photon id ``id0`` makes for photons 0..n-1 (unused slots: None)."""
    import torch
    from . import _lib
    lib = _lib.load(strict)
    out = [None] * len(prog.slot_kinds)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        for kind, slot, partner in device_draw_plan(prog):
            if kind is None:
                continue
            out[slot] = torch.empty(n, dtype=torch.float64, device=device)
            p1 = None
            if kind == 2:
                out[partner] = torch.empty(n, dtype=torch.float64, device=device)
                p1 = out[partner].data_ptr()
            rc = lib.mxb_debug_draws(int(seed) & 0xFFFFFFFFFFFFFFFF, int(id0), int(n), int(slot), int(kind),
                                     out[slot].data_ptr(), p1, stream)
            _lib.check(lib, rc, 'mxb_debug_draws')
    return out


def philox4x32_10_numpy(counter, key):
    """Philox4x32-10 (Salmon et al. 2011) in numpy: counter (n, 4) uint32, key (2,) uint32 -> (n, 4) uint32.
    Host restatement of mxb_device.cuh philox4x32_10, used by the tests to pin the counter / key layout."""
    import numpy as np
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy(), c[:, 3].copy()
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def philox_uniform_numpy(seed, ids, slot):
    """The uniform [0,1) the device draws for global photon ids ``ids`` and ``slot`` under ``seed``."""
    import numpy as np
    ids = np.asarray(ids, dtype=np.uint64)
    ctr = np.stack([ids & np.uint64(0xFFFFFFFF), ids >> np.uint64(32), np.full_like(ids, slot), np.zeros_like(ids)], axis=1)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10_numpy(ctr, (seed & 0xFFFFFFFF, seed >> 32)).astype(np.uint64)
    bits = (r[:, 0] << np.uint64(32)) | r[:, 1]
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
