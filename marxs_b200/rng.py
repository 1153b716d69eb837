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
