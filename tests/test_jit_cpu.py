"""Kernel specialisation on a CPU-only host: the op list of a lowered program turns into CUDA
source and compiles with NVRTC for sm_100a (no GPU needed for either); the cache key depends on the
program's STRUCTURE only, so moving elements never recompiles.  No launches here."""
import ctypes
import os

import numpy as np
import pytest

from marxs_b200 import _lib, optics, simulator
from marxs_b200.program import Lowering

CORE = ['pos', 'dir', 'polarization', 'energy', 'probability']


def lower(elements, existing=CORE, meta=None):
    lw = Lowering(existing, meta=meta if meta is not None else {'ROLL_PNT': (0., 'roll')})
    for e in elements:
        e._lower(lw)
    return lw.finish()


def fake_columns(prog, injected=False):
    """Only the NULL-ness of the pointers matters for specialisation."""
    cols = _lib.MxbColumns()
    for k in range(11 + len(prog.out_f64)):
        cols.f64[k] = 0x1000 + 16 * k
    for k in range(len(prog.out_i64)):
        cols.i64[k] = 0x8000 + 16 * k
    if injected:
        for k in range(len(prog.slot_kinds)):
            cols.draws[k] = 0x9000 + 16 * k
    return cols


def c1():
    return [optics.RectangleAperture(position=[100., 0, 0], zoom=2), optics.PerfectLens(focallength=1000., zoom=400),
            optics.FlatDetector(pixsize=0.01, position=[-1000., 0, 0], zoom=1e5)]


def c2():
    from marxs_b200.missions import chandra
    return [chandra.HRMA(), chandra.HETG(),
            chandra.ACIS(chips=[4, 5, 6, 7, 8, 9], aimpoint=chandra.AIMPOINTS['ACIS-S'])]


def test_source_c1_has_no_shared_memory_stage():
    prog = lower(c1(), existing=['dir', 'polarization', 'energy', 'probability'])
    src = _lib.jit_source(prog, fake_columns(prog))
    assert 'op_aperture(' in src and 'op_lens(' in src and 'op_detpix(' in src
    assert '#define JIT_STAGE_WORDS 0' in src          # single elements: parameters live in the constant bank
    assert 'array_search' not in src
    assert 'device_draw' not in src and 'draw_value(((const double*)0)' in src    # Philox path baked in


def test_source_c2_structure():
    prog = lower(c2())
    src = _lib.jit_source(prog, fake_columns(prog))
    assert src.count('array_open(') == 2 and src.count('array_search(') == 2
    assert 'select_order_fixed<7>' in src              # HETG: 7 equally likely orders, unrolled
    assert 'op_acis(' in src and 'op_grating(' in src and 'op_rscatter<1, 1>(' in src   # both scatter widths non-zero: baked
    assert '#define JIT_STAGE_WORDS {0}'.format(prog.stage_words) in src
    # FlatStack layers repeat the loc-coos commit of their stack: emitted once per column
    assert src.count('ph.pos = ph.ip') == 3
    inj = _lib.jit_source(prog, fake_columns(prog, injected=True))
    assert 'P.d[0]' in inj and 'P.d[0]' not in src


def test_structure_key_ignores_numbers():
    """Re-positioned elements (tolerancing loops) give byte-identical source: nothing is recompiled."""
    a = lower([optics.FlatDetector(pixsize=0.1, position=[0., 1., 2.], zoom=[1, 5, 5])])
    b = lower([optics.FlatDetector(pixsize=0.2, position=[3., -1., 0.5], zoom=[1, 7, 2])])
    assert not np.array_equal(a.blob, b.blob)
    assert _lib.jit_source(a, fake_columns(a)) == _lib.jit_source(b, fake_columns(b))
    # a different column set is a different kernel
    cols = fake_columns(a)
    cols.f64[11] = None
    assert _lib.jit_source(a, cols) != _lib.jit_source(a, fake_columns(a))


@pytest.mark.parametrize('strict', [False, True])
def test_nvrtc_compiles_c1_c2_for_sm100a(strict, tmp_path, monkeypatch):
    monkeypatch.setenv('MXB_CACHE_DIR', str(tmp_path))
    lib = _lib.load(strict)
    for elems, existing in ((c1(), ['dir', 'polarization', 'energy', 'probability']), (c2(), CORE)):
        prog = lower(elems, existing=existing)
        cols = fake_columns(prog)
        n = lib.mxb_jit_compile(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cols))
        if n < 0 and b'NVRTC unavailable' in lib.mxb_last_error():
            pytest.skip('no NVRTC on this host')
        assert n > 10000, lib.mxb_last_error()
        assert lib.mxb_jit_info().startswith(b'jit ') and b'(compiled)' in lib.mxb_jit_info()
        # second request is served from the disk cache
        assert lib.mxb_jit_compile(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cols)) == n
        assert b'(disk cache)' in lib.mxb_jit_info()
    names = os.listdir(str(tmp_path))
    assert sum(x.endswith('.cubin') for x in names) == 2 and sum(x.endswith('.cu') for x in names) == 2


def test_jit_mode_switch():
    lib = _lib.load(False)
    for m in (0, 1, 2):
        lib.mxb_set_jit(m)
        assert lib.mxb_get_jit() == m
    lib.mxb_set_jit(-1)
    assert lib.mxb_get_jit() in (0, 1, 2)


def test_malformed_program_is_rejected():
    prog = lower([optics.FlatDetector(pixsize=0.1, zoom=[1, 5, 5])])
    bad = prog.blob.copy()
    bad[16 + 2] = 1e6          # pg of op 0 far outside the blob
    cols = fake_columns(prog)
    lib = _lib.load(False)
    n = lib.mxb_jit_source(bad.ctypes.data, bad.size, ctypes.byref(cols), None, 0)
    assert n == -5 and b'outside the blob' in lib.mxb_last_error()


def test_born_program_compiles():
    """source -> pointing -> aperture -> detector: header flag 'born', no loads of the core planes."""
    from marxs_b200 import source
    src = source.PointSource(coords=(30., 10.), flux=100., energy={'energy': np.array([0.3, 1., 2.]), 'fluxdensity': np.array([0., 1., 2.])})
    pnt = source.JitterPointing(coords=(30., 10.), jitter=1e-6)
    prog = lower([src, pnt] + c1()[:1] + c1()[2:], existing=[])
    assert prog.blob[5] == 1 and prog.slot_kinds == ['uniform', 'uniform', 'uniform', 'uniform', 'normal', 'uniform', 'uniform']
    cols = fake_columns(prog)
    text = _lib.jit_source(prog, cols)
    assert 'op_generate(' in text and 'op_pointing(' in text and 'pipe_load(' not in text
    lib = _lib.load(False)
    n = lib.mxb_jit_compile(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cols))
    if n < 0 and b'NVRTC unavailable' in lib.mxb_last_error():
        pytest.skip('no NVRTC on this host')
    assert n > 10000, lib.mxb_last_error()
