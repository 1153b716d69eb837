"""The C-ABI shared libraries load on a CPU-only host and export every symbol that
include/mxb.h declares; the Python mirror of the header constants is in sync.
No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from marxs_b200 import _lib, program

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, 'include', 'mxb.h')).read()


def header_functions():
    names = re.findall(r'^\s*(?:int|void|const char\*|long long|size_t)\s+(mxb_\w+)\s*\(', HEADER, flags=re.M)
    return sorted(set(names))


def header_define(name):
    m = re.search(r'#define\s+{0}\s+(\S+)'.format(name), HEADER)
    assert m, name
    return int(m.group(1), 0)


@pytest.mark.parametrize('strict', [False, True])
def test_library_exports_header_symbols(strict):
    path = _lib.lib_path(strict)
    assert os.path.exists(path), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(path)
    funcs = header_functions()
    assert set(funcs) == set(_lib.EXPORTED_SYMBOLS)
    for f in funcs:
        assert hasattr(lib, f), f
    lib.mxb_version.restype = ctypes.c_int
    assert lib.mxb_version() == header_define('MXB_ABI_VERSION') == _lib.MXB_ABI_VERSION
    lib.mxb_build_info.restype = ctypes.c_char_p
    info = lib.mxb_build_info().decode()
    assert 'sm_100a' in info and (('strict' in info) == strict)


def test_python_constants_match_header():
    for py, c in ((program.MXB_MAGIC, 'MXB_MAGIC'), (program.HEADER_WORDS, 'MXB_HEADER_WORDS'),
                  (program.OP_WORDS, 'MXB_OP_WORDS'), (program.MAX_OPS, 'MXB_MAX_OPS'),
                  (program.COL_INIT, 'MXB_COL_INIT'), (program.FIRST_OUT, 'MXB_COL_FIRST_OUT'),
                  (program.ARRAY_HEADER_WORDS, 'MXB_ARRAY_HEADER_WORDS'),
                  (_lib.MXB_MAX_F64_COLS, 'MXB_MAX_F64_COLS'), (_lib.MXB_MAX_I64_COLS, 'MXB_MAX_I64_COLS'),
                  (_lib.MXB_MAX_SLOTS, 'MXB_MAX_SLOTS'), (_lib.MXB_STATUS_WORDS, 'MXB_STATUS_WORDS'),
                  (_lib.MXB_ST_OPHITS, 'MXB_ST_OPHITS')):
        assert py == header_define(c), c
    for name, code in program.OP.items():
        assert code == header_define('MXB_OP_' + name), name
    assert ctypes.sizeof(_lib.MxbColumns) == 8 * (_lib.MXB_MAX_F64_COLS + _lib.MXB_MAX_I64_COLS + _lib.MXB_MAX_SLOTS)


def test_no_device_fails_loudly():
    """No CPU fallback: tracing a CPU-resident table raises instead of computing something."""
    import torch
    import marxs_b200 as mb
    from marxs_b200 import optics
    if torch.cuda.is_available():
        pytest.skip('needs a host without GPU')
    p = mb.generate_test_photons(4, device='cpu')
    with pytest.raises(_lib.MxbError, match='CUDA'):
        optics.FlatDetector(pixsize=0.1)(p)
    with pytest.raises(_lib.MxbError):
        optics.FlatDetector(pixsize=0.1).geometry.intersect(p['dir'], p['pos'])


def test_malformed_program_rejected_without_gpu():
    """Argument validation happens before any CUDA call."""
    lib = _lib.load(False)
    cols = _lib.MxbColumns()
    bad = np.zeros(32)
    rc = lib.mxb_trace(None, bad.size, bad.ctypes.data, ctypes.byref(cols), 10, 0, 0, None, None)
    assert rc == -1 and b'magic' in lib.mxb_last_error()
    status = np.zeros(_lib.MXB_STATUS_WORDS, dtype=np.uint64)
    rc = lib.mxb_trace_host(bad.ctypes.data, bad.size, ctypes.byref(cols), ctypes.byref(cols), 10, 0, 0, 0,
                            status.ctypes.data)
    assert rc == -1
