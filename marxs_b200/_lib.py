"""ctypes binding of libmxb (the C ABI declared in include/mxb.h).

There is NO fallback: if the shared library is missing or there is no CUDA
device, calls raise.  ``MXB_STRICT=1`` selects the bit-parity build
(``libmxb_strict.so``: -fmad=false, IEEE division), the default is the fast
build (``libmxb.so``: FMA contraction, reciprocal normalisation).
"""
import ctypes
import os

MXB_ABI_VERSION = 12
MXB_MAX_F64_COLS = 56
MXB_MAX_I64_COLS = 8
MXB_MAX_SLOTS = 16
MXB_STATUS_WORDS = 64
MXB_ST_PROB_RANGE, MXB_ST_MULTI_HIT, MXB_ST_BRUTE, MXB_ST_FILTER_BOUNDS, MXB_ST_INTENSITY = range(5)
MXB_ST_OPHITS = 8

_HERE = os.path.dirname(os.path.abspath(__file__))


class MxbColumns(ctypes.Structure):
    _fields_ = [('f64', ctypes.c_void_p * MXB_MAX_F64_COLS),
                ('i64', ctypes.c_void_p * MXB_MAX_I64_COLS),
                ('draws', ctypes.c_void_p * MXB_MAX_SLOTS)]


class MxbHostOptions(ctypes.Structure):
    _fields_ = [('i32_out', ctypes.c_void_p * MXB_MAX_I64_COLS)]


class MxbError(RuntimeError):
    pass


_libs = {}
_default_strict = [os.environ.get('MXB_STRICT', '0') not in ('', '0')]


def set_strict(flag):
    """Select the bit-parity build (True) or the fast build (False) for subsequent calls."""
    _default_strict[0] = bool(flag)


def lib_path(strict=None):
    if strict is None:
        strict = _default_strict[0]
    if os.environ.get('MXB_LIB_PATH') and not strict:      # kernel-variant experiments
        return os.environ['MXB_LIB_PATH']
    return os.path.join(_HERE, 'libmxb_strict.so' if strict else 'libmxb.so')


def load(strict=None):
    """Load (once) and return the ctypes handle of libmxb."""
    path = lib_path(strict)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise MxbError('{0} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       '(marxs_b200 has no CPU fallback)'.format(path))
    lib = ctypes.CDLL(path)
    vp, i64, u64, sz, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    lib.mxb_version.restype = ci
    lib.mxb_build_info.restype = ctypes.c_char_p
    lib.mxb_last_error.restype = ctypes.c_char_p
    lib.mxb_device_count.restype = ci
    lib.mxb_host_release.restype = None
    lib.mxb_trace.restype = ci
    lib.mxb_trace.argtypes = [vp, sz, vp, ctypes.POINTER(MxbColumns), i64, i64, u64, vp, vp]
    lib.mxb_trace_from.restype = ci
    lib.mxb_trace_from.argtypes = [vp, sz, vp, vp, ctypes.POINTER(MxbColumns), i64, i64, u64, vp, vp]
    lib.mxb_trace_host.restype = ci
    lib.mxb_trace_host.argtypes = [vp, sz, ctypes.POINTER(MxbColumns), ctypes.POINTER(MxbColumns), i64, i64, i64, u64, vp]
    lib.mxb_trace_host_opts.restype = ci
    lib.mxb_trace_host_opts.argtypes = [vp, sz, ctypes.POINTER(MxbColumns), ctypes.POINTER(MxbColumns), i64, i64, i64, u64, vp,
                                        ctypes.POINTER(MxbHostOptions)]
    lib.mxb_debug_draws.restype = ci
    lib.mxb_debug_draws.argtypes = [u64, i64, i64, ci, ci, vp, vp, vp]
    lib.mxb_sigma_clip_workspace.restype = ctypes.c_size_t
    lib.mxb_sigma_clip_workspace.argtypes = []
    lib.mxb_sigma_clip_stats.restype = ci
    lib.mxb_sigma_clip_stats.argtypes = [vp, i64, ctypes.c_double, ci, vp, vp, vp]
    lib.mxb_debug_math.restype = ci
    lib.mxb_debug_math.argtypes = [ci, vp, vp, vp, i64, ctypes.c_double, ctypes.c_double, vp]
    lib.mxb_plane_intersect.restype = ci
    lib.mxb_plane_intersect.argtypes = [vp, ci, vp, vp, vp, vp, vp, i64, vp]
    lib.mxb_parallel_transport.restype = ci
    lib.mxb_parallel_transport.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.mxb_polarization_vectors.restype = ci
    lib.mxb_polarization_vectors.argtypes = [vp, vp, vp, i64, vp]
    lib.mxb_hist2d.restype = ci
    lib.mxb_hist2d.argtypes = [vp, vp, vp, vp, ctypes.c_longlong, ci, ctypes.c_double, ctypes.c_double, i64, ci, ci, vp, vp, vp]
    lib.mxb_set_jit.restype = None
    lib.mxb_set_jit.argtypes = [ci]
    lib.mxb_get_jit.restype = ci
    lib.mxb_jit_info.restype = ctypes.c_char_p
    lib.mxb_jit_source.restype = ctypes.c_longlong
    lib.mxb_jit_source.argtypes = [vp, sz, ctypes.POINTER(MxbColumns), vp, sz]
    lib.mxb_compact_workspace.restype = sz
    lib.mxb_compact_workspace.argtypes = [i64]
    lib.mxb_compact_events.restype = ci
    lib.mxb_compact_events.argtypes = [vp, vp, ci, vp, ctypes.c_longlong, vp, i64, vp, vp, sz, vp]
    lib.mxb_compact_append.restype = ci
    lib.mxb_compact_append.argtypes = [vp, vp, ci, vp, ctypes.c_longlong, vp, i64, vp, i64, vp, sz, vp]
    lib.mxb_jit_compile.restype = ctypes.c_longlong
    lib.mxb_jit_compile.argtypes = [vp, sz, ctypes.POINTER(MxbColumns)]
    if lib.mxb_version() != MXB_ABI_VERSION:
        raise MxbError('libmxb ABI {0} != python binding {1}: rebuild'.format(lib.mxb_version(), MXB_ABI_VERSION))
    _libs[path] = lib
    return lib


def jit_source(program, cols, strict=None):
    """CUDA source of the kernel libmxb specialises for ``program`` with the columns ``cols``."""
    lib = load(strict)
    n = lib.mxb_jit_source(program.blob.ctypes.data, program.blob.size, ctypes.byref(cols), None, 0)
    if n < 0:
        raise MxbError('mxb_jit_source failed ({0}): {1}'.format(n, lib.mxb_last_error().decode()))
    buf = ctypes.create_string_buffer(n + 1)
    lib.mxb_jit_source(program.blob.ctypes.data, program.blob.size, ctypes.byref(cols), buf, n + 1)
    return buf.value.decode()


def check(lib, rc, what):
    if rc != 0:
        raise MxbError('{0} failed ({1}): {2}'.format(what, rc, lib.mxb_last_error().decode()))


EXPORTED_SYMBOLS = ['mxb_version', 'mxb_build_info', 'mxb_last_error', 'mxb_device_count', 'mxb_host_release', 'mxb_trace',
                    'mxb_trace_from', 'mxb_trace_host', 'mxb_trace_host_opts', 'mxb_debug_draws', 'mxb_debug_math', 'mxb_sigma_clip_workspace', 'mxb_sigma_clip_stats', 'mxb_plane_intersect', 'mxb_parallel_transport', 'mxb_polarization_vectors', 'mxb_hist2d',
                    'mxb_set_jit', 'mxb_get_jit', 'mxb_jit_info', 'mxb_jit_source', 'mxb_jit_compile', 'mxb_compact_workspace', 'mxb_compact_events', 'mxb_compact_append']
