#!/usr/bin/env python
"""Host-only inspection of a specialised kernel: lower a configuration, dump the generated CUDA source, compile it
with NVRTC (no GPU needed) and report registers / spills / code size from the cubin.

    python tools/jit_inspect.py c2|c3 [out_dir]        # honours MXB_JIT_* experiment knobs
"""
import ctypes
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import torch  # noqa: E402

from marxs_b200 import _lib, simulator  # noqa: E402
from marxs_b200.program import Lowering  # noqa: E402


def program(cfg):
    if cfg == 'c2':
        import bench
        inst = bench.c2_instrument()
        inst.elements[2].image = torch.zeros((6, 1024, 1024), dtype=torch.float64)
        meta = {'ROLL_PNT': (0., 'roll')}
    elif cfg == 'c3':
        import bench_configs
        elements, _ = bench_configs.c3_elements()
        inst = simulator.Sequence(elements=elements)
        meta = {}
    else:
        raise SystemExit('unknown config ' + cfg)
    lw = Lowering(['pos', 'dir', 'polarization', 'energy', 'probability'], meta=meta)
    inst._lower(lw)
    return lw.finish()


def main():
    cfg = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else '/tmp/jit_' + cfg
    os.makedirs(out, exist_ok=True)
    os.environ['MXB_CACHE_DIR'] = out
    for f in glob.glob(os.path.join(out, '*.cubin')):
        os.unlink(f)
    prog = program(cfg)
    cols = _lib.MxbColumns()
    for k in range(11 + len(prog.out_f64)):
        cols.f64[k] = 0x1000 + 16 * k
    for k in range(len(prog.out_i64)):
        cols.i64[k] = 0x8000 + 16 * k
    lib = _lib.load(False)
    buf = ctypes.create_string_buffer(1 << 20)
    n = lib.mxb_jit_source(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cols), buf, len(buf))
    assert n > 0, lib.mxb_last_error()
    open(os.path.join(out, 'kernel.cu'), 'wb').write(buf.value)
    n = lib.mxb_jit_compile(prog.blob.ctypes.data, prog.blob.size, ctypes.byref(cols))
    if n < 0:
        raise SystemExit(lib.mxb_last_error().decode())
    print(lib.mxb_jit_info().decode(), n, 'bytes; ops', prog.n_ops, 'stage words', prog.stage_words)
    for cubin in glob.glob(os.path.join(out, '*.cubin')):
        r = subprocess.run(['cuobjdump', '-res-usage', cubin], capture_output=True, text=True).stdout
        print(r.strip().splitlines()[-1])
        sass = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True).stdout
        import re
        lines = [l for l in sass.splitlines() if re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+\S', l)]
        print('static SASS instructions:', len(lines), '=', len(lines) * 16 // 1024, 'KB')
        open(os.path.join(out, 'kernel.sass'), 'w').write(sass)


if __name__ == '__main__':
    main()
