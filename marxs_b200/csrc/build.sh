#!/bin/bash
# Build libmxb (fast) and libmxb_strict (bit-parity) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# the device headers are embedded so the specialised kernels (mxb_jit.cpp, NVRTC) compile from the same source
python3 - <<'PY'
out = []
for name, path in (('kSrcMxbH', '../../include/mxb.h'), ('kSrcDeviceCuh', 'mxb_device.cuh'), ('kSrcOpsCuh', 'mxb_ops.cuh')):
    text = open(path).read()
    assert ')MXBRAW"' not in text
    # raw string literals are limited to 64 KiB pieces by some host compilers: split on lines
    pieces, cur = [], ''
    for line in text.splitlines(True):
        if len(cur) + len(line) > 12000:
            pieces.append(cur)
            cur = ''
        cur += line
    pieces.append(cur)
    out.append('static const char %s[] =\n%s;\n' % (name, '\n'.join('R"MXBRAW(%s)MXBRAW"' % p for p in pieces)))
new = ''.join(out)
try:
    old = open('mxb_embed.inc').read()
except OSError:
    old = None
if old != new:
    open('mxb_embed.inc', 'w').write(new)
PY
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v"
$NVCC $COMMON -DMXB_FAST -o ../libmxb.so mxb_trace.cu mxb_stats.cu mxb_jit.cpp -ldl 2> build_fast.log || { cat build_fast.log; exit 1; }
$NVCC $COMMON -fmad=false -o ../libmxb_strict.so mxb_trace.cu mxb_stats.cu mxb_jit.cpp -ldl 2> build_strict.log || { cat build_strict.log; exit 1; }
# (compile times would make the tracked logs change with every build)
sed -i '/Compile time/d' build_fast.log build_strict.log
grep -E "registers|spill|error|warning" build_fast.log build_strict.log | grep -v "^$" | head -40
