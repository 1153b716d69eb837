#!/usr/bin/env python
"""Static SASS size per source function of a cubin compiled with -lineinfo (nvdisasm -g): where the code of a
specialised kernel comes from.  Usage: python tools/sass_by_function.py <cubin> [csrc dir]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def funcs(path):
    out = []
    for n, l in enumerate(open(path), 1):
        m = re.match(r'^(?:MXB_DEV|MXB_LIBM|MXB_SCAN_ATTR|__device__|static|inline)[^;]*?\b(\w+)\s*\(', l)
        if m:
            out.append((n, m.group(1)))
    return out


def main():
    cubin = sys.argv[1]
    csrc = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'marxs_b200', 'csrc')
    fm = {f: funcs(os.path.join(csrc, f)) for f in ('mxb_device.cuh', 'mxb_ops.cuh')}
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    cur = ('?', 0)
    byfn, byline = collections.Counter(), collections.Counter()
    for l in txt.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+\S', l):
            f, n = cur
            name = 'line %d' % n
            if f in fm:
                name = '?'
                for s, nm in fm[f]:
                    if s <= n:
                        name = nm
                    else:
                        break
            else:
                f = 'kernel.cu'
            byfn[(f, name)] += 1
            byline[(f, n)] += 1
    tot = sum(byfn.values())
    print('total', tot, 'instructions =', tot * 16 // 1024, 'KB')
    for k, v in byfn.most_common(40):
        print('%-16s %-28s %6d %5.1f%%' % (k[0], k[1], v, 100. * v / tot))


if __name__ == '__main__':
    main()
