#!/usr/bin/env python
"""PCIe ceiling of the box: pinned H2D, D2H and both at once (what bounds bench.py's e2e)."""
import json
import torch

n = 1 << 28                      # 256 Mi doubles = 2 GiB
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_a = torch.empty(n, dtype=torch.float64, device='cuda')
d_b = torch.ones(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = n * 8


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {}
for name, fn, nb in (('h2d', h2d, nbytes), ('d2h', d2h, nbytes), ('both', both, 2 * nbytes)):
    fn()
    ms = timed(fn)
    out[name + '_gbs'] = nb / ms / 1e6
print(json.dumps(out))
