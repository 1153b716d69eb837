#!/usr/bin/env python
"""End-to-end (host buffers) throughput of the C2 trace against the host-pipeline chunk size.
The number of D2H streams is read once per process: run with MXB_HOST_OUT_STREAMS=1 / 2."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from marxs_b200 import host as mhost  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
host_in = mhost.HostPhotonTable.from_columns(bench.synth_c2_numpy(n, 777))
host_in.meta['ROLL_PNT'] = (0., 'roll')
host_out = mhost.HostPhotonTable(n)
inst = bench.c2_instrument()
prog = mhost.lower([inst], host_in.colnames, host_in.meta)
mhost.trace_host(inst, host_in, out=host_out, program=prog)
h2d, d2h = mhost.h2d_d2h_bytes(prog, n, in_place=False)
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22):
    mhost.trace_host(inst, host_in, out=host_out, program=prog, chunk=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        mhost.trace_host(inst, host_in, out=host_out, program=prog, chunk=chunk)
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps(dict(out_streams=os.environ.get('MXB_HOST_OUT_STREAMS', '2'), chunk=chunk, ms=dt * 1e3,
                          photons_per_s=n / dt, pcie_gbs=(h2d + d2h) / dt / 1e9)), flush=True)
