#!/usr/bin/env python
"""BASELINE.json config 5: Chandra HETG observation of N photons sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5.py --photons 1e9

Every rank runs ONE launch: its shard of the photons is born on the device (PointSource ->
JitterPointing -> Chandra aperture), traced through HRMA -> HETG -> ACIS-S with the detector image
fused, then the epilogue: NCCL all-reduce of the image, device-side compaction of the detected
events (mxb_compact_events) and a gather of the compacted event columns on rank 0.  Photon ids are
global, so the union of the shards is bit-identical to a single-GPU run of all N photons.
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import marxs_b200 as mb  # noqa: E402
from marxs_b200 import dist as mdist, events, source, _lib  # noqa: E402
from marxs_b200.missions import chandra  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--photons', type=float, default=1e9)
    ap.add_argument('--gather', default='energy,order,CCD_ID,chipx,chipy,probability', help='event columns gathered on rank 0')
    ap.add_argument('--max-gather', type=float, default=2e7, help='gather only if the event list is at most this long')
    a = ap.parse_args()
    rank, world, local = mdist.init_from_env('nccl' if int(os.environ.get('WORLD_SIZE', '1')) > 1 else None)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    n_total = int(a.photons)
    lo, hi = mdist.shard_range(n_total, rank, world)
    n = hi - lo
    src = source.PointSource(coords=(30., 10.), flux=float(n_total), geomarea=1.,
                             energy={'energy': np.array([0.3, 0.5, 1., 2., 4., 8.]), 'fluxdensity': np.array([0., 5., 3., 2., 1., .5])})
    pnt = source.JitterPointing(coords=(30., 10.), jitter=np.deg2rad(0.2 / 3600.))
    acis = chandra.ACIS(chips=[4, 5, 6, 7, 8, 9], aimpoint=chandra.AIMPOINTS['ACIS-S'])
    image = torch.zeros((6, 1024, 1024), dtype=torch.float64, device=dev)
    acis.image = image
    elements = [chandra.Aperture(), chandra.HRMA(), chandra.HETG(), acis]
    mb.set_seed(20261017)
    # warm-up: kernel specialisation, NCCL communicator, and the result table (36 GB per 1.25e8 photons:
    # allocated once and reused, like a caller that processes observation after observation)
    photons = source.observe(src, pnt, elements, 1., device=dev, n=n, id0=lo, check=False)
    mdist.allreduce_images([torch.zeros(8, dtype=torch.float64, device=dev)])
    image.zero_()
    mb.set_seed(20261017)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ev0.record()
    photons = source.observe(src, pnt, elements, 1., device=dev, n=n, id0=lo, check=False, out=photons)
    ev1.record()
    mdist.allreduce_images([image])
    cols = [c for c in a.gather.split(',') if c in photons.colnames]
    ev = events.compact(photons, columns=cols, sel='CCD_ID', sel_min=0, weight='probability')
    ev2.record()
    n_ev = torch.tensor([len(ev)], dtype=torch.int64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(n_ev)
    gathered = None
    if int(n_ev.item()) <= a.max_gather:
        gathered = mdist.gather_events({c: ev.storage(c) for c in cols})
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    wall = mdist.max_over_ranks(time.perf_counter() - t0, dev)
    trace_ms = mdist.max_over_ranks(ev0.elapsed_time(ev1), dev)
    epi_ms = mdist.max_over_ranks(ev1.elapsed_time(ev2), dev)
    if rank == 0:
        line = dict(config='C5: Chandra HETG observation, photons born on the device, sharded by global photon id',
                    photons=n_total, n_gpus=world, photons_per_gpu=n, wall_s=wall, photons_per_s=n_total / wall,
                    trace_ms=trace_ms, epilogue_ms=epi_ms, events=int(n_ev.item()),
                    image_sum=float(image.sum()), gathered_on_rank0=(None if gathered is None else len(gathered[cols[0]])),
                    kernel=_lib.load().mxb_jit_info().decode(), columns=len(photons.colnames))
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
