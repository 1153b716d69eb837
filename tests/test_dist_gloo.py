"""world_size-2 tests (gloo, CPU) of the multi-GPU host logic: photon-range sharding
and the reduction epilogue.  The trace itself needs no collective."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from marxs_b200 import dist as mdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n, w in ((10, 3), (1_000_000_000, 8), (7, 8), (0, 4)):
        spans = [mdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch
    import torch.distributed as dist
    from marxs_b200 import dist as mdist
    rank, world, local = mdist.init_from_env('gloo')
    assert world == 2 and dist.get_backend() == 'gloo'
    lo, hi = mdist.shard_range(1001, rank, world)
    # each rank "detects" its own shard: event pixel = global photon id mod 16, weight 0.5
    ids = torch.arange(lo, hi)
    img = torch.zeros(16, dtype=torch.float64)
    img.index_add_(0, ids % 16, torch.full((hi - lo,), 0.5, dtype=torch.float64))
    cnt = torch.bincount(ids % 16, minlength=16)
    mdist.allreduce_images([img, cnt])
    ref = np.bincount(np.arange(1001) % 16, minlength=16)
    assert np.array_equal(cnt.numpy(), ref) and np.allclose(img.numpy(), 0.5 * ref)
    ev = mdist.gather_events({{'id': ids[ids % 7 == 0], 'w': (ids[ids % 7 == 0]).double() * 2}})
    if rank == 0:
        # rows land in rank order = global photon-id order, no padding, no sort needed
        assert np.array_equal(ev['id'].numpy(), np.arange(0, 1001, 7))
        assert np.allclose(ev['w'].numpy(), ev['id'].numpy() * 2.)
    else:
        assert ev is None
    # ragged to the extreme: rank 1 has no events at all; destination buffers preallocated and reused
    mine = ids[:5] if rank == 0 else ids[:0]
    out = {{'id': torch.full((64,), -7, dtype=torch.int64)}} if rank == 0 else None
    ev = mdist.gather_events({{'id': mine}}, out=out)
    if rank == 0:
        assert ev['id'].tolist() == [0, 1, 2, 3, 4] and ev['id'].data_ptr() == out['id'].data_ptr()
    # gather on another rank, counts known beforehand
    cnt2 = [3, 4]
    ev = mdist.gather_events({{'x': torch.arange(cnt2[rank], dtype=torch.float64) + 10 * rank}}, dst=1, counts=cnt2)
    if rank == 1:
        assert ev['x'].tolist() == [0., 1., 2., 10., 11., 12., 13.]
    else:
        assert ev is None
    assert mdist.gather_counts(3 + rank, torch.device('cpu')) == [3, 4]
    t = mdist.max_over_ranks(1.0 + rank, torch.device('cpu'))
    assert t == 2.0
    dist.barrier()
    dist.destroy_process_group()
    print('rank', rank, 'ok')
''')


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_epilogue_two_ranks_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT))
    port = free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert 'rank {0} ok'.format(r) in o
