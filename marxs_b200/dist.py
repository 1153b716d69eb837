"""Multi-GPU: static sharding of independent photons + one reduction epilogue.

Photons never interact, so the trace shards with NO data-path collective: rank r
of W traces the contiguous global photon-id range ``shard_range(N, r, W)`` with
replicated element tables, and the Philox counter is the GLOBAL photon id, so
per-photon results are bit-identical for any GPU count or batch size.  The only
exchange is the final all-reduce of detector images / order histograms (fp64
sum; integer count images are bit-reproducible) and, on request, an
all-gather-v of compacted event lists.  Backend: ``nccl`` on GPUs, ``gloo`` in
the CPU tests of this host logic."""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of global photon ids owned by ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa(device_index):
    """Best effort: pin this process (and so the first-touch placement of the pinned host buffers it
    allocates afterwards) to the CPUs of the NUMA node the GPU hangs off.  With one process per GPU the
    host<->device copies of all ranks then run against local memory instead of crossing the socket
    interconnect.  Returns the node number or None."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = '{0:04x}:{1:02x}:{2:02x}.0'.format(props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open('/sys/bus/pci/devices/{0}/numa_node'.format(bus)) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node{0}/cpulist'.format(node)) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


def allreduce_images(tensors):
    """In-place sum over ranks of detector images / histograms (one collective per tensor)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def gather_events(columns, dst=0):
    """All ranks send their compacted event columns (dict name -> 1-D tensor, same names,
    ragged lengths) to ``dst``; returns the concatenated dict there, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return columns
    world, rank = dist.get_world_size(), dist.get_rank()
    names = sorted(columns.keys())
    n_local = torch.tensor([len(columns[names[0]])], dtype=torch.int64, device=columns[names[0]].device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    out = {} if rank == dst else None
    for name in names:
        t = columns[name]
        pad = torch.zeros(nmax, dtype=t.dtype, device=t.device)
        pad[:len(t)] = t
        bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
        if rank == dst:
            out[name] = torch.cat([b[:c] for b, c in zip(bufs, counts)])
    return out


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing is the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
