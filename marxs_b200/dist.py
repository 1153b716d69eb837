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


def gather_counts(n_local, device):
    """All ranks' event counts as a python list (one small all-gather, one host read)."""
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[dist.get_rank()] = int(n_local)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return [int(c) for c in counts.tolist()]


def gather_events(columns, dst=0, counts=None, out=None):
    """Gather-v of compacted event lists on rank ``dst``.

    ``columns``: dict name -> 1-D tensor (same names and dtypes on every rank, ragged lengths).  One small
    collective exchanges the counts, then ONE grouped point-to-point exchange (``batch_isend_irecv`` =
    ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on NCCL) moves every plane of every rank straight
    into its final place in the destination columns: rank r's rows land at offset sum(counts[:r]) of each
    column, so nothing is padded to the longest list and nothing is re-packed.  Returns the dict of
    concatenated columns on ``dst`` (rank order = global photon-id order for contiguous shards), None on
    the other ranks.  ``counts``: the per-rank lengths if the caller already knows them; ``out``: dict of
    preallocated destination tensors of at least sum(counts) rows (dst only) to reuse memory."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return columns
    world, rank = dist.get_world_size(), dist.get_rank()
    names = sorted(columns.keys())
    first = columns[names[0]]
    n_local = int(first.shape[0])
    for name in names:
        if columns[name].dim() != 1 or columns[name].shape[0] != n_local:
            raise ValueError('event columns must be 1-D and of equal length on a rank')
    if counts is None:
        counts = gather_counts(n_local, first.device)
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    total = offsets[-1]
    ops = []
    result = None
    if rank == dst:
        result = {}
        for name in names:
            t = columns[name]
            buf = out[name] if out is not None and name in out else torch.empty(total, dtype=t.dtype, device=t.device)
            if buf.shape[0] < total or buf.dtype != t.dtype:
                raise ValueError('destination buffer of column {0} is too small or of the wrong dtype'.format(name))
            buf[offsets[rank]:offsets[rank + 1]].copy_(t)
            result[name] = buf[:total]
            for r in range(world):
                if r != dst and counts[r] > 0:
                    ops.append(dist.P2POp(dist.irecv, buf[offsets[r]:offsets[r + 1]], r))
    elif n_local > 0:
        for name in names:
            ops.append(dist.P2POp(dist.isend, columns[name].contiguous(), dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return result


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing is the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
