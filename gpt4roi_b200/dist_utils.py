"""torch.distributed plumbing for the sample-sharded (replica) path.

The region-token forward has no data-path collective: every stage is per-sample,
so multi-GPU inference is N replicas with the batch split across ranks -- RoIs travel with their
image and `roi_batch_ind` is re-based to the local shard.  NCCL (or gloo in the CPU tests) is used
only for the barrier and the max-over-ranks reduction of the measured time.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return (int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')),
            int(os.environ.get('LOCAL_RANK', '0')))


def init(backend=None, device=None):
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {}
        if backend == 'nccl' and device is not None:
            kw['device_id'] = device
        dist.init_process_group(backend, **kw)
    return world, rank, local


def barrier(device=None):
    if dist.is_initialized():
        dist.barrier()
    if device is not None and torch.device(device).type == 'cuda':
        torch.cuda.synchronize(device)


def max_over_ranks(value, device='cpu'):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch(input_ids, images, bboxes, world, rank):
    """Contiguous split of a global batch over ranks (sizes differ by at most one).  Returns the
    local (input_ids, images, bboxes); boxes stay attached to their images, so the local
    roi_batch_ind produced downstream (layers.py:294-302 semantics) is already re-based."""
    B = input_ids.shape[0]
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    bb = None if bboxes is None else list(bboxes[lo:hi])
    return input_ids[lo:hi], images[lo:hi], bb, (lo, hi)


def gather_rows(local, world, rank, sizes, device='cpu'):
    """All-gather per-sample result rows (e.g. last-token logits) back into global batch order."""
    if world == 1:
        return local
    outs = [torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=device) for n in sizes]
    dist.all_gather(outs, local.contiguous()) if len(set(sizes)) == 1 else _uneven_gather(outs, local, sizes, device)
    return torch.cat(outs, 0)


def _uneven_gather(outs, local, sizes, device):
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)
    for o, b, n in zip(outs, bufs, sizes):
        o.copy_(b[:n])
