"""Batch sharding across the GPUs of one box.

Images are independent (frozen ranges, eval-mode BN, per-image top-K), so the forward path has NO collective:
rank r of G takes a contiguous slice of the batch, weights are replicated.  The only optional exchange is one
all_gather of the detections ([B/G, K, 6] fp32 = 2.4 KB per image), which runs over NCCL (NVLink 5 / NVSwitch) on the
GPU box and over gloo in the CPU tests.
"""
from typing import List, Tuple


def shard_bounds(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) slices: the first n_items % world ranks take one extra item."""
    if world < 1 or n_items < 0:
        raise ValueError("shard_bounds: need world >= 1 and n_items >= 0")
    base, extra = divmod(n_items, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def my_slice(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return shard_bounds(n_items, world)[rank]


def gather_detections(local_dets, n_items: int, group=None):
    """all_gather of per-rank detections [n_local, K, 6] -> [n_items, K, 6] on every rank (uneven shards padded).

    `local_dets` lives on the device of the process group's backend (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_items, world)
    b, e = bounds[rank]
    if local_dets.shape[0] != e - b:
        raise ValueError("rank %d holds %d images, its shard is %d" % (rank, local_dets.shape[0], e - b))
    cap = max(x[1] - x[0] for x in bounds)
    pad = torch.zeros((cap,) + tuple(local_dets.shape[1:]), dtype=local_dets.dtype, device=local_dets.device)
    pad[: e - b] = local_dets
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: x[1] - x[0]] for p, x in zip(parts, bounds)], 0)
