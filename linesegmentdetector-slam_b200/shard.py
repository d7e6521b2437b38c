"""Map / scan-frame sharding across GPUs: independent units, contiguous split, NO data-path collective.

One process per GPU (torchrun).  Every rank runs the whole LSD pipeline on its own contiguous slice of the batch;
the only communication is control-plane: a barrier around timed regions and small reductions of result counts /
timings (NCCL on GPUs, gloo in the CPU tests).  Segments stay on the rank that produced them unless the caller asks
for `gather_segments`, which concatenates the per-rank tables in batch order on rank 0."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced split: the first n_items % world ranks get one extra item.  Returns (first, count)."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n={n_items} rank={rank} world={world}")
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(item, n_items, world):
    """Rank that owns `item` under shard_range."""
    if not (0 <= item < n_items):
        raise ValueError("item out of range")
    base, extra = divmod(n_items, world)
    cut = extra * (base + 1)
    return item // (base + 1) if item < cut else extra + (item - cut) // max(base, 1)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def all_reduce_scalar(value, op="sum", device=None):
    """sum / max of a python scalar over the ranks (identity without a process group)."""
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_counts(local_counts, n_items, device=None):
    """Per-map segment counts of the whole batch, in batch order, on every rank."""
    dist = _dist()
    local = np.asarray(local_counts, np.int64)
    if dist is None:
        return local.copy()
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    first, cnt = shard_range(n_items, rank, world)
    assert cnt == len(local), (cnt, len(local))
    full = torch.zeros(n_items, dtype=torch.int64, device=device or "cpu")
    full[first:first + cnt] = torch.as_tensor(local, device=full.device)
    dist.all_reduce(full)        # disjoint slices: the sum is the concatenation
    return full.cpu().numpy()


def gather_segments(local_tables, n_items, dst=0):
    """Concatenate per-map segment tables (list of (k_i, ncol) float arrays) in batch order on rank `dst`."""
    dist = _dist()
    if dist is None:
        return list(local_tables)
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == dst else None
    dist.gather_object(list(local_tables), out, dst=dst)
    if rank != dst:
        return None
    merged = []
    for r in range(world):
        merged.extend(out[r])
    assert len(merged) == n_items
    return merged
