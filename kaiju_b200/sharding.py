"""Read sharding across GPUs: contiguous ranges of read indices per rank, full index replica per GPU, no
collective during search; ONE all-gather of the per-read taxon array at the end, and/or ONE all-reduce of the dense
per-taxon count vector (the kaiju2table-style summary).  Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of the item indices owned by `rank`."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_packed(seq, off, lo, hi):
    """Sub-batch [lo, hi) of a packed (bytes, offsets) read set, offsets rebased to 0."""
    b, e = int(off[lo]), int(off[hi])
    return seq[b:e], (off[lo:hi + 1] - off[lo]).astype(np.uint64)


def all_gather_taxa(local_taxa, n_items, rank, world, dist):
    """All-gather per-read results of unequal shard sizes (pads to the largest shard).  `local_taxa` is a torch int64
    tensor on the backend's device; returns the full [n_items] tensor in input order."""
    import torch
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    pad = torch.zeros(mx, dtype=local_taxa.dtype, device=local_taxa.device)
    pad[: local_taxa.numel()] = local_taxa
    out = torch.empty(mx * world, dtype=local_taxa.dtype, device=local_taxa.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (h - l)] for r, (l, h) in enumerate(sizes)])


class _DevArray:
    """A device buffer owned by the library, exposed through __cuda_array_interface__ so torch can wrap it without a copy."""
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def all_reduce_counts(counts, dist, clf=None):
    """Sum the per-taxon read counts of all ranks.  `counts` is a torch int64 tensor (gloo tests), or None to reduce the
    classifier's own count vector in HBM in place (NCCL; `clf.counts()` then returns the job-wide table on every rank)."""
    import torch
    if counts is None:
        ptr, n = clf.counts_device_ptr
        counts = torch.as_tensor(_DevArray(ptr, n, "<i8"), device="cuda:%d" % clf.device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
