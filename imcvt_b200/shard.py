"""Frame sharding across GPUs: every frame is an independent unit (JLSencodeScan / HEVCImageEncoder
re-initialise all state per call, src/imageio_jls.c:256-267, src/HEVCe/HEVCe.c:1577-1599), so ranks encode disjoint
contiguous frame blocks and the only exchange is an all-gather of the per-frame stream sizes, from which every rank
derives the global output offsets.  Host logic only; works with any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests)."""
import torch


def frame_block(rank, world, n_total):
    """Contiguous block [first, first + count) of rank `rank`: ceil(n/world) frames per rank, the tail may be short."""
    per = (n_total + world - 1) // world
    first = min(n_total, rank * per)
    return first, min(per, n_total - first)


def gather_sizes(local_sizes, n_total, world, group=None):
    """local_sizes: int64 tensor [count] of this rank's stream sizes (any device the backend supports).
    Returns the int64 tensor [n_total] of all sizes in frame order, on every rank."""
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    padded = torch.zeros(per, dtype=torch.int64, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    if world == 1:
        return padded[:n_total]
    out = torch.empty(world * per, dtype=torch.int64, device=local_sizes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]


def global_offsets(sizes):
    """Exclusive prefix sum: byte offset of every frame's stream in the concatenated output."""
    s = sizes.to(torch.int64)
    return torch.cumsum(s, 0) - s
