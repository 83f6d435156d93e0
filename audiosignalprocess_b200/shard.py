"""Stream sharding across the GPUs of one box: every stream is
independent, so rank r owns the contiguous slice [lo, hi) of streams and no collective sits on
the data path; the only exchange is the host-side gather of outputs."""
import numpy as np


def shard_range(n_streams, world_size, rank):
    """Contiguous slice of streams owned by `rank`: stream s -> rank s // ceil(N / G)."""
    per = -(-n_streams // world_size)
    lo = min(n_streams, rank * per)
    hi = min(n_streams, lo + per)
    return lo, hi


def process_sharded(process_fn, pcm, dist=None, dst=0):
    """Runs process_fn on this rank's slice of pcm [n_streams, samples] and gathers the int16
    outputs on rank `dst` (returns the full array there, None elsewhere).  dist: an initialised
    torch.distributed module, or None for a single process."""
    n = pcm.shape[0]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return process_fn(pcm)
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n, world, rank)
    local = process_fn(pcm[lo:hi]) if hi > lo else np.zeros((0, pcm.shape[1]), np.int16)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([g for g in gathered if g.shape[0] > 0], axis=0)
