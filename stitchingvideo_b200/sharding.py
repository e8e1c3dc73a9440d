"""Frame sharding across the GPUs of one box (SURVEY.md §8e "Throughput" mode).

Video frames are independent once calibration is fixed, so GPU g of G takes the frame sets
f = g (mod G); the per-sequence tables are replicated on every GPU and there is NO collective on
the data path.  The only exchanges are control-plane: an optional one-time broadcast of the
calibration (K, R, gains, seam masks — a few KB to a few MB) from rank 0, and restoring output
order on the consumer.  torch.distributed is plumbing only (gloo on CPU in the tests, NCCL on the box).
"""
import numpy as np


def frames_for_rank(n_frames, rank, world):
    """Frame indices rank `rank` of `world` composes: rank, rank + world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_frames, world))


def owner_of(frame_idx, world):
    return frame_idx % world


def interleave(per_rank):
    """Restore stream order from per-rank result lists (rank r holds frames r, r+G, ...)."""
    world = len(per_rank)
    total = sum(len(p) for p in per_rank)
    out = []
    for f in range(total):
        r, k = f % world, f // world
        if k >= len(per_rank[r]):
            raise ValueError("rank %d is missing its result %d" % (r, k))
        out.append(per_rank[r][k])
    return out


def broadcast_calibration(cal, src=0):
    """Replicate the calibration dict (numpy arrays / plain values) from rank `src` to every rank.
    A no-op without an initialised process group (single GPU)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return cal
    box = [cal if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def checksum(a):
    """Order-sensitive 64-bit checksum of an array (cheap panorama identity across ranks)."""
    b = np.ascontiguousarray(a).view(np.uint8).ravel().astype(np.uint64)
    idx = (np.arange(b.size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    return int((b * idx).sum() % np.uint64(0xFFFFFFFFFFFFFFC5))
