"""Multi-GPU host logic: voices are sharded over the ranks of one box, every rank mixes its own voices
into a private copy of the bus buffers, and the partial buffers are summed.

The reference has no analogue (one audio thread, audio_spatializer.cpp:353).  Sharding is static and
contiguous by *instance*, so that all voices of one AudioSpatializerInstance stay on one GPU (Mode A
pre-sums per instance, Q15) and a voice's persistent state never migrates.
"""
import numpy as np


def instance_range(n_instances, world_size, rank):
    """[lo, hi) of the instances owned by `rank`: contiguous, sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(int(n_instances), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of_instance(instance, n_instances, world_size):
    """Rank that owns `instance` (inverse of instance_range)."""
    base, extra = divmod(int(n_instances), world_size)
    instance = np.asarray(instance)
    split = extra * (base + 1)
    return np.where(instance < split, instance // max(base + 1, 1), extra + (instance - split) // max(base, 1)).astype(np.int32)


def shard_voices(voices, n_instances, world_size, rank):
    """Sub-list of an abi.voice array owned by `rank`, with voice / instance / src_row re-based to the
    rank-local slot numbering (local instance = instance - lo; local voice = position in the shard;
    src_row = position in the shard, i.e. the rank only uploads its own source rows).
    Returns (local_voices, global_index) where global_index[k] is the position of local voice k in `voices`."""
    lo, hi = instance_range(n_instances, world_size, rank)
    v = np.asarray(voices)
    mine = np.nonzero((v["instance"] >= lo) & (v["instance"] < hi))[0]
    out = v[mine].copy()
    out["instance"] -= lo
    out["voice"] = np.arange(len(mine), dtype=np.int32)
    has_src = out["src_row"] >= 0
    out["src_row"] = np.where(has_src, np.arange(len(mine), dtype=np.int32), -1)
    return out, mine


def reduce_bus(bus, dist=None, root=None):
    """Sum of the per-rank partial bus buffers (a torch tensor, in place): all-reduce, or reduce to `root`.
    With NCCL this runs over NVLink / NVSwitch; the CPU tests run it over gloo."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bus
    if root is None:
        dist.all_reduce(bus)
    else:
        dist.reduce(bus, dst=root)
    return bus
