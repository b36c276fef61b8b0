"""Multi-GPU plumbing of the inference path: one process per GPU, clips sharded round-robin over ranks,
NO data-path collective -- a clip is the unit of independence (tracker memory and the ref-frame chain are per
clip, panoptic_fusetrack.py:393-406).  torch.distributed is used only to agree on timings / gather small results."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None, device=None):
    """Initialise the process group from the torchrun environment (no-op for a single process)."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_clips(clip_ids, rank, world):
    """Round-robin assignment of whole clips to ranks; every clip lands on exactly one rank."""
    return [c for i, c in enumerate(clip_ids) if i % world == rank]


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (device timings -> job time)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_objects(obj):
    """Gather small python results (e.g. per-clip track-id tables) on every rank."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
