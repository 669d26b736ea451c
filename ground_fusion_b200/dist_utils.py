"""Multi-GPU plumbing for the front end: one process per GPU, rank r tracks stream r.

The data path has no collective (streams are independent); the only exchanges are (a) the max-over-ranks of the
timed region for the benchmark and (b) an optional gather of every rank's `gf_obs` array on rank 0 when a single
consumer wants all tracks (<= max_cnt * 72 B per frame per rank: latency-, not bandwidth-bound).
torch.distributed is plumbing only: backend "nccl" on the GPU box, "gloo" in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def stream_seed_for_rank(rank, base_seed=0):
    """rank r <-> camera stream r (weak scaling: per-GPU work is fixed as N grows)."""
    return base_seed + rank


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_tracks(obs, max_cnt, dst=0, device=None):
    """Gathers each rank's structured gf_obs array (dtype ground_fusion_b200._lib.OBS_DTYPE) on rank `dst`.
    Returns a list with one array per rank on dst, None elsewhere."""
    from ._lib import OBS_DTYPE
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obs]
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = np.zeros(max_cnt * OBS_DTYPE.itemsize + 8, np.uint8)
    raw = np.ascontiguousarray(obs).view(np.uint8).ravel()
    buf[:8] = np.frombuffer(np.int64(len(obs)).tobytes(), np.uint8)
    buf[8:8 + raw.size] = raw
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, outs, dst=dst)
    if rank != dst:
        return None
    res = []
    for o in outs:
        b = o.cpu().numpy()
        n = int(np.frombuffer(b[:8].tobytes(), np.int64)[0])
        res.append(b[8:8 + n * OBS_DTYPE.itemsize].view(OBS_DTYPE).copy())
    return res
