"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing used by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ground_fusion_b200 import dist_utils
    from ground_fusion_b200._lib import OBS_DTYPE
    mx = dist_utils.max_over_ranks([1.0 + rank, 5.0 - rank])
    obs = np.zeros(3 + rank, OBS_DTYPE)
    obs["id"] = np.arange(3 + rank) + 100 * rank
    obs["v"][:, 3] = rank
    got = dist_utils.gather_tracks(obs, max_cnt=150)
    if rank == 0:
        q.put((mx, [(len(g), int(g["id"][-1]), float(g["v"][0, 3])) for g in got], dist_utils.stream_seed_for_rank(1)))
    else:
        assert got is None
        q.put((mx, None, None))
    dist.destroy_process_group()


def test_max_over_ranks_and_gather_tracks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for mx, _, _ in res:
        assert mx == [2.0, 5.0]
    gathered = [r for r in res if r[1] is not None][0]
    assert gathered[1] == [(3, 2, 0.0), (4, 103, 1.0)] and gathered[2] == 1


def test_single_process_is_identity():
    from ground_fusion_b200 import dist_utils
    from ground_fusion_b200._lib import OBS_DTYPE
    assert dist_utils.max_over_ranks([3.0]) == [3.0]
    o = np.zeros(2, OBS_DTYPE)
    assert dist_utils.gather_tracks(o, 150)[0] is o
