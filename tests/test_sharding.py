"""CPU: frame sharding — host logic, and a world_size-2 gloo run in which each rank
composes its own frames (through the oracle here, there being no GPU) and the interleaved stream
equals the unsharded one."""
import os
import socket

import numpy as np
import pytest

from stitchingvideo_b200 import sharding


def test_frames_for_rank_partition():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 33):
            parts = [sharding.frames_for_rank(n, r, world) for r in range(world)]
            assert sorted(f for p in parts for f in p) == list(range(n))
            assert all(sharding.owner_of(f, world) == r for r, p in enumerate(parts) for f in p)
            assert sharding.interleave(parts) == list(range(n))
    with pytest.raises(ValueError):
        sharding.frames_for_rank(4, 2, 2)
    with pytest.raises(ValueError):
        sharding.interleave([[0, 2], []])


def test_checksum_is_order_sensitive():
    a = np.arange(24, dtype=np.uint8).reshape(2, 4, 3)
    assert sharding.checksum(a) == sharding.checksum(a.copy())
    assert sharding.checksum(a) != sharding.checksum(a[::-1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from oracle import pipeline as P
    from stitchingvideo_b200 import rigs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 owns the calibration; every rank receives a replica (control plane only)
        Ks, Rs, spec = rigs.cameras("mini_cyl")
        cal_in = {"Ks": Ks, "Rs": Rs, "scale": spec["scale"]} if rank == 0 else None
        cal_in = sharding.broadcast_calibration(cal_in, src=0)
        cal = P.Calibration((spec["W"], spec["H"]), cal_in["Ks"], cal_in["Rs"], "cylindrical", cal_in["scale"])
        mine = []
        for f in sharding.frames_for_rank(n_frames, rank, world):
            frames = [rigs.frame("mini_cyl", f, i) for i in range(spec["n_used"])]
            pano, _ = P.compose(cal, frames, blender="feather")
            mine.append(sharding.checksum(pano))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            q.put(sharding.interleave(gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_unsharded():
    import torch.multiprocessing as mp
    from oracle import pipeline as P
    from stitchingvideo_b200 import rigs
    n_frames, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    cal = P.Calibration((spec["W"], spec["H"]), Ks, Rs, "cylindrical", spec["scale"])
    want = []
    for f in range(n_frames):
        pano, _ = P.compose(cal, [rigs.frame("mini_cyl", f, i) for i in range(spec["n_used"])], blender="feather")
        want.append(sharding.checksum(pano))
    assert got == want
