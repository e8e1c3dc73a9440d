"""Latency ("strip") mode — SURVEY.md §8e: column strips with pyramid-halo exchange.

-m gpu: strips produced by G handles in lock step on one GPU (halo messages handed over through device
buffers, exactly the bytes NCCL would carry) reassemble the unsplit panorama bit for bit.
CPU: the send/recv schedule itself under a world_size-3 gloo group with a host-memory stand-in for the
compositor (every rank receives exactly what its neighbours packed, no deadlock, sizes agree)."""
import os
import socket

import numpy as np
import pytest

from stitchingvideo_b200 import rigs, strips


def test_schedule_shape():
    s = strips.schedule(5)
    assert s[0] == ("warp",)
    assert [x for x in s if x[0] == "down"] == [("down", l) for l in range(5)]
    assert [x for x in s if x[0] == "band"] == [("band", l) for l in range(5, -1, -1)]
    # every Gaussian level is exchanged before the pyrDown that reads it; every restored band right after its band
    for l in range(5):
        assert s.index(("exchange", strips.GAUSS, l)) < s.index(("down", l))
    for l in range(1, 6):
        assert s.index(("exchange", strips.RESTORED, l)) == s.index(("band", l)) + 1
    assert strips.neighbour(0, 4, strips.LEFT) is None and strips.neighbour(3, 4, strips.RIGHT) is None
    assert strips.neighbour(1, 4, strips.LEFT) == 0 and strips.neighbour(1, 4, strips.RIGHT) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("weight_type", ["f32", "s16"])
def test_strips_reassemble_the_panorama(gpu, world, weight_type):
    Ks, Rs, spec = rigs.cameras("mini")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    wt = gpu.CV_32F if weight_type == "f32" else gpu.CV_16S
    mk = lambda: gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5,
                                weight_type=wt, gains=spec["gain_values"])
    whole = mk()
    # separate handles per halo mode: a mode must not be able to lean on halo columns the other one left behind
    comps = {strips.run_local: [mk() for _ in range(world)], strips.run_local_recompute: [mk() for _ in range(world)]}
    ranges = [whole.strip_range(r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == whole.pano_size[0]
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    for fi in range(2):
        frames = [rigs.frame("mini", fi, i) for i in range(n)]
        pano, mask = whole.compose(frames)
        for run in (strips.run_local, strips.run_local_recompute):       # halo exchange / halo recompute
            parts = run(comps[run], frames)
            got = np.concatenate([p[0] for p in parts], axis=1)
            gmask = np.concatenate([p[1] for p in parts], axis=1)
            assert got.shape == pano.shape
            assert np.array_equal(got, pano), "%s: %d differing values" % (run.__name__, int((got != pano).sum()))
            assert np.array_equal(gmask, mask)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4])
def test_peer_memory_halo_exchange(gpu, world):
    """the halo exchange without the host in the loop - each handle's kernels write its edge columns into the
    neighbours' receive areas and raise their flag words; one C call per rank and frame.  Here the "ranks" are handles on one
    device (plain pointers instead of IPC handles); frame after frame the strips reassemble the unsplit panorama bit for bit."""
    Ks, Rs, spec = rigs.cameras("mini")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    mk = lambda: gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5, gains=spec["gain_values"])
    whole = mk()
    comps = [mk() for _ in range(world)]
    for fi in range(3):                                      # the sequence numbers keep counting over the frames
        frames = [rigs.frame("mini", fi, i) for i in range(n)]
        pano, mask = whole.compose(frames)
        before = gpu.kernel_launch_count()
        parts = strips.run_local_peer(comps, frames, connected=fi > 0)
        assert gpu.kernel_launch_count() > before
        assert np.array_equal(np.concatenate([p[0] for p in parts], axis=1), pano), "frame %d" % fi
        assert np.array_equal(np.concatenate([p[1] for p in parts], axis=1), mask)


@pytest.mark.gpu
def test_strips_written_into_views_of_one_panorama(gpu):
    """strip results copied into column VIEWS of one full-size panorama buffer (pitch = the panorama's) must
    only touch their own columns - a linear copy over such a view would overwrite the neighbouring strips."""
    Ks, Rs, spec = rigs.cameras("mini")
    size, n, world = (spec["W"], spec["H"]), spec["n_used"], 3
    mk = lambda: gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5, gains=spec["gain_values"])
    whole = mk()
    frames = [rigs.frame("mini", 1, i) for i in range(n)]
    pano, mask = whole.compose(frames)
    got, gmask = np.full_like(pano, 77), np.full_like(mask, 77)
    for r in reversed(range(world)):                      # last strip first: an overrun to the right would be seen
        c = mk()
        c.set_strip(r, world)
        c.set_strip_halo(True)
        c.strip_compose(frames)
        c.strip_result(r, world, got, gmask)
    assert np.array_equal(got, pano) and np.array_equal(gmask, mask)


@pytest.mark.gpu
def test_strips_full_size_c3(gpu):
    Ks, Rs, spec = rigs.cameras("c3")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    mk = lambda: gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5,
                                gains=spec["gain_values"])
    whole = mk()
    frames = [rigs.frame("c3", 2, i, smooth=1) for i in range(n)]
    pano, mask = whole.compose(frames)
    for run in (strips.run_local, strips.run_local_recompute):
        parts = run([mk() for _ in range(4)], frames)
        assert np.array_equal(np.concatenate([p[0] for p in parts], axis=1), pano), run.__name__
        assert np.array_equal(np.concatenate([p[1] for p in parts], axis=1), mask), run.__name__


@pytest.mark.gpu
def test_strip_mode_errors(gpu):
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    comp = gpu.Compositor((spec["W"], spec["H"]), Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")
    with pytest.raises(gpu.StitchError) as e:
        comp.set_strip(0, 2)                    # multi-band path only
    assert e.value.code == -213
    Ks, Rs, spec = rigs.cameras("mini")
    comp = gpu.Compositor((spec["W"], spec["H"]), Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband")
    with pytest.raises(gpu.StitchError):
        comp.strip_band(0)                      # before set_strip
    with pytest.raises(gpu.StitchError):
        comp.set_strip(0, 64)                   # strips narrower than 2 * 2^num_bands columns


# ------------------------------------------------------------------ CPU: the exchange schedule under gloo
class FakeComp:
    """Host-memory stand-in: each (what, level, side) message is a deterministic byte pattern of the sender."""
    num_bands = 3

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.received = {}
        self.keep = []

    @staticmethod
    def size(what, level, lo_rank):                 # message size depends on the boundary (lo_rank | lo_rank + 1), not the sender
        return 16 * (1 + level) * (2 if what == strips.GAUSS else 1) + lo_rank

    def strip_halo_bytes(self, what, level, side):
        lo = self.rank - 1 if side == strips.LEFT else self.rank
        n = self.size(what, level, lo)
        return n, n

    def strip_pack(self, what, level, side, ptr):
        import ctypes
        n, _ = self.strip_halo_bytes(what, level, side)
        pat = (np.arange(n) * 7 + self.rank * 31 + what * 5 + level * 3 + side) % 251
        ctypes.memmove(ptr, pat.astype(np.uint8).ctypes.data, n)

    def strip_unpack(self, what, level, side, ptr):
        import ctypes
        _, n = self.strip_halo_bytes(what, level, side)
        buf = np.empty(n, np.uint8)
        ctypes.memmove(buf.ctypes.data, ptr, n)
        self.received[(what, level, side)] = buf

    def set_strip(self, rank, world):
        pass

    def set_strip_halo(self, recompute):
        assert not recompute


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comp = FakeComp(rank, world)
        sc = strips.StripCompositor(comp, rank, world, transport=strips.TorchTransport(comp, rank, world, device=None))
        for step in sc.steps:
            if step[0] == "exchange":
                sc.transport.exchange(step[1], step[2])
        ok = True
        for (what, level, side), got in comp.received.items():
            nbr = strips.neighbour(rank, world, side)
            other = strips.RIGHT if side == strips.LEFT else strips.LEFT
            want = (np.arange(got.size) * 7 + nbr * 31 + what * 5 + level * 3 + other) % 251
            ok = ok and np.array_equal(got, want.astype(np.uint8))
        n_expected = sum(1 for s in sc.steps if s[0] == "exchange") * sum(strips.neighbour(rank, world, sd) is not None for sd in (0, 1))
        q.put((rank, ok and len(comp.received) == n_expected))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_schedule_under_gloo():
    import torch.multiprocessing as mp
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True, 2: True}
