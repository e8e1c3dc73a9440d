"""Latency ("strip") mode — SURVEY.md §8e, BASELINE.json configs[4]: ONE very wide panorama split into
column strips, one rank (GPU) per strip, with the pyramid halo columns exchanged between neighbouring
ranks by send/recv (NCCL over NVLink on the box; gloo in the CPU plumbing test).

The arithmetic lives in libstitchb200 (sb_compositor_strip_*): every stage of the multi-band path runs on
this rank's columns only.  This module is the host-side schedule:

    strip_warp(frames)
    for l = 0 .. num_bands:   exchange(GAUSS, l);   if l < num_bands: strip_down(l)
    for l = num_bands .. 0:   strip_band(l);        if l >= 1: exchange(RESTORED, l)

exchange(what, l) packs this rank's edge columns (2 per camera Gaussian level, 1 per restored band) for
each neighbour, posts all sends and receives of the step as ONE batch (ncclGroupStart/End underneath, so
the two directions cannot deadlock), and unpacks what arrived.  Message sizes follow from the calibration,
which every rank holds, so no sizes are negotiated.  Results are bit-identical to the unsplit panorama.
"""
import numpy as np

GAUSS, RESTORED = 0, 1
LEFT, RIGHT = 0, 1


def schedule(num_bands):
    """The per-frame stage list: ("warp",) | ("exchange", what, level) | ("down", level) | ("band", level)."""
    steps = [("warp",)]
    for l in range(num_bands + 1):
        steps.append(("exchange", GAUSS, l))
        if l < num_bands:
            steps.append(("down", l))
    for l in range(num_bands, -1, -1):
        steps.append(("band", l))
        if l >= 1:
            steps.append(("exchange", RESTORED, l))
    return steps


def neighbour(rank, world, side):
    n = rank - 1 if side == LEFT else rank + 1
    return n if 0 <= n < world else None


def run_stage(comp, step, frames=None):
    if step[0] == "warp":
        comp.strip_warp(frames)
    elif step[0] == "down":
        comp.strip_down(step[1])
    elif step[0] == "band":
        comp.strip_band(step[1])
    else:
        raise ValueError(step)


class TorchTransport:
    """Halo exchange over torch.distributed point-to-point ops (NCCL: device buffers, ordered with the
    compositor through the current CUDA stream; gloo: host buffers, for the CPU plumbing test)."""

    def __init__(self, comp, rank, world, device=None):
        import torch
        self.torch, self.comp, self.rank, self.world = torch, comp, rank, world
        self.device = device
        self.bufs = {}
        if device is not None and str(device) != "cpu":
            comp.set_stream(torch.cuda.current_stream(device).cuda_stream)     # kernels, copies and NCCL share one stream order
        for step in schedule(comp.num_bands):
            if step[0] != "exchange":
                continue
            for side in (LEFT, RIGHT):
                if neighbour(rank, world, side) is None:
                    continue
                ns, nr = comp.strip_halo_bytes(step[1], step[2], side)
                self.bufs[(step[1], step[2], side)] = (torch.empty(ns, dtype=torch.uint8, device=device or "cpu"),
                                                       torch.empty(nr, dtype=torch.uint8, device=device or "cpu"))

    def bytes_per_frame(self):
        return sum(s.numel() for s, _ in self.bufs.values()), sum(r.numel() for _, r in self.bufs.values())

    def exchange(self, what, level):
        import torch.distributed as dist
        ops = []
        for side in (LEFT, RIGHT):
            nbr = neighbour(self.rank, self.world, side)
            if nbr is None:
                continue
            sb, rb = self.bufs[(what, level, side)]
            if sb.numel():
                self.comp.strip_pack(what, level, side, sb.data_ptr())
                ops.append(dist.P2POp(dist.isend, sb, nbr))
            if rb.numel():
                ops.append(dist.P2POp(dist.irecv, rb, nbr))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for side in (LEFT, RIGHT):
            if neighbour(self.rank, self.world, side) is None:
                continue
            sb, rb = self.bufs[(what, level, side)]
            if rb.numel():
                self.comp.strip_unpack(what, level, side, rb.data_ptr())


def connect_peers_distributed(comp, rank, world):
    """Peer-memory halo exchange between processes: every rank exports its two receive areas as CUDA IPC handles, the handles
    travel once through torch.distributed (control plane), each rank maps its neighbours' areas.  After this no collective
    library call and no host synchronisation is on the frame path."""
    import torch.distributed as dist
    mine = {}
    for side in (LEFT, RIGHT):
        if neighbour(rank, world, side) is not None:
            mine[side] = comp.strip_peer_export(side)[0]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    for side in (LEFT, RIGHT):
        nbr = neighbour(rank, world, side)
        if nbr is not None:
            comp.strip_peer_connect(side, ipc_handle=everyone[nbr][RIGHT if side == LEFT else LEFT])


def connect_peers_local(comps):
    """The same inside one process (all handles on one device or on peer-accessible devices): plain pointers."""
    world = len(comps)
    areas = [{side: c.strip_peer_export(side)[1] for side in (LEFT, RIGHT) if neighbour(r, world, side) is not None} for r, c in enumerate(comps)]
    for r, c in enumerate(comps):
        for side in (LEFT, RIGHT):
            nbr = neighbour(r, world, side)
            if nbr is not None:
                c.strip_peer_connect(side, same_process_ptr=areas[nbr][RIGHT if side == LEFT else LEFT])


class StripCompositor:
    """One rank of the strip-mode panorama: a Compositor restricted to its columns + a transport.

    halo="recompute" (default): no communication at all — every rank recomputes the ~94 level-0 halo columns per side
    itself (SURVEY.md §8e "alternative with zero comms"); the lowest-latency choice when strips are wide compared with
    the halo.  halo="exchange": pyramid halo columns travel between neighbouring ranks (send/recv, 2 * num_bands + 1
    exchanges per frame; host-driven, slower than one GPU on B200).  halo="peer": see __init__."""

    def __init__(self, comp, rank, world, transport=None, device=None, halo=None):
        """halo="peer": the exchange mode with the halo columns written straight into the neighbours' HBM by this rank's
        kernels (sb_compositor_strip_frame_peer) - no collective call, no host in the loop, one C call per frame.
        Default: "recompute" (measured fastest: 0.17 ms against 0.41 ms for "peer" and 1.3-2.3 ms for host-driven NCCL
        exchange on 8 B200, one GPU 0.57 ms), "exchange" when a transport is handed in."""
        if halo is None:
            halo = "exchange" if transport is not None else "recompute"
        comp.set_strip(rank, world)
        self.comp, self.rank, self.world, self.halo = comp, rank, world, halo
        comp.set_strip_halo(halo == "recompute")
        if halo == "peer":
            self.transport = None
            if device is not None and str(device) != "cpu":
                import torch
                comp.set_stream(torch.cuda.current_stream(device).cuda_stream)
            connect_peers_distributed(comp, rank, world)
        elif halo == "recompute":
            self.transport = None
            if device is not None and str(device) != "cpu":
                import torch
                comp.set_stream(torch.cuda.current_stream(device).cuda_stream)
        else:
            self.transport = transport or TorchTransport(comp, rank, world, device)
        self.steps = schedule(comp.num_bands)

    def enqueue(self, frames):
        """All stages of one frame, asynchronous on the compositor's stream."""
        if self.halo == "recompute":
            self.comp.strip_compose(frames)
            return
        if self.halo == "peer":
            self.comp.strip_frame_peer(frames)
            return
        for step in self.steps:
            if step[0] == "exchange":
                self.transport.exchange(step[1], step[2])
            else:
                run_stage(self.comp, step, frames)

    def compose(self, frames):
        """-> (strip, strip_mask): this rank's columns of the panorama (host arrays)."""
        self.enqueue(frames)
        return self.comp.strip_result(self.rank, self.world)

    def gather(self, strip, mask=None, dst=0):
        """Assemble the panorama on rank `dst` from the per-rank strips (host arrays; control plane)."""
        import torch.distributed as dist
        parts = [None] * self.world if self.rank == dst else None
        dist.gather_object((strip, mask), parts, dst=dst)
        if self.rank != dst:
            return None, None
        pano = np.concatenate([p[0] for p in parts], axis=1)
        pmask = None if mask is None else np.concatenate([p[1] for p in parts], axis=1)
        return pano, pmask


def run_local_peer(comps, frames, connected=False):
    """Peer-memory exchange on one device: every handle runs its whole frame on its own stream; the ranks meet only through
    the flag words in each other's receive areas."""
    world = len(comps)
    if not connected:
        for r, c in enumerate(comps):
            c.set_strip(r, world)
            c.set_strip_halo(False)
        connect_peers_local(comps)
        # Inside ONE process the ranks share a device: a cudaMalloc by one handle (first-frame staging buffers) synchronises
        # the whole device and would wait for ever behind another handle's kernel that is waiting for its neighbour's flag.
        # So every handle stages a frame once, alone, before the ranks start to depend on each other.
        for r, c in enumerate(comps):
            c.strip_warp(frames)
            c.strip_result(r, world)
    # (one process, one device: the streams of the handles may share a hardware queue, so every rank's push of a step is
    # enqueued before any rank's pull of it; with one rank per process / GPU each rank simply calls strip_frame_peer)
    for step in schedule(comps[0].num_bands):
        if step[0] != "exchange":
            for c in comps:
                run_stage(c, step, frames)
        else:
            for c in comps:
                c.strip_peer_push(step[1], step[2])
            for c in comps:
                c.strip_peer_pull(step[1], step[2])
    return [c.strip_result(r, world) for r, c in enumerate(comps)]


def run_local_recompute(comps, frames):
    """Recompute-halo mode on one device: every handle composes its strip independently."""
    world = len(comps)
    out = []
    for r, c in enumerate(comps):
        c.set_strip(r, world)
        c.set_strip_halo(True)
        c.strip_compose(frames)
        out.append(c.strip_result(r, world))
    return out


def run_local(comps, frames):
    """Single-process simulation used by the 1-GPU parity test: `comps[r]` plays rank r (all on one device),
    the stages run in lock step and the halo messages are handed over through device buffers.
    -> list of (strip, strip_mask) per rank."""
    import torch
    world = len(comps)
    for r, c in enumerate(comps):
        c.set_strip(r, world)
        c.set_strip_halo(False)
    dev = torch.device("cuda", comps[0].device)
    for step in schedule(comps[0].num_bands):
        if step[0] != "exchange":
            for c in comps:
                run_stage(c, step, frames)
            continue
        _, what, level = step
        sent = {}
        for r, c in enumerate(comps):
            for side in (LEFT, RIGHT):
                if neighbour(r, world, side) is None:
                    continue
                ns, nr = c.strip_halo_bytes(what, level, side)
                buf = torch.empty(ns, dtype=torch.uint8, device=dev)
                if ns:
                    c.strip_pack(what, level, side, buf.data_ptr())
                sent[(r, side)] = (buf, nr)
        torch.cuda.synchronize(dev)            # the ranks run on different streams here
        for r, c in enumerate(comps):
            for side in (LEFT, RIGHT):
                nbr = neighbour(r, world, side)
                if nbr is None:
                    continue
                buf, _ = sent[(nbr, RIGHT if side == LEFT else LEFT)]      # what the neighbour sent towards me
                _, nr = sent[(r, side)]
                assert buf.numel() == nr, "halo size mismatch between ranks %d and %d: %d vs %d" % (r, nbr, buf.numel(), nr)
                if nr:
                    c.strip_unpack(what, level, side, buf.data_ptr())
        torch.cuda.synchronize(dev)
    return [c.strip_result(r, world) for r, c in enumerate(comps)]
