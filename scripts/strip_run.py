"""Latency ("strip") mode across the GPUs of one box: one rank per GPU, NCCL send/recv halo exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        scripts/strip_run.py --rig c5 --frames 20

Every rank builds the (replicated) calibration, produces its column strip of each panorama and the strips
are compared, bit for bit, with the unsplit panorama composed on rank 0's GPU.  Latency = device time of one
frame (CUDA events on the stream shared by the kernels and NCCL), max over ranks; printed as one JSON line
together with the single-GPU latency of the same panorama.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import stitchingvideo_b200 as sv  # noqa: E402
from stitchingvideo_b200 import capi, rigs, strips  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rig", default="c5")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--halo", default="exchange", choices=["exchange", "recompute", "peer"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Ks, Rs, spec = rigs.cameras(args.rig)
    n, size = spec["n_used"], (spec["W"], spec["H"])
    mk = lambda: sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender="multiband", num_bands=5,
                               gains=spec["gain_values"], device=local)
    comp = mk()
    sets = [[torch.from_numpy(rigs.frame(args.rig, s, i, smooth=0)).to(dev) for i in range(n)] for s in range(2)]
    dsets = [[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
    sc = strips.StripCompositor(comp, rank, world, device=dev, halo=args.halo)
    sent, recvd = sc.transport.bytes_per_frame() if sc.transport else (0, 0)

    # ---- parity: strips vs the unsplit panorama (rank 0 composes it on its own GPU)
    strip, smask = sc.compose(dsets[0])
    pano, pmask = sc.gather(strip, smask, dst=0)
    ok = True
    single_ms = None
    if rank == 0:
        whole = mk()
        ref, rmask = whole.compose(dsets[0])
        ok = bool(np.array_equal(pano, ref) and np.array_equal(pmask, rmask))
        for it in range(args.warmup):
            whole.wait(whole.enqueue(dsets[it % 2], None, None))
        ts = []
        for it in range(args.frames):
            whole.wait(whole.enqueue(dsets[it % 2], None, None))
            ts.append(whole.last_gpu_ms(0))
        single_ms = float(np.median(ts))
        del whole
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)

    # ---- latency: device time per frame, max over ranks
    for it in range(args.warmup):
        sc.enqueue(dsets[it % 2])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []
    for it in range(args.frames):
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        sc.enqueue(dsets[it % 2])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lat.append(float(t.item()))
    if rank == 0:
        pw, ph = comp.pano_size
        print(json.dumps({"mode": "strip", "halo": args.halo, "rig": args.rig, "n_gpus": world, "panorama": "%dx%d" % (pw, ph), "bit_exact_vs_unsplit": ok,
                          "strip_latency_ms": {"median": float(np.median(lat)), "min": float(np.min(lat))},
                          "single_gpu_latency_ms": single_ms, "halo_bytes_per_frame_rank0": {"sent": sent, "received": recvd},
                          "exchanges_per_frame": sum(1 for s in sc.steps if s[0] == "exchange") if args.halo in ("exchange", "peer") else 0, "frames": args.frames}))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
