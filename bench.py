#!/usr/bin/env python
"""bench.py — 5x1080p -> panorama frames/s on N B200s (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the CPU arm: oracle/_ref or the oracle port)

A "step" is one pass of the hot path over one batch of synthetic frame sets (--batch frame sets of n cameras
each -> that many panoramas; value and e2e are frames/s, ms_per_step is per batch).
  value   : whole-job frames/s with the frames already resident in HBM (device pointers in, device
            panorama out), timed on the device between two marks that span every in-flight slot.
  e2e     : the same metric through the reference-facing C ABI with HOST buffers: each step copies
            its n source frames host->device from pinned memory and the panorama device->host.
  roofline: dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json.
  cpu_baseline: the CPU oracle (single core, "port") on a bounded sample, rank 0 at N=1 only.
Frames are independent once calibration is fixed: rank r processes its own frame sets, no
collective on the data path ("scaling": "weak").
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "5x1080p->panorama frames/s"
WORKLOADS = {
    "c2": "C2: 5-camera 1080p 360deg, CylindricalWarper + FeatherBlender(0.02), fixed calibration (BASELINE.json configs[1])",
    "c3": "C3: 5-camera 1080p 360deg, SphericalWarper + GainCompensator + MultiBandBlender(5 bands, CV_32F weights) (configs[2])",
    "c4": "C4: 8-camera 4K VR, SphericalWarper + GainCompensator + MultiBandBlender(5 bands) (configs[3])",
    "app6": "APP6: the live app's own per-frame case (BASELINE.md 1; APP64:748-759): 6-camera 1920x1088, cached-map cylindrical remap + BlockApply "
            "(block gain maps) + look-up composite without blending, cropped by the app's margins (0.1 / 0.1 / 10 / 10), panorama ~8020x883",
    "c5-strip": "C5 latency mode: ONE 8-camera 4K frame set -> 16384-wide panorama (SphericalWarper + GainCompensator + MultiBandBlender, 5 bands), the "
                "panorama cut into one column strip per GPU (BASELINE.json configs[4]; SURVEY 8e)",
}
# the only per-frame figure the reference publishes (REL32/resultTime-at.txt, mean 43.6 ms per frame set,
# hardware unknown): frames/s for exactly the APP6 workload
PUBLISHED_FPS = {"app6": 1000.0 / 43.6}


# SURVEY.md §8d "ALGORITHMIC bytes per frame" of the reference-shaped (fused-by-stage) dataflow, GB
SURVEY_DATAFLOW_GB = {"c2": 0.71, "c3": 1.30}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` ("workload:name", else "name") from the
    committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        e = t.get(kernel) or t[kernel.split(":", 1)[-1]]
        return e["dram_bytes_per_launch"], e["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        try:                                             # NVML directly (what nvidia-smi reads): no process spawn per sample
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = [("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop_flag:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                  "%.1f" % (pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)] + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.05)
            return
        except Exception:                                # noqa: BLE001 - fall back to the nvidia-smi command line of the recipe
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_frames(rig, n_sets, n_cams, rank):
    from stitchingvideo_b200 import rigs
    return [[rigs.frame(rig, rank * 1000 + s, i) for i in range(n_cams)] for s in range(n_sets)]


def bind_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned buffers) on the CPU cores of the NUMA node its GPU hangs off, so that
    the host<->device copies of N ranks do not all cross the socket interconnect.  Best effort; returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa_node unknown"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "node %d (%d cpus)" % (node, len(cpus))
        return "node %d has no usable cpus" % node
    except Exception as e:      # noqa: BLE001 - tuning only
        return "unavailable (%s)" % type(e).__name__


LAP = 32      # frame sets per recorded lap (one CUDA graph launch = one host call)


def measure(workload, args, rank, world, local, barrier, max_over_ranks, sampler=None):
    """All the numbers of one workload on this rank's GPU: value (device-resident), e2e (host buffers), roofline."""
    import torch
    import stitchingvideo_b200 as sv
    from stitchingvideo_b200 import capi, rigs

    Ks, Rs, spec = rigs.cameras(workload)
    n, size = spec["n_used"], (spec["W"], spec["H"])
    kw = {}
    if spec.get("block_gains"):          # the live app's BlockApply (APP64:310-331): block gain maps sized from the warped images
        probe = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], device=local)
        kw["gain_maps"] = rigs.block_gain_maps(workload, [probe.camera_roi(i)[2:] for i in range(n)])
        del probe
    if spec.get("crop"):
        kw["crop"], kw["crop_app_fill"] = spec["crop"], spec.get("crop_app_fill", False)
    comp = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], num_bands=5,
                         weight_type=sv.CV_32F, sharpness=0.02, gains=spec["gain_values"], output_type=sv.CV_8UC3, device=local, **kw)
    if args.variant is not None:
        comp.set_fused(10 + args.variant)      # kernel variant of the fused path (tuning hook)
    pw, ph = comp.pano_size
    depth = args.depth if args.depth is not None else (8 if workload in ("c3", "c4") else 4)
    n_sets = args.frame_sets
    host_sets = make_frames(workload, n_sets, n, rank)
    laps = max(1, args.batch // LAP)                        # graph launches per step
    frames_per_step = laps * LAP

    # ---------------- value: inputs resident in HBM, output stays on the device ----------------
    # One lap of the input ring (LAP frame sets over `depth` in-flight slots) is recorded once as a CUDA graph and
    # replayed: one host call per lap (sb_compositor_batch_*), the way a capture loop with a fixed buffer ring runs.
    dev_tensors = [[torch.from_numpy(f).cuda() for f in s] for s in host_sets]
    dev_sets = [[capi.DeviceImage.from_torch(t) for t in s] for s in dev_tensors]
    comp.set_depth(depth)
    # every frame set of the lap has its own device panorama (row pitch aligned for the kernel's vector stores): with device
    # sources AND device outputs a feather / no-blend lap is ONE persistent launch of the frame kernel (sb_batch_mode 1)
    wp = (pw + 7) & ~7
    out_t = [torch.empty((ph, wp, 3), dtype=torch.uint8, device="cuda") for _ in range(LAP)]
    dev_out = [capi.DeviceImage(t.data_ptr(), ph, pw, sv.CV_8UC3, wp * 3, local, owner=t) for t in out_t]
    lap = comp.batch([dev_sets[f % n_sets] for f in range(LAP)], dev_out)
    lap_mode = lap.mode
    for _ in range(args.warmup * laps):
        lap.launch()
    lap.wait()
    barrier()
    if sampler is not None:
        sampler.start()
    launches0 = sv.kernel_launch_count()
    comp.mark(0)
    for _ in range(args.steps * laps):
        lap.launch()
    comp.mark(1)
    ms_dev = comp.marked_ms()
    launches = sv.kernel_launch_count() - launches0
    barrier()
    ms_dev = max_over_ranks(ms_dev)
    value = world * args.steps * frames_per_step / (ms_dev / 1e3)
    lap.close()

    # ... and one frame set at a time on ONE stream (depth 1, a lap of single launches back to back, never overlapping):
    # frame time here = the sum of the frame's kernel durations as the launching stream sees them (inputs still rotate over
    # n_sets frame sets > L2), the denominator of the per-kernel roofline below.
    comp.set_depth(1)
    lap1 = comp.batch([dev_sets[f % n_sets] for f in range(LAP)], [None] * LAP)
    for _ in range(3):
        lap1.launch()
    lap1.wait()
    comp.mark(0)
    for _ in range(max(2, laps // 2)):
        lap1.launch()
    comp.mark(1)
    serial_frame_ms = comp.marked_ms() / (max(2, laps // 2) * LAP)
    lap1.close()
    del out_t, dev_out

    # ---------------- e2e: host buffers through the C ABI (H2D + kernels + D2H per frame set) ----------------
    # each frame set lives in one pinned block (n x H x W x 3), the way a capture pipeline delivers it; the library
    # recognises the contiguous set and moves it with a single DMA
    comp.set_depth(depth)
    pin_in = [torch.from_numpy(np.stack(s)).pin_memory() for s in host_sets]
    pin_sets = [[t[i].numpy() for i in range(n)] for t in pin_in]
    pin_out = [torch.empty((ph, pw, 3), dtype=torch.uint8).pin_memory() for _ in range(depth + 1)]
    e2e_steps, e2e_laps = args.steps, max(1, laps // args.e2e_divisor)     # (PCIe-bound: fewer laps per step keep the run short)
    elap = comp.batch([pin_sets[f % n_sets] for f in range(LAP)], [pin_out[f % (depth + 1)].numpy() for f in range(LAP)])
    for _ in range(min(args.warmup, 3) * e2e_laps):
        elap.launch()
    elap.wait()
    barrier()
    comp.mark(0)
    t0 = time.perf_counter()
    for _ in range(e2e_steps * e2e_laps):
        elap.launch()
    comp.mark(1)
    ms_e2e = comp.marked_ms()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))
    clocks = sampler.summary() if sampler is not None else None
    # ... and what this box can do with the copies ALONE (no kernels): the same pinned buffers, the same bytes per frame set, the
    # same number of streams, all ranks at once - the ceiling the end-to-end figure is to be read against (PCIe + host memory)
    copy_streams = [torch.cuda.Stream() for _ in range(depth)]
    dev_in = [torch.empty_like(pin_in[0], device="cuda") for _ in range(depth)]
    dev_o = [torch.empty((ph, pw, 3), dtype=torch.uint8, device="cuda") for _ in range(depth)]

    def copy_only(n_frames):
        for f in range(n_frames):
            k = f % depth
            with torch.cuda.stream(copy_streams[k]):
                dev_in[k].copy_(pin_in[f % n_sets], non_blocking=True)
                pin_out[f % (depth + 1)].copy_(dev_o[k], non_blocking=True)
    copy_only(2 * LAP)
    barrier()
    tc = time.perf_counter()
    copy_only(e2e_steps * e2e_laps * LAP)
    torch.cuda.synchronize()
    ms_copy = max_over_ranks((time.perf_counter() - tc) * 1e3)
    barrier()
    copy_value = world * e2e_steps * e2e_laps * LAP / (ms_copy / 1e3)
    del dev_in, dev_o, copy_streams
    e2e_frames_per_step = e2e_laps * LAP
    e2e_value = world * e2e_steps * e2e_frames_per_step / (ms_e2e / 1e3)
    h2d_frame, d2h_frame = n * size[0] * size[1] * 3, pw * ph * 3
    elap.close()

    # ---------------- roofline: per-kernel CUDA-event timing (separate pass, never the reported fps) ----------------
    comp.set_depth(1)
    if args.variant is None and depth > 1 and workload in ("c3", "c4"):
        comp.set_fused(12)      # profile the kernels the pipelined timed region launched (one launch per pyramid level)
    agg = {}
    for it in range(args.profile_frames + 1):
        recs = comp.profile_frame(dev_sets[it % n_sets])
        if it == 0:
            continue                      # warm-up
        for r in recs:
            a = agg.setdefault(r["name"], {"ms": 0.0, "bytes": 0.0, "launches": 0})
            a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["launches"] += 1
    peak, peak_src = peaks()
    total_ms = sum(a["ms"] for a in agg.values())
    kernels = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        share = a["ms"] / total_ms if total_ms else 0.0
        per_frame_launches = a["launches"] / args.profile_frames
        # isolated: events around ONE launch with nothing else in flight (carries ~5-10 us of event / launch latency);
        # in_stream: the kernel's share of the frame x the frame time of the back-to-back single-stream lap above
        iso_ms = a["ms"] / args.profile_frames
        ins_ms = share * serial_frame_ms
        mb = a["bytes"] / args.profile_frames / 1e6
        kernels.append({"name": name, "launches_per_frame": per_frame_launches, "share": share, "algorithmic_mb_per_frame": mb,
                        "isolated_ms_per_frame": iso_ms, "in_stream_ms_per_frame": ins_ms,
                        "achieved_gbs": mb / 1e3 / (ins_ms / 1e3) if ins_ms > 0 else 0.0,
                        "frac": (mb / 1e3 / (ins_ms / 1e3) / peak) if ins_ms > 0 else 0.0,
                        "frac_isolated": (mb / 1e3 / (iso_ms / 1e3) / peak) if iso_ms > 0 else 0.0})
    dom = kernels[0]
    traffic, traffic_src = ncu_traffic(workload + ":" + dom["name"])
    survey_gb = SURVEY_DATAFLOW_GB.get(workload)
    frame_ms = ms_dev / (args.steps * frames_per_step)
    total_mb = sum(k["algorithmic_mb_per_frame"] for k in kernels)
    roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "share_of_step": dom["share"],
                "how": "achieved = the kernel's algorithmic bytes per launch / its average launch duration when %d frame sets are launched one "
                       "kernel at a time, back to back on ONE stream (CUDA events around the laps; inputs rotate over %d frame sets = %.0f MB > L2); a frame of "
                       "several kernels is split by the per-kernel CUDA-event shares of the isolated pass" % (LAP, n_sets, n_sets * h2d_frame / 1e6),
                "avg_launch_us": dom["in_stream_ms_per_frame"] / dom["launches_per_frame"] * 1e3,
                "isolated_launch_us": dom["isolated_ms_per_frame"] / dom["launches_per_frame"] * 1e3,
                "frac_isolated": dom["frac_isolated"],
                "algorithmic_bytes_per_launch": dom["algorithmic_mb_per_frame"] * 1e6 / dom["launches_per_frame"],
                # SURVEY.md §8d's reference-shaped dataflow (accumulators in HBM) over the measured frame time, for
                # comparison with the survey's 60 % bar; the fused kernels move far fewer bytes than that
                "survey_dataflow": None if survey_gb is None else {
                    "gb_per_frame": survey_gb, "achieved_gbs": survey_gb / (frame_ms / 1e3), "frac": survey_gb / (frame_ms / 1e3) / peak},
                # The timed region itself (frame sets in flight on several streams, launches overlap): all algorithmic
                # bytes of the frame over the device time per frame.
                "timed_region": {"algorithmic_mb_per_frame": total_mb, "us_per_frame": frame_ms * 1e3,
                                 "achieved_gbs": total_mb / 1e3 / (frame_ms / 1e3), "frac": total_mb / 1e3 / (frame_ms / 1e3) / peak},
                "single_stream": {"us_per_frame": serial_frame_ms * 1e3, "achieved_gbs": total_mb / 1e3 / (serial_frame_ms / 1e3),
                                  "frac": total_mb / 1e3 / (serial_frame_ms / 1e3) / peak}}
    if lap_mode == 1 and len(kernels) == 1 and dom["launches_per_frame"] == 1:
        # The timed region launched this kernel as ONE persistent launch per lap of LAP frame sets: its average launch duration
        # over the timed region is (device time between the marks) / (number of launches), its algorithmic bytes per launch
        # are LAP x the frame's.  The single-frame launch figures of the separate pass stay beside it.
        n_launch = args.steps * laps
        launch_us = ms_dev * 1e3 / n_launch
        bytes_launch = dom["algorithmic_mb_per_frame"] * 1e6 * LAP
        roofline["single_frame_launch"] = {"avg_launch_us": roofline["avg_launch_us"], "achieved": roofline["achieved"], "frac": roofline["frac"],
                                           "isolated_launch_us": roofline["isolated_launch_us"], "frac_isolated": roofline["frac_isolated"],
                                           "how": roofline["how"]}
        roofline.update({"achieved": bytes_launch / (launch_us * 1e-6) / 1e9, "avg_launch_us": launch_us, "algorithmic_bytes_per_launch": bytes_launch,
                         "frame_sets_per_launch": LAP,
                         "how": "achieved = the kernel's algorithmic bytes per launch / its average launch duration over the timed region itself: %d "
                                "launches of the persistent frame kernel, %d frame sets each, CUDA events (marks) around the region on the launching "
                                "streams; single_frame_launch = the same kernel launched once per frame set in a separate pass" % (n_launch, LAP)})
        roofline["frac"] = roofline["achieved"] / peak
    out = {
        "value": value, "unit": "frames/s", "ms_per_step": ms_dev / args.steps,
        "config": {"workload": WORKLOADS[workload], "frame_sets_per_step": frames_per_step, "frame_sets_per_rank": args.steps * frames_per_step,
                   "cameras": n, "frame": "%dx%d" % size, "panorama": "%dx%d" % (pw, ph), "in_flight_slots": depth,
                   "host_calls": "one per lap of %d frame sets (sb_compositor_batch_*: %s)" % (
                       LAP, "one persistent launch of the frame kernel per lap" if lap_mode == 1 else "frame sets enqueued back to back on the slots' streams"),
                   "kernel_plan": comp.kernel_plan(),
                   "l2_policy": "inputs rotate over %d frame sets (%.0f MB) and every frame set streams its tables; the working set "
                                "exceeds the 126 MB L2" % (n_sets, n_sets * h2d_frame / 1e6)},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_frame * e2e_frames_per_step,
                "d2h_bytes_per_step": d2h_frame * e2e_frames_per_step, "frame_sets_per_step": e2e_frames_per_step,
                "ms_per_step": ms_e2e / e2e_steps, "host_buffers": "pinned, one block per frame set",
                "copy_only_ceiling": {"value": copy_value, "unit": "frames/s", "what": "the same host<->device copies with no kernel in between, all ranks at once",
                                      "frac": e2e_value / copy_value},
                "pcie_gbs": {"h2d": h2d_frame * e2e_frames_per_step / (ms_e2e / e2e_steps) / 1e6,
                             "d2h": d2h_frame * e2e_frames_per_step / (ms_e2e / e2e_steps) / 1e6}},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
    }
    del comp
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    head = measure(args.workload, args, rank, world, local, barrier, max_over_ranks, ClockSampler(local))
    # the other workloads of BASELINE.json in the same run (N=1 by default): multi-band C3 and the live app's own case
    extra = {}
    for w in args.also:
        m = measure(w, args, rank, world, local, barrier, max_over_ranks, ClockSampler(local))
        extra[w] = {"metric": METRIC if w != "app6" else "6x1088p->composite frame sets/s", "value": m["value"], "unit": m["unit"],
                    "ms_per_step": m["ms_per_step"], "vs_baseline": (m["value"] / PUBLISHED_FPS[w]) if w in PUBLISHED_FPS else None,
                    "config": m["config"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "clocks": m["clocks"],
                    "roofline": m["roofline"], "kernels": m["kernels"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    head["e2e"]["numa_binding_rank0"] = numa
    line = {
        "metric": METRIC, "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": (head["value"] / PUBLISHED_FPS[args.workload]) if args.workload in PUBLISHED_FPS else None,
        "dtype": "u8/s16 integer + f32 weights", "data": "synthetic", "config": head["config"], "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"], "kernels": head["kernels"],
    }
    if extra:
        line["workloads"] = extra
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, sample_frames=args.cpu_frames)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_strip(args):
    """--workload c5-strip: latency of ONE very wide panorama split into column strips, one rank (GPU) per strip.  A step is
    one frame set; value = 1 / (device latency of a frame, max over ranks) - total work is fixed as N grows ("strong").  The
    strips are compared bit for bit with the unsplit panorama before anything is timed."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import stitchingvideo_b200 as sv
    from stitchingvideo_b200 import capi, rigs, strips

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rig = "c5"
    Ks, Rs, spec = rigs.cameras(rig)
    n, size = spec["n_used"], (spec["W"], spec["H"])
    mk = lambda: sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender="multiband", num_bands=5,
                               gains=spec["gain_values"], device=local)
    comp = mk()
    sets = [[torch.from_numpy(rigs.frame(rig, s, i, smooth=0)).to(dev) for i in range(n)] for s in range(2)]
    dsets = [[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
    sampler = ClockSampler(local)
    launches0 = sv.kernel_launch_count()
    if world == 1:
        for it in range(args.warmup):
            comp.wait(comp.enqueue(dsets[it % 2], None, None))
        sampler.start()
        launches0 = sv.kernel_launch_count()
        lat = []
        for it in range(args.steps):
            comp.wait(comp.enqueue(dsets[it % 2], None, None))
            lat.append(comp.last_gpu_ms(0))
        ok, halo = True, "none (one strip)"
    else:
        sc = strips.StripCompositor(comp, rank, world, device=dev, halo=args.halo)
        strip, smask = sc.compose(dsets[0])
        pano, pmask = sc.gather(strip, smask, dst=0)
        ok = True
        if rank == 0:
            whole = mk()
            ref, rmask = whole.compose(dsets[0])
            ok = bool(np.array_equal(pano, ref) and np.array_equal(pmask, rmask))
            del whole
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = int(flag.item()) == 1
        for it in range(args.warmup):
            sc.enqueue(dsets[it % 2])
        torch.cuda.synchronize()
        dist.barrier()
        sampler.start()
        launches0 = sv.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lat = []
        for it in range(args.steps):
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            sc.enqueue(dsets[it % 2])
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lat.append(float(t.item()))
        halo = args.halo
    launches = sv.kernel_launch_count() - launches0
    clocks = sampler.summary()
    if rank == 0:
        ms = float(np.median(lat))
        pw, ph = comp.pano_size
        print(json.dumps({
            "metric": "8x4K->16384-wide panorama frames/s (one frame set in flight)", "value": 1e3 / ms, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8/s16 integer + f32 weights", "data": "synthetic",
            "config": {"workload": WORKLOADS["c5-strip"], "cameras": n, "frame": "%dx%d" % size, "panorama": "%dx%d" % (pw, ph), "strips": world,
                       "halo": halo, "timing": "CUDA events around one frame set on the stream the kernels (and the halo traffic) run on, "
                                               "median of the steps, max over ranks per step"},
            "bit_exact_vs_unsplit": ok, "latency_ms": {"median": ms, "min": float(np.min(lat)), "max": float(np.max(lat))},
            "gpu_launches": launches, "clocks": clocks, "e2e": None, "roofline": None}))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


# ------------------------------------------------------------------------------------------- CPU arms
def _oracle_ctx(workload):
    from stitchingvideo_b200 import rigs
    from oracle import pipeline as P
    Ks, Rs, spec = rigs.cameras(workload)
    size = (spec["W"], spec["H"])
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    return cal, spec


def _have_ref():
    """oracle/_ref: the reference's own blenders.cpp / warpers.cpp compiled against the OpenCV stand-in."""
    try:
        from oracle import ref as RF
        return RF.available()
    except Exception:
        return False


REF_NOTE = ("blending = the reference's own blenders.cpp (oracle/_ref, compiled where it lies on an OpenCV stand-in); "
            "remap / gain / pyramids = the stand-in's primitives (C restatement of OpenCV 2.4.11)")


def _oracle_frame(cal, spec, workload, idx):
    from stitchingvideo_b200 import rigs
    from oracle import pipeline as P
    frames = [rigs.frame(workload, idx, i, smooth=0) for i in range(spec["n_used"])]
    if spec.get("crop"):                 # the live app's loop: warp + BlockApply + feedSizeRemap (APP64:748-759)
        gm = rigs.block_gain_maps(workload, cal.sizes) if spec.get("block_gains") else None
        t = time.perf_counter()
        P.compose_app(cal, frames, gain_maps=gm, crop=spec["crop"], fill=spec.get("crop_app_fill", False))
        return time.perf_counter() - t
    t = time.perf_counter()
    P.compose(cal, frames, blender=spec["blender"], num_bands=5, gains=spec["gain_values"], use_ref=_have_ref())
    return time.perf_counter() - t


def cpu_baseline(workload, sample_frames=4):
    """The CPU oracle (single-threaded C restatement of the reference path) on a bounded sample."""
    cal, spec = _oracle_ctx(workload)
    _oracle_frame(cal, spec, workload, 0)
    ts = [_oracle_frame(cal, spec, workload, 1 + i) for i in range(sample_frames)]
    best = min(ts)
    out = {"value": 1.0 / best, "unit": "frames/s", "cores": 1, "kind": "reference" if _have_ref() else "port",
           "sample": "%d full frame sets of the same workload after 1 warm-up, best of %d (%.2f s each); "
                     "per-sequence maps/masks excluded like the GPU arm" % (sample_frames, sample_frames, best),
           "implementation": REF_NOTE if _have_ref() else "oracle port (C restatement of the reference path)",
           "host_cpu": _cpu_model(), "host_cores": os.cpu_count()}
    try:
        out["cv2_all_cores"] = cv2_baseline(workload)
    except Exception as e:      # cv2 is optional context, not the baseline
        out["cv2_all_cores"] = {"unavailable": str(e)[:100]}
    return out


def cv2_baseline(workload, frames=3):
    """Context only: OpenCV 4.13's own cv::detail blenders + cv::remap with all host cores
    (a stronger CPU baseline than the reference's 2.4.11 build)."""
    import cv2
    from stitchingvideo_b200 import rigs
    cal, spec = _oracle_ctx(workload)
    cv2.setNumThreads(os.cpu_count())
    best = 1e9
    for it in range(frames + 1):
        fr = [rigs.frame(workload, it, i, smooth=0) for i in range(spec["n_used"])]
        t = time.perf_counter()
        b = (cv2.detail_MultiBandBlender(0, 5, cv2.CV_32F) if spec["blender"] == "multiband" else
             cv2.detail_FeatherBlender(0.02) if spec["blender"] == "feather" else cv2.detail.Blender_createDefault(cv2.detail.Blender_NO))
        from oracle import oracle as O
        b.prepare(O.result_roi(cal.corners, cal.sizes))
        for i, f in enumerate(fr):
            w = cv2.remap(f, cal.maps[i][0], cal.maps[i][1], cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
            if spec["gain_values"]:
                w = cv2.convertScaleAbs(w, alpha=spec["gain_values"][i])
            b.feed(w.astype(np.int16), cal.masks[i], cal.corners[i])
        d, m = b.blend(None, None)
        cv2.convertScaleAbs(d)
        if it:
            best = min(best, time.perf_counter() - t)
    return {"value": 1.0 / best, "unit": "frames/s", "cores": os.cpu_count(), "impl": "cv2 %s" % cv2.__version__}


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("model name"):
                    return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


_W = {}


def _worker_init(workload):
    _W["ctx"] = _oracle_ctx(workload)
    _W["workload"] = workload


def _worker_frame(idx):
    cal, spec = _W["ctx"]
    return _oracle_frame(cal, spec, _W["workload"], idx)


def run_reference(args):
    """CPU arm: the reference path on all host cores.  One step = one batch of `cores` independent
    frame sets, one per worker process (frames are independent, so this is how the CPU path scales)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind = "reference" if _have_ref() else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(args.workload,)) as pool:
        for w in range(args.warmup):
            pool.map(_worker_frame, range(cores))
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_worker_frame, range(k * cores, (k + 1) * cores))
        dt = time.perf_counter() - t0
    value = args.steps * cores / dt
    _, spec = _oracle_ctx(args.workload)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/s16 integer + f32 weights", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "cameras": spec["n_used"], "frame": "%dx%d" % (spec["W"], spec["H"])},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "each step = %d full frame sets, one per worker process (%d processes); "
                                       "per-sequence maps/masks built once per worker and excluded" % (cores, cores),
                             "implementation": REF_NOTE if kind == "reference" else "oracle port (C restatement of the reference path)",
                             "host_cpu": _cpu_model()},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=960, help="frame sets per step (a multiple of the 32-frame lap; 20 steps x 960 = 19 200 frame sets, "
                    "~0.5 s of device time for the one-kernel feather workload)")
    ap.add_argument("--also", default=None, help="comma-separated extra workloads measured in the same run and reported under \"workloads\" "
                    "(default: c3,app6 at N=1, none for N>1)")
    ap.add_argument("--e2e-divisor", type=int, default=8, help="the PCIe-bound end-to-end leg runs batch / this many frame sets per step")
    ap.add_argument("--depth", type=int, default=None, help="frame sets in flight (slots); default 4 for the one-kernel feather workload, 8 for the "
                    "multi-band ones (lets the launch-latency-bound coarse pyramid levels of several frames overlap: C3 +9 %)")
    ap.add_argument("--frame-sets", type=int, default=8, help="distinct synthetic frame sets rotated through (8 x 31 MB > L2)")
    ap.add_argument("--profile-frames", type=int, default=5)
    ap.add_argument("--cpu-frames", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="recompute", choices=["recompute", "peer", "exchange"], help="c5-strip: how a strip gets the columns next to "
                    "its boundaries (recompute: no communication; peer: in-kernel peer-memory writes + flags; exchange: NCCL send/recv)")
    ap.add_argument("--variant", type=int, default=None, choices=[0, 1, 2, 3, 4, 5], help="fused kernel variant: 10 + v is passed to set_fused (default: library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.also is None:
        args.also = [w for w in ("c3", "app6") if w != args.workload] if (world == 1 and args.workload == "c2") else []
    else:
        args.also = [w for w in args.also.split(",") if w and w != args.workload]
    if args.workload == "c5-strip":
        if args.impl == "reference":      # (the CPU arm of the 16384-wide panorama takes minutes per frame: not run; C2 is the reference arm's workload)
            if int(os.environ.get("RANK", 0)) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "c5-strip has no CPU arm (one 8x4K frame set takes minutes on the host); run --workload c2"}))
            return
        run_strip(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
