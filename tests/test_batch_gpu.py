"""sb_compositor_batch_*: a lap of frame sets recorded as a CUDA graph gives the panoramas enqueue/wait gives (and the
oracle gives), replay after replay, for host and device buffers."""
import numpy as np
import pytest

from oracle import pipeline as P
from stitchingvideo_b200 import rigs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rig,blender", [("mini_cyl", "feather"), ("mini", "multiband"), ("mini_cyl", "no")])
def test_batch_equals_per_frame_calls(gpu, rig, blender):
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    gains = ([0.95, 1.02, 1.0, 0.98, 1.05] * 2)[:n]
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=gains)
    comp.set_depth(3)
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    sets = [[rigs.frame(rig, f, i) for i in range(n)] for f in range(4)]
    refs = [P.compose(cal, s, blender=blender, gains=gains) for s in sets]
    lap = [sets[f % 4] for f in range(7)]                       # 7 frame sets over 3 slots: slots are reused inside the lap
    w, h = comp.pano_size
    panos = [np.zeros((h, w, 3), np.uint8) for _ in lap]
    masks = [np.zeros((h, w), np.uint8) for _ in lap]
    b = comp.batch(lap, panos, masks)
    assert b.n_frames == 7
    for rep in range(3):
        for p in panos:
            p[:] = 0
        b.launch()
        b.wait()
        for f in range(7):
            assert np.array_equal(panos[f], refs[f % 4][0]), "replay %d frame %d" % (rep, f)
            assert np.array_equal(masks[f], refs[f % 4][1])
    assert b.last_gpu_ms() > 0
    b.close()
    # the handle still serves per-frame calls afterwards
    pano, mask = comp.compose(sets[1])
    assert np.array_equal(pano, refs[1][0]) and np.array_equal(mask, refs[1][1])


@pytest.mark.parametrize("rig,blender,out16", [("mini_cyl", "feather", False), ("mini_cyl", "no", False), ("mini_cyl_n9", "feather", True),
                                               ("mini_sph_n7", "feather", False)])
def test_persistent_lap_on_device_buffers(gpu, rig, blender, out16):
    """Device sources + device panoramas: the whole lap is ONE launch of the frame kernel (sb_batch_mode 1), bit-exact per
    frame set against the oracle, replay after replay."""
    import torch
    from stitchingvideo_b200 import capi
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    gains = ([0.95, 1.02, 1.0, 0.98, 1.05] * 2)[:n]
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=gains,
                          output_type=gpu.CV_16SC3 if out16 else gpu.CV_8UC3)
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    sets = [[rigs.frame(rig, f, i) for i in range(n)] for f in range(3)]
    refs = [P.compose(cal, s, blender=blender, gains=gains, output_8u=not out16) for s in sets]
    dev = [[capi.DeviceImage.from_torch(torch.from_numpy(a).cuda()) for a in s] for s in sets]
    w, h = comp.pano_size
    wp = (w + 7) & ~7                                                     # row pitch aligned for the kernel's vector stores
    F = 5
    pt = [torch.zeros((h, wp, 3), dtype=torch.int16 if out16 else torch.uint8, device="cuda") for _ in range(F)]
    mt = [torch.zeros((h, wp), dtype=torch.uint8, device="cuda") for _ in range(F)]
    esz = 6 if out16 else 3
    panos = [capi.DeviceImage(t.data_ptr(), h, w, gpu.CV_16SC3 if out16 else gpu.CV_8UC3, wp * esz, 0, owner=t) for t in pt]
    masks = [capi.DeviceImage(t.data_ptr(), h, w, gpu.CV_8UC1, wp, 0, owner=t) for t in mt]
    b = comp.batch([dev[f % 3] for f in range(F)], panos, masks)
    assert comp.kernel_plan() == 2 and b.mode == 1, "expected the persistent multi-frame launch"
    for rep in range(2):
        for t in pt:
            t.zero_()
        before = gpu.kernel_launch_count()
        b.launch()
        b.wait()
        assert gpu.kernel_launch_count() - before == 1
        for f in range(F):
            assert np.array_equal(pt[f][:, :w].cpu().numpy(), refs[f % 3][0]), "replay %d frame %d" % (rep, f)
            assert np.array_equal(mt[f][:, :w].cpu().numpy(), refs[f % 3][1])
    b.close()


def test_batch_counts_its_kernel_launches(gpu):
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    n = spec["n_used"]
    comp = gpu.Compositor((spec["W"], spec["H"]), Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")
    comp.set_depth(2)
    frames = [rigs.frame("mini_cyl", 0, i) for i in range(n)]
    b = comp.batch([frames] * 4, [None] * 4)
    before = gpu.kernel_launch_count()
    b.launch(); b.launch()
    b.wait()
    assert gpu.kernel_launch_count() - before == 2 * 4          # one frame kernel per frame set and replay


def test_output_views_must_be_writable_in_place(gpu):
    """an output array with non-contiguous pixels used to be replaced by a temporary and never filled."""
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    n = spec["n_used"]
    comp = gpu.Compositor((spec["W"], spec["H"]), Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")
    frames = [rigs.frame("mini_cyl", 0, i) for i in range(n)]
    w, h = comp.pano_size
    wide = np.zeros((h, w, 4), np.uint8)
    with pytest.raises(gpu.StitchError):
        comp.compose(frames, wide[:, :, :3])                     # pixel stride 4, 3 channels: cannot be written in place
    good = np.zeros((h, 2 * w, 3), np.uint8)[:, :w]              # a column view with contiguous pixels is fine
    pano, _ = comp.compose(frames)
    comp.compose(frames, good)
    assert np.array_equal(good, pano)


def test_multi_device_driver(gpu):
    """sb_multi_*: one process, one host thread per device, frame f on devices[f % n] (here two handles on device 0 - a box
    with several GPUs passes their indices): the panoramas come back in frame order and equal the oracle's."""
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    m = gpu.MultiCompositor([0, 0], 2, size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")
    sets = [[rigs.frame("mini_cyl", f, i) for i in range(n)] for f in range(7)]
    out = m.run(sets)
    for f, (pano, mask) in enumerate(out):
        ref, rmask = P.compose(cal, sets[f], blender="feather")
        assert np.array_equal(pano, ref) and np.array_equal(mask, rmask), "frame %d" % f
    with pytest.raises(gpu.StitchError):
        gpu.MultiCompositor([0, 99], 1, size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")     # no such device: an error, not a downgrade


@pytest.mark.parametrize("rig,blender,variants", [("mini", "multiband", (0, 10, 11, 12, 13, 14, 16, 17, 18, 19)), ("mini_cyl", "feather", (0, 10, 11, 15)),
                                                  ("mini_cyl", "no", (0, 11, 15))])
@pytest.mark.parametrize("out16", [False, True])
def test_device_panorama_is_written_in_place(gpu, rig, blender, variants, out16):
    """enqueue() with a panorama that already lives on the device: the frame's last kernel writes it directly (aligned pitch)
    or it is copied (odd pitch) - the same pixels as the oracle either way, for every kernel variant, and nothing outside the
    panorama's columns is touched."""
    import torch
    from stitchingvideo_b200 import capi
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    gains = ([0.95, 1.02, 1.0, 0.98, 1.05] * 2)[:n]
    otype = gpu.CV_16SC3 if out16 else gpu.CV_8UC3
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=blender, gains=gains, output_type=otype)
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    frames = [rigs.frame(rig, 1, i) for i in range(n)]
    ref, ref_mask = P.compose(cal, frames, blender=blender, gains=gains, output_8u=not out16)
    w, h = comp.pano_size
    esz = 6 if out16 else 3
    for v in variants:
        comp.set_fused(v)
        for wp in ((w + 7) & ~7, w + 3):                      # in place / through the copy
            t = torch.full((h, wp, 3), 77, dtype=torch.int16 if out16 else torch.uint8, device="cuda")
            m = torch.full((h, wp), 77, dtype=torch.uint8, device="cuda")
            pano = capi.DeviceImage(t.data_ptr(), h, w, otype, wp * esz, 0, owner=t)
            mask = capi.DeviceImage(m.data_ptr(), h, w, gpu.CV_8UC1, wp, 0, owner=m)
            comp.wait(comp.enqueue(frames, pano, mask))
            assert np.array_equal(t[:, :w].cpu().numpy(), ref), "variant %d pitch %d" % (v, wp)
            assert np.array_equal(m[:, :w].cpu().numpy(), ref_mask), "variant %d pitch %d (mask)" % (v, wp)
            assert (t[:, w:] == 77).all() and (m[:, w:] == 77).all(), "variant %d pitch %d: wrote outside the panorama" % (v, wp)


def test_graph_mode_lap_still_matches(gpu):
    """SB_BATCH_GRAPH=1 (opt-in: a lap recorded as a CUDA graph; kernels launched while a stream is being captured fall back
    from programmatic dependent launch to plain kernel nodes) gives the same panoramas - run in a fresh process, the switch is
    read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import rigs
from oracle import pipeline as P
rig = "mini_cyl"
Ks, Rs, spec = rigs.cameras(rig)
size, n = (spec["W"], spec["H"]), spec["n_used"]
comp = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender="feather")
comp.set_depth(2)
cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
sets = [[rigs.frame(rig, f, i) for i in range(n)] for f in range(3)]
refs = [P.compose(cal, s, blender="feather") for s in sets]
w, h = comp.pano_size
panos = [np.zeros((h, w, 3), np.uint8) for _ in range(5)]
b = comp.batch([sets[f % 3] for f in range(5)], panos)
assert b.mode == 2, b.mode
for rep in range(2):
    b.launch(); b.wait()
    for f in range(5):
        assert np.array_equal(panos[f], refs[f % 3][0]), (rep, f)
print("graph ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SB_BATCH_GRAPH="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert r.returncode == 0 and "graph ok" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]
