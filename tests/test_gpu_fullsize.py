"""-m gpu: BASELINE.json's full-size configurations.

C2 / C3 (5 x 1080p): one frame set against the oracle, bit for bit (the oracle needs ~1-2 s per frame).
C4 (8 x 4K, multi-band): too slow for the oracle in a test; checked through size-independent properties of
the path instead: (i) the three GPU paths (fast fused, CV_16S fused, staged reference-shaped) agree bit for
bit; (ii) constant frames give a panorama that never exceeds the constant and whose median is within 2 of
it (bilinear remap and the integer pyramids preserve a constant exactly, so every Laplacian band is 0 and
only the top band carries v; the truncating casts of the weighted accumulate / normalise,
blenders.cpp:321-323,400-402, only ever lose counts — most where the coarse weights are small, the dark rim
OpenCV's multi-band blender is known for); (iii) the panorama mask equals the union of the warped masks;
(iv) frame sharding: shard g of G run alone reproduces the unsharded stream."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pipeline as P
from stitchingvideo_b200 import rigs, sharding

pytestmark = pytest.mark.gpu


def same(got, ref, what):
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    d = got != ref
    assert not d.any(), "%s: %d of %d values differ, first at %s" % (what, int(d.sum()), d.size, np.argwhere(d)[:3].tolist())


@pytest.mark.parametrize("rig", ["c2", "c3", "c1"])
def test_full_size_frame_matches_oracle(gpu, rig):
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], num_bands=5,
                          gains=spec["gain_values"])
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    assert comp.kernel_plan() == 2, "the BASELINE configs must run on the newest frame kernels, not on a fallback"
    for i in range(n):
        roi = comp.camera_roi(i)
        assert (roi[0], roi[1]) == cal.corners[i] and (roi[2], roi[3]) == cal.sizes[i]
    frames = [rigs.frame(rig, 1, i, smooth=1) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender=spec["blender"], num_bands=5, gains=spec["gain_values"])
    for fused in (11, 15, 12, 13, 14, 10):
        comp.set_fused(fused)
        pano, mask = comp.compose(frames)
        same(pano, ref, "%s panorama (variant %d)" % (rig, fused))
        same(mask, rmask, "%s mask (variant %d)" % (rig, fused))


def test_c4_frame_matches_oracle(gpu):
    """8 x 4K, spherical + gain + multi-band(5): one frame set against the oracle (about half a minute of CPU)."""
    Ks, Rs, spec = rigs.cameras("c4")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    comp = gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5,
                          gains=spec["gain_values"])
    cal = P.Calibration(size, Ks, Rs, "spherical", spec["scale"])
    frames = [rigs.frame("c4", 0, i, smooth=1) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender="multiband", num_bands=5, gains=spec["gain_values"])
    pano, mask = comp.compose(frames)
    same(pano, ref, "C4 panorama")
    same(mask, rmask, "C4 mask")


def test_c4_properties(gpu):
    Ks, Rs, spec = rigs.cameras("c4")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    comp = gpu.Compositor(size, Ks, Rs, warper="spherical", scale=spec["scale"], blender="multiband", num_bands=5, gains=None)
    pw, ph = comp.pano_size
    assert pw > 18000 and ph > 2000
    # (ii) constant frames -> constant panorama inside the mask
    frames = [np.full((size[1], size[0], 3), (37, 141, 250), np.uint8) for _ in range(n)]
    pano, mask = comp.compose(frames)
    inside = mask == 255
    assert inside.mean() > 0.9
    for ch, v in enumerate((37, 141, 250)):
        vals = pano[..., ch][inside]
        assert vals.max() <= v and np.median(vals) >= v - 2, "channel %d: max %d median %d for constant %d" % (ch, vals.max(), np.median(vals), v)
    assert not pano[~inside].any()                      # Blender::blend zeroes unmasked pixels
    # (iii) panorama mask = union of the warped masks
    w = gpu.SphericalWarper(spec["scale"])
    union = np.zeros((ph, pw), np.uint8)
    ones = np.full((size[1], size[0]), 255, np.uint8)
    x0 = min(comp.camera_roi(i)[0] for i in range(n))
    y0 = min(comp.camera_roi(i)[1] for i in range(n))
    for i in range(n):
        tl, m = w.warp(ones, Ks[i], Rs[i], O.INTER_NEAREST, O.BORDER_CONSTANT)
        union[tl[1] - y0:tl[1] - y0 + m.shape[0], tl[0] - x0:tl[0] - x0 + m.shape[1]] |= m
    same(mask, union, "C4 panorama mask")
    # (i) the three GPU paths agree on a textured frame set
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (size[1] // 8, size[0] // 8, 3), dtype=np.uint8)
    tex = [np.ascontiguousarray(np.roll(np.kron(base, np.ones((8, 8, 1), np.uint8)), 97 * i, axis=1)) for i in range(n)]
    outs = []
    for fused in (11, 15, 10, 0):
        comp.set_fused(fused)
        outs.append(comp.compose(tex)[0].copy())
    same(outs[1], outs[0], "C4 fused CV_16S path vs fast path")
    same(outs[2], outs[0], "C4 staged path vs fast path")


def test_sharded_stream_equals_unsharded(gpu):
    """Frame sharding: shard g of G, run alone, reproduces its part of the unsharded stream."""
    Ks, Rs, spec = rigs.cameras("mini_cyl")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    comp = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")
    n_frames, G = 6, 2
    whole = [sharding.checksum(comp.compose([rigs.frame("mini_cyl", f, i) for i in range(n)])[0]) for f in range(n_frames)]
    per_rank = []
    for g in range(G):
        shard = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="feather")   # one handle per GPU
        per_rank.append([sharding.checksum(shard.compose([rigs.frame("mini_cyl", f, i) for i in range(n)])[0])
                         for f in sharding.frames_for_rank(n_frames, g, G)])
    assert sharding.interleave(per_rank) == whole


def test_full_size_no_blend_matches_oracle(gpu):
    """The live app's shape (APP64:724-770): cylindrical warp + gain + Blender::NO composite of 5 x 1080p, one fused launch."""
    Ks, Rs, spec = rigs.cameras("c2")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    gains = [0.95, 1.02, 1.0, 0.98, 1.05]
    comp = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="no", gains=gains)
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    assert comp.kernel_plan() == 2
    frames = [rigs.frame("c2", 4, i, smooth=1) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender="no", gains=gains)
    for fused in (11, 15, 10, 0):
        comp.set_fused(fused)
        pano, mask = comp.compose(frames)
        same(pano, ref, "no-blend panorama (variant %d)" % fused)
        same(mask, rmask, "no-blend mask (variant %d)" % fused)


def test_app6_as_the_app_runs_it(gpu):
    """BASELINE.md §1's case at full size: 6 x 1920x1088, cylindrical cached-map remap + BlockApply + the look-up composite
    cropped by the app's margins with its unconditional gather (APP64:47, 150-177, 310-331, 748-759) - one launch, bit-exact."""
    Ks, Rs, spec = rigs.cameras("app6")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, "cylindrical", spec["scale"])
    gm = rigs.block_gain_maps("app6", cal.sizes)
    comp = gpu.Compositor(size, Ks, Rs, warper="cylindrical", scale=spec["scale"], blender="no", gain_maps=gm, crop=spec["crop"],
                          crop_app_fill=True)
    assert comp.kernel_plan() == 2
    frames = [rigs.frame("app6", 2, i, smooth=1) for i in range(n)]
    ref, rmask = P.compose_app(cal, frames, gain_maps=gm, crop=spec["crop"], fill=True)
    before = gpu.kernel_launch_count()
    pano, mask = comp.compose(frames)
    assert gpu.kernel_launch_count() - before == 1
    same(pano, ref, "app6 composite")
    w, h = comp.pano_size
    from oracle import oracle as O
    roi = O.result_roi(cal.corners, cal.sizes)
    ow, oh, xx, yy = P.crop_geometry(roi[2], roi[3], *spec["crop"])
    assert (w, h) == (ow, oh)
    same(mask, rmask[yy:yy + oh, xx:xx + ow], "app6 mask")


def test_c5_strips_match_the_oracle(gpu):
    """BASELINE.json configs[4] against the ORACLE (round 1 only compared strips with the unsplit GPU panorama): the 16K-wide
    panorama of 8 x 4K cameras composed as 8 column strips - both halo policies - equals the CPU oracle's panorama bit for bit."""
    from stitchingvideo_b200 import strips
    Ks, Rs, spec = rigs.cameras("c5")
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    cal = P.Calibration(size, Ks, Rs, spec["warper"], spec["scale"])
    frames = [rigs.frame("c5", 0, i, smooth=1) for i in range(n)]
    ref, rmask = P.compose(cal, frames, blender="multiband", num_bands=5, gains=spec["gain_values"])
    assert ref.shape[1] > 16000
    mk = lambda: gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender="multiband", num_bands=5, gains=spec["gain_values"])
    whole = mk()
    pano, mask = whole.compose(frames)
    same(pano, ref, "c5 unsplit panorama")
    same(mask, rmask, "c5 unsplit mask")
    del whole
    comps = [mk() for _ in range(8)]
    for run in (strips.run_local_recompute, strips.run_local):
        parts = run(comps, frames)
        same(np.concatenate([p[0] for p in parts], axis=1), ref, "c5 strips %s" % run.__name__)
        same(np.concatenate([p[1] for p in parts], axis=1), rmask, "c5 strip masks %s" % run.__name__)
