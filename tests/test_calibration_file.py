"""Calibration-table serialization (include/stitchb200.h sb_calibration_*).
CPU part: the file round-trips every field bit for bit, is written atomically and detects corruption (host-only code
of the C-ABI library, no device needed).  GPU part: a compositor resumed from the file composes the same panorama."""
import ctypes as C
import os

import numpy as np
import pytest

from stitchingvideo_b200 import capi, rigs


def _config(n=3, with_masks=True, with_maps=False, seed=0):
    rng = np.random.default_rng(seed)
    K = rng.normal(size=(n, 3, 3)).astype(np.float32)
    R = rng.normal(size=(n, 3, 3)).astype(np.float32)
    cfg = capi.SbCompositorConfig()
    cfg.n_cameras, cfg.src_size = n, capi.SbSize(320, 200)
    cfg.warper_kind, cfg.warper_scale = capi.WARP_SPHERICAL, 123.456
    cfg.K, cfg.R = K.ctypes.data_as(C.POINTER(C.c_float)), R.ctypes.data_as(C.POINTER(C.c_float))
    cfg.blender_kind, cfg.num_bands, cfg.weight_type, cfg.sharpness = capi.BLEND_MULTI_BAND, 4, capi.CV_16S, 0.05
    cfg.output_type = capi.CV_16SC3
    keep = [K, R]
    if with_maps:
        maps = [rng.uniform(0.8, 1.2, (4 + i, 7)).astype(np.float32) for i in range(n)]
        arr = (capi.SbImage * n)(*[capi._image(m)[0] for m in maps])
        cfg.comp_kind, cfg.gain_maps = capi.COMP_GAIN_BLOCKS, arr
        keep += [maps, arr]
    else:
        g = rng.uniform(0.9, 1.1, n)
        cfg.comp_kind, cfg.gains = capi.COMP_GAIN, g.ctypes.data_as(C.POINTER(C.c_double))
        keep += [g]
    if with_masks:
        masks = [np.ascontiguousarray((rng.random((50 + i, 80 - i)) > 0.3).astype(np.uint8) * 255) for i in range(n)]
        marr = (capi.SbImage * n)(*[capi._image(m)[0] for m in masks])
        cfg.seam_masks = marr
        keep += [masks, marr]
    return cfg, keep


@pytest.mark.parametrize("with_masks,with_maps", [(True, False), (False, True), (True, True), (False, False)])
def test_calibration_file_round_trip(tmp_path, with_masks, with_maps):
    cfg, keep = _config(3, with_masks, with_maps, seed=7)
    path = str(tmp_path / "rig.sbcal")
    capi.save_calibration(cfg, path)
    assert not os.path.exists(path + ".tmp")
    d = capi.load_calibration(path)
    assert d["n_cameras"] == 3 and d["src_size"] == (320, 200) and d["warper_kind"] == capi.WARP_SPHERICAL
    assert np.float32(d["warper_scale"]) == np.float32(123.456) and np.float32(d["sharpness"]) == np.float32(0.05)
    assert (d["blender_kind"], d["num_bands"], d["weight_type"], d["output_type"]) == (capi.BLEND_MULTI_BAND, 4, capi.CV_16S, capi.CV_16SC3)
    assert np.array_equal(d["K"].view(np.uint32), keep[0].view(np.uint32)) and np.array_equal(d["R"].view(np.uint32), keep[1].view(np.uint32))
    if with_maps:
        assert d["comp_kind"] == capi.COMP_GAIN_BLOCKS and d["gains"] is None
        assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(d["gain_maps"], keep[2]))
    else:
        assert d["comp_kind"] == capi.COMP_GAIN and np.array_equal(d["gains"], keep[2]) and d["gain_maps"] is None
    if with_masks:
        assert all(np.array_equal(a, b) for a, b in zip(d["seam_masks"], keep[-2]))
    else:
        assert d["seam_masks"] is None
    # same content -> same bytes (the file is a pure function of the configuration)
    capi.save_calibration(cfg, path + "2")
    assert open(path, "rb").read() == open(path + "2", "rb").read()


def test_calibration_file_rejects_corruption(tmp_path):
    cfg, keep = _config(2, True, False, seed=1)
    path = str(tmp_path / "rig.sbcal")
    capi.save_calibration(cfg, path)
    raw = bytearray(open(path, "rb").read())
    for mutate in (lambda b: b[:len(b) // 2], lambda b: b[:60] + bytes([b[60] ^ 1]) + b[61:], lambda b: b"NOTACAL!" + b[8:], lambda b: b""):
        bad = str(tmp_path / "bad.sbcal")
        open(bad, "wb").write(bytes(mutate(bytes(raw))))
        with pytest.raises(capi.StitchError) as e:
            capi.load_calibration(bad)
        assert e.value.code == capi.SB_ERR_BAD_ARG
    with pytest.raises(capi.StitchError):
        capi.load_calibration(str(tmp_path / "missing.sbcal"))
    cfg.n_cameras = 0
    with pytest.raises(capi.StitchError):
        capi.save_calibration(cfg, path)


@pytest.mark.gpu
@pytest.mark.parametrize("rig", ["mini", "mini_cyl"])
def test_compositor_resumes_from_calibration_file(gpu, tmp_path, rig):
    Ks, Rs, spec = rigs.cameras(rig)
    size, n = (spec["W"], spec["H"]), spec["n_used"]
    comp = gpu.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], num_bands=5,
                          gains=spec["gain_values"])
    frames = [rigs.frame(rig, 0, i) for i in range(n)]
    pano, mask = comp.compose(frames)
    path = str(tmp_path / "rig.sbcal")
    comp.save_calibration(path)
    del comp
    again = gpu.Compositor.from_calibration(path)
    assert again.n == n and again.src_size == size
    pano2, mask2 = again.compose(frames)
    assert np.array_equal(pano, pano2) and np.array_equal(mask, mask2)
    again.save_calibration(path + "2")                               # and the resumed one saves the same file
    assert open(path, "rb").read() == open(path + "2", "rb").read()


@pytest.mark.parametrize("n,deg", [(5, 2), (40, 3), (512, 4), (513, 4), (3000, 6)])
def test_gain_solve_matches_the_reference_dense_solve(n, deg):
    """sb_gain_solve (host-only): the reference's normal equations + dense LU up to 512 unknowns, sparse conjugate gradients
    beyond (the block compensator's sizes) — against the oracle's dense cv::solve restatement on a random overlap graph."""
    from oracle import oracle as O
    rng = np.random.default_rng(n)
    pairs = {(i, i) for i in range(n)}
    for i in range(n):
        for j in rng.integers(0, n, deg):
            if i != j:
                pairs.add((min(i, int(j)), max(i, int(j))))
    pi, pj = np.array(sorted(pairs), np.int32).T
    cnt = rng.integers(0, 1024, len(pi))
    cnt[rng.random(len(pi)) < 0.1] = 0                                   # empty intersections count as 1 (max(1, countNonZero))
    Nk = np.maximum(cnt, 1).astype(np.float64)
    Iij, Iji = rng.uniform(20, 300, len(pi)), rng.uniform(20, 300, len(pi))
    N, I = np.zeros((n, n), np.int32), np.zeros((n, n), np.float64)
    N[pi, pj] = Nk; N[pj, pi] = Nk
    I[pi, pj] = Iij; I[pj, pi] = Iji
    ref = O.gain_solve(N, I)
    got = capi.gain_solve(n, pi, pj, Nk, Iij, Iji)
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=0)
