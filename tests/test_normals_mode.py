"""A2c -- the RGB + surface-normals patch mode (surface_normals.cu:11-73, patch_extractor.cu:12-111, HFTest.cpp:322-363 and
:443-470; kept commented out in the reference's test path, live in its training-patch generator).

CPU: the C oracle against the independent numpy restatement, bit for bit.  GPU: the CUDA kernels (normals map, 6-channel
gather + quantisation, 384-input encoder, then the unchanged forest / vote / pose stages) against the oracle.
"""
import os

import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import oracle as O
from tests import npref


def _frame(seed=3, small=True):
    cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5) if small else synth.Camera()
    bgr, depth = synth.render_frame(seed, cam, n_objects=4)
    p = O.default_params(W=cam.W, H=cam.H, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, patch_mode=1,
                         normals_focal=287.5 if small else 575.0, fill_seed=77)
    return cam, bgr, depth, p


def test_normals_map_oracle_equals_numpy():
    cam, bgr, depth, p = _frame()
    n_c = O.normals(depth, p.normals_focal)
    n_np = npref.surface_normals(depth, p.normals_focal)
    assert n_c.shape == (cam.H, cam.W, 3)
    assert np.array_equal(n_c.view(np.uint32), n_np.view(np.uint32))      # bit-exact, NaNs included
    valid = np.abs(np.linalg.norm(n_c, axis=2) - 1) < 1e-5
    assert valid.mean() > 0.3 and not n_c[0].any() and not n_c[:, 0].any()  # unit normals on surfaces, zero border
    # a fronto-parallel plane: a = (left - right)/2 runs along -x, b = (down - up)/2 along +y, n = -(a x b) = (0, 0, +1),
    # the half space the reference draws its random fill normals from (z_rand >= 0, patch_extractor.cu:30-31)
    flat = np.full((40, 50), 800, np.uint16)
    nf = O.normals(flat, 575.0)
    assert np.allclose(nf[5:-5, 5:-5], [0, 0, 1], atol=1e-6)


@pytest.mark.parametrize("fill_random", [0, 1])
def test_gather_and_quantise_oracle_equals_numpy(fill_random):
    cam, bgr, depth, p = _frame()
    p.fill_random = fill_random
    locs = O.scan_centres(depth, p)
    assert len(locs) > 3000
    sel = np.linspace(0, len(locs) - 1, 1500).astype(np.int64)
    nrm = O.normals(depth, p.normals_focal)
    pc = O.gather_normals(bgr, depth, nrm, p, locs[sel])
    if fill_random:   # the fill is keyed on the patch index, so restate on the same indices
        pn = np.zeros_like(pc)
        full = npref.gather_normals(bgr, depth, nrm, locs[:sel.max() + 1], cam.W, cam.H, 8, p.voxel_m, p.normals_focal,
                                    fill_random=1, fill_seed=p.fill_seed)
        pc = O.gather_normals(bgr, depth, nrm, p, locs[:sel.max() + 1])
        pn = full
    else:
        pn = npref.gather_normals(bgr, depth, nrm, locs[sel], cam.W, cam.H, 8, p.voxel_m, p.normals_focal)
    assert np.array_equal(pc.view(np.uint32), pn.view(np.uint32))
    assert (pc[..., 3:] != 0).any() and (np.abs(np.linalg.norm(pc[..., 3:], axis=-1) - 1) < 1e-4).mean() > 0.5
    qc, qn = O.quantise_normals(pc), npref.quantise_normals(pn)
    assert qc.shape[1] == 384 and np.array_equal(qc, qn)
    # CHW layout and the two quantisation rules on hand-made values
    one = np.zeros((1, 8, 8, 6), np.float32)
    one[0, 2, 5] = [0.5, 1.0, 0.25, -1.0, 0.0, 1.0]
    q1 = O.quantise_normals(one).reshape(6, 8, 8)
    assert list(q1[:, 2, 5]) == [127, 255, 63, 0, 127, 255]


def test_scan_uses_the_normals_focal_length():
    """HFTest.cpp:356 hands the literal 575.0f to the 7-channel extractor, :394 hands fx to the RGB-D one."""
    cam, bgr, depth, p = _frame()
    p.normals_focal = 400.0
    a = O.scan_centres(depth, p)
    p.patch_mode = 0
    b = O.scan_centres(depth, p)
    ref = npref.scan_centres(depth, cam.W, cam.H, 2, 8, p.voxel_m, 400.0, p.distance_threshold_m)
    assert np.array_equal(a, ref) and len(a) != len(b)


@pytest.fixture(scope="module")
def gpu_case(tmp_path_factory):
    from object_detector_6d_b200 import api
    from tests.helpers import to_api_params
    d = str(tmp_path_factory.mktemp("normals"))
    cam, bgr, depth, p = _frame(seed=5)
    p.fill_random = 1
    layers = synth.make_encoder_weights(11, dims=(384, 1500, 1000, 800))
    locs = O.scan_centres(depth, p)
    sel = np.linspace(0, len(locs) - 1, 3000).astype(np.int64)
    nrm = O.normals(depth, p.normals_focal)
    calib = O.encode(O.quantise_normals(O.gather_normals(bgr, depth, nrm, p, locs[sel])), layers)
    forest_dir = os.path.join(d, "forest")
    synth.write_forest(forest_dir, calib, T=3, K=3, max_depth=10, votes_per_leaf=6, seed=21)
    wpath = os.path.join(d, "weights384.bin")
    synth.write_weights_raw(wpath, layers)
    det = api.Detector(forest_dir, wpath, to_api_params(p), device=0, n_slots=1)
    det.set_debug_capture(True)
    yield dict(det=det, bgr=bgr, depth=depth, p=p, layers=layers, forest_dir=forest_dir, locs=locs, nrm=nrm, wpath=wpath)
    det.close()


@pytest.mark.gpu
def test_cuda_normals_mode_matches_the_oracle(gpu_case):
    from object_detector_6d_b200 import api
    g = gpu_case
    det, p = g["det"], g["p"]
    assert tuple(det.model.dims) == (384, 1500, 1000, 800)
    hyp = det.detect(g["bgr"], g["depth"])
    P, Pp = det.counts(0)
    assert P == len(g["locs"]) and np.array_equal(det.fetch(api.BUF_LOCS)[:P], g["locs"])
    # normals map: bit-exact (the kernel evaluates the oracle's operations one for one)
    n_gpu = det.fetch(api.BUF_NORMALS)
    assert np.array_equal(n_gpu[..., :3].view(np.uint32), g["nrm"].view(np.uint32)) and not n_gpu[..., 3].any()
    # quantised 6-channel patches: bit-exact, random fill included
    q_ref = O.quantise_normals(O.gather_normals(g["bgr"], g["depth"], g["nrm"], p, g["locs"][:Pp]))
    q_gpu = det.fetch(api.BUF_PATCH_U8)
    assert q_gpu.shape == (Pp, 384) and np.array_equal(q_gpu, q_ref)
    # encoder with a 384-wide first layer: tolerance; everything after it on the GPU's own features: bit-exact
    feat = det.fetch(api.BUF_FEATURES)
    ref = O.encode(q_ref, g["layers"])
    err = np.abs(feat - ref)
    assert err.max() < 3e-2 and err.mean() < 2e-3
    forest = O.Forest(g["forest_dir"])
    _, ords = O.traverse(forest, feat)
    assert np.array_equal(det.fetch(api.BUF_LEAF_ORD), ords)
    hyp_ref, counts, _ = O.detect(forest, g["bgr"], g["depth"], p, g["layers"], features_override=feat)
    assert counts == (P, Pp) and len(hyp_ref) == len(hyp) and len(hyp) > 0
    for name in hyp.dtype.names:
        assert np.array_equal(hyp[name], hyp_ref[name]), name


@pytest.mark.gpu
def test_encoder_width_must_match_the_patch_mode(gpu_case):
    from object_detector_6d_b200 import api
    from tests.helpers import to_api_params
    g = gpu_case
    p0 = to_api_params(g["p"])
    p0.patch_mode = 0
    with pytest.raises(api.Hf6dError, match="encoder input 384"):
        api.Detector(g["forest_dir"], g["wpath"], p0, device=0)
