"""Committed golden vectors (tests/golden/, generator: tests/golden/make_golden.py).

CPU: the oracle reproduces every stored intermediate of the tiny case bit for bit (so a change in the oracle or in the
seeded generators is caught).  GPU: the CUDA path is checked against the same stored numbers through the C ABI without
executing the oracle at all."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from object_detector_6d_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights_sha(layers) -> str:
    h = hashlib.sha256()
    for Wm, b in layers:
        h.update(np.ascontiguousarray(Wm).tobytes())
        h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()


def _params(p, raw: bytes):
    """The fixture stores the parameter block as it was when the vectors were made; fields appended to the struct since
    (patch_mode, normals_focal) keep their defaults."""
    import ctypes as C
    n = min(len(raw), type(p).patch_mode.offset)  # the stored block ends with alignment padding
    C.memmove(C.byref(p), raw, n)
    return p


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    g = dict(np.load(os.path.join(GOLD, "tiny_frame.npz")))
    d = str(tmp_path_factory.mktemp("tiny"))
    fdir = os.path.join(d, "forest")
    os.makedirs(fdir)
    for name in g["forest_names"]:
        with open(os.path.join(fdir, str(name)), "wb") as f:
            f.write(g["forest__" + str(name)].tobytes())
    layers = synth.make_encoder_weights(3)
    assert _weights_sha(layers) == str(g["weights_sha"]), "seeded encoder weights changed: regenerate tests/golden"
    wpath = os.path.join(d, "weights.bin")
    synth.write_weights_raw(wpath, layers)
    K, H, W = 2, 120, 160
    maps = np.zeros((K, H, W), np.uint64)
    maps[tuple(g["maps_idx"])] = g["maps_val"]
    g.update(forest_dir=fdir, weights=wpath, layers=layers, maps=maps, params_bytes=g["params"].tobytes())
    return g


def _same_hyps(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for name in a.dtype.names:
        assert np.array_equal(a[name], b[name]), name


def test_oracle_reproduces_golden(tiny):
    from oracle import oracle as O
    p = _params(O.default_params(), tiny["params_bytes"])
    bgr, depth = tiny["bgr"], tiny["depth"]
    locs = O.scan_centres(depth, p)
    assert np.array_equal(locs, tiny["locs"])
    Pp = (len(locs) // p.batch_size) * p.batch_size
    assert (len(locs), Pp) == (266, 200)
    q = O.normalise(O.gather(bgr, depth, p, locs[:Pp]))
    assert np.array_equal(q, tiny["q"])
    feats = O.encode(q, tiny["layers"])
    assert np.array_equal(feats, tiny["features"])
    forest = O.Forest(tiny["forest_dir"])
    ids, ords = O.traverse(forest, feats)
    assert np.array_equal(ids, tiny["leaf_id"]) and np.array_equal(ords, tiny["leaf_ord"])
    maps, n_cast = O.vote(forest, ords, locs[:Pp], depth, p)
    assert n_cast == int(tiny["n_cast"]) and np.array_equal(maps, tiny["maps"])
    for k in range(forest.K):
        assert np.array_equal(O.blur(maps[k], p.centers_blur_size, p.centers_blur_size), tiny["blurred"][k])
    hyp, counts, _ = O.detect(forest, bgr, depth, p, tiny["layers"])
    assert counts == (266, 200)
    _same_hyps(hyp, tiny["hyp"])


def test_independent_restatement_reproduces_golden(tiny):
    """tests/npref.py (numpy / pure Python, written from the reference lines) against the stored vectors."""
    from oracle import oracle as O
    from tests import npref
    p = _params(O.default_params(), tiny["params_bytes"])
    bgr, depth = tiny["bgr"], tiny["depth"]
    locs = npref.scan_centres(depth, p.W, p.H, p.stride, p.patch_vox, p.voxel_m, p.fx, p.distance_threshold_m)
    assert np.array_equal(locs, tiny["locs"])
    patches = npref.gather(bgr, depth, locs[:200], p.W, p.H, p.patch_vox, p.voxel_m, p.fx, p.max_depth_range_m,
                           p.fill_random, p.fill_seed)
    assert np.array_equal(npref.normalise(patches), tiny["q"])
    assert np.abs(npref.encode(tiny["q"], tiny["layers"]) - tiny["features"]).max() < 2e-5
    pf = npref.read_forest(tiny["forest_dir"])
    assert np.array_equal(npref.traverse(pf, tiny["features"]), tiny["leaf_ord"])
    maps = npref.cast_votes(pf, tiny["leaf_ord"], locs[:200], depth, p.W, p.H, p.fx, p.fy, p.cx, p.cy)
    assert np.array_equal(maps, tiny["maps"])
    for k in range(pf["K"]):
        ref = npref.nms(tiny["blurred"][k], p.centers_nms_wsize, p.centers_nms_wsize)
        got = sorted({(int(h["cx"]), int(h["cy"])) for h in tiny["hyp"] if h["cls"] == k})
        assert set(got) <= {(x, y) for _, x, y in ref}


# ------------------------------------------------------------------------------------------------ GPU (no oracle)
@pytest.mark.gpu
def test_cuda_path_against_golden(tiny):
    from object_detector_6d_b200 import api
    p = _params(api.default_params(), tiny["params_bytes"])
    det = api.Detector(tiny["forest_dir"], tiny["weights"], p, device=0)
    det.set_debug_capture(True)
    try:
        det.upload(0, tiny["bgr"], tiny["depth"])
        det.run(0)
        assert det.counts(0) == (266, 200)
        assert np.array_equal(det.fetch(api.BUF_LOCS), tiny["locs"])
        assert np.array_equal(det.fetch(api.BUF_PATCH_U8), tiny["q"])
        feat = det.fetch(api.BUF_FEATURES)
        err = np.abs(feat - tiny["features"])
        assert err.max() < 3e-2 and err.mean() < 2e-3  # bf16 tensor-core encoder vs the fp32 oracle features
        # everything after the encoder on the stored fp32 features: bit-exact
        det.inject(api.BUF_FEATURES, tiny["features"])
        det.run(0, api.STAGE_TRAVERSE, api.STAGE_POSE)
        assert np.array_equal(det.fetch(api.BUF_LEAF_ORD), tiny["leaf_ord"])
        assert np.array_equal(det.fetch(api.BUF_MAPS), tiny["maps"])
        assert np.array_equal(det.fetch(api.BUF_BLURRED), tiny["blurred"])
        _same_hyps(det.collect(0), tiny["hyp"])
    finally:
        det.close()


@pytest.mark.gpu
def test_cuda_blur_against_opencv_golden(tiny):
    """The device box filter against vectors produced by cv::blur (tests/golden/cv_blur.npz)."""
    from object_detector_6d_b200 import api
    g = np.load(os.path.join(GOLD, "cv_blur.npz"))
    p = _params(api.default_params(), tiny["params_bytes"])
    for k in (13, 35):
        p.centers_blur_size = k
        det = api.Detector(tiny["forest_dir"], tiny["weights"], p, device=0)
        try:
            maps = np.zeros((2, 120, 160), np.uint64)
            maps[0] = g["acc"]
            det.upload(0, tiny["bgr"], tiny["depth"])
            det.run(0, api.STAGE_SCAN, api.STAGE_SCAN)
            det.inject(api.BUF_MAPS, maps)
            det.run(0, api.STAGE_CENTRES, api.STAGE_CENTRES)
            ours = det.fetch(api.BUF_BLURRED)[0]
        finally:
            det.close()
        assert np.allclose(ours, g[f"blur{k}"], rtol=2e-6, atol=1e-7)
