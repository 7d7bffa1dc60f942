"""CPU tests of the oracle: the C restatement (oracle/hf6d_oracle.c) against the independent numpy / pure-Python
restatement in tests/npref.py, against OpenCV's own cv::blur where it is available, and against the domain's
size-independent properties.  No GPU, no libhf6d compute."""
import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import oracle as O
from tests import npref
from tests.helpers import make_case


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("small"))
    cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
    cs = make_case(d, K=2, T=2, seed=5, max_depth=8, votes_per_leaf=3, cam=cam, calib_patches=2000)
    cs["forest"] = O.Forest(cs["forest_dir"])
    p = cs["params"]
    cs["locs_all"] = O.scan_centres(cs["depth"], p)
    cs["Pp"] = (len(cs["locs_all"]) // p.batch_size) * p.batch_size
    return cs


def _geom(p):
    return dict(W=p.W, H=p.H, ps=p.patch_vox, vox=p.voxel_m, fx=p.fx)


def test_scan_matches_independent_restatement(small):
    p = small["params"]
    ref = npref.scan_centres(small["depth"], p.W, p.H, p.stride, p.patch_vox, p.voxel_m, p.fx, p.distance_threshold_m)
    assert len(ref) > 1000
    assert np.array_equal(small["locs_all"], ref)
    # row-major push order: y non-decreasing, x increasing inside a row
    y, x = ref[:, 1].astype(np.int64), ref[:, 0].astype(np.int64)
    assert (np.diff(y * 100000 + x) > 0).all()


@pytest.mark.parametrize("fill_random", [0, 1])
def test_gather_matches_independent_restatement(small, fill_random):
    p = small["params"]
    q = O.Params.from_buffer_copy(p)
    q.fill_random, q.fill_seed = fill_random, 77
    locs = small["locs_all"][:: max(1, len(small["locs_all"]) // 1500)]
    a = O.gather(small["bgr"], small["depth"], q, locs)
    b = npref.gather(small["bgr"], small["depth"], locs, p.W, p.H, p.patch_vox, p.voxel_m, p.fx, p.max_depth_range_m,
                     fill_random, 77)
    assert a.shape == b.shape
    assert np.array_equal(a, b), f"{(a != b).sum()} of {a.size} values differ"
    if fill_random:
        assert (a[..., 3] > 0).any()


def test_normalise_matches_independent_restatement(small):
    p = small["params"]
    locs = small["locs_all"][:: max(1, len(small["locs_all"]) // 3000)]
    patches = O.gather(small["bgr"], small["depth"], p, locs)
    a = O.normalise(patches)
    b = npref.normalise(patches)
    assert np.array_equal(a, b), f"{(a != b).any(1).sum()} of {len(a)} patches differ"
    live = a[(a[:, :192] != 0).any(1)]
    assert live[:, :192].min() >= 25 and live[:, :192].max() <= 229  # (x+1)*0.4+0.1 in [0.1, 0.9]


def test_flat_patch_quantises_to_zero(small):
    """variance 0 -> 0/0 = NaN -> (unsigned char)NaN == 0 (HFTest.cpp:538-565 on x86)."""
    patches = np.full((4, 8, 8, 4), 0.25, np.float32)
    a, b = O.normalise(patches), npref.normalise(patches)
    assert np.array_equal(a, b)
    assert (a[:, 192:] == 0).all()  # depth: 64 exact terms 0.25/64 -> mean exact -> variance 0 -> NaN -> 0
    # colour: 0.25/192 is inexact, the 192-term sequential sum misses 0.25 by an ulp or two, the variance is a
    # denormal-sized positive number and every element clips to the same bound: all 25 or all 229, never NaN
    assert set(np.unique(a[:, :192]).tolist()) <= {25, 229}
    patches = np.full((2, 8, 8, 4), 0.5, np.float32)
    patches[..., :3] = 0.75  # 0.75/192 = 2^-8 exactly: colour mean exact as well -> everything NaN -> 0
    assert (O.normalise(patches) == 0).all() and (npref.normalise(patches) == 0).all()


def test_encoder_matches_float64_restatement(small):
    p = small["params"]
    locs = small["locs_all"][:600]
    q = O.normalise(O.gather(small["bgr"], small["depth"], p, locs))
    a = O.encode(q, small["layers"])
    b = npref.encode(q, small["layers"])
    assert a.shape == (600, 800)
    # fp32 accumulation (fixed order) vs float64: a few ulp of the pre-activation
    assert np.abs(a - b).max() < 2e-5
    assert (a > 0).all() and (a < 1).all()


def test_forest_file_roundtrip_three_readers(small):
    """Writer (synth, Python) -> readers: oracle (C), npref (Python), libhf6d's own loader (C++, host-only entry)."""
    from object_detector_6d_b200 import api
    fo = small["forest"]
    pf = npref.read_forest(small["forest_dir"])
    assert (pf["T"], pf["K"], pf["F"], pf["ps"]) == (fo.T, fo.K, fo.F, fo.patch_vox)
    n_leaves = [len(lv) for _, lv in pf["trees"]]
    assert n_leaves == [fo.leaf_count(t) for t in range(fo.T)] == small["stats"]["leaves"]
    mi = api.inspect_forest(small["forest_dir"])
    assert (mi.T, mi.K, mi.F, mi.patch_vox) == (fo.T, fo.K, fo.F, fo.patch_vox)
    assert mi.n_leaves == sum(n_leaves) == fo.n_leaves
    assert mi.n_internal == fo.n_internal == sum(n_leaves) - fo.T  # full binary trees
    gated = sum(len(lf.votes[c]) for _, lv in pf["trees"] for lf in lv for c in range(pf["K"])
                if lf.class_prob[c] >= np.float32(0.5))
    assert mi.n_votes == gated


def test_traverse_matches_independent_restatement(small):
    p = small["params"]
    locs = small["locs_all"][:2000]
    feats = O.encode(O.normalise(O.gather(small["bgr"], small["depth"], p, locs)), small["layers"])
    ids, ords = O.traverse(small["forest"], feats)
    pf = npref.read_forest(small["forest_dir"])
    ref = npref.traverse(pf, feats)
    assert np.array_equal(ords, ref)
    for t, (_, leaves) in enumerate(pf["trees"]):
        lid = np.array([lf.leaf_id for lf in leaves], np.int32)
        assert np.array_equal(ids[:, t], lid[ref[:, t]])
    assert len(np.unique(ords[:, 0])) > 4  # the synthetic splits actually spread the patches


def test_traverse_nan_goes_right(small):
    feats = np.full((3, small["forest"].F), np.nan, np.float32)
    _, ords = O.traverse(small["forest"], feats)
    pf = npref.read_forest(small["forest_dir"])
    assert np.array_equal(ords, npref.traverse(pf, feats))
    for t, (root, _) in enumerate(pf["trees"]):
        n = root
        while not n.leaf:
            n = n.right
        assert (ords[:, t] == n.ordinal).all()


def test_votes_match_independent_restatement(small):
    p = small["params"]
    n = 1500
    locs = small["locs_all"][:n]
    feats = O.encode(O.normalise(O.gather(small["bgr"], small["depth"], p, locs)), small["layers"])
    _, ords = O.traverse(small["forest"], feats)
    maps, cast = O.vote(small["forest"], ords, locs, small["depth"], p)
    pf = npref.read_forest(small["forest_dir"])
    ref = npref.cast_votes(pf, ords, locs, small["depth"], p.W, p.H, p.fx, p.fy, p.cx, p.cy)
    assert cast > 0 and maps.sum() > 0
    assert np.array_equal(maps, ref), f"{(maps != ref).sum()} map cells differ"
    # linearity / sharding property: votes of disjoint tree subsets add up exactly (Q16 integers)
    parts = np.zeros_like(maps)
    for r in range(2):
        o = ords.copy()
        o[:, [t for t in range(pf["T"]) if t % 2 != r]] = -1
        parts += O.vote(small["forest"], o, locs, small["depth"], p)[0]
    assert np.array_equal(parts, maps)
    # should_detect switches a whole class off
    sd = np.array([0, 1], np.uint8)
    m2, _ = O.vote(small["forest"], ords, locs, small["depth"], p, should_detect=sd)
    assert m2[0].sum() == 0 and np.array_equal(m2[1], maps[1])


def test_blur_matches_float64_box_filter(small):
    rng = np.random.default_rng(0)
    acc = np.zeros((60, 90), np.uint64)
    ys, xs = rng.integers(0, 60, 400), rng.integers(0, 90, 400)
    np.add.at(acc, (ys, xs), rng.integers(32768, 65537, 400).astype(np.uint64))
    for k in (13, 35):
        a = O.blur(acc, k, k)
        b = npref.blur_reference(acc, k)
        assert np.allclose(a, b, rtol=3e-7, atol=1e-9)


def test_blur_matches_opencv():
    """cv::blur itself (OpenCV 4.x here; the reference used 2.4.10): same normalised box, BORDER_REFLECT_101."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    acc = np.zeros((120, 160), np.uint64)
    ys, xs = rng.integers(0, 120, 3000), rng.integers(0, 160, 3000)
    np.add.at(acc, (ys, xs), rng.integers(32768, 65537, 3000).astype(np.uint64))
    img = (acc.astype(np.float64) / 65536.0).astype(np.float32)
    for k in (13, 35):
        ours = O.blur(acc, k, k)
        theirs = cv2.blur(img, (k, k))
        assert np.allclose(ours, theirs, rtol=2e-6, atol=1e-7), np.abs(ours - theirs).max()


def test_blur_golden_from_opencv():
    """The same comparison against vectors generated once with cv2.blur and committed (tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv_blur.npz"))
    for k in (13, 35):
        ours = O.blur(g["acc"], k, k)
        assert np.allclose(ours, g[f"blur{k}"], rtol=2e-6, atol=1e-7)


def test_nms_matches_literal_deque_restatement():
    rng = np.random.default_rng(2)
    img = np.zeros((70, 110), np.float32)
    ys, xs = rng.integers(0, 70, 300), rng.integers(0, 110, 300)
    img[ys, xs] = rng.integers(1, 40, 300).astype(np.float32)  # many exact ties
    img = O.blur((img * 65536).astype(np.uint64), 5, 5)
    for wx, wy in ((9, 9), (1, 8), (20, 12)):
        s, xs_, ys_ = O.nms(img, wx, wy)
        ref = npref.nms(img, wx, wy)
        assert len(ref) == len(s)
        assert sorted(zip(s.tolist(), xs_.tolist(), ys_.tolist())) == sorted((float(v), x, y) for v, x, y in ref)
        assert (np.diff(s) <= 0).all()
        if len(ys_):  # the reference's loop-bound quirk: window tops stop at rows - 2*wy + 1
            assert ys_.max() <= img.shape[0] - 2 * wy + 1 + wy // 2


def test_hypotheses_are_consistent(small):
    """End to end on the oracle: hypothesis tuples are quantised as the reference does and poses follow the tuple."""
    p = small["params"]
    hyp, (P, Pp), st = O.detect(small["forest"], small["bgr"], small["depth"], p, small["layers"])
    assert P == len(small["locs_all"]) and Pp == small["Pp"]
    assert len(hyp) > 0
    assert ((hyp["yaw_deg"] >= -180) & (hyp["yaw_deg"] <= 180)).all()  # peaks are kept in [180, 540] - 360
    assert ((hyp["roll_deg"] >= -180) & (hyp["roll_deg"] <= 180)).all()
    assert (hyp["z"] > 0).all() and (hyp["z"] < 3.0).all()
    for h in hyp[:20]:
        R = npref.xtion_rotmat(np.float32(np.float32(h["yaw_deg"]) / np.float32(180.0) * np.pi),
                               np.float32(np.float32(h["pitch_deg"]) / np.float32(180.0) * np.pi),
                               np.float32(np.float32(h["roll_deg"]) / np.float32(180.0) * np.pi))
        pose = h["pose"].reshape(4, 4)
        assert np.allclose(pose[:3, :3], R[:3, :3], atol=1e-6)
        assert np.isclose(pose[2, 3], h["z"])
        assert np.isclose(pose[0, 3], (h["cx"] - p.cx) * h["z"] / p.fx, atol=1e-6)
        assert np.allclose(pose[:3, :3] @ pose[:3, :3].T, np.eye(3), atol=1e-5)


def test_pose_seeking_matches_independent_restatement(small):
    """A10 - A12 (HFTest.cpp:694-925): centre selection, window walk over the back-map with its per-vote multiplicity, z
    histogram and mode, yaw/pitch map with wrap copies, roll histogram, 7-degree separation -- the C oracle against the
    pure-Python restatement in tests/npref.py on a whole 320x240 frame: every hypothesis tuple and every score, bit for bit."""
    p = small["params"]
    locs = small["locs_all"][:small["Pp"]]
    feats = O.encode(O.normalise(O.gather(small["bgr"], small["depth"], p, locs)), small["layers"])
    _, ords = O.traverse(small["forest"], feats)
    max_loc = [5, 3]
    hyp = O.hypotheses(small["forest"], ords, locs, small["depth"], p, max_loc=max_loc)
    pf = npref.read_forest(small["forest_dir"])
    maps, entries = npref.cast_votes(pf, ords, locs, small["depth"], p.W, p.H, p.fx, p.fy, p.cx, p.cy, want_entries=True)
    ref = npref.seek_poses(pf, entries, maps, small["depth"], p.W, p.H, p.fx, p.fy, p.cx, p.cy,
                           centers_blur=p.centers_blur_size, centers_nms=p.centers_nms_wsize, pose_blur=p.pose_blur_size,
                           pose_nms=p.pose_nms_wsize, max_loc=max_loc, max_yaw_pitch=p.max_yaw_pitch_hypotheses,
                           max_roll=p.max_roll_hypotheses, min_loc_ratio=p.min_location_score_ratio,
                           min_yp_ratio=p.min_yaw_pitch_drop_ratio)
    assert len(ref) == len(hyp) > 50
    for a, h in zip(ref, hyp):
        b = (int(h["cls"]), int(h["cx"]), int(h["cy"]), float(h["z"]), int(h["yaw_deg"]), int(h["pitch_deg"]), int(h["roll_deg"]),
             float(h["loc_score"]), float(h["yawpitch_score"]), float(h["roll_score"]))
        assert a == b
    # a class switched off produces nothing, the other class is untouched (HFTest.cpp:695)
    sd = np.array([0, 1], np.uint8)
    hyp1 = O.hypotheses(small["forest"], ords, locs, small["depth"], p, should_detect=sd, max_loc=max_loc)
    maps1, entries1 = npref.cast_votes(pf, ords, locs, small["depth"], p.W, p.H, p.fx, p.fy, p.cx, p.cy, should_detect=sd,
                                       want_entries=True)
    ref1 = npref.seek_poses(pf, entries1, maps1, small["depth"], p.W, p.H, p.fx, p.fy, p.cx, p.cy,
                            centers_blur=p.centers_blur_size, centers_nms=p.centers_nms_wsize, pose_blur=p.pose_blur_size,
                            pose_nms=p.pose_nms_wsize, max_loc=max_loc, max_yaw_pitch=p.max_yaw_pitch_hypotheses,
                            max_roll=p.max_roll_hypotheses, min_loc_ratio=p.min_location_score_ratio,
                            min_yp_ratio=p.min_yaw_pitch_drop_ratio, should_detect=sd)
    assert len(ref1) == len(hyp1) > 0 and all(h["cls"] == 1 for h in hyp1)
    assert [r[:7] for r in ref1] == [r[:7] for r in ref if r[0] == 1]


def test_empty_and_far_frames(small):
    p = small["params"]
    zero = np.zeros_like(small["depth"])
    assert len(O.scan_centres(zero, p)) == 0
    hyp, (P, Pp), _ = O.detect(small["forest"], small["bgr"], zero, p, small["layers"])
    assert (P, Pp, len(hyp)) == (0, 0, 0)
    far = np.full_like(small["depth"], 2000)  # beyond distance_threshold 1.5 m
    assert len(O.scan_centres(far, p)) == 0
    few = zero.copy()
    few[100:110, 100:110] = 700  # fewer valid centres than one batch: the reference drops the partial batch
    hyp, (P, Pp), _ = O.detect(small["forest"], small["bgr"], few, p, small["layers"])
    assert 0 < P < p.batch_size and Pp == 0 and len(hyp) == 0


def test_cpu_baseline_uses_all_cores_even_under_torchrun(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1 into every rank; the reference arm of bench.py must still time the oracle on all
    host cores (it sets the thread count through the OpenMP runtime, not the environment)."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; print(O.set_threads())"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
    assert int(out.strip().splitlines()[-1]) == (os.cpu_count() or 1)
