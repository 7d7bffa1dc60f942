"""Pins of the CPU oracle (oracle/hf6d_oracle.c) to the REFERENCE'S OWN CODE.

oracle/_ref/libhf6d_refsrc.so is HoughForest/src/HFBase.cpp + HFTest.cpp (unmodified, compiled where they lie under
/root/reference), PatchGen/src/cuda/patch_extractor.cu + surface_normals.cu (kernel launches rewritten to a host loop,
nothing else) and the pre-ICP head of MeshUtils::icp, built against stand-in headers for the absent libraries
(oracle/build_ref.py, oracle/ref_shim/).  Everything below compares what that code computes with what the oracle computes
on the same seeded inputs -- bit for bit wherever the reference's result is defined bit for bit:

  HFBase::loadForestFromFolder / loadNodeFromFile   HFBase.cpp:58-145      forest reader (A5)
  HFTest::get_leaf                                  HFTest.cpp:144-163     leaf assignment, NaN -> right (A6)
  get_obj_center_vote_from_6dof + Point3DToImage    HFTest.cpp:21-102      vote geometry (A7)
  HFTest::non_max_suppression                       HFTest.cpp:219-268     sliding-window NMS and its loop-bound quirk (A9)
  extract_patches_rgbd + kernel extract_rgbd        patch_extractor.cu:230-433   centre scan + gather (A2a, A2b)
  generate_normals, extract_patches + kernel extract   surface_normals.cu, patch_extractor.cu:12-221   (A2c)
  HFTest::test_image                                HFTest.cpp:296-1024    texture, normalise + quantise, voting, merge, blur
                                                                           + NMS, z / yaw-pitch / roll seeking, pre-ICP pose
                                                                           (A1, A3, A7-A12) -- the whole per-frame path

What is NOT pinned by this (the stand-ins decide it, see their headers): the encoder arithmetic (Caffe: the stand-in net
calls the oracle's fp32 encoder), the texture unit's filter (checked against the B200's texture unit in the GPU tests),
the fill RNG (clock-seeded in the reference), cv::blur's accumulation (checked against cv2 in tests/test_golden.py).

The reference accumulates votes in float (order = patch order per batch, HFTest.cpp:201-203, :649); the oracle in Q16
integers.  With class probabilities that are multiples of 1/16 both are exact, so every score must agree bit for bit; with
arbitrary probabilities the tuples must agree and the scores to float rounding.
"""
import ctypes as C
import os

import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import oracle as O
from oracle import refsrc as R
from tests import npref
from tests.helpers import make_case

if os.path.isdir("/root/reference"):
    assert R.available(), "the reference tree is present, so the reference-source library must build"
pytestmark = pytest.mark.skipif(not R.available(), reason="no oracle/_ref/libhf6d_refsrc.so (reference tree absent and no prebuilt library)")

CAM = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)


def _params(p, **kw):
    q = O.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    for k, v in kw.items():
        setattr(q, k, v)
    return q


@pytest.fixture(scope="module")
def dyadic(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("pin_dyadic"))
    cs = make_case(d, K=2, T=2, seed=5, max_depth=10, votes_per_leaf=4, cam=CAM, calib_patches=3000, prob_quantum=16)
    cs["forest"] = O.Forest(cs["forest_dir"])
    cs["ref"] = R.Reference(cs["forest_dir"], cs["weights"])
    cs["ref_hyp"] = cs["ref"].test_image(cs["bgr"], cs["depth"], cs["params"], capture=True)
    cs["ref_net_input"] = R.captured_net_input()
    cs["ref_maps"] = R.captured_maps()
    cs["ref_blurred"] = R.captured_blurred()
    cs["oracle_hyp"], cs["counts"], _ = O.detect(cs["forest"], cs["bgr"], cs["depth"], cs["params"], cs["layers"])
    return cs


# ------------------------------------------------------------------------------------------------ A5 forest loader
def test_loader_agrees_with_both_readers(dyadic):
    """HFBase::loadNodeFromFile's trees == the independent python reader's, and the product's inspect counts."""
    from object_detector_6d_b200 import api
    ref = dyadic["ref"]
    assert (ref.T, ref.K, ref.F, ref.patch_vox) == (dyadic["forest"].T, dyadic["forest"].K, dyadic["forest"].F, 8)
    assert ref.voxel_m == dyadic["forest"].voxel_m
    n_leaves = n_internal = n_gated = 0
    for t in range(ref.T):
        dump = ref.tree_dump(t)
        root, _ = npref.read_tree(os.path.join(dyadic["forest_dir"], f"tree{t}.dat"), ref.K)
        ids, probs, counts, votes, tests, thr = [], [], [], [], [], []

        def walk(n):
            if n.leaf:
                ids.append(n.leaf_id)
                probs.append(n.class_prob)
                counts.append([len(v) for v in n.votes])
                votes.extend(v for cv in n.votes for v in cv)
            else:
                tests.append((n.mode, n.f1, n.f2))
                thr.append(n.thr)
                walk(n.left)
                walk(n.right)
        walk(root)
        assert np.array_equal(dump["leaf_id"], np.array(ids, np.int32))
        assert np.array_equal(dump["class_prob"], np.array(probs, np.float32))
        assert np.array_equal(dump["vote_count"], np.array(counts, np.int32))
        assert np.array_equal(dump["votes"], np.array(votes, np.float32).reshape(-1, 6))
        assert np.array_equal(dump["tests"], np.array(tests, np.int32).reshape(-1, 3))
        assert np.array_equal(dump["thresholds"], np.array(thr, np.float32))
        assert dyadic["forest"].leaf_count(t) == len(ids)
        n_leaves += len(ids)
        n_internal += len(thr)
        n_gated += int(dump["vote_count"][dump["class_prob"] >= 0.5].sum())
    info = api.inspect_forest(dyadic["forest_dir"])  # libhf6d's own reader (host only, no GPU)
    assert (info.n_leaves, info.n_internal, info.n_votes) == (n_leaves, n_internal, n_gated)


# ------------------------------------------------------------------------------------------------ A6 get_leaf
def test_get_leaf_bitexact_including_nans(dyadic):
    rng = np.random.default_rng(3)
    feats = O.encode(O.normalise(O.gather(dyadic["bgr"], dyadic["depth"], dyadic["params"], dyadic["locs"][:3000])),
                     dyadic["layers"])
    feats = np.concatenate([feats, rng.normal(0.5, 0.3, (1000, feats.shape[1])).astype(np.float32)])
    feats[rng.random(feats.shape) < 0.01] = np.nan  # NaN compares false: descend right (HFTest.cpp:158)
    ids, _ = O.traverse(dyadic["forest"], feats)
    assert np.array_equal(dyadic["ref"].get_leaves(feats), ids)


# ------------------------------------------------------------------------------------------------ A7 vote geometry
def test_vote_geometry_bitexact():
    """get_obj_center_vote_from_6dof + Point3DToImage against the literal restatement the oracle is checked with
    (npref.xtion_rotmat: 4x4 products in Eigen order), including z == 0 and projections that truncate toward zero."""
    rng = np.random.default_rng(11)
    n = 4000
    dof = np.zeros((n, 6), np.float32)
    dof[:, 0] = rng.uniform(-np.pi, np.pi, n)
    dof[:, 1] = rng.uniform(-np.pi / 2, np.pi / 2, n)
    dof[:, 2] = rng.uniform(-np.pi, np.pi, n)
    dof[:, 3:] = rng.uniform(-0.3, 0.3, (n, 3))
    px = rng.integers(0, 640, n).astype(np.int32)
    py = rng.integers(0, 480, n).astype(np.int32)
    dm = rng.integers(300, 1500, n).astype(np.uint16)
    dof[:50, 3:] = 0
    dm[:50] = 0  # centre at the camera: z == 0 -> (0, 0), HFTest.cpp:25-28
    c3, uv = R.vote_pixels(dof, px, py, dm)
    fx = fy = np.float32(575.0)
    cx, cy = np.float32(319.5), np.float32(239.5)
    for i in range(n):
        Rm = npref.xtion_rotmat(dof[i, 0], dof[i, 1], dof[i, 2])
        z = np.float32(dm[i]) / np.float32(1000.0)
        t = np.array([(np.float32(px[i]) - cx) * z / fx, (np.float32(py[i]) - cy) * z / fy, z], np.float32)
        v = -dof[i, 3:]
        want = np.array([np.float32(np.float32(np.float32(Rm[r, 0] * v[0]) + np.float32(Rm[r, 1] * v[1])) + np.float32(Rm[r, 2] * v[2])) + t[r]
                         for r in range(3)], np.float32)
        assert np.array_equal(c3[i], want), i
        if want[2] == 0:
            assert tuple(uv[i]) == (0, 0)
        else:
            u = int(npref._f2i(np.float32(np.float32(want[0] / want[2]) * fx + cx) + np.float32(0.5)))
            w = int(npref._f2i(np.float32(np.float32(want[1] / want[2]) * fy + cy) + np.float32(0.5)))
            assert tuple(uv[i]) == (u, w), i


# ------------------------------------------------------------------------------------------------ A9 NMS
@pytest.mark.parametrize("rows,cols,wx,wy,seed", [(96, 128, 40, 40, 1), (300, 1, 1, 20, 2), (200, 180, 35, 35, 3),
                                                  (720, 1, 1, 35, 4), (60, 70, 7, 5, 5), (40, 40, 40, 40, 6), (30, 50, 40, 40, 7)])
def test_nms_bitexact(rows, cols, wx, wy, seed):
    """HFTest::non_max_suppression == the oracle's on smooth random maps (distinct scores: order is defined)."""
    rng = np.random.default_rng(seed)
    acc = (rng.random((rows, cols)) < 0.05) * rng.integers(1, 1 << 20, (rows, cols))
    img = O.blur(acc.astype(np.uint64), min(5, cols) | 1 if cols > 1 else 1, min(5, rows) | 1)
    s0, x0, y0 = R.nms(img, wx, wy)
    s1, x1, y1 = O.nms(img, wx, wy)
    assert len(s0) == len(s1)
    k0 = sorted(zip(-s0, x0, y0))
    k1 = sorted(zip(-s1, x1, y1))
    assert k0 == k1
    if len(set(s0.tolist())) == len(s0):  # no equal scores: std::sort's order is the oracle's
        assert np.array_equal(s0, s1) and np.array_equal(x0, x1) and np.array_equal(y0, y1)


def test_nms_ties_and_plateaus():
    """Plateaus of equal values: the two monotonic deques keep the earliest element; same SET of maxima as the oracle
    (the order of equal scores is std::sort's, i.e. unspecified in the reference: choice C10)."""
    rng = np.random.default_rng(21)
    for _ in range(20):
        img = rng.integers(0, 4, (50, 64)).astype(np.float32)
        a = set(zip(*[v.tolist() for v in R.nms(img, 9, 7)]))
        b = set(zip(*[v.tolist() for v in O.nms(img, 9, 7)]))
        assert a == b


# ------------------------------------------------------------------------------------------------ A2a + A2b gather
@pytest.mark.parametrize("fill_random", [0, 1])
def test_scan_and_gather_bitexact(dyadic, fill_random):
    """patch_extractor_gpu::extract_patches_rgbd (host scan loop + the extract_rgbd kernel, run on the host with the
    software texture filter) == oracle scan + gather: patch order, adaptive size, sample positions, depth clamp, fill."""
    p = _params(dyadic["params"], fill_random=fill_random, fill_seed=77)
    locs_ref, patches_ref = R.extract_rgbd(dyadic["bgr"], dyadic["depth"], p)
    locs = O.scan_centres(dyadic["depth"], p)
    assert np.array_equal(locs_ref, locs)
    patches = O.gather(dyadic["bgr"], dyadic["depth"], p, locs)
    assert np.array_equal(patches_ref.view(np.uint32), patches.view(np.uint32))
    if fill_random:
        assert (patches_ref[..., :3].reshape(len(locs), -1).std(1) > 0).any()


def test_scan_honours_distance_threshold_and_stride(dyadic):
    p = _params(dyadic["params"], stride=3, distance_threshold_m=0.9)
    locs_ref, _ = R.extract_rgbd(dyadic["bgr"], dyadic["depth"], p)
    assert np.array_equal(locs_ref, O.scan_centres(dyadic["depth"], p))
    assert 0 < len(locs_ref) < len(dyadic["locs"])


# ------------------------------------------------------------------------------------------------ A2c normals variant
@pytest.mark.parametrize("fill_random", [0, 1])
def test_normals_variant_bitexact(dyadic, fill_random):
    p = _params(dyadic["params"], patch_mode=1, fill_random=fill_random, fill_seed=5, normals_focal=287.5)
    nrm_ref = R.normals(dyadic["depth"], p.normals_focal)
    nrm = O.normals(dyadic["depth"], p.normals_focal)
    assert np.array_equal(nrm_ref.view(np.uint32), nrm.view(np.uint32))
    locs_ref, patches_ref = R.extract_normals(dyadic["bgr"], dyadic["depth"], nrm_ref, p)
    locs = O.scan_centres(dyadic["depth"], p)
    assert np.array_equal(locs_ref, locs)
    patches = O.gather_normals(dyadic["bgr"], dyadic["depth"], nrm, p, locs)
    assert np.array_equal(patches_ref.view(np.uint32), patches.view(np.uint32))


# ------------------------------------------------------------------------------------------------ A1 + A3 inside test_image
def test_net_input_is_the_oracles_quantised_patch(dyadic):
    """What test_image hands to the net (HFTest.cpp:370-398 texture + gather, :500-570 normalise / quantise, batches of
    100, tail dropped) == oracle normalise(gather) / 255, bit for bit -- including the NaN -> 0 patches."""
    P, Pp = dyadic["counts"]
    x = dyadic["ref_net_input"].reshape(-1, 256)
    assert x.shape[0] == Pp == (P // 100) * 100
    q = O.normalise(O.gather(dyadic["bgr"], dyadic["depth"], dyadic["params"], dyadic["locs"][:Pp]))
    assert np.array_equal(x, q.astype(np.float32) / np.float32(255.0))


def test_flat_patches_quantise_to_zero_in_the_reference_too(dyadic):
    """Constant colour and depth: variance 0 -> 0/0 = NaN -> (unsigned char)NaN.  The reference's own code, compiled for
    x86-64, gives 0 (cvttss2si), which is what the oracle (choice C4) and the CUDA kernel reproduce."""
    bgr = np.full_like(dyadic["bgr"], 200)
    depth = np.full_like(dyadic["depth"], 800)
    dyadic["ref"].test_image(bgr, depth, dyadic["params"], capture=True)
    x = R.captured_net_input().reshape(-1, 256)
    locs = O.scan_centres(depth, dyadic["params"])
    Pp = (len(locs) // 100) * 100
    q = O.normalise(O.gather(bgr, depth, dyadic["params"], locs[:Pp]))
    assert Pp > 0 and x.shape[0] == Pp
    assert (q[:, 192:] == 0).all() and (x[:, 192:] == 0).all()  # depth plane: 0/0 (colour: 200/255 sums do not cancel exactly)
    assert np.array_equal(x, q.astype(np.float32) / np.float32(255.0))


# ------------------------------------------------------------------------------------------------ A7-A12 whole frame
def _assert_same_hypotheses(ref, ora, exact_scores):
    """Same hypotheses, same values.  Roll modes with EQUAL scores come out of std::sort in an unspecified order in the
    reference (choice C10: the oracle keeps emission order), so both lists are put in tuple order first; every tuple
    (class, centre, yaw, pitch, roll) occurs once."""
    assert len(ref) == len(ora) and len(ora) > 0
    deg = lambda r: np.rint(r.astype(np.float64) * 180.0 / np.pi).astype(np.int64)  # noqa: E731
    kr = np.lexsort((deg(ref["roll"]), deg(ref["pitch"]), deg(ref["yaw"]), ref["row"], ref["col"], ref["obj_id"]))
    ko = np.lexsort((ora["roll_deg"], ora["pitch_deg"], ora["yaw_deg"], ora["cy"], ora["cx"], ora["cls"]))
    moved = np.nonzero(kr != ko)[0]
    ref, ora = ref[kr], ora[ko]
    assert np.array_equal(ref["obj_id"], ora["cls"])
    assert np.array_equal(ref["col"], ora["cx"]) and np.array_equal(ref["row"], ora["cy"])
    assert np.array_equal(ref["z"], ora["z"])
    for a, b in (("yaw", "yaw_deg"), ("pitch", "pitch_deg"), ("roll", "roll_deg")):
        # HFTest.cpp:922-924: (float)(bin - 360) / 180.0f * M_PI narrowed to float -- the same expression here
        want = (ora[b].astype(np.float32) / np.float32(180.0)).astype(np.float64) * np.pi
        assert np.array_equal(ref[a], want.astype(np.float32)), a
    assert np.array_equal(ref["rotmat"], ora["pose"]), "pre-ICP pose (MeshUtils.cpp:29-59, 423-440)"
    pose_score = (ora["yawpitch_score"] + ora["roll_score"]) / np.float32(2.0)  # HFTest.cpp:934
    if exact_scores:
        assert np.array_equal(ref["location_score"], ora["loc_score"])
        assert np.array_equal(ref["pose_score"], pose_score)
    else:
        assert np.allclose(ref["location_score"], ora["loc_score"], rtol=2e-5, atol=0)
        assert np.allclose(ref["pose_score"], pose_score, rtol=2e-5, atol=0)
    # hypotheses that changed places between the two lists can only be roll modes of equal score
    for i in moved:
        same = (ora["cls"] == ora["cls"][i]) & (ora["cx"] == ora["cx"][i]) & (ora["cy"] == ora["cy"][i]) & \
               (ora["yaw_deg"] == ora["yaw_deg"][i]) & (ora["pitch_deg"] == ora["pitch_deg"][i]) & \
               (ora["roll_score"] == ora["roll_score"][i])
        assert same.sum() >= 2, f"hypothesis {i} is ordered differently without a score tie"


def test_whole_frame_bitexact_with_dyadic_probabilities(dyadic):
    """HFTest::test_image end to end == oracle detect: every hypothesis tuple, every score, every pose entry."""
    _assert_same_hypotheses(dyadic["ref_hyp"], dyadic["oracle_hyp"], exact_scores=True)


def test_vote_maps_and_blur_bitexact(dyadic):
    """The reference's float centre maps (before and after cv::blur) == the oracle's Q16 maps / 65536 and their blur."""
    K, H, W = dyadic["forest"].K, CAM.H, CAM.W
    P, Pp = dyadic["counts"]
    feats = O.encode(O.normalise(O.gather(dyadic["bgr"], dyadic["depth"], dyadic["params"], dyadic["locs"][:Pp])), dyadic["layers"])
    _, ords = O.traverse(dyadic["forest"], feats)
    maps, n_cast = O.vote(dyadic["forest"], ords, dyadic["locs"][:Pp], dyadic["depth"], dyadic["params"])
    ref_maps = dyadic["ref_maps"].reshape(K, H, W)
    assert n_cast > 0 and np.array_equal(ref_maps, (maps.astype(np.float64) / 65536.0).astype(np.float32))
    ref_blur = dyadic["ref_blurred"].reshape(K, H, W)
    for c in range(K):
        assert np.array_equal(ref_blur[c], O.blur(maps[c], 13, 13)), c


def test_whole_frame_with_openmp_threads(dyadic):
    """The reference votes inside an OpenMP loop and merges per-thread maps (HFTest.cpp:601-656): with exact sums the
    result cannot depend on the schedule.  Hypotheses arrive in thread order, so compare as sets."""
    hyp = dyadic["ref"].test_image(dyadic["bgr"], dyadic["depth"], dyadic["params"], n_threads=4)
    a = sorted(h.tobytes() for h in hyp)
    b = sorted(h.tobytes() for h in dyadic["ref_hyp"])
    assert a == b


def test_whole_frame_object_switches(dyadic):
    sd, ml = [0, 1], [12, 3]
    ref = dyadic["ref"].test_image(dyadic["bgr"], dyadic["depth"], dyadic["params"], should_detect=sd, max_loc=ml)
    ora, _, _ = O.detect(dyadic["forest"], dyadic["bgr"], dyadic["depth"], dyadic["params"], dyadic["layers"],
                         should_detect=sd, max_loc=ml)
    assert (ora["cls"] == 1).all()
    _assert_same_hypotheses(ref, ora, exact_scores=True)


def test_whole_frame_with_arbitrary_probabilities(tmp_path):
    """Class probabilities that are not dyadic: the reference's float sums round, the oracle's Q16 sums do not (choice
    C8).  Tuples and poses must still be identical, scores equal to float rounding."""
    cs = make_case(str(tmp_path), K=3, T=3, seed=9, max_depth=9, votes_per_leaf=3, cam=CAM, calib_patches=3000, fill_random=1)
    ref = R.Reference(cs["forest_dir"], cs["weights"])
    rh = ref.test_image(cs["bgr"], cs["depth"], cs["params"])
    oh, _, _ = O.detect(O.Forest(cs["forest_dir"]), cs["bgr"], cs["depth"], cs["params"], cs["layers"])
    _assert_same_hypotheses(rh, oh, exact_scores=False)
