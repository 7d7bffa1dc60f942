"""Stage REFINE (ICP + hypothesis scoring + joint optimisation, MeshUtils.cpp:341-464, 629-793, 864-1168).

CPU: the oracle (oracle/refine.py) against independent restatements and known answers.  GPU: libhf6d's refine path against the
oracle through the C ABI.  The reference leaves this step's arithmetic to PCL (not vendored): parity here is tolerance parity,
and the tolerances are stated next to every comparison.
"""
import itertools
import os
import tempfile

import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import refine as R

CAM = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
OBJECT_SEED = 1000


def refine_params(cam=CAM, **kw):
    return R.RefineParams(fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, **kw)


@pytest.fixture(scope="module")
def scene_case():
    bgr, depth, truth = synth.render_scene(3, OBJECT_SEED, CAM, n_objects=3)
    clouds = synth.object_models(OBJECT_SEED, 3)
    p = refine_params()
    models = [R.ObjectModel(x, c, p, nn_search_radius=0.015, icp_iterations=30) for x, c in clouds]
    scene = R.Scene(bgr, depth, p)
    return dict(bgr=bgr, depth=depth, truth=truth, clouds=clouds, p=p, models=models, scene=scene)


def truth_pose(truth, k):
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = truth["R"][k]
    m[:3, 3] = truth["centre"][k]
    return m


def perturbed(pose, seed, ang=0.06, shift=0.006):
    rng = np.random.default_rng(seed)
    d = np.eye(4, dtype=np.float32)
    d[:3, :3] = synth._rot_axis(rng.normal(size=3), ang)
    d[:3, 3] = rng.uniform(-shift, shift, 3)
    return (pose @ d).astype(np.float32)


def make_hyps(case, seeds=(1, 2), wrong=True):
    """Hypothesis tuples around the true poses (what the Hough stage hands over), plus clearly wrong ones."""
    from oracle.oracle import HYP_DTYPE
    rows = []
    for k in range(len(case["models"])):
        for s in seeds:
            rows.append((k, perturbed(truth_pose(case["truth"], k), 10 * k + s), 1.0 - 0.1 * s, 0.9, 0.8))
        if wrong:
            bad = truth_pose(case["truth"], k).copy()
            bad[:3, 3] += np.array([0.12, -0.05, 0.1], np.float32)
            rows.append((k, bad, 0.5, 0.5, 0.5))
    h = np.zeros(len(rows), HYP_DTYPE)
    for i, (k, pose, loc, yp, ro) in enumerate(rows):
        h[i]["cls"] = k
        h[i]["pose"] = pose.reshape(-1)
        h[i]["loc_score"], h[i]["yawpitch_score"], h[i]["roll_score"] = loc, yp, ro
        h[i]["z"] = pose[2, 3]
    return h


# ----------------------------------------------------------------------------------------------------------- CPU: the oracle
def test_voxel_grid_matches_a_dictionary_restatement():
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-0.05, 0.05, (4000, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (4000, 3)).astype(np.uint8)
    vx, vc, ijk = R.voxel_grid(xyz, rgb, 0.005)
    cells = {}
    inv = np.float32(1.0) / np.float32(0.005)
    for q, c in zip(xyz, rgb):
        key = tuple(np.floor(q * inv).astype(int)[::-1])  # (k, j, i): PCL's index order
        cells.setdefault(key, []).append((q, c))
    keys = sorted(cells)
    assert len(keys) == len(vx)
    for n, key in enumerate(keys):
        q = np.array([a for a, _ in cells[key]], np.float64)
        c = np.array([b for _, b in cells[key]], np.float64)
        assert tuple(ijk[n][::-1]) == key
        np.testing.assert_allclose(vx[n], q.mean(0), atol=1e-7)
        assert np.array_equal(vc[n], np.floor(c.mean(0)).astype(np.uint8))


def test_normals_of_a_plane_and_isolated_points():
    g = np.arange(-0.04, 0.04, 0.005, dtype=np.float32)
    X, Y = np.meshgrid(g, g)
    plane = np.stack([X.ravel(), Y.ravel(), 0.7 + 0.2 * X.ravel()], 1).astype(np.float32)
    lonely = np.array([[0.5, 0.5, 0.5]], np.float32)
    pts = np.concatenate([plane, lonely])
    nrm, curv = R.estimate_normals(pts, 0.03)
    n_true = np.array([0.2, 0.0, -1.0]) / np.linalg.norm([0.2, 0.0, -1.0])  # towards the camera at the origin
    np.testing.assert_allclose(nrm[:len(plane)], np.tile(n_true, (len(plane), 1)), atol=1e-4)
    assert np.all(curv[:len(plane)] < 1e-6)
    assert np.isnan(nrm[-1]).all()  # fewer than 3 neighbours
    xyz, _, _, _, ok = R.normals_not_nan(pts, np.zeros((len(pts), 3), np.uint8), 0.03)
    assert len(xyz) == len(plane) and not ok[-1]


def test_ply_round_trip_and_bounding_box(tmp_path):
    xyz, rgb = synth.object_models(OBJECT_SEED, 1, spacing=0.01)[0]
    path = str(tmp_path / "obj.ply")
    synth.write_ply(path, xyz, rgb)
    x2, c2, mcl = R.read_ply(path)
    np.testing.assert_allclose(x2, xyz, atol=1e-6)
    assert np.array_equal(c2, rgb)
    corners = np.array(list(itertools.product(*zip(xyz.min(0), xyz.max(0)))))
    assert abs(mcl - np.linalg.norm(corners - xyz.mean(0), axis=1).max()) < 1e-4


def test_icp_recovers_a_known_perturbation(scene_case):
    case = scene_case
    for k in range(len(case["models"])):
        pose = truth_pose(case["truth"], k)
        start = perturbed(pose, 7 + k)
        out, ok, its = R.icp(case["scene"], case["models"][k], case["p"], start)
        assert ok and 0 < its <= 30
        pts = case["models"][k].xyz
        e0 = np.linalg.norm(R._transform(pts, start) - R._transform(pts, pose), axis=1).mean()
        e1 = np.linalg.norm(R._transform(pts, out) - R._transform(pts, pose), axis=1).mean()
        d1, _ = case["scene"].tree.query(R._transform(pts, out).astype(np.float64))
        d0, _ = case["scene"].tree.query(R._transform(pts, start).astype(np.float64))
        assert np.median(d1) <= np.median(d0) + 1e-4, (k, np.median(d0), np.median(d1))  # closer to the surface
        assert e1 < max(e0, 0.02)  # and not further from the truth (symmetric solids may slide)


def test_icp_without_an_object_radius_keeps_the_hough_pose(scene_case):
    case = scene_case
    m = R.ObjectModel(*case["clouds"][0], case["p"])  # nn_search_radius = -1: obj_nn_search_radius_[id] is 0
    start = perturbed(truth_pose(case["truth"], 0), 3)
    out, ok, _ = R.icp(case["scene"], m, case["p"], start)
    assert not ok and np.array_equal(out, start)


def test_scoring_prefers_the_true_pose(scene_case):
    case = scene_case
    for k in range(len(case["models"])):
        good = R.evaluate_hypothesis(case["scene"], case["models"][k], case["p"], truth_pose(case["truth"], k), 1.0, 1.0)
        bad_pose = truth_pose(case["truth"], k).copy()
        bad_pose[:3, 3] += np.array([0.12, -0.05, 0.1], np.float32)
        bad = R.evaluate_hypothesis(case["scene"], case["models"][k], case["p"], bad_pose, 1.0, 1.0)
        assert good.inliers_ratio > 0.9 and good.final_score > bad.final_score
        assert not bad.accepted
        assert good.explained.sum() > 0


def test_solution_enumeration_is_every_independent_set():
    group = [0, 1, 2, 3, 4]
    excl = {(0, 1), (1, 0), (1, 2), (2, 1), (3, 4), (4, 3)}
    sol = [False] * 5
    seen = []
    while R._next_solution(sol, group, excl, False):
        seen.append(tuple(sol))
    want = [s for s in itertools.product((False, True), repeat=5)
            if any(s) and not any(s[a] and s[b] for a, b in excl)]
    assert sorted(seen) == sorted(want) and len(seen) == len(set(seen))
    sol = [False] * 3
    seen = []
    while R._next_solution(sol, [0, 1, 2], set(), True):
        seen.append(tuple(sol))
    assert seen == [(True, False, False), (False, True, False), (False, False, True)]


def test_joint_optimisation_keeps_one_pose_per_object(scene_case):
    case = scene_case
    out = R.refine_frame(case["scene"], case["models"], case["p"], make_hyps(case))
    hy = make_hyps(case)
    chosen_cls = [int(hy[out["accepted"][i]]["cls"]) for i in out["chosen"]]
    assert len(chosen_cls) == len(set(chosen_cls)) >= 1  # poses of one object explain the same scene points: mutually exclusive
    for i in out["chosen"]:
        assert out["evals"][out["accepted"][i]].accepted


# ----------------------------------------------------------------------------------------------------------- GPU vs oracle
@pytest.fixture(scope="module")
def gpu_case(scene_case):
    from object_detector_6d_b200 import api
    from tests.helpers import make_case, to_api_params
    case = scene_case
    d = tempfile.mkdtemp(prefix="hf6d_refine_")
    cs = make_case(d, K=3, T=2, seed=5, max_depth=8, votes_per_leaf=4, cam=CAM, calib_patches=2000)
    det = api.Detector(cs["forest_dir"], cs["weights"], to_api_params(cs["params"]), device=0)
    for k, (x, c) in enumerate(case["clouds"]):
        if k == 0:  # one model through the PLY reader
            path = os.path.join(d, "obj0.ply")
            synth.write_ply(path, x, c)
            det.load_object_ply(0, path, 0.015, 30)
        else:
            det.set_object_model(k, x, c, 0.015, 30)
    det.upload(0, case["bgr"], case["depth"])
    det.sync(0)
    hyps = make_hyps(case)
    dets = det.refine(hyps)
    yield dict(det=det, hyps=hyps, dets=dets, api=api)
    det.close()


def _angle(a, b):
    return np.arccos(np.clip(np.abs(np.sum(a * b, 1)), 0, 1))


@pytest.mark.gpu
def test_gpu_models_match_the_oracle(scene_case, gpu_case):
    api, det = gpu_case["api"], gpu_case["det"]
    for k, m in enumerate(scene_case["models"]):
        pts = det.refine_fetch(api.RBUF_MODEL_POINTS, k)
        nrm = det.refine_fetch(api.RBUF_MODEL_NORMALS, k)
        assert len(pts) == len(m.xyz)
        np.testing.assert_allclose(pts[:, :3], m.xyz, atol=2e-6)  # fixed-point vs double means
        rgb = pts[:, 3].copy().view(np.uint32)
        assert np.array_equal(np.stack([rgb & 255, (rgb >> 8) & 255, (rgb >> 16) & 255], 1).astype(np.uint8), m.rgb)
        n2, _ = R.estimate_normals(m.xyz, scene_case["p"].normals_radius)  # what evaluate_hypothesis re-estimates
        ok = np.isfinite(n2).all(1)
        assert np.array_equal(ok, np.isfinite(nrm[:, :3]).all(1))
        flat = m.curvature[ok] < 0.02  # away from edges the normal is well conditioned
        assert np.percentile(_angle(nrm[ok][flat][:, :3], n2[ok][flat]), 99) < 2e-3


@pytest.mark.gpu
def test_gpu_scene_matches_the_oracle(scene_case, gpu_case):
    api, det, sc = gpu_case["api"], gpu_case["det"], scene_case["scene"]
    pts = det.refine_fetch(api.RBUF_SCENE_POINTS)
    nrm = det.refine_fetch(api.RBUF_SCENE_NORMALS)
    lab = det.refine_fetch(api.RBUF_SCENE_LABELS)
    sizes = det.refine_fetch(api.RBUF_CLUSTER_SIZES)
    assert len(pts) == len(sc.xyz)  # same voxels survive normals_not_nan
    np.testing.assert_allclose(pts[:, :3], sc.xyz, atol=2e-6)
    rgb = pts[:, 3].copy().view(np.uint32)
    assert np.array_equal(np.stack([rgb & 255, (rgb >> 8) & 255, (rgb >> 16) & 255], 1).astype(np.uint8), sc.rgb)
    flat = sc.curvature < 0.02
    assert np.percentile(_angle(nrm[flat][:, :3], sc.normals[flat]), 99) < 2e-3
    assert np.all(np.sum(nrm[:, :3] * -pts[:, :3], 1) >= -1e-6)  # every normal faces the camera
    np.testing.assert_allclose(nrm[flat][:, 3], sc.curvature[flat], atol=2e-4)
    # clusters: the same partition up to points whose join test sits on the eps_angle / curvature threshold
    assert abs(len(sizes) - len(sc.cluster_sizes)) <= max(2, len(sc.cluster_sizes) // 10)
    both = (lab >= 0) & (sc.cluster >= 0)
    assert both.mean() > 0.9 * (sc.cluster >= 0).mean()
    pairs = set(zip(lab[both].tolist(), sc.cluster[both].tolist()))
    big = [c for c in range(len(sc.cluster_sizes)) if sc.cluster_sizes[c] >= 50]
    for c in big:  # every large oracle cluster is (almost) one GPU cluster
        ids, cnt = np.unique(lab[sc.cluster == c], return_counts=True)
        assert cnt.max() >= 0.95 * sc.cluster_sizes[c], (c, ids, cnt)
    assert len(pairs) <= len(sizes) + len(sc.cluster_sizes)


@pytest.mark.gpu
def test_gpu_icp_and_scores_match_the_oracle(scene_case, gpu_case):
    case, hyps, dets = scene_case, gpu_case["hyps"], gpu_case["dets"]
    ref = R.refine_frame(case["scene"], case["models"], case["p"], hyps)
    assert len(dets) == len(hyps)
    for i, d in enumerate(dets):
        ev = ref["evals"][i]
        assert bool(d["icp_converged"]) == bool(ref["converged"][i])
        m = case["models"][int(hyps[i]["cls"])]
        # poses: compared where they act, on the model's points (1 mm; the north star's translation tolerance)
        a = R._transform(m.xyz, d["pose"].reshape(4, 4))
        b = R._transform(m.xyz, ref["poses"][i])
        assert np.abs(a - b).max() < 1e-3, (i, np.abs(a - b).max())
        # scores: the poses differ by a fraction of a voxel, so a few boundary points may change sides
        assert abs(int(d["visible"]) - ev.visible) <= max(3, 0.01 * ev.visible)
        assert abs(int(d["inliers"]) - ev.inliers) <= max(3, 0.01 * max(ev.inliers, 1))
        if ev.inliers > 0:
            assert abs(d["similarity"] - ev.similarity_score) < 5e-3
            assert abs(d["inliers_ratio"] - ev.inliers_ratio) < 1e-2
            assert abs(d["clutter"] - ev.clutter_score) < 3e-2
            assert abs(d["final_score"] - ev.final_score) < 0.1
        if abs(ev.final_score - case["p"].final_score_threshold) > 0.2 and abs(ev.inliers_ratio - case["p"].inliers_threshold) > 0.02 \
                and abs(ev.clutter_score - case["p"].clutter_threshold) > 0.05:
            assert bool(d["accepted"]) == ev.accepted, i
    acc = [i for i, d in enumerate(dets) if d["accepted"]]
    assert sorted(acc) == sorted(ref["accepted"])
    chosen_ref = sorted(ref["accepted"][i] for i in ref["chosen"])
    chosen_gpu = sorted(int(d["hypothesis"]) for d in dets if d["selected"])
    # the same objects are reported; which of two near-identical poses of one object wins may differ by a score ulp
    assert sorted(int(hyps[i]["cls"]) for i in chosen_gpu) == sorted(int(hyps[i]["cls"]) for i in chosen_ref)
    ranks = sorted(int(d["rank"]) for d in dets if d["rank"] >= 0)
    assert ranks == list(range(len(ranks))) and len(ranks) <= len(chosen_gpu)


@pytest.mark.gpu
def test_gpu_scoring_at_the_oracle_pose_is_tight(scene_case, gpu_case):
    """With ICP out of the way (0 iterations allowed -> not converged -> the pose is the input pose) scoring sees identical
    poses on both sides: counts equal, scores to float accuracy."""
    case, det, api = scene_case, gpu_case["det"], gpu_case["api"]
    hyps = make_hyps(case, seeds=(1,), wrong=False)
    for k in range(len(case["models"])):
        hyps[k]["pose"] = truth_pose(case["truth"], k).reshape(-1)
    p = det.refine_params()
    for k, (x, c) in enumerate(case["clouds"]):
        det.set_object_model(k, x, c, -1.0, 30)  # no object radius: ICP finds no correspondences
    dets = det.refine(hyps)
    models = [R.ObjectModel(x, c, case["p"], nn_search_radius=-1.0, icp_iterations=30) for x, c in case["clouds"]]
    for i, d in enumerate(dets):
        assert not d["icp_converged"]
        np.testing.assert_array_equal(d["pose"], hyps[i]["pose"])
        ev = R.evaluate_hypothesis(case["scene"], models[i], case["p"], hyps[i]["pose"].reshape(4, 4), hyps[i]["loc_score"],
                                   (hyps[i]["yawpitch_score"] + hyps[i]["roll_score"]) / 2)
        assert int(d["visible"]) == ev.visible and int(d["inliers"]) == ev.inliers
        assert int(d["explained"]) == int(ev.explained.sum())
        assert abs(d["similarity"] - ev.similarity_score) < 1e-3  # normals agree to ~1e-3 rad away from edges, less on them
        assert abs(d["inliers_ratio"] - ev.inliers_ratio) < 1e-6
        assert abs(d["clutter"] - ev.clutter_score) < 2e-2       # cluster borders (see the scene test)
    for k, (x, c) in enumerate(case["clouds"]):
        det.set_object_model(k, x, c, 0.015, 30)
    assert p.default_icp_iterations == 60


@pytest.mark.gpu
def test_gpu_refine_edge_cases(scene_case, gpu_case):
    det, api = gpu_case["det"], gpu_case["api"]
    from oracle.oracle import HYP_DTYPE
    assert len(det.refine(np.zeros(0, HYP_DTYPE))) == 0  # no hypotheses: the scene is still prepared
    far = make_hyps(scene_case, seeds=(1,), wrong=False)[:1]
    pose = far[0]["pose"].reshape(4, 4).copy()
    pose[2, 3] = 1.6  # `if (h.rotmat(2,3) > 1.5f) return false`
    far[0]["pose"] = pose.reshape(-1)
    d = det.refine(far)[0]
    assert not d["accepted"] and d["rank"] == -1
    bad = far.copy()
    bad[0]["cls"] = 7
    with pytest.raises(api.Hf6dError):
        det.refine(bad)
    # an empty frame: no scene points, ICP cannot converge, nothing is accepted
    det.upload(0, np.full_like(scene_case["bgr"], 255), np.zeros_like(scene_case["depth"]))
    det.sync(0)
    d = det.refine(make_hyps(scene_case, seeds=(1,), wrong=False))
    assert not d["accepted"].any() and not d["icp_converged"].any()
    assert len(det.refine_fetch(api.RBUF_SCENE_POINTS)) == 0
    det.upload(0, scene_case["bgr"], scene_case["depth"])
    det.sync(0)
