"""Stage REFINE at BASELINE configs[1] size (640x480, six objects, the bench's scene family): the CUDA path against
oracle/refine.py on the whole frame -- the scene's 70 000 down-sampled points, every cluster, and two hypotheses per object
through ICP, scoring and the joint optimisation.  Tolerances as in tests/test_refine.py (the reference leaves this arithmetic to
PCL; parity is tolerance parity)."""
import tempfile

import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import refine as R
from tests import test_refine as TR

pytestmark = pytest.mark.gpu
CAM = synth.Camera()


def test_gpu_refine_matches_the_oracle_at_full_size():
    from object_detector_6d_b200 import api
    from tests.helpers import make_case, to_api_params
    bgr, depth, truth = synth.render_scene(2, TR.OBJECT_SEED, CAM, n_objects=6)
    clouds = synth.object_models(TR.OBJECT_SEED, 6)
    p = TR.refine_params(CAM)
    models = [R.ObjectModel(x, c, p, nn_search_radius=0.015, icp_iterations=20) for x, c in clouds]
    scene = R.Scene(bgr, depth, p)
    case = dict(bgr=bgr, depth=depth, truth=truth, clouds=clouds, p=p, models=models, scene=scene)
    hyps = TR.make_hyps(case, seeds=(1,), wrong=True)
    with tempfile.TemporaryDirectory() as d:
        cs = make_case(d, K=6, T=1, seed=5, max_depth=6, votes_per_leaf=2, cam=CAM, calib_patches=1500)
        det = api.Detector(cs["forest_dir"], cs["weights"], to_api_params(cs["params"]), device=0)
        for k, (x, c) in enumerate(clouds):
            det.set_object_model(k, x, c, 0.015, 20)
        det.upload(0, bgr, depth)
        det.sync(0)
        dets = det.refine(hyps)
        pts = det.refine_fetch(api.RBUF_SCENE_POINTS)
        nrm = det.refine_fetch(api.RBUF_SCENE_NORMALS)
        lab = det.refine_fetch(api.RBUF_SCENE_LABELS)
        sizes = det.refine_fetch(api.RBUF_CLUSTER_SIZES)
        ms = det.refine_ms()
        det.close()
    # scene: the same points survive, in PCL's order
    assert len(pts) == len(scene.xyz) > 50000
    np.testing.assert_allclose(pts[:, :3], scene.xyz, atol=2e-6)
    flat = scene.curvature < 0.02
    assert np.percentile(TR._angle(nrm[flat][:, :3], scene.normals[flat]), 99) < 2e-3
    assert abs(len(sizes) - len(scene.cluster_sizes)) <= max(2, len(scene.cluster_sizes) // 10)
    for c in [c for c in range(len(scene.cluster_sizes)) if scene.cluster_sizes[c] >= 100]:
        _, cnt = np.unique(lab[scene.cluster == c], return_counts=True)
        assert cnt.max() >= 0.95 * scene.cluster_sizes[c]
    # hypotheses: poses to 1 mm on the model's points, the same acceptances away from the thresholds, the same objects chosen
    ref = R.refine_frame(scene, models, p, hyps)
    for i, dd in enumerate(dets):
        m = models[int(hyps[i]["cls"])]
        a = R._transform(m.xyz, dd["pose"].reshape(4, 4))
        b = R._transform(m.xyz, ref["poses"][i])
        assert np.abs(a - b).max() < 1e-3, (i, np.abs(a - b).max())
        ev = ref["evals"][i]
        assert abs(int(dd["inliers"]) - ev.inliers) <= max(3, 0.01 * max(ev.inliers, 1))
        if abs(ev.final_score - p.final_score_threshold) > 0.2 and abs(ev.inliers_ratio - p.inliers_threshold) > 0.02 \
                and abs(ev.clutter_score - p.clutter_threshold) > 0.05:
            assert bool(dd["accepted"]) == ev.accepted, i
    chosen_ref = sorted(int(hyps[ref["accepted"][i]]["cls"]) for i in ref["chosen"])
    chosen_gpu = sorted(int(dd["cls"]) for dd in dets if dd["selected"])
    assert chosen_gpu == chosen_ref and len(chosen_gpu) >= 3
    print("refine ms at 640x480:", ms)
