"""The whole loop on the GPU: rendered training views -> patch features (the detector's own scan / gather / encode kernels) ->
training vectors (the reference's patches.forest format) -> hf6d_train_forest -> detection in an unseen frame -> ICP + scoring
-> poses against the ground truth.  What patch_generator + train_patch_generator + `HoughForest --train` + `HoughForest --test`
do in the reference (PatchGen/src/patch_generator.cpp, train_patch_generator.cpp:60-150, HoughForest/src/main.cpp:41-76).

  python tools/train_and_detect.py [--objects 3] [--views 8] [--trees 3] [--scale 1]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detector_6d_b200 import api, synth  # noqa: E402

OBJECT_SEED = 1000


def training_vectors(det, bgr, depth, truth, cam):
    """Labelled feature vectors of one view: every processed patch whose centre lies on an object (the reference's generator
    renders the objects alone, so all of its patches do), with the vote train_patch_generator stores for it."""
    det.upload(0, bgr, depth)
    det.run(0, api.STAGE_SCAN, api.STAGE_ENCODE)
    P, Pp = det.counts(0)
    locs = det.fetch(api.BUF_LOCS)[:Pp]
    feat = det.fetch(api.BUF_FEATURES)[:Pp]
    xs, ys = locs[:, 0], locs[:, 1]
    oid = truth["obj_id"][ys, xs].astype(np.int64)
    keep = oid >= 0
    xs, ys, oid, feat = xs[keep], ys[keep], oid[keep], feat[keep]
    z = depth[ys, xs].astype(np.float64) / 1000.0
    t = np.stack([(xs - cam.cx) * z / cam.fx, (ys - cam.cy) * z / cam.fy, z], 1)
    dof = np.zeros((len(xs), 6), np.float32)
    for k in range(truth["R"].shape[0]):
        m = oid == k
        if m.any():
            dof[m, :3] = synth.euler_from_rotation(truth["R"][k])
            dof[m, 3:] = (t[m] - truth["centre"][k]) @ truth["R"][k]  # the patch in the object frame (HFTest.cpp:41-102)
    return oid.astype(np.int32), dof, feat.astype(np.float32)


def run(objects=3, views=8, trees=3, scale=1, test_seed=500, verbose=True, tests_per_node=30, thresholds_per_test=10):
    cam = synth.Camera.scaled(scale) if scale != 1 else synth.Camera()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        layers = synth.make_encoder_weights(3)
        wpath = os.path.join(d, "weights.bin")
        synth.write_weights_raw(wpath, layers)
        # any forest lets a context run its encoder; this one is never traversed
        boot = os.path.join(d, "boot")
        synth.write_forest(boot, np.random.default_rng(0).uniform(0, 1, (64, 800)).astype(np.float32), T=1, K=objects, max_depth=2,
                           votes_per_leaf=1, seed=1)
        p = api.default_params(W=cam.W, H=cam.H, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, fill_random=0, batch_size=1)
        det = api.Detector(boot, wpath, p, device=0)
        t0 = time.time()
        parts = []
        for v in range(views):
            bgr, depth, truth = synth.render_scene(100 + v, OBJECT_SEED, cam, n_objects=objects)
            parts.append(training_vectors(det, bgr, depth, truth, cam))
        det.close()
        cls = np.concatenate([q[0] for q in parts])
        dof = np.concatenate([q[1] for q in parts])
        feat = np.concatenate([q[2] for q in parts])
        out["training_vectors"] = len(cls)
        out["feature_s"] = time.time() - t0
        forest = os.path.join(d, "forest")
        st = api.train_forest(forest, cls, dof, feat, K=objects, trees=trees, seed=1, tests_per_node=tests_per_node,
                              thresholds_per_test=thresholds_per_test)
        out.update(train_ms=st.train_ms, leaves=int(st.leaves), depth=int(st.max_depth))

        det = api.Detector(forest, wpath, api.default_params(W=cam.W, H=cam.H, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy,
                                                              fill_random=0), device=0)
        models = synth.object_models(OBJECT_SEED, objects)
        for k, (xyz, rgb) in enumerate(models):
            det.set_object_model(k, xyz, rgb, 0.015, 60)
        bgr, depth, truth = synth.render_scene(test_seed, OBJECT_SEED, cam, n_objects=objects)
        hyp = det.detect(bgr, depth)
        dets = det.refine(hyp)
        out["hypotheses"] = len(hyp)
        out["refine_ms"] = det.refine_ms()
        det.close()
        res = []
        for k in range(objects):
            mine = dets[(dets["cls"] == k) & (dets["rank"] >= 0)]
            c_true = truth["centre"][k]
            # Hough stage alone: the strongest centre of the class, back-projected
            hk = hyp[hyp["cls"] == k]
            hough_err = float(np.linalg.norm(hk[0]["pose"].reshape(4, 4)[:3, 3] - c_true)) if len(hk) else None
            if len(mine) == 0:
                res.append(dict(obj=k, found=False, hough_centre_err_m=hough_err))
                continue
            pose = mine[0]["pose"].reshape(4, 4)
            pts = models[k][0][::7]
            a = pts @ pose[:3, :3].T + pose[:3, 3]
            b = pts @ truth["R"][k].T.astype(np.float32) + truth["centre"][k].astype(np.float32)
            # distance of every posed model point to the nearest truly posed one: blind to the solids' symmetries
            from scipy.spatial import cKDTree
            add_s = float(np.mean(cKDTree(b).query(a)[0]))
            res.append(dict(obj=k, found=True, centre_err_m=float(np.linalg.norm(pose[:3, 3] - c_true)), add_s_m=add_s,
                            hough_centre_err_m=hough_err, final_score=float(mine[0]["final_score"]),
                            inliers_ratio=float(mine[0]["inliers_ratio"])))
        out["objects"] = res
    if verbose:
        print(f"{out['training_vectors']} training vectors from {views} views in {out['feature_s']:.2f} s; "
              f"{trees} trees in {out['train_ms']:.0f} ms ({out['leaves']} leaves, depth {out['depth']}); "
              f"{out['hypotheses']} hypotheses in the test frame; refine ms {out['refine_ms']}")
        for r in res:
            print("  ", r)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=3)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--trees", type=int, default=3)
    ap.add_argument("--scale", type=int, default=1)
    a = ap.parse_args()
    run(a.objects, a.views, a.trees, a.scale)
