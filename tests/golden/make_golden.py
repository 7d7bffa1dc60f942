"""Generates the committed fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

* cv_blur.npz   -- cv::blur itself (python cv2, OpenCV 4.x in this image; the reference used 2.4.10) on a seeded Q16
                   accumulator: the one third-party piece of the path that can be executed here, pinned as vectors.
* tiny_frame.npz -- a complete tiny case (160x120 frame, stride 8, K=2, T=3, P=266, P'=200) with every intermediate of the
                   CPU oracle: centres, quantised patches, fp32 features, leaf ordinals, Q16 vote maps, blurred maps and
                   hypothesis tuples.  The reference has no golden vectors, so these pin the ORACLE (a
                   change in oracle/hf6d_oracle.c or in the synthetic generators shows up as a diff here) and let the
                   GPU suite check the CUDA path against committed numbers without executing the oracle.
The encoder weights are regenerated from their seed (10.7 MB otherwise); their SHA-256 is stored and checked.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from object_detector_6d_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

TINY = dict(W=160, H=120, f=143.75, stride=8, K=2, T=3, max_depth=6, votes=3, frame_seed=11, weights_seed=3,
            forest_seed=21)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def weights_sha(layers) -> str:
    h = hashlib.sha256()
    for Wm, b in layers:
        h.update(np.ascontiguousarray(Wm).tobytes())
        h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()


def tiny_inputs():
    c = TINY
    cam = synth.Camera(c["W"], c["H"], c["f"], c["f"], c["W"] / 2 - 0.5, c["H"] / 2 - 0.5)
    bgr, depth = synth.render_frame(c["frame_seed"], cam, n_objects=3)
    # stride 8 on 160x120: 266 valid centres -> two batches of 100 are processed, the partial batch of 66 is dropped
    p = O.default_params(W=cam.W, H=cam.H, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, stride=c["stride"],
                         fill_random=1, fill_seed=4242)
    layers = synth.make_encoder_weights(c["weights_seed"])
    return cam, bgr, depth, p, layers


def make_tiny():
    c = TINY
    cam, bgr, depth, p, layers = tiny_inputs()
    locs = O.scan_centres(depth, p)
    Pp = (len(locs) // p.batch_size) * p.batch_size
    assert Pp == 200 and len(locs) > 200, (len(locs), Pp)
    patches = O.gather(bgr, depth, p, locs[:Pp])
    q = O.normalise(patches)
    feats = O.encode(q, layers)
    with tempfile.TemporaryDirectory() as d:
        fdir = os.path.join(d, "forest")
        synth.write_forest(fdir, feats, T=c["T"], K=c["K"], max_depth=c["max_depth"], votes_per_leaf=c["votes"],
                           seed=c["forest_seed"], min_samples=2)
        forest_files = {}
        for name in sorted(os.listdir(fdir)):
            forest_files[name] = np.frombuffer(open(os.path.join(fdir, name), "rb").read(), np.uint8)
        forest = O.Forest(fdir)
        ids, ords = O.traverse(forest, feats)
        maps, n_cast = O.vote(forest, ords, locs[:Pp], depth, p)
        blurred = np.stack([O.blur(maps[k], p.centers_blur_size, p.centers_blur_size) for k in range(c["K"])])
        hyp, counts, _ = O.detect(forest, bgr, depth, p, layers)
        hyp_f, _, _ = O.detect(forest, bgr, depth, p, layers, features_override=feats)
        assert len(hyp) == len(hyp_f) and all(np.array_equal(hyp[n], hyp_f[n]) for n in hyp.dtype.names)
    assert len(hyp) > 0, "tiny case produces no hypotheses; pick other seeds"
    nz = np.nonzero(maps)
    out = dict(
        bgr=bgr, depth=depth, params=np.frombuffer(bytes(p), np.uint8), weights_sha=np.array(weights_sha(layers)),
        locs=locs, q=q, features=feats, leaf_id=ids, leaf_ord=ords,
        maps_idx=np.stack(nz).astype(np.int32), maps_val=maps[nz], n_cast=np.int64(n_cast), blurred=blurred, hyp=hyp,
        forest_names=np.array(list(forest_files.keys())),
    )
    for name, data in forest_files.items():
        out["forest__" + name] = data
    np.savez_compressed(os.path.join(HERE, "tiny_frame.npz"), **out)
    print(f"tiny_frame.npz: P={len(locs)} P'={Pp} votes cast={n_cast} hypotheses={len(hyp)} "
          f"leaves={[forest.leaf_count(t) for t in range(forest.T)]}")


def make_cv_blur():
    import cv2
    rng = np.random.default_rng(1)
    acc = np.zeros((120, 160), np.uint64)
    ys, xs = rng.integers(0, 120, 3000), rng.integers(0, 160, 3000)
    np.add.at(acc, (ys, xs), rng.integers(32768, 65537, 3000).astype(np.uint64))
    img = (acc.astype(np.float64) / 65536.0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cv_blur.npz"), acc=acc, blur13=cv2.blur(img, (13, 13)),
                        blur35=cv2.blur(img, (35, 35)), cv_version=np.array(cv2.__version__))
    print("cv_blur.npz: OpenCV", cv2.__version__)


if __name__ == "__main__":
    make_cv_blur()
    make_tiny()
