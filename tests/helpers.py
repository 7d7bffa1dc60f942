"""Shared workload construction for the tests (CPU and GPU).  The oracle is used here only as the checker / as the
source of calibration features for the synthetic forest."""
from __future__ import annotations

import os

import numpy as np

from object_detector_6d_b200 import synth
from oracle import oracle as O


def make_case(tmpdir: str, *, K=3, T=4, seed=1, max_depth=14, votes_per_leaf=8, n_objects=6, cam=None, stride=2,
              calib_patches=6000, fill_random=0, weights_seed=3, min_samples=2, prob_quantum=0):
    """Frame + encoder weights + forest (written to tmpdir) + oracle parameter block."""
    cam = cam or synth.Camera()
    bgr, depth = synth.render_frame(seed, cam, n_objects=n_objects)
    layers = synth.make_encoder_weights(weights_seed)
    p = O.default_params(W=cam.W, H=cam.H, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, stride=stride,
                         fill_random=fill_random, fill_seed=1234)
    locs = O.scan_centres(depth, p)
    # calibration batch: an evenly spaced subset of this frame's patches through the oracle encoder
    sel = np.linspace(0, len(locs) - 1, min(calib_patches, len(locs))).astype(np.int64)
    calib = O.encode(O.normalise(O.gather(bgr, depth, p, locs[sel])), layers)
    forest_dir = os.path.join(tmpdir, "forest")
    stats = synth.write_forest(forest_dir, calib, T=T, K=K, max_depth=max_depth, votes_per_leaf=votes_per_leaf,
                               seed=7 + seed, min_samples=min_samples, prob_quantum=prob_quantum)
    wpath = os.path.join(tmpdir, "weights.bin")
    synth.write_weights_raw(wpath, layers)
    return dict(bgr=bgr, depth=depth, layers=layers, params=p, forest_dir=forest_dir, weights=wpath, stats=stats,
                cam=cam, locs=locs)


def to_api_params(p):
    """Oracle Params -> api.Params (identical layout)."""
    import ctypes as C
    from object_detector_6d_b200 import api
    q = api.Params()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q
