"""End-to-end parity at BASELINE configs[1] size: the CUDA path from raw frame to hypotheses against the CPU oracle running
its OWN fp32 encoder (the reference's Caffe forward is fp32 sgemm, HoughForest/src/HFTest.cpp:585-596).

Every other GPU test isolates the encoder by feeding the oracle the GPU's feature matrix; this one does not.  It answers
the north star's acceptance clause -- leaf indices and final poses against the reference CPU path on identical inputs --
for the three encoder modes (bf16 operands, split bf16, fp16 operands), and writes the numbers to gpurun_out/r02_parity.json (committed under profiles/):

  leaf agreement          fraction of (patch, tree) pairs that reach the same leaf
  tuples reproduced       fraction of the oracle's hypothesis tuples (class, centre px, z cm, yaw/pitch/roll deg) that the
                          GPU run produces identically.  Poses are quantised to 1 px / 1 cm / 1 degree, so "translation
                          <= 1 mm, rotation <= 0.5 deg" holds for a hypothesis exactly when its tuple is identical.
  strongest per class     the hypothesis at rank 0 of every level (centre, yaw/pitch, roll) of each class

and, as the yardstick for what "equal to the fp32 reference" can mean when the reference's BLAS summation order is
unknown, the same numbers for an fp32 encoder that merely sums in a different order (numpy / BLAS on the host).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TUPLE = ("cls", "cx", "cy", "z", "yaw_deg", "pitch_deg", "roll_deg")
# stated bars (measured values are in profiles/r02_parity.json)
SPLIT_LEAF_AGREEMENT = 0.999      # mode 1: as good as an fp32 evaluation in another summation order
SPLIT_FEATURE_TOL = 1e-4          # |f - f_oracle|, sigmoid outputs in (0, 1)
BF16_LEAF_AGREEMENT = 0.85        # mode 0: bf16 operands (8-bit mantissa) move several % of the leaves


def _tuples(h):
    return {tuple(float(h[n][i]) if n == "z" else int(h[n][i]) for n in TUPLE) for i in range(len(h))}


def _strongest(h):
    """First hypothesis of every class: rank 0 centre, rank 0 yaw/pitch peak, rank 0 roll mode (lists are in rank order)."""
    out = {}
    for i in range(len(h)):
        out.setdefault(int(h["cls"][i]), tuple(float(h[n][i]) if n == "z" else int(h[n][i]) for n in TUPLE))
    return out


def _compare(name, feat, leaf, hyp, ref):
    err = np.abs(feat - ref["feat"])
    t_ref, t = _tuples(ref["hyp"]), _tuples(hyp)
    s_ref, s = _strongest(ref["hyp"]), _strongest(hyp)
    # centres: (class, cx, cy) triples
    c_ref = {x[:3] for x in t_ref}
    c = {x[:3] for x in t}
    # oracle hypotheses that have a GPU hypothesis of the same class within one bin of every coordinate
    # (1 px, 1 cm, 1 degree; angles modulo 360)
    g = np.array(sorted(t), np.float64).reshape(-1, 7)
    near = 0
    for x in t_ref:
        d = np.abs(g - np.array(x, np.float64))
        d[:, 4:] = np.minimum(d[:, 4:], 360.0 - d[:, 4:])
        near += bool(((d[:, 0] == 0) & (d[:, 1] <= 1) & (d[:, 2] <= 1) & (d[:, 3] <= 0.0101) & (d[:, 4:] <= 1).all(1)).any())
    return {
        "encoder": name,
        "feature_max_abs_err": float(err.max()), "feature_mean_abs_err": float(err.mean()),
        "leaf_agreement": float((leaf == ref["leaf"]).mean()),
        "patches_with_every_tree_equal": float((leaf == ref["leaf"]).all(1).mean()),
        "hypotheses_oracle": len(t_ref), "hypotheses": len(t),
        "tuples_reproduced": len(t_ref & t) / max(1, len(t_ref)),
        "tuples_within_one_bin": near / max(1, len(t_ref)),
        "centres_reproduced": len(c_ref & c) / max(1, len(c_ref)),
        "strongest_per_class_reproduced": sum(1 for k in s_ref if s.get(k) == s_ref[k]) / max(1, len(s_ref)),
    }


@pytest.fixture(scope="module")
def workload(tmp_path_factory):
    import bench
    from oracle import oracle as O
    d = str(tmp_path_factory.mktemp("e2e"))
    frames, layers, forest_dir, wpath, stats = bench.make_workload(d, 2)
    forest = O.Forest(forest_dir)
    p = O.default_params(fill_random=1, fill_seed=1)
    refs = []
    for bgr, depth in frames:
        locs = O.scan_centres(depth, p)
        Pp = (len(locs) // 100) * 100
        q = O.normalise(O.gather(bgr, depth, p, locs[:Pp]))
        feat = O.encode(q, layers)
        _, leaf = O.traverse(forest, feat)
        hyp, _, _ = O.detect(forest, bgr, depth, p, layers)
        refs.append(dict(q=q, feat=feat, leaf=leaf, hyp=hyp))
    return dict(frames=frames, layers=layers, forest_dir=forest_dir, weights=wpath, stats=stats, forest=forest, params=p, refs=refs)


def _gpu_run(workload, mode, storage=None):
    from object_detector_6d_b200 import api
    from tests.helpers import to_api_params
    det = api.Detector(workload["forest_dir"], workload["weights"], to_api_params(workload["params"]), device=0, n_slots=1)
    det.set_debug_capture(True)
    out = []
    try:
        det.set_encoder_mode(mode)
        assert det.encoder_mode() == mode
        if storage is not None:
            det.set_feature_storage(storage)
            assert det.feature_storage() == storage
        for bgr, depth in workload["frames"]:
            hyp = det.detect(bgr, depth)
            out.append(dict(hyp=hyp, feat=det.fetch(api.BUF_FEATURES), leaf=det.fetch(api.BUF_LEAF_ORD), q=det.fetch(api.BUF_PATCH_U8),
                            enc_ms=det.encoder_layer_ms(0)))
    finally:
        det.close()
    return out


@pytest.fixture(scope="module")
def report(workload):
    from oracle import oracle as O
    rep = {"workload": "BASELINE configs[1]: 6-object forest trained on labelled synthetic patches (T=4, depth<=20, ~16 votes per "
                       "leaf, coherent votes), 640x480, fill random; %d frames" % len(workload["frames"]), "forest": workload["stats"],
           "oracle": "oracle/hf6d_oracle.c with its own fp32 encoder (8 interleaved partial sums, expf sigmoid)", "frames": []}
    # modes 0 / 2 as they run by default (fp16 feature rows, hf6d_set_feature_storage 1) and with the reference's fp32 rows
    runs = {0: _gpu_run(workload, 0), 1: _gpu_run(workload, 1), 2: _gpu_run(workload, 2),
            3: _gpu_run(workload, 0, storage=0), 4: _gpu_run(workload, 2, storage=0)}
    rep["feature_storage"] = "gpu_bf16 / gpu_fp16: fp16 feature rows (the default); *_fp32_rows: hf6d_set_feature_storage(ctx, 0)"
    for i, ref in enumerate(workload["refs"]):
        row = {"patches": int(ref["feat"].shape[0])}
        for mode, name in ((0, "gpu_bf16"), (1, "gpu_split_bf16"), (2, "gpu_fp16"), (3, "gpu_bf16_fp32_rows"), (4, "gpu_fp16_fp32_rows")):
            g = runs[mode][i]
            assert np.array_equal(g["q"], ref["q"]), "quantised patches must be bit-exact before the encoder"
            row[name] = _compare(name, g["feat"], g["leaf"], g["hyp"], ref)
            row[name]["encoder_layer_ms"] = [float(x) for x in g["enc_ms"]]
        # yardstick: fp32 on the host in BLAS summation order, everything downstream by the oracle
        x = ref["q"].astype(np.float32) / np.float32(255.0)
        h = x
        for W, b in workload["layers"]:
            h = (np.float32(1.0) / (np.float32(1.0) + np.exp(-(h @ W.T + b)))).astype(np.float32)
        _, leaf = O.traverse(workload["forest"], h)
        bgr, depth = workload["frames"][i]
        hyp, _, _ = O.detect(workload["forest"], bgr, depth, workload["params"], workload["layers"], features_override=h)
        row["fp32_other_summation_order"] = _compare("fp32, numpy/BLAS summation order (host)", h, leaf, hyp, ref)
        rep["frames"].append(row)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "r02_parity.json"), "w") as f:
            json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))
    return rep


def test_split_encoder_matches_the_fp32_oracle_end_to_end(report):
    for row in report["frames"]:
        r = row["gpu_split_bf16"]
        assert r["feature_max_abs_err"] < SPLIT_FEATURE_TOL
        assert r["leaf_agreement"] >= SPLIT_LEAF_AGREEMENT
        assert r["strongest_per_class_reproduced"] == 1.0
        # no worse than what a different fp32 summation order does to the same frame (plus a small allowance)
        assert r["tuples_reproduced"] >= row["fp32_other_summation_order"]["tuples_reproduced"] - 0.02


def test_bf16_encoder_end_to_end_numbers_are_published(report):
    for row in report["frames"]:
        r = row["gpu_bf16"]
        assert r["leaf_agreement"] >= BF16_LEAF_AGREEMENT
        assert r["feature_max_abs_err"] < 3e-2
        assert r["tuples_within_one_bin"] > 0.5


def test_fp16_encoder_end_to_end_numbers_are_published(report):
    """Mode 2 (fp16 operands, the same kernel and rate as bf16): 11-bit significands, so it must sit between the two."""
    for row in report["frames"]:
        r, b = row["gpu_fp16"], row["gpu_bf16"]
        assert r["feature_max_abs_err"] < 4e-3 and r["feature_mean_abs_err"] < b["feature_mean_abs_err"] / 4
        assert r["leaf_agreement"] > b["leaf_agreement"]
        assert r["tuples_reproduced"] >= b["tuples_reproduced"]


def test_fp16_feature_rows_cost_no_parity(report):
    """Feature storage 1 (the default of modes 0 / 2) against storage 0 on the same frames: rounding the sigmoid once to fp16
    (<= 2.5e-4) sits two orders below mode 0's operand error, so leaves and tuples against the fp32 oracle must not move
    by more than noise."""
    for row in report["frames"]:
        for a, b in (("gpu_bf16", "gpu_bf16_fp32_rows"), ("gpu_fp16", "gpu_fp16_fp32_rows")):
            assert abs(row[a]["feature_max_abs_err"] - row[b]["feature_max_abs_err"]) < 5e-4
            assert row[a]["leaf_agreement"] > row[b]["leaf_agreement"] - 0.005
            assert row[a]["tuples_within_one_bin"] > row[b]["tuples_within_one_bin"] - 0.1


def test_split_mode_is_deterministic_and_switchable(workload):
    """Mode 1 then mode 0 on the same context: the bf16 result is the one a fresh context gives (nothing leaks)."""
    from object_detector_6d_b200 import api
    from tests.helpers import to_api_params
    bgr, depth = workload["frames"][0]
    det = api.Detector(workload["forest_dir"], workload["weights"], to_api_params(workload["params"]), device=0, n_slots=2)
    try:
        h0 = det.detect(bgr, depth)
        f0 = det.fetch(api.BUF_FEATURES)
        det.set_encoder_mode(1)
        h1a = det.detect(bgr, depth)
        f1a = det.fetch(api.BUF_FEATURES)
        t = det.submit(bgr, depth)
        t2 = det.submit(bgr, depth)
        h1b, h1c = det.wait(t), det.wait(t2)
        det.set_encoder_mode(0)
        h0b = det.detect(bgr, depth)
        f0b = det.fetch(api.BUF_FEATURES)
    finally:
        det.close()
    assert np.array_equal(f0, f0b) and not np.array_equal(f0, f1a)
    for a, b in ((h0, h0b), (h1a, h1b), (h1a, h1c)):
        assert len(a) == len(b) and all(np.array_equal(a[n], b[n]) for n in a.dtype.names)
