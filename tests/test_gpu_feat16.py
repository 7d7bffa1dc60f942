"""GPU tests of feature storage 1 (hf6d_set_feature_storage, the default; HF6D_FEATURES=fp32 / fp16): the feature layer of the bf16 / fp16 operand modes writes fp16 rows
and the traversal reads them -- half the HBM bytes of both kernels.  What must hold:

  * the stored value is the fp32 feature of storage 0 rounded once to fp16 (same accumulators, same sigmoid);
  * hf6d_fetch hands out the exact widening of the stored halves, and every later stage is bit-exact against the oracle fed
    those values (the bar of tests/test_gpu_parity.py, unchanged);
  * injected fp32 features are traversed as fp32 (the stage-isolated parity tests do not change meaning);
  * the split-bf16 mode (the near-fp32 accuracy mode) keeps fp32 rows.
"""
import numpy as np
import pytest

from tests.helpers import make_case, to_api_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    from oracle import oracle as O
    d = str(tmp_path_factory.mktemp("case16"))
    cs = make_case(d, K=3, T=4, seed=1, max_depth=14, votes_per_leaf=8)
    cs["forest"] = O.Forest(cs["forest_dir"])
    return cs


def _run(case, n_slots, storage, mode=0, monkeypatch=None):
    from object_detector_6d_b200 import api
    monkeypatch.setenv("HF6D_FEATURES", storage or "fp32")
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=n_slots)
    try:
        if mode:
            det.set_encoder_mode(mode)
        hyp = det.detect(case["bgr"], case["depth"])
        return dict(hyp=hyp, feat=det.fetch(api.BUF_FEATURES), leaf=det.fetch(api.BUF_LEAF_ORD), counts=det.counts(0))
    finally:
        det.close()


@pytest.mark.parametrize("n_slots", [1, 2])
@pytest.mark.parametrize("mode", [0, 2])
def test_fp16_rows_are_the_fp32_features_rounded_once(case, n_slots, mode, monkeypatch):
    a = _run(case, n_slots, None, mode, monkeypatch)
    b = _run(case, n_slots, "fp16", mode, monkeypatch)
    assert a["feat"].shape == b["feat"].shape and a["feat"].size > 0
    assert np.array_equal(b["feat"], a["feat"].astype(np.float16).astype(np.float32))
    agree = (a["leaf"] == b["leaf"]).mean()
    print(f"mode {mode}, {n_slots} slot(s): leaves equal between fp32 and fp16 feature rows: {agree * 100:.3f} %")
    assert agree > 0.98


@pytest.mark.parametrize("n_slots", [1, 2])
def test_every_later_stage_is_bit_exact_on_the_stored_values(case, n_slots, monkeypatch):
    from oracle import oracle as O
    r = _run(case, n_slots, "fp16", 0, monkeypatch)
    _, ords = O.traverse(case["forest"], r["feat"])
    assert np.array_equal(r["leaf"], ords), f"{(r['leaf'] != ords).sum()} of {ords.size} leaves differ"
    hyp_ref, (P, Pp), _ = O.detect(case["forest"], case["bgr"], case["depth"], case["params"], case["layers"],
                                   features_override=r["feat"])
    assert (P, Pp) == r["counts"] and len(hyp_ref) == len(r["hyp"]) > 0
    for name in hyp_ref.dtype.names:
        assert np.array_equal(r["hyp"][name], hyp_ref[name]), name


def test_injected_fp32_features_are_traversed_as_fp32(case, monkeypatch):
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    monkeypatch.setenv("HF6D_FEATURES", "fp16")
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=1)
    try:
        det.set_debug_capture(True)
        det.upload(0, case["bgr"], case["depth"])
        det.run(0, api.STAGE_SCAN, api.STAGE_ENCODE)
        q = det.fetch(api.BUF_PATCH_U8)
        feat_ref = O.encode(q, case["layers"])  # fp32 values that are not representable in fp16
        det.inject(api.BUF_FEATURES, feat_ref)
        det.run(0, api.STAGE_TRAVERSE, api.STAGE_TRAVERSE)
        leaf = det.fetch(api.BUF_LEAF_ORD)
        assert np.array_equal(det.fetch(api.BUF_FEATURES), feat_ref)
        # and the next encoded frame goes back to the fp16 rows
        det.run(0, api.STAGE_ENCODE, api.STAGE_TRAVERSE)
        feat2, leaf2 = det.fetch(api.BUF_FEATURES), det.fetch(api.BUF_LEAF_ORD)
    finally:
        det.close()
    _, ords = O.traverse(case["forest"], feat_ref)
    assert np.array_equal(leaf, ords)
    assert np.array_equal(feat2, feat2.astype(np.float16).astype(np.float32))
    _, ords2 = O.traverse(case["forest"], feat2)
    assert np.array_equal(leaf2, ords2)


def test_split_mode_keeps_fp32_rows(case, monkeypatch):
    a = _run(case, 1, None, 1, monkeypatch)
    b = _run(case, 1, "fp16", 1, monkeypatch)
    assert np.array_equal(a["feat"], b["feat"]) and np.array_equal(a["leaf"], b["leaf"])
    assert not np.array_equal(b["feat"], b["feat"].astype(np.float16).astype(np.float32))


def test_storage_is_selectable_at_run_time(case, monkeypatch):
    from object_detector_6d_b200 import api
    monkeypatch.delenv("HF6D_FEATURES", raising=False)
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=2)
    try:
        assert det.feature_storage() == 1  # the default where the feature layer has the kernel (F = 800)
        h1 = det.detect(case["bgr"], case["depth"])
        f1 = det.fetch(api.BUF_FEATURES)
        det.set_feature_storage(0)
        assert det.feature_storage() == 0
        h0 = det.detect(case["bgr"], case["depth"])
        f0 = det.fetch(api.BUF_FEATURES)
        det.set_feature_storage(1)
        t = det.submit(case["bgr"], case["depth"])
        h1b = det.wait(t)
        with pytest.raises(api.Hf6dError):
            det.set_feature_storage(2)
    finally:
        det.close()
    assert np.array_equal(f1, f0.astype(np.float16).astype(np.float32)) and not np.array_equal(f1, f0)
    assert len(h1) == len(h1b) and all(np.array_equal(h1[n], h1b[n]) for n in h1.dtype.names)
    assert len(h0) > 0
