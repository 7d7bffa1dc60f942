"""GPU parity tests: every stage of libhf6d (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: integer / index work bit-exact -- patch centres, quantised patches, leaf assignment, Q16 vote maps,
centre lists, hypothesis tuples; floating point derived from integers (blurred maps, scores, poses) bit-exact too,
because both sides evaluate the same IEEE operations; the bf16 tensor-core encoder within a stated tolerance.
"""
import numpy as np
import pytest

from tests.helpers import make_case, to_api_params

pytestmark = pytest.mark.gpu

# bf16 operands (8-bit mantissa) and bf16 hidden activations through three sigmoid layers, vs the fp32 oracle:
# measured on B200 max 1.6e-2 / mean 8.4e-4 over 5.7e7 features
ENC_ABS_TOL = 3e-2
ENC_MEAN_TOL = 2e-3


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    d = str(tmp_path_factory.mktemp("case"))
    cs = make_case(d, K=3, T=4, seed=1, max_depth=14, votes_per_leaf=8)
    det = api.Detector(cs["forest_dir"], cs["weights"], to_api_params(cs["params"]), device=0, n_slots=2)
    det.set_debug_capture(True)
    cs["det"] = det
    cs["forest"] = O.Forest(cs["forest_dir"])
    yield cs
    det.close()


@pytest.fixture(scope="module")
def full_run(case):
    """One whole-frame run on the GPU, every intermediate fetched."""
    from object_detector_6d_b200 import api
    det = case["det"]
    det.upload(0, case["bgr"], case["depth"])
    det.run(0)
    hyp = det.collect(0)
    out = dict(hyp=hyp, counts=det.counts(0), locs=det.fetch(api.BUF_LOCS), q=det.fetch(api.BUF_PATCH_U8),
               feat=det.fetch(api.BUF_FEATURES), leaf=det.fetch(api.BUF_LEAF_ORD), maps=det.fetch(api.BUF_MAPS),
               blurred=det.fetch(api.BUF_BLURRED), centres=det.fetch(api.BUF_CENTRES), ms=det.stage_ms(0),
               launches=det.launch_count(0))
    return out


def test_library_is_native():
    from object_detector_6d_b200 import api
    L = api.load()
    for name in api.EXPORTS:
        assert hasattr(L, name), name


def test_scan_bitexact(case, full_run):
    from oracle import oracle as O
    locs = O.scan_centres(case["depth"], case["params"])
    P, Pp = full_run["counts"]
    assert P == len(locs) and Pp == (P // 100) * 100
    assert np.array_equal(full_run["locs"], locs)


@pytest.mark.parametrize("kernel", ["tiled", "direct"])
@pytest.mark.parametrize("fill_random", [0, 1])
def test_gather_normalise_bitexact(case, fill_random, kernel, monkeypatch):
    """uint8 CHW patches: software texture filter + sequential mean / variance + NaN->0 quantisation.  Both gather kernels:
    the shared-memory staged tiles (default) and the per-tap global loads (HF6D_GATHER=direct, also every tile's fallback)."""
    import ctypes as C
    monkeypatch.setenv("HF6D_GATHER", kernel)
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    p = O.Params()
    C.memmove(C.byref(p), C.byref(case["params"]), C.sizeof(p))
    p.fill_random = fill_random
    p.fill_seed = 99
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(p), device=0)
    det.set_debug_capture(True)
    det.upload(0, case["bgr"], case["depth"])
    det.run(0, api.STAGE_SCAN, api.STAGE_GATHER)
    q_gpu = det.fetch(api.BUF_PATCH_U8)
    det.close()
    locs = O.scan_centres(case["depth"], p)
    Pp = (len(locs) // 100) * 100
    q_ref = O.normalise(O.gather(case["bgr"], case["depth"], p, locs[:Pp]))
    assert q_gpu.shape == q_ref.shape
    bad = np.nonzero((q_gpu != q_ref).any(1))[0]
    assert bad.size == 0, f"{bad.size} of {Pp} patches differ, first {bad[:5]}"


def test_flat_frame_hits_nan_quantisation(case):
    """Constant colour + constant depth: variance 0 -> 0/0 = NaN -> (unsigned char)NaN == 0 (x86), HFTest.cpp:538-565."""
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    det = case["det"]
    bgr = np.full_like(case["bgr"], 200)
    depth = np.full_like(case["depth"], 800)
    det.upload(1, bgr, depth)
    det.run(1, api.STAGE_SCAN, api.STAGE_GATHER)
    q_gpu = det.fetch(api.BUF_PATCH_U8, slot=1)
    locs = O.scan_centres(depth, case["params"])
    Pp = (len(locs) // 100) * 100
    q_ref = O.normalise(O.gather(bgr, depth, case["params"], locs[:Pp]))
    assert Pp > 0 and (q_ref[:, 192:] == 0).all()  # depth channel: 0/0
    assert np.array_equal(q_gpu, q_ref)


def test_encoder_within_tolerance(case, full_run):
    from oracle import oracle as O
    feat_ref = O.encode(full_run["q"], case["layers"])
    err = np.abs(full_run["feat"] - feat_ref)
    print(f"encoder: max abs err {err.max():.3e}, mean {err.mean():.3e}")
    assert err.max() < ENC_ABS_TOL and err.mean() < ENC_MEAN_TOL


@pytest.mark.parametrize("variants,n3", [("1,1,1", "160"), ("2,2,2", "160"), ("3,3,3", "160"), ("4,0,4", "160"),
                                         ("5,4,0", "160"), ("2,2,2", "208"), ("1,1,1", "208")])
def test_encoder_kernel_variants_agree(case, full_run, variants, n3, monkeypatch):
    """Every row of HF6D_ENC_CONFIGS (stand-alone CTAs, CTA pairs, ring depths, epilogue shapes) is the same arithmetic in a
    different schedule: K is accumulated in the same order by the same MMA shape per output element, so the features
    must be BIT-identical to the default variant's."""
    from object_detector_6d_b200 import api
    monkeypatch.setenv("HF6D_ENC_VARIANT", variants)
    monkeypatch.setenv("HF6D_ENC_N3", n3)  # feature-layer tile width: 160 (also the stand-alone-CTA fallback) or 208
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=1)
    try:
        det.upload(0, case["bgr"], case["depth"])
        det.run(0, api.STAGE_SCAN, api.STAGE_ENCODE)
        feat = det.fetch(api.BUF_FEATURES)
    finally:
        det.close()
    assert feat.shape == full_run["feat"].shape
    assert np.array_equal(feat, full_run["feat"])


def test_latency_and_throughput_contexts_agree(case, full_run):
    """A one-slot context (the encoder kernels that are fastest alone) and a multi-slot one (smaller footprints, so that
    other frames' CTAs co-reside) must give the same frame result, bit for bit."""
    from object_detector_6d_b200 import api
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=1)
    try:
        hyp = det.detect(case["bgr"], case["depth"])
        feat = det.fetch(api.BUF_FEATURES)
    finally:
        det.close()
    assert np.array_equal(feat, full_run["feat"])
    assert len(hyp) == len(full_run["hyp"]) and all(np.array_equal(hyp[n], full_run["hyp"][n]) for n in hyp.dtype.names)


def test_traverse_bitexact_on_oracle_features(case, full_run):
    """Stage-isolated: the fp32 oracle features injected -> every (patch, tree) leaf identical."""
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    det = case["det"]
    feat_ref = O.encode(full_run["q"], case["layers"])
    det.upload(1, case["bgr"], case["depth"])
    det.run(1, api.STAGE_SCAN, api.STAGE_GATHER)
    det.inject(api.BUF_FEATURES, feat_ref, slot=1)
    det.run(1, api.STAGE_TRAVERSE, api.STAGE_TRAVERSE)
    leaf_gpu = det.fetch(api.BUF_LEAF_ORD, slot=1)
    _, ords = O.traverse(case["forest"], feat_ref)
    assert leaf_gpu.shape == ords.shape
    assert np.array_equal(leaf_gpu, ords), f"{(leaf_gpu != ords).sum()} of {ords.size} leaves differ"


def test_traverse_bitexact_on_gpu_features(case, full_run):
    from oracle import oracle as O
    _, ords = O.traverse(case["forest"], full_run["feat"])
    assert np.array_equal(full_run["leaf"], ords)


def test_leaf_agreement_end_to_end(case, full_run):
    """bf16 encoder vs fp32 oracle encoder: leaves cannot be bit-identical end to end; report and bound."""
    from oracle import oracle as O
    feat_ref = O.encode(full_run["q"], case["layers"])
    _, ords = O.traverse(case["forest"], feat_ref)
    agree = (full_run["leaf"] == ords).mean()
    print(f"end-to-end leaf agreement (bf16 encoder vs fp32 oracle): {agree * 100:.2f}%")
    assert agree > 0.90


def test_vote_maps_bitexact(case, full_run):
    from oracle import oracle as O
    Pp = full_run["counts"][1]
    maps_ref, n_cast = O.vote(case["forest"], full_run["leaf"], full_run["locs"][:Pp], case["depth"], case["params"])
    assert n_cast > 0
    assert np.array_equal(full_run["maps"], maps_ref)


def test_blur_bitexact(case, full_run):
    from oracle import oracle as O
    for c in range(case["forest"].K):
        ref = O.blur(full_run["maps"][c], 13, 13)
        assert np.array_equal(full_run["blurred"][c], ref), f"class {c}"


def test_centres_match_oracle_nms(case, full_run):
    from oracle import oracle as O
    for c in range(case["forest"].K):
        s, xs, ys = O.nms(full_run["blurred"][c], 40, 40)
        n = min(12, len(s))
        got = full_run["centres"][c]
        assert got["n"] == n
        assert np.array_equal(got["c"]["score"][:n], s[:n])
        assert np.array_equal(got["c"]["x"][:n], xs[:n])
        assert np.array_equal(got["c"]["y"][:n], ys[:n])


def _same_hyps(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for name in a.dtype.names:
        assert np.array_equal(a[name], b[name]), name


def test_hypotheses_match_oracle_on_gpu_features(case, full_run):
    """Everything after the encoder, end to end: the oracle is fed the GPU's own feature matrix."""
    from oracle import oracle as O
    hyp_ref, (P, Pp), _ = O.detect(case["forest"], case["bgr"], case["depth"], case["params"], case["layers"],
                                   features_override=full_run["feat"])
    assert (P, Pp) == full_run["counts"]
    assert len(hyp_ref) > 0
    _same_hyps(full_run["hyp"], hyp_ref)


def test_should_detect_and_max_loc(case, full_run):
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    det = case["det"]
    K = case["forest"].K
    sd = [1] * K
    sd[0] = 0
    ml = [5] * K
    det.set_objects(should_detect=sd, max_loc=ml)
    try:
        det.inject(api.BUF_FEATURES, full_run["feat"], slot=1)
        det.upload(1, case["bgr"], case["depth"])
        det.run(1, api.STAGE_SCAN, api.STAGE_GATHER)
        det.run(1, api.STAGE_TRAVERSE, api.STAGE_POSE)
        hyp = det.collect(1)
    finally:
        det.set_objects()
    hyp_ref, _, _ = O.detect(case["forest"], case["bgr"], case["depth"], case["params"], case["layers"],
                             should_detect=sd, max_loc=ml, features_override=full_run["feat"])
    assert (hyp["cls"] != 0).all()
    _same_hyps(hyp, hyp_ref)


def test_pipelined_submit_wait_is_deterministic(case, full_run):
    det = case["det"]
    t0 = det.submit(case["bgr"], case["depth"])
    t1 = det.submit(case["bgr"], case["depth"])
    h0, h1 = det.wait(t0), det.wait(t1)
    _same_hyps(h0, full_run["hyp"])
    _same_hyps(h1, full_run["hyp"])
    assert det.launch_count(0) >= 18


def test_tree_shards_sum_to_full_maps(case, full_run):
    """Tree sharding (one context per simulated rank): Q16 maps add up exactly, leaf tables merge by max."""
    from object_detector_6d_b200 import api
    det = case["det"]
    world = 2
    maps = np.zeros_like(full_run["maps"])
    leaf = np.full_like(full_run["leaf"], -1)
    try:
        for r in range(world):
            det.set_tree_shard(r, world)
            det.upload(1, case["bgr"], case["depth"])
            det.run(1, api.STAGE_SCAN, api.STAGE_VOTE)
            maps += det.fetch(api.BUF_MAPS, slot=1)
            part = det.fetch(api.BUF_LEAF_ORD, slot=1)
            assert (part[:, [t for t in range(det.T) if t % world != r]] == -1).all()
            leaf = np.maximum(leaf, part)
    finally:
        det.set_tree_shard(0, 1)
    assert np.array_equal(maps, full_run["maps"])
    assert np.array_equal(leaf, full_run["leaf"])


def test_empty_frame(case):
    det = case["det"]
    hyp = det.detect(case["bgr"], np.zeros_like(case["depth"]))
    assert len(hyp) == 0
    assert det.counts(0) == (0, 0)


def test_software_filter_equals_the_texture_unit(case):
    """the reference's gather is the hardware texture filter (patch_extractor.cu:339-343).  Rebuild that
    exact fetch with a cudaTextureObject_t on this GPU and compare it with the oracle's software filter (choice C1:
    fraction in 1.8 fixed point, border 0, (w00*T00 + w10*T10) + w01*T01 + w11*T11 in fp32)."""
    import ctypes as C
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    p = O.Params()
    C.memmove(C.byref(p), C.byref(case["params"]), C.sizeof(p))
    p.fill_random = 0
    det = case["det"]
    det.upload(1, case["bgr"], case["depth"])
    det.run(1, api.STAGE_SCAN, api.STAGE_SCAN)
    hw = det.texture_gather(slot=1)
    locs = O.scan_centres(case["depth"], p)
    Pp = (len(locs) // 100) * 100
    sw = O.gather(case["bgr"], case["depth"], p, locs[:Pp])
    assert hw.shape == sw.shape
    diff = np.abs(hw - sw)
    n_bad = int((hw != sw).sum())
    print(f"texture unit vs software filter: {n_bad} of {hw.size} values differ, max abs diff {diff.max():.3e}")
    # ps = 8: every sample fraction is a multiple of 1/8, exactly representable in the unit's 8 fractional bits, so
    # the only freedom left is the order of the fp32 blend; the quantised patches downstream must not move
    assert diff.max() <= 2e-6
    q_hw, q_sw = O.normalise(hw), O.normalise(sw)
    frac = float((q_hw != q_sw).mean())
    print(f"quantised uint8 values that differ between hardware and software filter: {frac * 100:.4f}%")
    assert frac < 1e-3


@pytest.mark.parametrize("cap", [None, 0, 777])
def test_enumeration_path_of_the_pose_stage_is_exact(case, full_run, cap, monkeypatch):
    """The pose stage has two implementations of its first pass: reading the vote stream the vote kernel wrote (default) and
    enumerating the votes again from the leaf table (sharded contexts, forests beyond the stream budget; HF6D_POSE_STREAM=0
    forces it).  The enumeration path lists window entries in a fixed-capacity buffer and accumulates in place what does not
    fit.  Neither the path nor the capacity may change a single bit of the result."""
    from object_detector_6d_b200 import api
    monkeypatch.setenv("HF6D_POSE_STREAM", "0")
    if cap is not None:
        monkeypatch.setenv("HF6D_ENTRY_CAP", str(cap))
    det = api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=1)
    try:
        hyp = det.detect(case["bgr"], case["depth"])
    finally:
        det.close()
    _same_hyps(hyp, full_run["hyp"])


def test_pose_rerun_and_replaced_leaf_table(case, full_run):
    """Stage-isolated runs around the vote stream: POSE alone twice (the stream and the zeroed counters are reused), then a
    leaf table injected after voting (the stream no longer describes the slot: the enumeration path must take over)."""
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    det = case["det"]
    det.upload(1, case["bgr"], case["depth"])
    det.run(1)
    for _ in range(2):
        det.run(1, api.STAGE_CENTRES, api.STAGE_POSE)
        _same_hyps(det.collect(1), full_run["hyp"])
    # another leaf table: tree 0's leaves of every patch shifted by one patch
    leaf = full_run["leaf"].copy()
    leaf[1:, 0] = full_run["leaf"][:-1, 0]
    det.inject(api.BUF_LEAF_ORD, leaf, slot=1)
    det.run(1, api.STAGE_VOTE, api.STAGE_CENTRES)
    det.inject(api.BUF_LEAF_ORD, leaf, slot=1)  # same table again, but injected AFTER the vote stage
    det.run(1, api.STAGE_POSE, api.STAGE_POSE)
    hyp = det.collect(1)
    Pp = full_run["counts"][1]
    ref = O.hypotheses(case["forest"], leaf, full_run["locs"][:Pp], case["depth"], case["params"])
    _same_hyps(hyp, ref)
    det.run(1, api.STAGE_VOTE, api.STAGE_POSE)  # and the stream path on the same table
    _same_hyps(det.collect(1), ref)


def test_class_shards_concatenate_to_the_full_list(case, full_run):
    """hf6d_set_class_shard: centres + pose for classes k % N == r only; the shards' lists merge to the unsharded one."""
    from object_detector_6d_b200 import api
    det = case["det"]
    try:
        for world in (2, 3):
            parts = []
            for rank in range(world):
                det.set_class_shard(rank, world)
                h = det.detect(case["bgr"], case["depth"])
                assert set(np.unique(h["cls"])) <= {k for k in range(det.K) if k % world == rank}
                parts.append(h)
            allh = np.concatenate(parts)
            allh = allh[np.argsort(allh["cls"], kind="stable")]
            _same_hyps(allh, full_run["hyp"])
    finally:
        det.set_class_shard(0, 1)


@pytest.mark.parametrize("seed,T,depth", [(1, 1, 0), (2, 3, 1), (3, 5, 7), (4, 9, 12)])
def test_traverse_random_forests_with_nans(case, seed, T, depth, tmp_path):
    """Stage-isolated traversal on random forest files (a root that is a leaf, odd depths, single-level trees -- the
    two-levels-per-record layout has to pad those) and features that contain NaNs (NaN compares false -> right)."""
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    from tests.test_properties import _write_random_tree
    rng = np.random.default_rng(seed)
    d = str(tmp_path)
    K, F = 2, 800
    for t in range(T):
        _write_random_tree(rng, f"{d}/tree{t}.dat", K, F, depth)
    with open(f"{d}/forest.txt", "w") as f:
        f.write(f"{T} {K} {F} 8 0.005\n")
    det = api.Detector(d, case["weights"], to_api_params(case["params"]), device=0, n_slots=1)
    try:
        det.upload(0, case["bgr"], case["depth"])
        det.run(0, api.STAGE_SCAN, api.STAGE_SCAN)
        det.sync(0)
        P, Pp = det.counts(0)
        feats = rng.normal(0, 0.4, (Pp, F)).astype(np.float32)
        feats[rng.random(feats.shape) < 0.02] = np.nan
        det.inject(api.BUF_FEATURES, feats)
        det.run(0, api.STAGE_TRAVERSE, api.STAGE_TRAVERSE)
        det.sync(0)
        _, ords = O.traverse(O.Forest(d), feats)
        assert np.array_equal(det.fetch(api.BUF_LEAF_ORD), ords)
    finally:
        det.close()
