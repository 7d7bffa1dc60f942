"""GPU parity at the sizes and forest shapes BASELINE.json names (configs[0], [1], [3], [4]) -- the per-stage tests in
test_gpu_parity.py run one small case; these run the whole path at full size.

For every configuration the CUDA path (through the C ABI) must agree with the oracle BIT FOR BIT on: patch centres,
quantised uint8 patches (a sample), leaf ordinals on the GPU's own features, Q16 vote maps, centre lists and the final
hypothesis tuples + pre-ICP poses (the oracle is fed the GPU's feature matrix).  Plus size-independent
properties at full size: P' = floor(P/100)*100, tree-shard linearity of the vote maps, slot-to-slot determinism.
"""
import numpy as np
import pytest

from object_detector_6d_b200 import synth
from tests.helpers import make_case, to_api_params

pytestmark = pytest.mark.gpu


def _check_case(cs, n_slots=2, sample=4000):
    from object_detector_6d_b200 import api
    from oracle import oracle as O
    det = api.Detector(cs["forest_dir"], cs["weights"], to_api_params(cs["params"]), device=0, n_slots=n_slots)
    try:
        det.set_debug_capture(True)
        forest = O.Forest(cs["forest_dir"])
        hyp = det.detect(cs["bgr"], cs["depth"])
        P, Pp = det.counts(0)
        locs = det.fetch(api.BUF_LOCS)
        feat = det.fetch(api.BUF_FEATURES)
        leaf = det.fetch(api.BUF_LEAF_ORD)
        maps = det.fetch(api.BUF_MAPS)
        q = det.fetch(api.BUF_PATCH_U8)
        # scan + P' rule (HFTest.cpp:426-433)
        ref_locs = O.scan_centres(cs["depth"], cs["params"])
        assert P == len(ref_locs) and Pp == (P // 100) * 100 and Pp > 0
        assert np.array_equal(locs[:P], ref_locs)
        # quantised patches on an evenly spaced sample (the oracle gather is single-threaded)
        sel = np.unique(np.linspace(0, Pp - 1, min(sample, Pp)).astype(np.int64))
        q_ref = O.normalise(O.gather(cs["bgr"], cs["depth"], cs["params"], ref_locs[sel]))
        assert np.array_equal(q[sel], q_ref)
        # traversal, votes, centres, pose: oracle on the GPU's own features
        _, ords = O.traverse(forest, feat)
        assert np.array_equal(leaf, ords)
        assert leaf.min() >= 0 and all(leaf[:, t].max() < forest.leaf_count(t) for t in range(forest.T))
        maps_ref, n_cast = O.vote(forest, leaf, locs[:Pp], cs["depth"], cs["params"])
        assert n_cast > 0 and np.array_equal(maps, maps_ref)
        hyp_ref, _, _ = O.detect(forest, cs["bgr"], cs["depth"], cs["params"], cs["layers"], features_override=feat)
        assert len(hyp_ref) == len(hyp) and len(hyp) > 0
        for name in hyp.dtype.names:
            assert np.array_equal(hyp[name], hyp_ref[name]), name
        # determinism: a second slot gives the same bits
        if n_slots > 1:
            det.upload(1, cs["bgr"], cs["depth"])
            det.run(1)
            hyp1 = det.collect(1)
            assert np.array_equal(det.fetch(api.BUF_MAPS, slot=1), maps)
            for name in hyp.dtype.names:
                assert np.array_equal(hyp[name], hyp1[name]), name
        return dict(P=P, Pp=Pp, n_cast=n_cast, n_hyp=len(hyp), maps=maps, feat=feat)
    finally:
        det.close()


def test_config_c1_one_object_segmented(tmp_path):
    """configs[0] (the reference's own CPU-runnable case): a 1-object forest, one object in the frame,
    `are_objects_segmented: true` -- samples without depth keep the constant fill (HFTest.cpp:1235, patch_extractor.cu:238-249),
    K = 1 exercises the single-class paths of the vote / centre / pose kernels (one map, 16 accumulator slots)."""
    cs = make_case(str(tmp_path), K=1, T=4, seed=4, max_depth=16, votes_per_leaf=12, n_objects=1, calib_patches=6000, fill_random=0)
    # segmented input: no depth (and black) outside an ellipse around the image centre, so patches on its rim mix measured
    # samples with the constant fill
    H, W = cs["depth"].shape
    yy, xx = np.mgrid[0:H, 0:W]
    outside = ((xx - W / 2) / (0.36 * W)) ** 2 + ((yy - H / 2) / (0.40 * H)) ** 2 > 1.0
    cs["depth"] = np.where(outside, 0, cs["depth"]).astype(np.uint16)
    cs["bgr"] = np.where(outside[..., None], 0, cs["bgr"]).astype(np.uint8)
    out = _check_case(cs)
    assert 5000 < out["Pp"] < 60000 and out["n_hyp"] > 0


def test_config_c2_six_objects_full_frame(tmp_path):
    """configs[1]: 6-object forest, 4 trees, depth ~20, 16 votes per leaf, 640x480 at stride 2."""
    cs = make_case(str(tmp_path), K=6, T=4, seed=2, max_depth=20, votes_per_leaf=16, calib_patches=8000, fill_random=1)
    out = _check_case(cs)
    assert out["Pp"] > 50000 and out["n_hyp"] > 100


def test_config_c2_tree_shard_linearity_full_size(tmp_path):
    """configs[2] property at full size: the maps of the tree shards sum to the unsharded maps, for 2 and 4 shards."""
    from object_detector_6d_b200 import api
    cs = make_case(str(tmp_path), K=6, T=4, seed=3, max_depth=18, votes_per_leaf=16, calib_patches=8000)
    det = api.Detector(cs["forest_dir"], cs["weights"], to_api_params(cs["params"]), device=0, n_slots=1)
    try:
        det.detect(cs["bgr"], cs["depth"])
        full = det.fetch(api.BUF_MAPS).astype(np.uint64)
        leaf_full = det.fetch(api.BUF_LEAF_ORD)
        for world in (2, 4):
            acc = np.zeros_like(full)
            leaf = np.full_like(leaf_full, -1)
            for rank in range(world):
                det.set_tree_shard(rank, world)
                det.upload(0, cs["bgr"], cs["depth"])
                det.run(0, api.STAGE_SCAN, api.STAGE_VOTE)
                det.sync(0)
                acc += det.fetch(api.BUF_MAPS).astype(np.uint64)
                leaf = np.maximum(leaf, det.fetch(api.BUF_LEAF_ORD))
            assert np.array_equal(acc, full), world
            assert np.array_equal(leaf, leaf_full), world
        det.set_tree_shard(0, 1)
    finally:
        det.close()


def test_config_c4_large_frame_stride2(tmp_path):
    """configs[3]: 1280x960 depth, fx = fy = 1150, stride 2 (~280k patches per frame)."""
    cam = synth.Camera(1280, 960, 1150.0, 1150.0, 639.5, 479.5)
    cs = make_case(str(tmp_path), K=6, T=4, seed=4, max_depth=16, votes_per_leaf=8, cam=cam, calib_patches=6000)
    out = _check_case(cs, n_slots=1, sample=3000)
    assert out["Pp"] > 200000


@pytest.mark.parametrize("T,depth,votes", [(10, 15, 1), (20, 20, 4), (40, 15, 64), (80, 25, 16)])
def test_config_c5_forest_sweep(tmp_path, T, depth, votes):
    """configs[4]: forest-scale sweep (10-80 trees, depth 15-25, 1-64 votes per leaf, 6 objects) on a 320x240 frame so the
    oracle stays quick; exercises the interleaved-descent variants of the traversal kernel and both lane-group widths of
    the pose kernels."""
    cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
    cs = make_case(str(tmp_path), K=6, T=T, seed=10 + T, max_depth=depth, votes_per_leaf=votes, cam=cam,
                   calib_patches=4000, min_samples=1 if depth >= 20 else 2)
    out = _check_case(cs, n_slots=1, sample=1500)
    assert out["n_cast"] > 0
