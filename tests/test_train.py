"""Forest training (HoughForest/src/HFTrain.cpp:14-1265).

CPU: the oracle (oracle/train.py) against known answers and the reference's file formats.  GPU: hf6d_train_forest through the C
ABI against the oracle with the same counter-based draws -- the trees must be the same FILES, byte for byte (every statistic
that decides a split is an integer or a double sum whose rounding to float both sides share; see the oracle header for the two
places where a near-tie could legitimately differ).
"""
import os
import struct

import numpy as np
import pytest

from oracle import train as T


def make_samples(n=4000, K=3, F=48, seed=0):
    """Feature vectors with class- and pose-dependent structure, so that all three objectives have something to find."""
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, K, n).astype(np.int32)
    dof = np.zeros((n, 6), np.float32)
    dof[:, :3] = rng.uniform(-np.pi, np.pi, (n, 3))
    dof[:, 3:] = rng.uniform(-0.1, 0.1, (n, 3))
    feat = rng.uniform(0, 1, (n, F)).astype(np.float32)
    feat[:, :8] += cls[:, None] * 0.35                      # class signal
    feat[:, 8:16] += dof[:, 3:4] * 4.0                      # location signal
    feat[:, 16:24] += np.cos(dof[:, 0:1]) * 0.4             # pose signal
    return cls, dof, feat.astype(np.float32)


def parse_tree(raw: bytes, K: int):
    """HFBase::loadNodeFromFile (HFBase.cpp:58-108) into nested dicts."""
    pos = [0]

    def rd(fmt):
        v = struct.unpack_from("<" + fmt, raw, pos[0])
        pos[0] += struct.calcsize("<" + fmt)
        return v

    def node():
        (leaf,) = rd("B")
        if leaf:
            (lid,) = rd("i")
            prob = rd(f"{K}f")
            votes = []
            for _ in range(K):
                (m,) = rd("i")
                votes.append(np.array(rd(f"{6 * m}f"), np.float32).reshape(m, 6))
            return dict(leaf=True, id=lid, prob=np.array(prob, np.float32), votes=votes)
        mm, f1, f2, thr = rd("iiif")
        left = node()
        right = node()
        return dict(leaf=False, test=(mm, f1, f2, thr), left=left, right=right)
    root = node()
    assert pos[0] == len(raw)
    return root


def leaves_of(nd):
    return [nd] if nd["leaf"] else leaves_of(nd["left"]) + leaves_of(nd["right"])


def depth_of(nd):
    return 0 if nd["leaf"] else 1 + max(depth_of(nd["left"]), depth_of(nd["right"]))


# ----------------------------------------------------------------------------------------------------------- CPU: the oracle
def test_rng_is_a_pure_function_and_spreads():
    a = T.rng_u64(1, 0, 3, 5, 7, T.DRAW_F1)
    assert a == T.rng_u64(1, 0, 3, 5, 7, T.DRAW_F1)
    vals = {T.rng_u64(1, t, l, n, q, d) for t in range(2) for l in range(3) for n in range(3) for q in range(3) for d in range(4)}
    assert len(vals) == 2 * 3 * 3 * 3 * 4
    assert T.mix64(0) == 0 and T.mix64(1) == 0x5692161D100B05E5  # splitmix64 finaliser, known answer


def test_training_file_round_trip(tmp_path):
    cls, dof, feat = make_samples(200, 3, 16)
    path = str(tmp_path / "patches.forest")
    T.write_patches_file(path, 3, cls, dof, feat)
    K, F, c2, d2, f2 = T.read_patches_file(path)
    assert (K, F) == (3, 16) and np.array_equal(c2, cls) and np.array_equal(d2, dof) and np.array_equal(f2, feat)
    assert os.path.getsize(path) == 8 + 200 * (4 + 24 + 64)  # HFTrain.cpp:21-60


def test_oracle_tree_is_consistent(tmp_path):
    cls, dof, feat = make_samples(3000, 3, 32)
    K = 3
    roots = T.train_forest(str(tmp_path), cls, dof, feat, K, trees=1, seed=3, min_samples=25, tests_per_node=6, thresholds_per_test=4)
    raw = (tmp_path / "tree0.dat").read_bytes()
    tree = parse_tree(raw, K)
    lv = leaves_of(tree)
    n_train = int(np.float32(2.0) / np.float32(3.0) * np.float32(3000))
    assert sum(sum(len(v) for v in l["votes"]) for l in lv) == n_train  # every training sample votes in exactly one leaf
    assert [l["id"] for l in lv] == list(range(len(lv)))
    assert depth_of(tree) >= 4
    order, _ = T.shuffle(3000, 3, 0)
    train = np.sort(order[:n_train])
    spc = np.bincount(cls[train], minlength=K)
    for l in lv:  # make_leafs' normalisation (HFTrain.cpp:163-171)
        cnt = np.array([len(v) for v in l["votes"]], np.float64)
        if cnt.sum() == 0:
            continue
        want = cnt / np.array([np.sum(spc[c] / spc * cnt) for c in range(K)])
        np.testing.assert_allclose(l["prob"], want, rtol=1e-5)
    # the samples really descend to the leaf that holds their vote
    def descend(nd, f):
        while not nd["leaf"]:
            mm, a, b, thr = nd["test"]
            v = np.float32(f[a] - f[b]) if mm == 0 else f[a]
            nd = nd["left"] if v < np.float32(thr) else nd["right"]
        return nd
    for i in train[:200]:
        l = descend(tree, feat[i])
        assert any(np.array_equal(v, dof[i]) for v in l["votes"][cls[i]])
    assert (tmp_path / "forest.txt").read_text().split() == ["1", "3", "32", "8", "0.005"]
    assert roots[0].leaf is False


def test_oracle_forest_loads_in_the_detection_oracle(tmp_path):
    """The trainer's files are what HFBase::loadForestFromFolder reads (oracle/hf6d_oracle.c, pinned to the reference's reader)."""
    from oracle import oracle as O
    cls, dof, feat = make_samples(1500, 2, 800, seed=4)
    T.train_forest(str(tmp_path), cls, dof, feat, 2, trees=2, seed=9, min_samples=40, tests_per_node=4, thresholds_per_test=3)
    forest = O.Forest(str(tmp_path))
    assert (forest.T, forest.K, forest.F) == (2, 2, 800)
    _, ords = O.traverse(forest, feat[:64])
    assert ords.shape == (64, 2) and (ords >= 0).all()


def test_shuffle_takes_two_thirds_without_repeats():
    order, n_train = T.shuffle(1000, 5, 1)
    assert n_train == 666 and len(set(order.tolist())) == 1000
    o2, _ = T.shuffle(1000, 5, 2)
    assert not np.array_equal(order, o2)  # every tree draws its own subset


# ----------------------------------------------------------------------------------------------------------- GPU vs oracle
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(n=4000, K=3, F=48, trees=2, seed=7, min_samples=20, tests_per_node=8, thresholds_per_test=5),
                                 dict(n=2500, K=6, F=800, trees=1, seed=11, min_samples=30, tests_per_node=30, thresholds_per_test=10),
                                 dict(n=5000, K=2, F=64, trees=1, seed=2, min_samples=5, tests_per_node=3, thresholds_per_test=2)])
def test_gpu_trainer_writes_the_oracle_forest(tmp_path, cfg):
    from object_detector_6d_b200 import api
    cls, dof, feat = make_samples(cfg["n"], cfg["K"], cfg["F"], seed=cfg["seed"])
    kw = {k: cfg[k] for k in ("trees", "seed", "min_samples", "tests_per_node", "thresholds_per_test")}
    ref_dir, gpu_dir = tmp_path / "ref", tmp_path / "gpu"
    T.train_forest(str(ref_dir), cls, dof, feat, cfg["K"], **kw)
    st = api.train_forest(str(gpu_dir), cls, dof, feat, K=cfg["K"], **kw)
    assert (gpu_dir / "forest.txt").read_text().split() == (ref_dir / "forest.txt").read_text().split()
    total_leaves = 0
    for t in range(cfg["trees"]):
        a, b = (gpu_dir / f"tree{t}.dat").read_bytes(), (ref_dir / f"tree{t}.dat").read_bytes()
        ta, tb = parse_tree(a, cfg["K"]), parse_tree(b, cfg["K"])
        total_leaves += len(leaves_of(ta))
        assert depth_of(ta) == depth_of(tb) and len(leaves_of(ta)) == len(leaves_of(tb))
        assert a == b, f"tree {t} differs from the oracle's"
    assert st.leaves == total_leaves and st.training_samples == int(np.float32(2.0) / np.float32(3.0) * np.float32(cfg["n"]))
    assert st.train_ms > 0


@pytest.mark.gpu
def test_gpu_trainer_from_the_training_vector_file_and_into_the_detector(tmp_path):
    """The reference's flow: train_patch_generator's file in, forest files out, `HoughForest --test` reads them."""
    from object_detector_6d_b200 import api, synth
    cls, dof, feat = make_samples(3000, 2, 800, seed=5)
    path = str(tmp_path / "patches.forest")
    T.write_patches_file(path, 2, cls, dof, feat)
    out = tmp_path / "forest"
    st = api.train_forest(str(out), input_file=path, trees=2, seed=3, min_samples=30, tests_per_node=5, thresholds_per_test=4)
    mem = tmp_path / "forest_mem"
    api.train_forest(str(mem), cls, dof, feat, K=2, trees=2, seed=3, min_samples=30, tests_per_node=5, thresholds_per_test=4)
    for t in range(2):
        assert (out / f"tree{t}.dat").read_bytes() == (mem / f"tree{t}.dat").read_bytes()
    info = api.ModelInfo()
    api._ck_host(api.load().hf6d_inspect_forest(str(out).encode(), info))
    assert (info.T, info.K, info.F) == (2, 2, 800) and info.n_leaves == st.leaves
    # and the detector traverses it: leaves of the training vectors are the leaves the trainer put them in
    wpath = str(tmp_path / "w.bin")
    synth.write_weights_raw(wpath, synth.make_encoder_weights(3))
    det = api.Detector(str(out), wpath, api.default_params(W=320, H=240, fx=287.5, fy=287.5, cx=159.5, cy=119.5), device=0)
    assert det.T == 2 and det.K == 2
    det.close()


@pytest.mark.gpu
def test_gpu_trainer_errors_are_reported(tmp_path):
    from object_detector_6d_b200 import api
    cls, dof, feat = make_samples(300, 2, 16)
    with pytest.raises(api.Hf6dError, match="at least one tree"):
        api.train_forest(str(tmp_path), cls, dof, feat, K=2, trees=0)
    bad = cls.copy()
    bad[5] = 7
    with pytest.raises(api.Hf6dError, match="class 7"):
        api.train_forest(str(tmp_path), bad, dof, feat, K=2)
    with pytest.raises(api.Hf6dError, match="Could not open"):
        api.train_forest(str(tmp_path), input_file=str(tmp_path / "absent.forest"))
    # a node whose samples are all alike cannot split: one leaf, every vote in it
    same = np.ones((90, 16), np.float32)
    api.train_forest(str(tmp_path / "flat"), np.zeros(90, np.int32), dof[:90], same, K=1, trees=1)
    tree = parse_tree((tmp_path / "flat" / "tree0.dat").read_bytes(), 1)
    assert tree["leaf"] and len(tree["votes"][0]) == 60
