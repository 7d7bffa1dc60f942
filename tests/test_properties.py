"""Property tests (hypothesis) on the CPU side: the C oracle against the independent numpy / pure-Python restatement on
generated inputs that concentrate on the reference's quirks -- NMS ties and loop bounds (HFTest.cpp:219-268), the
NaN -> 0 quantisation of flat patches (HFTest.cpp:500-570), float -> int truncation toward zero in the vote projection
(HFTest.cpp:21-37), and the forest file format (HFBase.cpp:58-145)."""
import os
import struct
import tempfile

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O
from tests import npref

SET = dict(max_examples=40, deadline=None)


@settings(**SET)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 9), st.integers(1, 9), st.integers(2, 4), st.booleans())
def test_nms_equals_the_literal_two_deque_statement(seed, wx, wy, levels, sparse):
    """Few distinct values => many ties; windows from 1x1 up; emitted maxima are exactly the reference's."""
    rng = np.random.default_rng(seed)
    rows, cols = int(rng.integers(2 * wy, 2 * wy + 20)), int(rng.integers(wx, wx + 24))
    img = rng.integers(0, levels, (rows, cols)).astype(np.float32)
    if sparse:
        img *= rng.random((rows, cols)) < 0.15
    s, xs, ys = O.nms(img, wx, wy)
    ref = npref.nms(img, wx, wy)
    assert len(s) == len(ref)
    got = sorted(zip(s.tolist(), xs.tolist(), ys.tolist()), key=lambda t: (-t[0], t[1], t[2]))
    want = sorted(((float(a), int(x), int(y)) for a, x, y in ref), key=lambda t: (-t[0], t[1], t[2]))
    assert got == want
    for v, x, y in got:
        assert v != 0 and img[y, x] == v
        # the maximum of its window, and nothing equal before it in row-major order
        x0, y0 = x - wx // 2, y - wy // 2
        win = img[y0:y0 + wy, x0:x0 + wx]
        assert win.max() == v and np.argmax(win.ravel() == v) == (wy // 2) * wx + wx // 2
        assert y <= rows - 2 * wy + 1 + wy // 2  # the loop-bound quirk: bottom wy-1 window rows never produced


@settings(**SET)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from(["random", "flat", "flat_depth", "two_level", "tiny_var"]))
def test_normalise_quantise_edge_cases(seed, kind):
    rng = np.random.default_rng(seed)
    P = 7
    x = rng.random((P, 8, 8, 4)).astype(np.float32)
    if kind == "flat":
        x[:] = (rng.integers(0, 256, (P, 1, 1, 1)) / 256.0).astype(np.float32)  # 8 significant bits: see below
    elif kind == "flat_depth":
        x[..., 3] = np.float32(0.5)
    elif kind == "two_level":
        x = np.where(x > 0.5, np.float32(1), np.float32(0)).astype(np.float32)
    elif kind == "tiny_var":
        x = (np.float32(0.3) + x * np.float32(1e-4)).astype(np.float32)
    q = O.normalise(x)
    assert np.array_equal(q, npref.normalise(x))
    if kind == "flat":
        # x = m/256: x/64 is exact and the 64 partial sums k*x/64 are exact, so the depth plane has zero variance:
        # 0/0 = NaN -> (uchar) 0 on x86.  The colour planes divide by 192: their mean is only x up to rounding, so they
        # need not be NaN (the two statements still have to agree on them).
        assert not q[:, 192:].any()
    if kind == "flat_depth":
        assert not q[:, 192:].any() and q[:, :192].any()
    if kind == "random":
        assert q.min() >= 25 and q.max() <= 229             # [0.1, 0.9] * 255, truncated


def _write_random_tree(rng, path, K, F, depth):
    """A random tree in the reference's pre-order format; returns (number of leaves, bytes)."""
    out = bytearray()
    leaves = [0]

    def node(d):
        if d >= depth or (d > 0 and rng.random() < 0.3):
            out.extend(struct.pack("<B", 1))
            out.extend(struct.pack("<i", leaves[0]))
            leaves[0] += 1
            probs = rng.random(K).astype(np.float32)
            out.extend(probs.tobytes())
            for c in range(K):
                n = int(rng.integers(0, 4))
                out.extend(struct.pack("<i", n))
                out.extend(rng.normal(0, 0.5, (n, 6)).astype(np.float32).tobytes())
        else:
            out.extend(struct.pack("<B", 0))
            mode = int(rng.integers(0, 2))
            out.extend(struct.pack("<iiif", mode, int(rng.integers(0, F)), int(rng.integers(0, F)),
                                   float(np.float32(rng.normal(0, 0.3)))))
            node(d + 1)
            node(d + 1)

    node(0)
    with open(path, "wb") as f:
        f.write(out)
    return leaves[0]


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.integers(1, 4), st.integers(0, 6))
def test_forest_files_three_readers_and_traversal_agree(seed, T, K, depth):
    """Random forests (incl. a root that is a leaf): the oracle, the numpy reader and the product's loader agree on the
    structure, and oracle / numpy traversals agree on random features with NaNs."""
    from object_detector_6d_b200 import api
    rng = np.random.default_rng(seed)
    F = 16
    with tempfile.TemporaryDirectory() as d:
        n_leaves = [_write_random_tree(rng, os.path.join(d, f"tree{t}.dat"), K, F, depth) for t in range(T)]
        with open(os.path.join(d, "forest.txt"), "w") as f:
            f.write(f"{T} {K} {F} 8 0.005\n")
        fo = O.Forest(d)
        assert (fo.T, fo.K, fo.F) == (T, K, F) and [fo.leaf_count(t) for t in range(T)] == n_leaves
        mi = api.inspect_forest(d)
        assert (mi.T, mi.K, mi.F, mi.n_leaves) == (T, K, F, sum(n_leaves))
        assert mi.n_internal == sum(n - 1 for n in n_leaves)            # full binary trees
        feats = rng.normal(0, 0.4, (50, F)).astype(np.float32)
        feats[rng.random(feats.shape) < 0.05] = np.nan                   # NaN compares false -> right child
        _, ords = O.traverse(fo, feats)
        forest_np = npref.read_forest(d)
        assert np.array_equal(ords, npref.traverse(forest_np, feats))
        assert ords.min() >= 0 and all(ords[:, t].max() < n_leaves[t] for t in range(T))


@settings(**SET)
@given(st.floats(-3.0, 3.0, width=32), st.floats(-3.0, 3.0, width=32))
def test_projection_truncates_toward_zero_like_the_reference(dx, dy):
    """Point3DToImage (HFTest.cpp:21-37): u = (int)(x/z*fx + cx + 0.5f) -- values in (-1, 0) land in column/row 0."""
    fx, cx, z = np.float32(575.0), np.float32(0.25), np.float32(1.0)   # cx chosen so that small x straddle zero
    x = np.float32(dx / 575.0)
    u = np.float32(np.float32(np.float32(x / z) * fx) + cx) + np.float32(0.5)
    assert npref._f2i(u) == int(np.trunc(u))
    if -1 < u < 0:
        assert npref._f2i(u) == 0
