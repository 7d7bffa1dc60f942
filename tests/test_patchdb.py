"""the LMDB patch database + training-vector interop (hf6d_patchdb_*, hf6d_patch_annotation,
hf6d_generate_train_vectors) against oracle/patchdb.py -- an independent restatement of the LMDB file layout, the real
protobuf serialiser for caffe::Datum, numpy for the annotation arithmetic."""
from __future__ import annotations

import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detector_6d_b200 import api, synth  # noqa: E402
from oracle import patchdb as O  # noqa: E402


def patch_items(n, ps=8, ch=4, objs=2, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        o = i * objs // n
        out.append(("%04d_%08d" % (o, i), rng.integers(0, 256, (ch, ps, ps), dtype=np.uint8), o))
    return out


@pytest.mark.parametrize("n,ps,ch", [(1, 8, 4), (13, 8, 4), (14, 8, 4), (700, 8, 4), (3000, 4, 4), (40, 16, 6), (9, 24, 4), (5, 40, 6)])
def test_written_database_is_read_by_the_independent_reader(tmp_path, n, ps, ch):
    """C++ writer -> Python reader: keys, Datum bytes (== google.protobuf's serialisation), tree statistics of the meta page.
    ps 24 / 40 push the values past LMDB's node limit onto overflow pages; 3000 entries need two branch levels' worth of leaves."""
    items = patch_items(n, ps, ch)
    db = api.PatchDb(str(tmp_path), "w")
    for k, d, o in items:
        db.put(k, d, o)
    assert db.entries() == n
    db.close()
    got = O.read_lmdb(str(tmp_path))
    assert [k for k, _ in got] == [k.encode() for k, _, _ in items]
    for (k, v), (_, d, o) in zip(got, items):
        assert v == O.datum_bytes(ch, ps, ps, d.tobytes(), o)
    m = O.read_meta(str(tmp_path))
    assert m["txnid"] == 1 and m["pgno"] == 1 and m["mapsize"] == 1 << 40 and m["free"][7] == O.INVALID
    pad, flags, depth, branch, leaf, overflow, entries, root = m["main"]
    size = os.path.getsize(tmp_path / "data.mdb")
    assert size == (m["last_pg"] + 1) * O.PAGE == (2 + branch + leaf + overflow) * O.PAGE and entries == n
    assert (overflow > 0) == (8 + 13 + len(O.datum_bytes(ch, ps, ps, bytes(ch * ps * ps), 0)) > O.NODEMAX)


@pytest.mark.parametrize("n,ps", [(1, 8), (50, 8), (2000, 8), (30, 24)])
def test_independently_written_database_is_read(tmp_path, n, ps):
    """Python writer (half-full leaves, narrow branch pages, big values on overflow pages, current meta on page 0 with an
    older transaction on page 1) -> C++ reader."""
    items = patch_items(n, ps)
    O.write_lmdb(str(tmp_path), [(k.encode(), O.datum_bytes(4, ps, ps, d.tobytes(), o)) for k, d, o in items])
    db = api.PatchDb(str(tmp_path), "r")
    assert db.entries() == n
    got = list(db)
    db.close()
    assert len(got) == n
    for (k, dims, data), (k0, d0, o0) in zip(got, items):
        assert k == k0 and dims == (4, ps, ps, o0) and np.array_equal(data, d0.reshape(-1))


def test_reader_takes_any_valid_datum(tmp_path):
    """Fields in another order, unknown fields, float_data packed and unpacked, `encoded`: still a Datum."""
    D = O._datum_class()
    d = D()
    d.channels, d.height, d.width, d.label, d.encoded = 4, 2, 2, 7, False
    d.data = bytes(range(16))
    d.float_data.extend([1.5, -2.0])
    odd = b"\x28\x07" + b"\x3d\x00\x00\xc0\x3f" + b"\x7a\x03abc" + d.SerializeToString()[:-2]  # label first, float_data as fixed32, field 15
    O.write_lmdb(str(tmp_path), [(b"0000_00000000", d.SerializeToString()), (b"0000_00000001", odd)])
    got = list(api.PatchDb(str(tmp_path), "r"))
    assert [g[1] for g in got] == [(4, 2, 2, 7)] * 2 and all(np.array_equal(g[2], np.arange(16)) for g in got)


def test_database_errors(tmp_path):
    with pytest.raises(api.Hf6dError, match="cannot open"):
        api.PatchDb(str(tmp_path / "absent"), "r")
    db = api.PatchDb(str(tmp_path), "w")
    db.put("0000_00000001", np.zeros((4, 8, 8), np.uint8), 0)
    with pytest.raises(api.Hf6dError, match="does not sort after"):
        db.put("0000_00000001", np.zeros((4, 8, 8), np.uint8), 0)
    with pytest.raises(api.Hf6dError, match="does not sort after"):
        db.put("0000_00000000", np.zeros((4, 8, 8), np.uint8), 0)
    db.close()
    with pytest.raises(api.Hf6dError, match="already exists"):  # the reference: "mdb_open failed. Does the lmdb already exist?"
        api.PatchDb(str(tmp_path), "w")
    raw = bytearray((tmp_path / "data.mdb").read_bytes())
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "data.mdb").write_bytes(raw[:5000])
    with pytest.raises(api.Hf6dError, match="outside the file|shorter"):
        list(api.PatchDb(str(bad), "r"))
    raw[16:20] = b"\0\0\0\0"
    raw[4096 + 16:4096 + 20] = b"\0\0\0\0"
    (bad / "data.mdb").write_bytes(raw)
    with pytest.raises(api.Hf6dError, match="no valid LMDB meta page"):
        api.PatchDb(str(bad), "r")
    empty = tmp_path / "empty"
    empty.mkdir()
    api.PatchDb(str(empty), "w").close()
    assert O.read_lmdb(str(empty)) == [] and list(api.PatchDb(str(empty), "r")) == []


def test_annotation_matches_the_numpy_restatement():
    rng = np.random.default_rng(3)
    for _ in range(200):
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = R
        pose[:3, 3] = rng.uniform(-0.3, 0.3, 3) + [0, 0, -0.7]
        x, y, d = int(rng.integers(0, 640)), int(rng.integers(0, 480)), int(rng.integers(300, 1500))
        a = api.patch_annotation(640, 480, x, y, d, pose)
        b = O.annotation(640, 480, x, y, d, pose)
        np.testing.assert_array_equal(a[:3], b[:3])
        np.testing.assert_allclose(a[3:], b[3:], rtol=0, atol=2e-7)


def test_train_vectors_fail_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    items = patch_items(5)
    db = api.PatchDb(str(tmp_path), "w")
    for k, d, o in items:
        db.put(k, d, o)
    db.close()
    (tmp_path / "patch_annotation_lmdb.txt").write_text("2\n" + "".join(f"{k} 0 0 0 0 0 0\n" for k, _, _ in items))
    layers = synth.make_encoder_weights(3)
    w = str(tmp_path / "w.bin")
    synth.write_weights_raw(w, layers)
    with pytest.raises(api.Hf6dError, match="no CUDA device"):
        api.generate_train_vectors(w, str(tmp_path), str(tmp_path / "out.forest"))
    assert O.written_entries(5, 1) == 4 and O.written_entries(5, 2) == 4 and O.written_entries(4, 2) == 2 and O.written_entries(1, 1) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 7])
def test_gpu_train_vectors_from_a_patch_database(tmp_path, batch):
    """hf6d_generate_train_vectors: LMDB (written by the INDEPENDENT writer) + annotation file -> the training-vector file;
    which entries are written and the record layout against the restatement; features against the fp32 oracle encoder."""
    from oracle import oracle as OR
    n = 333
    items = patch_items(n, objs=3, seed=5)
    O.write_lmdb(str(tmp_path), [(k.encode(), O.datum_bytes(4, 8, 8, d.tobytes(), o)) for k, d, o in items])
    rng = np.random.default_rng(9)
    dof = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    with open(tmp_path / "patch_annotation_lmdb.txt", "w") as f:
        f.write("3\n")
        for (k, _, _), a in zip(items, dof):
            f.write(k + " " + " ".join("%.6g" % v for v in a) + "\n")
    layers = synth.make_encoder_weights(3)
    w = str(tmp_path / "w.bin")
    synth.write_weights_raw(w, layers)
    out = str(tmp_path / "patches.forest")
    st = api.generate_train_vectors(w, str(tmp_path), out, batch_size=batch, encoder_mode=1)
    nw = O.written_entries(n, batch)
    assert (st.entries, st.written, st.classes, st.feature_length) == (n, nw, 3, 800)
    raw = open(out, "rb").read()
    assert len(raw) == 8 + nw * (4 + 24 + 3200) and struct.unpack("<ii", raw[:8]) == (3, 800)
    rec = np.frombuffer(raw[8:], np.uint8).reshape(nw, -1)
    objs = rec[:, :4].copy().view("<i4")[:, 0]
    got_dof = rec[:, 4:28].copy().view("<f4")
    feats = rec[:, 28:].copy().view("<f4")
    assert np.array_equal(objs, [o for _, _, o in items[:nw]])
    np.testing.assert_array_equal(got_dof, np.array([[np.float32("%.6g" % v) for v in a] for a in dof[:nw]], np.float32))
    q = np.stack([d.reshape(-1) for _, d, _ in items[:nw]])
    ref = OR.encode(q, layers)
    assert np.abs(feats - ref).max() < 1e-4  # split-bf16 mode: the fp32 encoder to ~3e-5
    # and through a detection context's own entry point, default bf16 mode
    p = api.default_params(patch_vox=8, voxel_m=0.005)
    ex = api.Detector(weights_path=w, params=p, extractor=True)
    f2 = ex.encode_patches(q)
    assert np.abs(f2 - ref).max() < 3e-2 and ex.patch_capacity() >= 70000
    ex.close()
