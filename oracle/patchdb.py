"""TEST INFRASTRUCTURE (the checker, never the product): an independent restatement, in Python, of the two on-disk formats of
SURVEY.md 8(f)4 -- the LMDB data file patch_generator writes (PatchGen/src/patch_generator.cpp:479-493, 562-580) and the
caffe::Datum records inside it -- and of the arithmetic around them:

  * `read_lmdb` / `write_lmdb`: lmdb 0.9.x `data.mdb`, format version 1, 4096-byte pages (meta pages 0 / 1, B+tree of
    branch / leaf / overflow pages).  Restated from lmdb's published layout (mdb.c: MDB_page, MDB_node, MDB_meta, MDB_db)
    separately from csrc/patchdb.hpp: struct.unpack over the byte layout, recursive walk, a different bulk-load shape
    (half-full leaves, every value above 1 KiB pushed to overflow pages, two committed transactions so that the current meta
    page is page 0) so that each side reads files the other would never have written itself.
    PARITY UNPINNED for the page layout: liblmdb is not in this image, so no file here was ever opened by it.
  * `datum_bytes` / `parse_datum`: caffe.proto's Datum through google.protobuf with a descriptor built at run time (the real
    protobuf serialiser: this part IS pinned).
  * `annotation`: patch_generator::get_yaw_pitch_roll_from_rot_mat + get_object_coords (patch_generator.cpp:20-56) in numpy.
  * `train_vectors`: train_patch_generator::generate_train_patches (train_patch_generator.cpp:60-150) with the oracle's fp32
    encoder: which entries are written (the batch in which the cursor ends is dropped) and the record layout.
"""
from __future__ import annotations

import os
import struct

import numpy as np

PAGE = 4096
HDR = 16
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 1, 2, 4, 8
F_BIGDATA = 1
MAGIC = 0xBEEFC0DE
INVALID = 0xFFFFFFFFFFFFFFFF
NODEMAX = (((PAGE - HDR) // 2) & ~1) - 2


# ------------------------------------------------------------------------------------------------ Datum via google.protobuf
_DATUM = None


def _datum_class():
    global _DATUM
    if _DATUM is None:
        from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

        fd = descriptor_pb2.FileDescriptorProto()
        fd.name = "hf6d_oracle_caffe_datum.proto"
        fd.package = "hf6d_oracle_caffe"
        fd.syntax = "proto2"
        m = fd.message_type.add()
        m.name = "Datum"
        T = descriptor_pb2.FieldDescriptorProto
        for name, num, typ, label in (("channels", 1, T.TYPE_INT32, T.LABEL_OPTIONAL), ("height", 2, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                      ("width", 3, T.TYPE_INT32, T.LABEL_OPTIONAL), ("data", 4, T.TYPE_BYTES, T.LABEL_OPTIONAL),
                                      ("label", 5, T.TYPE_INT32, T.LABEL_OPTIONAL), ("float_data", 6, T.TYPE_FLOAT, T.LABEL_REPEATED),
                                      ("encoded", 7, T.TYPE_BOOL, T.LABEL_OPTIONAL)):
            f = m.field.add()
            f.name, f.number, f.type, f.label = name, num, typ, label
        pool = descriptor_pool.DescriptorPool()
        pool.Add(fd)
        _DATUM = message_factory.GetMessageClass(pool.FindMessageTypeByName("hf6d_oracle_caffe.Datum"))
    return _DATUM


def datum_bytes(channels, height, width, data: bytes, label) -> bytes:
    """What the reference's datum.SerializeToString gives (patch_generator.cpp:466-485: channels, height, width, label, data)."""
    d = _datum_class()()
    d.channels, d.height, d.width, d.label = int(channels), int(height), int(width), int(label)
    d.data = bytes(data)
    return d.SerializeToString()


def parse_datum(buf: bytes):
    d = _datum_class()()
    d.ParseFromString(bytes(buf))
    return {"channels": d.channels, "height": d.height, "width": d.width, "label": d.label, "data": bytes(d.data),
            "float_data": list(d.float_data), "encoded": d.encoded}


# ------------------------------------------------------------------------------------------------ LMDB data file
def _meta(page: bytes):
    pgno, _pad, flags, _lo, _hi = struct.unpack_from("<QHHHH", page, 0)
    magic, version, _addr, mapsize = struct.unpack_from("<IIQQ", page, HDR)
    dbs = [struct.unpack_from("<IHHQQQQQ", page, HDR + 24 + 48 * i) for i in range(2)]
    last_pg, txnid = struct.unpack_from("<QQ", page, HDR + 24 + 96)
    return {"pgno": pgno, "flags": flags, "magic": magic, "version": version, "mapsize": mapsize, "free": dbs[0], "main": dbs[1],
            "last_pg": last_pg, "txnid": txnid}


def read_meta(folder):
    with open(os.path.join(folder, "data.mdb"), "rb") as f:
        raw = f.read(2 * PAGE)
    metas = [_meta(raw[i * PAGE:(i + 1) * PAGE]) for i in range(2)]
    ok = [m for m in metas if m["magic"] == MAGIC and m["flags"] & P_META]
    if not ok:
        raise ValueError("no LMDB meta page")
    return max(ok, key=lambda m: m["txnid"])


def read_lmdb(folder):
    """[(key bytes, value bytes)] of the main database in key order."""
    with open(os.path.join(folder, "data.mdb"), "rb") as f:
        raw = f.read()
    m = read_meta(folder)
    if m["version"] != 1 or m["free"][0] != PAGE:
        raise ValueError("unsupported LMDB version / page size")
    pad, flags, depth, branch, leaf, overflow, entries, root = m["main"]
    out = []

    def walk(pgno, level):
        page = raw[pgno * PAGE:(pgno + 1) * PAGE]
        no, _pad, pflags, lower, upper = struct.unpack_from("<QHHHH", page, 0)
        assert no == pgno and HDR <= lower <= upper <= PAGE, (no, pgno, lower, upper)
        n = (lower - HDR) // 2
        ptrs = struct.unpack_from("<%dH" % n, page, HDR)
        for i, off in enumerate(ptrs):
            lo, hi, nflags, ksize = struct.unpack_from("<HHHH", page, off)
            key = page[off + 8:off + 8 + ksize]
            if pflags & P_BRANCH:
                assert i > 0 or ksize == 0, "leftmost branch key must be empty"
                walk(lo | hi << 16 | nflags << 32, level + 1)
            else:
                assert pflags & P_LEAF and level + 1 == depth, (pflags, level, depth)
                size = lo | hi << 16
                if nflags & F_BIGDATA:
                    (ov,) = struct.unpack_from("<Q", page, off + 8 + ksize)
                    ono, _p, oflags, pages = struct.unpack_from("<QHHI", raw, ov * PAGE)
                    assert ono == ov and oflags & P_OVERFLOW and pages == (HDR - 1 + size) // PAGE + 1
                    out.append((key, raw[ov * PAGE + HDR:ov * PAGE + HDR + size]))
                else:
                    out.append((key, page[off + 8 + ksize:off + 8 + ksize + size]))

    if root != INVALID:
        walk(root, 0)
    assert len(out) == entries, (len(out), entries)
    assert all(out[i][0] < out[i + 1][0] for i in range(len(out) - 1)), "keys out of order"
    return out


def write_lmdb(folder, items, big_threshold=1024, leaf_fill=0.5):
    """A valid data.mdb with a deliberately different shape from csrc/patchdb.hpp's writer (see the module docstring)."""
    items = sorted(items)
    pages = {}
    nxt = [2]

    def alloc(n=1):
        p = nxt[0]
        nxt[0] += n
        return p

    def build_page(flags, nodes, pgno):
        page = bytearray(PAGE)
        lower, upper = HDR, PAGE
        for nd in nodes:
            upper -= (len(nd) + 1) & ~1
            page[upper:upper + len(nd)] = nd
            struct.pack_into("<H", page, lower, upper)
            lower += 2
        struct.pack_into("<QHHHH", page, 0, pgno, 0, flags, lower, upper)
        assert lower <= upper
        return bytes(page)

    n_overflow = 0
    leaves, cur, cur_first, used = [], [], None, 0
    for key, val in items:
        big = 8 + len(key) + len(val) > NODEMAX or len(val) > big_threshold
        if big:
            n = (HDR - 1 + len(val)) // PAGE + 1
            ov = alloc(n)
            blob = bytearray(n * PAGE)
            struct.pack_into("<QHHI", blob, 0, ov, 0, P_OVERFLOW, n)
            blob[HDR:HDR + len(val)] = val
            for i in range(n):
                pages[ov + i] = bytes(blob[i * PAGE:(i + 1) * PAGE])
            n_overflow += n
            nd = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, F_BIGDATA, len(key)) + key + struct.pack("<Q", ov)
        else:
            nd = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 0, len(key)) + key + val
        need = ((len(nd) + 1) & ~1) + 2
        if cur and used + need > (PAGE - HDR) * leaf_fill:
            leaves.append((cur_first, cur))
            cur, used = [], 0
        if not cur:
            cur_first = key
        cur.append(nd)
        used += need
    if cur:
        leaves.append((cur_first, cur))
    level = []
    for first, nodes in leaves:
        p = alloc()
        pages[p] = build_page(P_LEAF, nodes, p)
        level.append((first, p))
    depth, n_branch = (1 if level else 0), 0
    while len(level) > 1:
        up = []
        for i in range(0, len(level), 7):  # narrow branch pages: a deep tree from few entries
            group = level[i:i + 7]
            if len(group) == 1 and up:  # a branch page needs two children: borrow from the previous page
                prev_first, prev_p, prev_group = up.pop()
                group = [prev_group.pop()] + group
                pages[prev_p] = _branch(prev_group, prev_p, build_page)
                up.append((prev_first, prev_p, prev_group))
            p = alloc()
            n_branch += 1
            pages[p] = _branch(group, p, build_page)
            up.append((group[0][0], p, group))
        level = [(f, p) for f, p, _ in up]
        depth += 1
    root = level[0][1] if level else INVALID
    last = nxt[0] - 1

    def meta(pgno, txnid, main, last_pg):
        page = bytearray(PAGE)
        struct.pack_into("<QHHHH", page, 0, pgno, 0, P_META, 0, 0)
        struct.pack_into("<IIQQ", page, HDR, MAGIC, 1, 0, 1 << 40)
        struct.pack_into("<IHHQQQQQ", page, HDR + 24, PAGE, 0x08, 0, 0, 0, 0, 0, INVALID)
        struct.pack_into("<IHHQQQQQ", page, HDR + 72, *main)
        struct.pack_into("<QQ", page, HDR + 120, last_pg, txnid)
        return bytes(page)

    main = (0, 0, depth, n_branch, len(leaves), n_overflow, len(items), root)
    pages[1] = meta(1, 1, (0, 0, 0, 0, 0, 0, 0, INVALID), 1)   # an older, empty transaction
    pages[0] = meta(0, 2, main, max(last, 1))                  # the current one lives on page 0
    with open(os.path.join(folder, "data.mdb"), "wb") as f:
        for p in range(max(last, 1) + 1):
            f.write(pages.get(p, bytes(PAGE)))


def _branch(group, pgno, build_page):
    nodes = []
    for i, (first, child) in enumerate(group):
        key = b"" if i == 0 else first
        nodes.append(struct.pack("<HHHH", child & 0xFFFF, child >> 16 & 0xFFFF, child >> 32 & 0xFFFF, len(key)) + key)
    return build_page(P_BRANCH, nodes, pgno)


# ------------------------------------------------------------------------------------------------ annotation + training vectors
def annotation(W, H, x, y, depth_mm, pose, view_angle_deg=45.3105):
    """(yaw, pitch, roll, x, y, z) of patch_annotation_lmdb.txt (patch_generator.cpp:20-56, 497-511)."""
    P = np.asarray(pose, np.float32).reshape(4, 4)
    yaw = np.float32(np.arctan2(np.float64(P[1, 0]), np.float64(P[0, 0])))
    a = np.float32(np.sqrt(np.float64(np.float32(P[2, 1] * P[2, 1]) + np.float32(P[2, 2] * P[2, 2]))))
    pitch = np.float32(np.arctan2(np.float64(-P[2, 0]), np.float64(a)))
    roll = np.float32(np.arctan2(np.float64(P[2, 1]), np.float64(P[2, 2])))
    ang = np.float32(np.float32(np.float32(np.float32(view_angle_deg) / np.float32(180.0)) * np.float32(3.141592)) / np.float32(2.0))
    focal = np.float32(np.float32(H) / np.float32(2.0)) / np.float32(np.tan(np.float64(ang)))
    cx, cy = np.float32(W / 2.0 - 0.5), np.float32(H / 2.0 - 0.5)
    z = np.float32(depth_mm) / np.float32(1000.0)
    px = (np.float32(x) - cx) * z / focal
    py = (np.float32(y) - cy) * z / focal
    corr = np.diag([1.0, -1.0, -1.0, 1.0])
    obj = np.linalg.inv(corr @ P.astype(np.float64)) @ np.array([px, py, z, 1.0], np.float64)
    return np.array([yaw, pitch, roll, obj[0], obj[1], obj[2]], np.float32)


def written_entries(n_entries, batch_size):
    """train_patch_generator.cpp:74-106: a batch is written only if MDB_NEXT succeeded after each of its entries."""
    return batch_size * ((n_entries - 1) // batch_size) if n_entries > 0 else 0


def train_vector_file(K, F, objs, dofs, feats) -> bytes:
    """train_patch_generator.cpp:1-9, 68-70, 122-150."""
    out = [struct.pack("<ii", K, F)]
    for o, d, f in zip(objs, dofs, feats):
        out.append(struct.pack("<i", int(o)) + np.asarray(d, "<f4").tobytes() + np.asarray(f, "<f4").tobytes())
    return b"".join(out)
