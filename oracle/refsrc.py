"""ctypes binding of oracle/_ref/libhf6d_refsrc.so -- the reference's own sources compiled against stand-in headers
(oracle/build_ref.py, oracle/ref_driver.cpp).  TEST INFRASTRUCTURE ONLY: the pin the C oracle is checked against.

In the build container the library is (re)built from /root/reference on first use; on the GPU box, where the reference
tree does not exist, the prebuilt library that travelled with the snapshot is loaded.  available() says whether there is
one at all.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref
from . import oracle as O

_lib = None


class Hyp(C.Structure):
    _fields_ = [("obj_id", C.c_int32), ("row", C.c_int32), ("col", C.c_int32), ("z", C.c_float), ("yaw", C.c_float),
                ("pitch", C.c_float), ("roll", C.c_float), ("location_score", C.c_float), ("pose_score", C.c_float),
                ("rotmat", C.c_float * 16)]


HYP_DTYPE = np.dtype([("obj_id", "<i4"), ("row", "<i4"), ("col", "<i4"), ("z", "<f4"), ("yaw", "<f4"), ("pitch", "<f4"),
                      ("roll", "<f4"), ("location_score", "<f4"), ("pose_score", "<f4"), ("rotmat", "<f4", (16,))])
assert HYP_DTYPE.itemsize == C.sizeof(Hyp)


def available() -> bool:
    try:
        return lib() is not None
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build(os.environ.get("HF6D_REFERENCE", "/root/reference"))
        if not path:
            return None
        L = C.CDLL(path)
        L.hf6d_refsrc_create.restype = C.c_void_p
        L.hf6d_refsrc_create.argtypes = [C.c_char_p, C.c_char_p]
        L.hf6d_refsrc_destroy.argtypes = [C.c_void_p]
        L.hf6d_refsrc_forest_info.restype = C.c_float
        L.hf6d_refsrc_forest_info.argtypes = [C.c_void_p, C.c_void_p]
        L.hf6d_refsrc_set_encoder.argtypes = [C.c_void_p]
        L.hf6d_refsrc_get_leaves.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.hf6d_refsrc_tree_dump.restype = C.c_longlong
        L.hf6d_refsrc_tree_dump.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hf6d_refsrc_vote_pixels.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        L.hf6d_refsrc_nms.restype = C.c_int32
        L.hf6d_refsrc_nms.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32]
        L.hf6d_refsrc_extract_rgbd.restype = C.c_int32
        L.hf6d_refsrc_extract_rgbd.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(O.Params), C.c_void_p, C.c_void_p, C.c_int32]
        L.hf6d_refsrc_normals.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
        L.hf6d_refsrc_extract_normals.restype = C.c_int32
        L.hf6d_refsrc_extract_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(O.Params), C.c_void_p,
                                                  C.c_void_p, C.c_int32]
        L.hf6d_refsrc_test_image.restype = C.c_int32
        L.hf6d_refsrc_test_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(O.Params), C.c_void_p, C.c_void_p,
                                             C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        for name in ("hf6d_refsrc_captured_net_input", "hf6d_refsrc_captured_maps", "hf6d_refsrc_captured_blurred"):
            getattr(L, name).restype = C.c_longlong
            getattr(L, name).argtypes = [C.c_void_p, C.c_longlong]
        # the stand-in Caffe net forwards through the oracle's fp32 encoder (the one stage this library does not pin)
        enc = C.cast(O.lib().hf6d_ref_encode_f32, C.c_void_p)
        L.hf6d_refsrc_set_encoder(enc)
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Reference:
    """An HFTest object of the reference with a forest loaded by HFBase::loadForestFromFolder."""

    def __init__(self, forest_dir: str, weights_path: str = ""):
        self.h = lib().hf6d_refsrc_create(forest_dir.encode(), weights_path.encode())
        if not self.h:
            raise IOError(f"reference loader rejected {forest_dir}")
        info = np.zeros(4, np.int32)
        self.voxel_m = lib().hf6d_refsrc_forest_info(self.h, _p(info))
        self.T, self.K, self.F, self.patch_vox = (int(x) for x in info)

    def close(self):
        if self.h:
            lib().hf6d_refsrc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def get_leaves(self, features):
        """HFTest::get_leaf for every row x every tree -> leaf_id [P][T]."""
        features = np.ascontiguousarray(features, np.float32)
        out = np.zeros((features.shape[0], self.T), np.int32)
        lib().hf6d_refsrc_get_leaves(self.h, _p(features), features.shape[0], _p(out))
        return out

    def tree_dump(self, t: int):
        """What HFBase::loadNodeFromFile built for tree t: leaves in file order and the internal tests in pre-order."""
        n_int = C.c_int32(0)
        n = lib().hf6d_refsrc_tree_dump(self.h, t, None, None, None, None, 0, None, None, C.byref(n_int))
        ids = np.zeros(n, np.int32)
        probs = np.zeros((n, self.K), np.float32)
        counts = np.zeros((n, self.K), np.int32)
        lib().hf6d_refsrc_tree_dump(self.h, t, _p(ids), _p(probs), _p(counts), None, 0, None, None, C.byref(n_int))
        votes = np.zeros((int(counts.sum()), 6), np.float32)
        tests = np.zeros((n_int.value, 3), np.int32)
        thr = np.zeros(n_int.value, np.float32)
        lib().hf6d_refsrc_tree_dump(self.h, t, _p(ids), _p(probs), _p(counts), _p(votes), votes.size, _p(tests), _p(thr),
                                    C.byref(n_int))
        return dict(leaf_id=ids, class_prob=probs, vote_count=counts, votes=votes, tests=tests, thresholds=thr)

    def test_image(self, bgr, depth, p: O.Params, should_detect=None, max_loc=None, n_threads=1, capture=False, cap=4096):
        """HFTest::test_image.  Returns the hypotheses it hands to MeshUtils (pre-ICP), in call order."""
        sd = None if should_detect is None else np.ascontiguousarray(should_detect, np.uint8)
        ml = None if max_loc is None else np.ascontiguousarray(max_loc, np.int32)
        out = np.zeros(cap, HYP_DTYPE)
        n = lib().hf6d_refsrc_test_image(self.h, _p(np.ascontiguousarray(bgr)), _p(np.ascontiguousarray(depth)), C.byref(p),
                                         _p(sd), _p(ml), n_threads, 1 if capture else 0, _p(out), cap)
        return out[:min(n, cap)].copy()


def _captured(fn):
    n = fn(None, 0)
    out = np.zeros(n, np.float32)
    fn(_p(out), n)
    return out


def captured_net_input():
    return _captured(lib().hf6d_refsrc_captured_net_input)


def captured_maps():
    return _captured(lib().hf6d_refsrc_captured_maps)


def captured_blurred():
    return _captured(lib().hf6d_refsrc_captured_blurred)


def vote_pixels(dof6, px, py, depth_mm, intr=(575.0, 575.0, 319.5, 239.5)):
    dof6 = np.ascontiguousarray(dof6, np.float32)
    n = dof6.shape[0]
    px = np.ascontiguousarray(px, np.int32)
    py = np.ascontiguousarray(py, np.int32)
    dm = np.ascontiguousarray(depth_mm, np.uint16)
    c3 = np.zeros((n, 3), np.float32)
    uv = np.zeros((n, 2), np.int32)
    lib().hf6d_refsrc_vote_pixels(_p(dof6), n, _p(px), _p(py), _p(dm), _p(np.asarray(intr, np.float32)), _p(c3), _p(uv))
    return c3, uv


def nms(img, wx, wy, cap=8192):
    img = np.ascontiguousarray(img, np.float32)
    rows, cols = img.shape
    s = np.zeros(cap, np.float32)
    xs = np.zeros(cap, np.int32)
    ys = np.zeros(cap, np.int32)
    n = min(lib().hf6d_refsrc_nms(_p(img), rows, cols, wx, wy, _p(s), _p(xs), _p(ys), cap), cap)
    return s[:n].copy(), xs[:n].copy(), ys[:n].copy()


def extract_rgbd(bgr, depth, p: O.Params):
    cap = ((p.W + p.stride - 1) // p.stride) * ((p.H + p.stride - 1) // p.stride)
    patches = np.zeros((cap, p.patch_vox, p.patch_vox, 4), np.float32)
    locs = np.zeros((cap, 2), np.int32)
    n = lib().hf6d_refsrc_extract_rgbd(_p(np.ascontiguousarray(bgr)), _p(np.ascontiguousarray(depth)), C.byref(p),
                                       _p(patches), _p(locs), cap)
    return locs[:n].copy(), patches[:n].copy()


def normals(depth, focal=575.0):
    H, W = depth.shape
    out = np.zeros((H, W, 3), np.float32)
    lib().hf6d_refsrc_normals(_p(np.ascontiguousarray(depth)), W, H, C.c_float(focal), _p(out))
    return out


def extract_normals(bgr, depth, nrm, p: O.Params):
    cap = ((p.W + p.stride - 1) // p.stride) * ((p.H + p.stride - 1) // p.stride)
    patches = np.zeros((cap, p.patch_vox, p.patch_vox, 6), np.float32)
    locs = np.zeros((cap, 2), np.int32)
    n = lib().hf6d_refsrc_extract_normals(_p(np.ascontiguousarray(bgr)), _p(np.ascontiguousarray(depth)),
                                          _p(np.ascontiguousarray(nrm, np.float32)), C.byref(p), _p(patches), _p(locs), cap)
    return locs[:n].copy(), patches[:n].copy()
