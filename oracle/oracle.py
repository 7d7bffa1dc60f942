"""ctypes binding of the CPU oracle (oracle/hf6d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product package
(object_detector_6d_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhf6d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hf6d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "hf6d_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("stride", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("patch_vox", C.c_int32), ("voxel_m", C.c_float),
                ("max_depth_range_m", C.c_float), ("distance_threshold_m", C.c_float), ("fill_random", C.c_int32),
                ("fill_seed", C.c_uint64), ("batch_size", C.c_int32), ("max_yaw_pitch_hypotheses", C.c_int32),
                ("max_roll_hypotheses", C.c_int32), ("min_location_score_ratio", C.c_float),
                ("min_yaw_pitch_drop_ratio", C.c_float), ("centers_blur_size", C.c_int32),
                ("centers_nms_wsize", C.c_int32), ("pose_blur_size", C.c_int32), ("pose_nms_wsize", C.c_int32),
                ("patch_mode", C.c_int32), ("normals_focal", C.c_float)]


class Hypothesis(C.Structure):
    _fields_ = [("cls", C.c_int32), ("cx", C.c_int32), ("cy", C.c_int32), ("z", C.c_float), ("yaw_deg", C.c_int32),
                ("pitch_deg", C.c_int32), ("roll_deg", C.c_int32), ("loc_score", C.c_float),
                ("yawpitch_score", C.c_float), ("roll_score", C.c_float), ("pose", C.c_float * 16)]


HYP_DTYPE = np.dtype([("cls", "<i4"), ("cx", "<i4"), ("cy", "<i4"), ("z", "<f4"), ("yaw_deg", "<i4"),
                      ("pitch_deg", "<i4"), ("roll_deg", "<i4"), ("loc_score", "<f4"), ("yawpitch_score", "<f4"),
                      ("roll_score", "<f4"), ("pose", "<f4", (16,))])
assert HYP_DTYPE.itemsize == C.sizeof(Hypothesis)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.hf6d_ref_forest_load.restype = C.c_void_p
        L.hf6d_ref_forest_load.argtypes = [C.c_char_p]
        L.hf6d_ref_forest_free.argtypes = [C.c_void_p]
        L.hf6d_ref_forest_info.restype = C.c_float
        L.hf6d_ref_forest_info.argtypes = [C.c_void_p, C.c_void_p]
        L.hf6d_ref_tree_leaf_count.restype = C.c_int32
        L.hf6d_ref_tree_leaf_count.argtypes = [C.c_void_p, C.c_int32]
        L.hf6d_ref_default_params.argtypes = [C.POINTER(Params)]
        L.hf6d_ref_scan_centres.restype = C.c_int32
        L.hf6d_ref_scan_centres.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32]
        L.hf6d_ref_gather.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32, C.c_void_p]
        L.hf6d_ref_normalise.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.hf6d_ref_normals.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
        L.hf6d_ref_gather_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32,
                                              C.c_void_p]
        L.hf6d_ref_quantise_normals.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.hf6d_ref_encode.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.hf6d_ref_traverse.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.hf6d_ref_vote.restype = C.c_int64
        L.hf6d_ref_vote.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Params),
                                    C.c_void_p, C.c_void_p]
        L.hf6d_ref_blur.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.hf6d_ref_nms.restype = C.c_int32
        L.hf6d_ref_nms.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32]
        L.hf6d_ref_hypotheses.restype = C.c_int32
        L.hf6d_ref_hypotheses.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Params),
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.hf6d_ref_detect.restype = C.c_int32
        L.hf6d_ref_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p]
        _lib = L
    return _lib


def set_threads(n: int | None = None) -> int:
    """OpenMP threads of the oracle (default: all host cores).  torchrun exports OMP_NUM_THREADS=1 into every rank, which
    would silently turn the 'all cores' CPU baseline into a single-threaded one; this goes through the OpenMP runtime the
    oracle library is linked against, so it works after the environment has been read.  Returns the threads in use."""
    lib()
    omp = C.CDLL("libgomp.so.1")
    omp.omp_set_num_threads(int(n or os.cpu_count() or 1))
    omp.omp_get_max_threads.restype = C.c_int
    return int(omp.omp_get_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**kw) -> Params:
    p = Params()
    lib().hf6d_ref_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Forest:
    def __init__(self, folder: str):
        self.h = lib().hf6d_ref_forest_load(folder.encode())
        if not self.h:
            raise IOError(f"oracle: cannot load forest from {folder}")
        info = np.zeros(6, np.int32)
        self.voxel_m = lib().hf6d_ref_forest_info(self.h, _p(info))
        self.T, self.K, self.F, self.patch_vox, self.n_leaves, self.n_internal = (int(x) for x in info)

    def leaf_count(self, t: int) -> int:
        return lib().hf6d_ref_tree_leaf_count(self.h, t)

    def close(self):
        if self.h:
            lib().hf6d_ref_forest_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def scan_centres(depth, p: Params):
    cap = ((p.W + p.stride - 1) // p.stride) * ((p.H + p.stride - 1) // p.stride)
    locs = np.zeros((cap, 2), np.int32)
    n = lib().hf6d_ref_scan_centres(_p(depth), C.byref(p), _p(locs), cap)
    return locs[:n].copy()


def gather(bgr, depth, p: Params, locs):
    P = locs.shape[0]
    out = np.zeros((P, p.patch_vox, p.patch_vox, 4), np.float32)
    lib().hf6d_ref_gather(_p(bgr), _p(depth), C.byref(p), _p(np.ascontiguousarray(locs)), P, _p(out))
    return out


def normals(depth, focal=575.0):
    """A2c: surface normals [H][W][3] (surface_normals.cu:11-73)."""
    H, W = depth.shape
    out = np.zeros((H, W, 3), np.float32)
    lib().hf6d_ref_normals(_p(depth), W, H, C.c_float(focal), _p(out))
    return out


def gather_normals(bgr, depth, nrm, p: Params, locs):
    """A2c: patches [P][ps][ps][6] = (B, G, R, nx, ny, nz)."""
    P = locs.shape[0]
    out = np.zeros((P, p.patch_vox, p.patch_vox, 6), np.float32)
    lib().hf6d_ref_gather_normals(_p(bgr), _p(depth), _p(np.ascontiguousarray(nrm)), C.byref(p),
                                  _p(np.ascontiguousarray(locs)), P, _p(out))
    return out


def quantise_normals(patches):
    P, ps = patches.shape[0], patches.shape[1]
    q = np.zeros((P, 6 * ps * ps), np.uint8)
    lib().hf6d_ref_quantise_normals(_p(np.ascontiguousarray(patches)), P, ps, _p(q))
    return q


def normalise(patches):
    P, ps = patches.shape[0], patches.shape[1]
    q = np.zeros((P, 4 * ps * ps), np.uint8)
    lib().hf6d_ref_normalise(_p(np.ascontiguousarray(patches)), P, ps, _p(q))
    return q


def encode(q, layers):
    (W1, b1), (W2, b2), (W3, b3) = layers
    P = q.shape[0]
    out = np.zeros((P, W3.shape[0]), np.float32)
    lib().hf6d_ref_encode(_p(np.ascontiguousarray(q)), P, W1.shape[1], _p(W1), _p(b1), W1.shape[0], _p(W2), _p(b2),
                          W2.shape[0], _p(W3), _p(b3), W3.shape[0], _p(out))
    return out


def traverse(forest: Forest, features):
    features = np.ascontiguousarray(features, np.float32)
    P = features.shape[0]
    ids = np.zeros((P, forest.T), np.int32)
    ords = np.zeros((P, forest.T), np.int32)
    lib().hf6d_ref_traverse(forest.h, _p(features), P, _p(ids), _p(ords))
    return ids, ords


def vote(forest: Forest, leaf_ord, locs, depth, p: Params, should_detect=None):
    maps = np.zeros((forest.K, p.H, p.W), np.uint64)
    sd = None if should_detect is None else np.ascontiguousarray(should_detect, np.uint8)
    n = lib().hf6d_ref_vote(forest.h, _p(np.ascontiguousarray(leaf_ord, np.int32)),
                            _p(np.ascontiguousarray(locs, np.int32)), _p(depth), leaf_ord.shape[0], C.byref(p), _p(sd),
                            _p(maps))
    return maps, int(n)


def blur(acc, kx, ky):
    acc = np.ascontiguousarray(acc, np.uint64)
    out = np.zeros(acc.shape, np.float32)
    rows, cols = acc.shape
    lib().hf6d_ref_blur(_p(acc), rows, cols, kx, ky, _p(out))
    return out


def nms(img, wx, wy, cap=4096):
    img = np.ascontiguousarray(img, np.float32)
    rows, cols = img.shape
    s = np.zeros(cap, np.float32)
    xs = np.zeros(cap, np.int32)
    ys = np.zeros(cap, np.int32)
    n = lib().hf6d_ref_nms(_p(img), rows, cols, wx, wy, _p(s), _p(xs), _p(ys), cap)
    n = min(n, cap)
    return s[:n].copy(), xs[:n].copy(), ys[:n].copy()


def hypotheses(forest: Forest, leaf_ord, locs, depth, p: Params, should_detect=None, max_loc=None, cap=4096):
    hy = np.zeros(cap, HYP_DTYPE)
    sd = None if should_detect is None else np.ascontiguousarray(should_detect, np.uint8)
    ml = None if max_loc is None else np.ascontiguousarray(max_loc, np.int32)
    n = lib().hf6d_ref_hypotheses(forest.h, _p(np.ascontiguousarray(leaf_ord, np.int32)),
                                  _p(np.ascontiguousarray(locs, np.int32)), _p(depth), leaf_ord.shape[0], C.byref(p),
                                  _p(sd), _p(ml), None, _p(hy), cap)
    return hy[:min(n, cap)].copy()


def detect(forest: Forest, bgr, depth, p: Params, layers, should_detect=None, max_loc=None, features_override=None,
           cap=4096):
    """Whole frame.  Returns (hypotheses, (P, P'), stage_seconds[6])."""
    hy = np.zeros(cap, HYP_DTYPE)
    flat = [a for Wb in layers for a in Wb]
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in flat])
    dims = np.array([layers[0][0].shape[1], layers[0][0].shape[0], layers[1][0].shape[0], layers[2][0].shape[0]],
                    np.int32)
    sd = None if should_detect is None else np.ascontiguousarray(should_detect, np.uint8)
    ml = None if max_loc is None else np.ascontiguousarray(max_loc, np.int32)
    npatch = np.zeros(2, np.int32)
    st = np.zeros(6, np.float64)
    fo = None if features_override is None else np.ascontiguousarray(features_override, np.float32)
    n = lib().hf6d_ref_detect(forest.h, _p(bgr), _p(depth), C.byref(p), ptrs, _p(dims), _p(sd), _p(ml), _p(fo), _p(hy),
                              cap, _p(npatch), _p(st))
    return hy[:min(n, cap)].copy(), (int(npatch[0]), int(npatch[1])), st
