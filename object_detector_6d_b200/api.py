"""ctypes binding of libhf6d.so (include/hf6d.h) -- the host-side mirror of the reference's HFTest interface.

The reference is compiled C++ (HoughForest/include/HFTest.h); its host side here is C++ too (csrc/hough_forest_main.cpp
is the `HoughForest --test` drop-in).  This module exists for the parity tests, the bench and the multi-GPU driver: it
loads the C ABI and nothing else -- no numpy implementation of any stage lives here, and importing it never touches the
oracle.  If the shared library is missing it is built in-tree with nvcc; if that fails the import raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

STAGE_SCAN, STAGE_GATHER, STAGE_ENCODE, STAGE_TRAVERSE, STAGE_VOTE, STAGE_CENTRES, STAGE_POSE = range(7)
STAGE_COUNT = 7
STAGE_NAMES = ("scan", "gather", "encode", "traverse", "vote", "centres", "pose")
(BUF_COUNTS, BUF_LOCS, BUF_PATCH_U8, BUF_FEATURES, BUF_LEAF_ORD, BUF_MAPS, BUF_BLURRED, BUF_CENTRES, BUF_FRAME_BGR,
 BUF_FRAME_DEPTH, BUF_NORMALS) = range(11)
MAX_CENTRES = 16

EXPORTS = (
    "hf6d_default_params", "hf6d_create", "hf6d_create_from_options", "hf6d_destroy", "hf6d_last_error",
    "hf6d_get_params", "hf6d_model", "hf6d_set_objects", "hf6d_get_objects", "hf6d_set_fill_seed",
    "hf6d_set_tree_shard", "hf6d_set_patch_shard", "hf6d_set_peer_split", "hf6d_set_class_shard", "hf6d_peer_blob_bytes", "hf6d_peer_export", "hf6d_peer_attach",
    "hf6d_peer_detach", "hf6d_peer_timed_out", "hf6d_set_encoder_mode", "hf6d_get_encoder_mode", "hf6d_set_feature_storage", "hf6d_get_feature_storage",
    "hf6d_set_debug_capture", "hf6d_detect", "hf6d_submit",
    "hf6d_wait", "hf6d_host_alloc", "hf6d_host_free", "hf6d_upload", "hf6d_run", "hf6d_sync", "hf6d_collect",
    "hf6d_fetch", "hf6d_inject", "hf6d_device_ptr", "hf6d_set_stream", "hf6d_stage_ms", "hf6d_launch_count",
    "hf6d_pose_from_tuple", "hf6d_count_cast_votes", "hf6d_bind_frame", "hf6d_encoder_layer_ms", "hf6d_result_bytes",
    "hf6d_parse_options", "hf6d_inspect_forest", "hf6d_inspect_weights", "hf6d_debug_texture_gather",
    "hf6d_default_refine_params", "hf6d_set_refine_params", "hf6d_get_refine_params", "hf6d_set_object_model",
    "hf6d_load_object_ply", "hf6d_load_option_models", "hf6d_refine", "hf6d_refine_ms", "hf6d_refine_fetch",
    "hf6d_default_train_params", "hf6d_train_forest", "hf6d_train_forest_mem",
    "hf6d_default_render_params", "hf6d_renderer_create_ply", "hf6d_renderer_create", "hf6d_renderer_destroy",
    "hf6d_renderer_view_count", "hf6d_renderer_view", "hf6d_render",
    "hf6d_patchdb_create", "hf6d_patchdb_open", "hf6d_patchdb_put", "hf6d_patchdb_entries", "hf6d_patchdb_next", "hf6d_patchdb_close",
    "hf6d_patch_annotation", "hf6d_create_extractor", "hf6d_patch_capacity", "hf6d_encode_patches", "hf6d_generate_train_vectors",
)


class Params(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("stride", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("patch_vox", C.c_int32), ("voxel_m", C.c_float),
                ("max_depth_range_m", C.c_float), ("distance_threshold_m", C.c_float), ("fill_random", C.c_int32),
                ("fill_seed", C.c_uint64), ("batch_size", C.c_int32), ("max_yaw_pitch_hypotheses", C.c_int32),
                ("max_roll_hypotheses", C.c_int32), ("min_location_score_ratio", C.c_float),
                ("min_yaw_pitch_drop_ratio", C.c_float), ("centers_blur_size", C.c_int32),
                ("centers_nms_wsize", C.c_int32), ("pose_blur_size", C.c_int32), ("pose_nms_wsize", C.c_int32),
                ("patch_mode", C.c_int32), ("normals_focal", C.c_float)]


class ObjectOptions(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("should_detect", C.c_int32), ("max_location_hypotheses", C.c_int32),
                ("instances", C.c_int32)]


class ModelInfo(C.Structure):
    _fields_ = [("T", C.c_int32), ("K", C.c_int32), ("F", C.c_int32), ("patch_vox", C.c_int32),
                ("voxel_m", C.c_float), ("n_leaves", C.c_int64), ("n_internal", C.c_int64), ("n_votes", C.c_int64),
                ("max_depth", C.c_int32), ("dims", C.c_int32 * 4)]


class Options(C.Structure):
    _fields_ = [("params", Params), ("gpu", C.c_int32), ("n_objects", C.c_int32), ("forest_folder", C.c_char * 1024),
                ("caffe_weights", C.c_char * 1024), ("caffe_definition", C.c_char * 1024),
                ("location_score_coeff", C.c_float), ("pose_score_coeff", C.c_float)]


class RefineParams(C.Structure):
    """hf6d_refine_params: MeshUtils' settings (HoughForest/include/MeshUtils.h:115-156, HFTest.cpp:1203-1225)."""
    _fields_ = [("scene_leaf_m", C.c_float), ("object_leaf_m", C.c_float), ("normals_radius_m", C.c_float),
                ("nn_search_radius_m", C.c_float), ("occlusion_threshold_m", C.c_float), ("similarity_coeff", C.c_float),
                ("inliers_coeff", C.c_float), ("clutter_coeff", C.c_float), ("location_score_coeff", C.c_float),
                ("pose_score_coeff", C.c_float), ("group_total_explain_coeff", C.c_float),
                ("group_common_explain_coeff", C.c_float), ("inliers_threshold", C.c_float), ("clutter_threshold", C.c_float),
                ("final_score_threshold", C.c_float), ("cluster_eps_angle_threshold", C.c_float),
                ("cluster_curvature_threshold", C.c_float), ("cluster_tolerance_near", C.c_float),
                ("cluster_tolerance_far", C.c_float), ("cluster_min_points", C.c_int32), ("use_color_similarity", C.c_int32),
                ("use_normal_similarity", C.c_int32), ("search_single_object_instance", C.c_int32),
                ("search_single_object_in_group", C.c_int32), ("default_icp_iterations", C.c_int32)]


class TrainParams(C.Structure):
    """hf6d_train_params: the --train flags of the reference's main.cpp:13-25."""
    _fields_ = [("trees", C.c_int32), ("min_samples", C.c_int32), ("tests_per_node", C.c_int32),
                ("thresholds_per_test", C.c_int32), ("start_tree_no", C.c_int32), ("patch_size_in_voxels", C.c_int32),
                ("voxel_size_in_m", C.c_float), ("seed", C.c_uint64), ("device", C.c_int32)]


class RenderParams(C.Structure):
    """hf6d_render_params: the --render flags of PatchGen/src/main.cpp:22-49."""
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("view_angle_deg", C.c_float), ("tesselation_level", C.c_int32),
                ("use_vertices", C.c_int32), ("in_place_rotations", C.c_int32), ("lightings", C.c_int32), ("heights", C.c_int32),
                ("height_step", C.c_float), ("start_height", C.c_float), ("above_z", C.c_int32), ("below_z", C.c_int32),
                ("render_around_0", C.c_int32), ("object_radius", C.c_float), ("device", C.c_int32)]


class TrainStats(C.Structure):
    _fields_ = [("nodes", C.c_int64), ("leaves", C.c_int64), ("max_depth", C.c_int32), ("training_samples", C.c_int32),
                ("train_ms", C.c_float)]


class TrainVecStats(C.Structure):
    _fields_ = [("entries", C.c_int64), ("written", C.c_int64), ("classes", C.c_int32), ("feature_length", C.c_int32),
                ("encode_ms", C.c_float)]


DETECTION_DTYPE = np.dtype([("hypothesis", "<i4"), ("cls", "<i4"), ("pose", "<f4", (16,)), ("similarity", "<f4"),
                            ("inliers_ratio", "<f4"), ("clutter", "<f4"), ("location_score", "<f4"), ("pose_score", "<f4"),
                            ("final_score", "<f4"), ("icp_converged", "<i4"), ("icp_iterations", "<i4"), ("visible", "<i4"),
                            ("inliers", "<i4"), ("explained", "<i4"), ("accepted", "<i4"), ("selected", "<i4"), ("rank", "<i4")])
RBUF_SCENE_POINTS, RBUF_SCENE_NORMALS, RBUF_SCENE_LABELS, RBUF_CLUSTER_SIZES, RBUF_MODEL_POINTS, RBUF_MODEL_NORMALS, \
    RBUF_MODEL_VERTICES = range(7)

HYP_DTYPE = np.dtype([("cls", "<i4"), ("cx", "<i4"), ("cy", "<i4"), ("z", "<f4"), ("yaw_deg", "<i4"),
                      ("pitch_deg", "<i4"), ("roll_deg", "<i4"), ("loc_score", "<f4"), ("yawpitch_score", "<f4"),
                      ("roll_score", "<f4"), ("pose", "<f4", (16,))])
CENTRE_DTYPE = np.dtype([("score", "<f4"), ("x", "<i4"), ("y", "<i4")])
CENTRE_LIST_DTYPE = np.dtype([("n", "<i4"), ("c", CENTRE_DTYPE, (MAX_CENTRES,))])

_lib = None


class Hf6dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hf6d error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building if necessary) libhf6d.so.  Raises if the CUDA extension cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_lib()
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.hf6d_default_params.argtypes = [C.POINTER(Params)]
    L.hf6d_default_params.restype = None
    L.hf6d_create.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, i32, i32, C.POINTER(vp)]
    L.hf6d_create_from_options.argtypes = [C.c_char_p, i32, i32, i32, i32, C.POINTER(vp)]
    L.hf6d_destroy.argtypes = [vp]
    L.hf6d_destroy.restype = None
    L.hf6d_last_error.argtypes = [vp]
    L.hf6d_last_error.restype = C.c_char_p
    L.hf6d_get_params.argtypes = [vp, C.POINTER(Params)]
    L.hf6d_model.argtypes = [vp, C.POINTER(ModelInfo)]
    L.hf6d_set_objects.argtypes = [vp, C.POINTER(ObjectOptions), i32]
    L.hf6d_get_objects.argtypes = [vp, C.POINTER(ObjectOptions), i32]
    L.hf6d_set_fill_seed.argtypes = [vp, C.c_uint64]
    L.hf6d_set_tree_shard.argtypes = [vp, i32, i32]
    L.hf6d_set_patch_shard.argtypes = [vp, i32, i32]
    L.hf6d_set_peer_split.argtypes = [vp, i32]
    L.hf6d_set_class_shard.argtypes = [vp, i32, i32]
    L.hf6d_peer_blob_bytes.argtypes = []
    L.hf6d_peer_blob_bytes.restype = C.c_size_t
    L.hf6d_peer_export.argtypes = [vp, vp, C.c_size_t]
    L.hf6d_peer_attach.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.hf6d_peer_detach.argtypes = [vp]
    L.hf6d_peer_timed_out.argtypes = [vp]
    L.hf6d_set_encoder_mode.argtypes = [vp, i32]
    L.hf6d_get_encoder_mode.argtypes = [vp]
    L.hf6d_set_feature_storage.argtypes = [vp, i32]
    L.hf6d_get_feature_storage.argtypes = [vp]
    L.hf6d_count_cast_votes.restype = C.c_int64
    L.hf6d_count_cast_votes.argtypes = [vp, i32]
    L.hf6d_set_debug_capture.argtypes = [vp, i32]
    L.hf6d_detect.argtypes = [vp, vp, vp, vp, i32, C.POINTER(i32)]
    L.hf6d_submit.argtypes = [vp, vp, vp, C.POINTER(i32)]
    L.hf6d_wait.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
    L.hf6d_host_alloc.argtypes = [C.c_size_t]
    L.hf6d_host_alloc.restype = vp
    L.hf6d_host_free.argtypes = [vp]
    L.hf6d_host_free.restype = None
    L.hf6d_upload.argtypes = [vp, i32, vp, vp]
    L.hf6d_run.argtypes = [vp, i32, i32, i32]
    L.hf6d_sync.argtypes = [vp, i32]
    L.hf6d_collect.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
    L.hf6d_fetch.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.hf6d_fetch.restype = i64
    L.hf6d_inject.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.hf6d_device_ptr.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.hf6d_set_stream.argtypes = [vp, i32, vp]
    L.hf6d_stage_ms.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.hf6d_launch_count.argtypes = [vp, i32]
    L.hf6d_bind_frame.argtypes = [vp, i32, vp, vp]
    L.hf6d_result_bytes.argtypes = [vp]
    L.hf6d_result_bytes.restype = i64
    L.hf6d_encoder_layer_ms.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.hf6d_pose_from_tuple.argtypes = [C.POINTER(Params), i32, i32, C.c_float, i32, i32, i32, C.POINTER(C.c_float)]
    L.hf6d_pose_from_tuple.restype = None
    L.hf6d_debug_texture_gather.argtypes = [vp, i32, vp, C.c_size_t]
    L.hf6d_debug_texture_gather.restype = i64
    L.hf6d_parse_options.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(ObjectOptions), i32]
    L.hf6d_inspect_forest.argtypes = [C.c_char_p, C.POINTER(ModelInfo)]
    L.hf6d_inspect_weights.argtypes = [C.c_char_p, C.POINTER(C.c_int32)]
    L.hf6d_default_render_params.argtypes = [C.POINTER(RenderParams)]
    L.hf6d_default_render_params.restype = None
    L.hf6d_renderer_create_ply.argtypes = [C.POINTER(RenderParams), C.c_char_p, C.POINTER(vp)]
    L.hf6d_renderer_create.argtypes = [C.POINTER(RenderParams), vp, vp, i32, vp, i32, C.POINTER(vp)]
    L.hf6d_renderer_destroy.argtypes = [vp]
    L.hf6d_renderer_destroy.restype = None
    L.hf6d_renderer_view_count.argtypes = [vp]
    L.hf6d_renderer_view.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.hf6d_render.argtypes = [vp, C.POINTER(C.c_double), C.c_float, vp, vp]
    L.hf6d_default_train_params.argtypes = [C.POINTER(TrainParams)]
    L.hf6d_default_train_params.restype = None
    L.hf6d_train_forest.argtypes = [C.POINTER(TrainParams), C.c_char_p, C.c_char_p, C.POINTER(TrainStats)]
    L.hf6d_train_forest_mem.argtypes = [C.POINTER(TrainParams), i32, i32, i32, vp, vp, vp, C.c_char_p, C.POINTER(TrainStats)]
    L.hf6d_default_refine_params.argtypes = [C.POINTER(RefineParams)]
    L.hf6d_default_refine_params.restype = None
    L.hf6d_set_refine_params.argtypes = [vp, C.POINTER(RefineParams)]
    L.hf6d_get_refine_params.argtypes = [vp, C.POINTER(RefineParams)]
    L.hf6d_set_object_model.argtypes = [vp, i32, vp, vp, i32, C.c_float, i32]
    L.hf6d_load_object_ply.argtypes = [vp, i32, C.c_char_p, C.c_float, i32]
    L.hf6d_load_option_models.argtypes = [vp]
    L.hf6d_refine.argtypes = [vp, i32, vp, i32, vp, i32, C.POINTER(i32)]
    L.hf6d_refine_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hf6d_refine_fetch.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.hf6d_refine_fetch.restype = i64
    L.hf6d_patchdb_create.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.hf6d_patchdb_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.hf6d_patchdb_put.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, vp]
    L.hf6d_patchdb_entries.argtypes = [vp]
    L.hf6d_patchdb_entries.restype = i64
    L.hf6d_patchdb_next.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32), vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hf6d_patchdb_close.argtypes = [vp]
    L.hf6d_patch_annotation.argtypes = [i32, i32, C.c_float, i32, i32, C.c_uint16, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.hf6d_patch_annotation.restype = None
    L.hf6d_create_extractor.argtypes = [C.POINTER(Params), C.c_char_p, i32, C.POINTER(vp)]
    L.hf6d_patch_capacity.argtypes = [vp]
    L.hf6d_encode_patches.argtypes = [vp, i32, vp, i32, vp]
    L.hf6d_generate_train_vectors.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, i32, C.POINTER(TrainVecStats)]
    _lib = L
    return L


def default_params(**kw) -> Params:
    p = Params()
    load().hf6d_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def train_forest(out_dir: str, cls=None, dof=None, features=None, K: int = 0, input_file: str | None = None, **kw) -> TrainStats:
    """HFTrain::train on the GPU (HoughForest/src/HFTrain.cpp:1199-1265): writes forest.txt + tree<N>.dat into out_dir.
    Either input_file (the reference's training-vector file) or cls / dof / features arrays.  kw: TrainParams fields."""
    L = load()
    p = TrainParams()
    L.hf6d_default_train_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    st = TrainStats()
    os.makedirs(out_dir, exist_ok=True)
    if input_file is not None:
        rc = L.hf6d_train_forest(C.byref(p), input_file.encode(), out_dir.encode(), C.byref(st))
    else:
        cls = np.ascontiguousarray(cls, np.int32)
        dof = np.ascontiguousarray(dof, np.float32).reshape(len(cls), 6)
        features = np.ascontiguousarray(features, np.float32).reshape(len(cls), -1)
        rc = L.hf6d_train_forest_mem(C.byref(p), len(cls), K, features.shape[1], cls.ctypes.data, dof.ctypes.data,
                                     features.ctypes.data, out_dir.encode(), C.byref(st))
    if rc < 0:
        raise Hf6dError(rc, (L.hf6d_last_error(None) or b"").decode())
    return st


class Renderer:
    """RenderViewsTesselatedSphere (PatchGen/src/render_views_tesselated_sphere_mod.cpp) on the GPU: hf6d_renderer_*."""

    def __init__(self, xyz=None, rgb=None, faces=None, ply_path: str | None = None, **kw):
        self._L = load()
        self.params = RenderParams()
        self._L.hf6d_default_render_params(C.byref(self.params))
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._h = C.c_void_p()
        if ply_path is not None:
            rc = self._L.hf6d_renderer_create_ply(C.byref(self.params), ply_path.encode(), C.byref(self._h))
        else:
            xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
            rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
            faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
            rc = self._L.hf6d_renderer_create(C.byref(self.params), xyz.ctypes.data, rgb.ctypes.data, len(xyz), faces.ctypes.data,
                                              len(faces), C.byref(self._h))
        if rc:
            raise Hf6dError(rc, (self._L.hf6d_last_error(None) or b"").decode())

    def view_count(self) -> int:
        return self._L.hf6d_renderer_view_count(self._h)

    def view(self, i: int) -> np.ndarray:
        m = (C.c_double * 16)()
        if self._L.hf6d_renderer_view(self._h, i, m):
            raise Hf6dError(-1, (self._L.hf6d_last_error(None) or b"").decode())
        return np.array(m, np.float64).reshape(4, 4)

    def render(self, pose, ambient: float = 0.0):
        m = (C.c_double * 16)(*np.asarray(pose, np.float64).reshape(-1))
        bgr = np.zeros((self.params.H, self.params.W, 3), np.uint8)
        depth = np.zeros((self.params.H, self.params.W), np.uint16)
        rc = self._L.hf6d_render(self._h, m, ambient, bgr.ctypes.data, depth.ctypes.data)
        if rc:
            raise Hf6dError(rc, (self._L.hf6d_last_error(None) or b"").decode())
        return bgr, depth

    def close(self):
        if getattr(self, "_h", None):
            self._L.hf6d_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ck_host(rc):
    if rc < 0:
        raise Hf6dError(rc, (load().hf6d_last_error(None) or b"").decode())


class PatchDb:
    """The reference's LMDB patch database (PatchGen/src/patch_generator.cpp:479-493, train_patch_generator.cpp:33-101):
    hf6d_patchdb_*.  mode "w": put() in ascending key order, close() commits; mode "r": iterate (key, dims, data)."""

    def __init__(self, folder: str, mode: str = "r"):
        self._L = load()
        self._h = C.c_void_p()
        fn = self._L.hf6d_patchdb_create if mode == "w" else self._L.hf6d_patchdb_open
        _ck_host(fn(folder.encode(), C.byref(self._h)))

    def put(self, key: str, data, label: int):
        data = np.ascontiguousarray(data, np.uint8)
        ch, h, w = data.shape
        _ck_host(self._L.hf6d_patchdb_put(self._h, key.encode(), ch, h, w, label, data.ctypes.data))

    def entries(self) -> int:
        return self._L.hf6d_patchdb_entries(self._h)

    def __iter__(self):
        key = C.create_string_buffer(64)
        dims = (C.c_int32 * 4)()
        buf = np.zeros(1 << 20, np.uint8)
        n = C.c_size_t()
        while True:
            rc = self._L.hf6d_patchdb_next(self._h, key, dims, buf.ctypes.data, buf.nbytes, C.byref(n))
            _ck_host(rc)
            if rc == 0:
                return
            yield key.value.decode(), tuple(dims), buf[:n.value].copy()

    def close(self):
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            _ck_host(self._L.hf6d_patchdb_close(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def patch_annotation(W, H, x, y, depth_mm, pose, view_angle_deg: float = 45.3105) -> np.ndarray:
    """One line of patch_annotation_lmdb.txt: (yaw, pitch, roll, x, y, z) (patch_generator.cpp:20-56).  Host arithmetic."""
    m = (C.c_float * 16)(*np.asarray(pose, np.float32).reshape(-1))
    out = (C.c_float * 6)()
    load().hf6d_patch_annotation(W, H, view_angle_deg, int(x), int(y), int(depth_mm), m, out)
    return np.array(out, np.float32)


def generate_train_vectors(weights_path: str, lmdb_folder: str, output_file: str, batch_size: int = 1, device: int = 0,
                           encoder_mode: int = 0) -> TrainVecStats:
    """train_patch_generator::generate_train_patches (PatchGen/src/train_patch_generator.cpp:22-200), encoder on the GPU."""
    st = TrainVecStats()
    _ck_host(load().hf6d_generate_train_vectors(weights_path.encode(), lmdb_folder.encode(), output_file.encode(), batch_size,
                                                device, encoder_mode, C.byref(st)))
    return st


def parse_options(path: str):
    """Host-only: a DetectorOptions text file -> (Options, [object dicts]).  Raises Hf6dError on a malformed file."""
    o = Options()
    objs = (ObjectOptions * 32)()
    _ck_host(load().hf6d_parse_options(path.encode(), C.byref(o), objs, 32))
    return o, [dict(name=objs[k].name.decode(), should_detect=bool(objs[k].should_detect),
                    max_location_hypotheses=objs[k].max_location_hypotheses, instances=objs[k].instances)
               for k in range(min(o.n_objects, 32))]


def inspect_forest(forest_dir: str) -> ModelInfo:
    """Host-only: load + flatten forest.txt / treeN.dat with the product's own reader; counts only."""
    mi = ModelInfo()
    _ck_host(load().hf6d_inspect_forest(forest_dir.encode(), C.byref(mi)))
    return mi


def inspect_weights(path: str):
    dims = (C.c_int32 * 4)()
    _ck_host(load().hf6d_inspect_weights(path.encode(), dims))
    return tuple(dims)


class PinnedArray:
    """numpy view over pinned host memory from hf6d_host_alloc."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = load().hf6d_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("hf6d_host_alloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load().hf6d_host_free(self.ptr)
            self.ptr = None


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch.as_tensor can wrap a libhf6d device buffer (for NCCL)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class Detector:
    """One libhf6d context = the model on one GPU plus `n_slots` frame workspaces.

    Mirrors what HFTest holds between frames (HoughForest/include/HFTest.h:25-50): forest, encoder, options.
    """

    def __init__(self, forest_dir=None, weights_path=None, params: Params | None = None, device: int = 0,
                 n_slots: int = 1, options_path: str | None = None, frame_size=None, extractor: bool = False):
        self._L = load()
        self._h = C.c_void_p()
        if extractor:  # no forest: patches (SCAN..GATHER) and, with weights, features (hf6d_create_extractor)
            rc = self._L.hf6d_create_extractor(C.byref(params), weights_path.encode() if weights_path else None, device,
                                               C.byref(self._h))
        elif options_path is not None:
            W, H = frame_size or (640, 480)
            rc = self._L.hf6d_create_from_options(options_path.encode(), W, H, device, n_slots, C.byref(self._h))
        else:
            p = params if params is not None else default_params()
            rc = self._L.hf6d_create(C.byref(p), forest_dir.encode(), weights_path.encode(), device, n_slots,
                                     C.byref(self._h))
        if rc:
            raise Hf6dError(rc, (self._L.hf6d_last_error(None) or b"").decode())
        self.params = Params()
        self._L.hf6d_get_params(self._h, C.byref(self.params))
        self.model = ModelInfo()
        self._L.hf6d_model(self._h, C.byref(self.model))
        self.n_slots = n_slots
        self.T, self.K, self.F = self.model.T, self.model.K, self.model.F
        self.W, self.H = self.params.W, self.params.H

    # ------------------------------------------------------------------ stage REFINE
    def refine_params(self) -> RefineParams:
        p = RefineParams()
        self._ck(self._L.hf6d_get_refine_params(self._h, C.byref(p)))
        return p

    def set_refine_params(self, p: RefineParams = None, **kw):
        """MeshUtils::setReg / setGroupReg / set*Threshold / setClusteringOptions / ... (HFTest.cpp:1203-1225)."""
        p = p if p is not None else self.refine_params()
        for k, v in kw.items():
            setattr(p, k, v)
        self._ck(self._L.hf6d_set_refine_params(self._h, C.byref(p)))

    def set_object_model(self, cls: int, xyz, rgb, nn_search_radius: float = -1.0, icp_iterations: int = -1):
        """MeshUtils::insertObjectFromPLY (MeshUtils.h:213-247) from vertices in memory."""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
        assert len(xyz) == len(rgb)
        self._ck(self._L.hf6d_set_object_model(self._h, cls, xyz.ctypes.data, rgb.ctypes.data, len(xyz), nn_search_radius,
                                               icp_iterations))

    def load_object_ply(self, cls: int, path: str, nn_search_radius: float = -1.0, icp_iterations: int = -1):
        self._ck(self._L.hf6d_load_object_ply(self._h, cls, path.encode(), nn_search_radius, icp_iterations))

    def load_option_models(self):
        """MeshUtils::insertObjectFromPLY for every detected object of the options file (HFTest.cpp:1227-1233)."""
        self._ck(self._L.hf6d_load_option_models(self._h))

    def refine(self, hyps, slot: int = 0):
        """ICP + evaluate_hypothesis + optimize_hypotheses for the hypotheses of the frame the slot holds
        (HFTest.cpp:922-994, 1261-1273).  Returns a DETECTION_DTYPE array, one row per hypothesis."""
        hyps = np.ascontiguousarray(hyps, HYP_DTYPE)
        out = np.zeros(max(len(hyps), 1), DETECTION_DTYPE)
        n = C.c_int(0)
        self._ck(self._L.hf6d_refine(self._h, slot, hyps.ctypes.data, len(hyps), out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def refine_ms(self):
        ms = (C.c_float * 4)()
        self._ck(self._L.hf6d_refine_ms(self._h, ms))
        return dict(zip(("scene", "icp", "score", "optimise"), (float(x) for x in ms)))

    def refine_fetch(self, what: int, arg: int = 0):
        cap = self._ck(self._L.hf6d_refine_fetch(self._h, what, arg, None, 0))
        buf = np.zeros(max(cap, 1), np.uint8)
        n = self._ck(self._L.hf6d_refine_fetch(self._h, what, arg, buf.ctypes.data, cap))
        raw = buf[:n]
        if what in (RBUF_SCENE_LABELS, RBUF_CLUSTER_SIZES):
            return raw.view(np.int32).copy()
        return raw.view(np.float32).reshape(-1, 3 if what == RBUF_MODEL_VERTICES else 4).copy()

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc < 0:
            raise Hf6dError(rc, (self._L.hf6d_last_error(self._h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._L.hf6d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def _check_frame(self, bgr, depth):
        if bgr.dtype != np.uint8 or bgr.shape != (self.H, self.W, 3) or not bgr.flags.c_contiguous:
            raise ValueError(f"bgr must be contiguous uint8 [{self.H},{self.W},3]")
        if depth.dtype != np.uint16 or depth.shape != (self.H, self.W) or not depth.flags.c_contiguous:
            raise ValueError(f"depth must be contiguous uint16 [{self.H},{self.W}]")

    # ------------------------------------------------------------------ configuration
    def set_objects(self, should_detect=None, max_loc=None, names=None, instances=None):
        objs = (ObjectOptions * self.K)()
        for k in range(self.K):
            objs[k].name = (names[k] if names else f"object{k}").encode()
            objs[k].should_detect = int(should_detect[k]) if should_detect is not None else 1
            objs[k].max_location_hypotheses = int(max_loc[k]) if max_loc is not None else 12
            objs[k].instances = int(instances[k]) if instances is not None else 1
        self._ck(self._L.hf6d_set_objects(self._h, objs, self.K))

    def objects(self):
        objs = (ObjectOptions * self.K)()
        self._ck(self._L.hf6d_get_objects(self._h, objs, self.K))
        return [dict(name=o.name.decode(), should_detect=bool(o.should_detect),
                     max_location_hypotheses=o.max_location_hypotheses, instances=o.instances) for o in objs]

    def set_fill_seed(self, seed: int):
        self._ck(self._L.hf6d_set_fill_seed(self._h, seed))

    def set_tree_shard(self, rank: int, world: int):
        self._ck(self._L.hf6d_set_tree_shard(self._h, rank, world))

    def set_patch_shard(self, rank: int, world: int):
        """Gather .. vote only this rank's share of the frame's patches (128-patch row blocks, reference patch order)."""
        self._ck(self._L.hf6d_set_patch_shard(self._h, rank, world))

    def set_peer_split(self, split: int):
        """What peer_attach shards: 0 = trees, 1 = patches."""
        self._ck(self._L.hf6d_set_peer_split(self._h, split))

    def set_class_shard(self, rank: int, world: int):
        self._ck(self._L.hf6d_set_class_shard(self._h, rank, world))

    # ------------------------------------------------------------------ peer exchange (tree-sharded mode over NVLink)
    def peer_export(self) -> bytes:
        """This rank's blob (CUDA IPC handles of its vote maps, leaf tables and flag block)."""
        n = int(self._L.hf6d_peer_blob_bytes())
        buf = C.create_string_buffer(n)
        self._ck(self._L.hf6d_peer_export(self._h, buf, n))
        return buf.raw

    def peer_attach(self, rank: int, world: int, blobs):
        """Map the peers' buffers; `blobs` = every rank's peer_export(), in rank order."""
        blobs = [bytes(b) for b in blobs]
        if len(blobs) != world or len({len(b) for b in blobs}) != 1:
            raise ValueError("need one blob of equal size per rank")
        joined = C.create_string_buffer(b"".join(blobs), len(blobs[0]) * world)
        self._ck(self._L.hf6d_peer_attach(self._h, rank, world, joined, len(blobs[0])))

    def peer_detach(self):
        self._ck(self._L.hf6d_peer_detach(self._h))

    def peer_timed_out(self) -> bool:
        return bool(self._L.hf6d_peer_timed_out(self._h))

    def set_encoder_mode(self, mode: int):
        """0 = bf16 operands (throughput), 1 = split bf16 hi + lo operands (~fp32 products, about 3x the encoder time),
        2 = fp16 operands (same kernel and rate as 0, 8x finer significands)."""
        self._ck(self._L.hf6d_set_encoder_mode(self._h, int(mode)))

    def encoder_mode(self) -> int:
        return int(self._L.hf6d_get_encoder_mode(self._h))

    def set_feature_storage(self, storage: int):
        """0 = fp32 feature rows, 1 = fp16 rows written by the feature layer and read by the traversal (modes 0 / 2)."""
        self._ck(self._L.hf6d_set_feature_storage(self._h, int(storage)))

    def feature_storage(self) -> int:
        return int(self._L.hf6d_get_feature_storage(self._h))

    def count_cast_votes(self, slot: int = 0) -> int:
        n = int(self._L.hf6d_count_cast_votes(self._h, slot))
        if n < 0:
            self._ck(n)
        return n

    def patch_capacity(self) -> int:
        return self._L.hf6d_patch_capacity(self._h)

    def encode_patches(self, patches, slot: int = 0) -> np.ndarray:
        """The encoder over caller-held quantised patches uint8[n][C*ps*ps] (Datum.data): float[n][F]."""
        patches = np.ascontiguousarray(patches, np.uint8).reshape(len(patches), -1)
        out = np.zeros((len(patches), self.F), np.float32)
        self._ck(self._L.hf6d_encode_patches(self._h, slot, patches.ctypes.data, len(patches), out.ctypes.data))
        return out

    def set_debug_capture(self, on: bool):
        self._ck(self._L.hf6d_set_debug_capture(self._h, int(on)))

    def set_stream(self, slot: int, cuda_stream: int | None):
        self._ck(self._L.hf6d_set_stream(self._h, slot, C.c_void_p(cuda_stream) if cuda_stream else None))

    # ------------------------------------------------------------------ whole frame (the call a user makes)
    def detect(self, bgr, depth, cap: int = 4096):
        self._check_frame(bgr, depth)
        out = np.zeros(cap, HYP_DTYPE)
        n = C.c_int(0)
        self._ck(self._L.hf6d_detect(self._h, self._ptr(bgr), self._ptr(depth), self._ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    def submit(self, bgr, depth) -> int:
        t = C.c_int(0)
        self._ck(self._L.hf6d_submit(self._h, self._ptr(bgr), self._ptr(depth), C.byref(t)))
        return t.value

    def wait(self, ticket: int, cap: int = 4096):
        out = np.zeros(cap, HYP_DTYPE)
        n = C.c_int(0)
        self._ck(self._L.hf6d_wait(self._h, ticket, self._ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    # ------------------------------------------------------------------ stage level
    def upload(self, slot, bgr, depth):
        self._check_frame(bgr, depth)
        self._ck(self._L.hf6d_upload(self._h, slot, self._ptr(bgr), self._ptr(depth)))

    def bind_frame(self, slot, d_bgr_ptr, d_depth_ptr):
        """Device-resident frame (raw device pointers, e.g. torch tensor .data_ptr()); (None, None) unbinds."""
        self._ck(self._L.hf6d_bind_frame(self._h, slot, C.c_void_p(d_bgr_ptr) if d_bgr_ptr else None,
                                         C.c_void_p(d_depth_ptr) if d_depth_ptr else None))

    def encoder_layer_ms(self, slot=0):
        ms = (C.c_float * 3)()
        self._ck(self._L.hf6d_encoder_layer_ms(self._h, slot, ms))
        return np.array(list(ms), np.float64)

    def run(self, slot=0, first=STAGE_SCAN, last=STAGE_POSE):
        self._ck(self._L.hf6d_run(self._h, slot, first, last))

    def sync(self, slot=0):
        self._ck(self._L.hf6d_sync(self._h, slot))

    def collect(self, slot=0, cap: int = 4096):
        out = np.zeros(cap, HYP_DTYPE)
        n = C.c_int(0)
        self._ck(self._L.hf6d_collect(self._h, slot, self._ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    def stage_ms(self, slot=0):
        ms = (C.c_float * STAGE_COUNT)()
        self._ck(self._L.hf6d_stage_ms(self._h, slot, ms))
        return np.array(list(ms), np.float64)

    def result_bytes(self) -> int:
        return int(self._L.hf6d_result_bytes(self._h))

    def launch_count(self, slot=0) -> int:
        return self._ck(self._L.hf6d_launch_count(self._h, slot))

    def counts(self, slot=0):
        a = np.zeros(2, np.int32)
        self._ck(self._L.hf6d_fetch(self._h, slot, BUF_COUNTS, self._ptr(a), a.nbytes))
        return int(a[0]), int(a[1])

    def fetch(self, what, slot=0):
        P, Pp = self.counts(slot)
        K, T, F, H, W = self.K, self.T, self.F, self.H, self.W
        shapes = {
            BUF_COUNTS: ((2,), np.int32), BUF_LOCS: ((P, 2), np.int32), BUF_PATCH_U8: ((Pp, self.model.dims[0]), np.uint8),
            BUF_FEATURES: ((Pp, F), np.float32), BUF_LEAF_ORD: ((Pp, T), np.int32), BUF_MAPS: ((K, H, W), np.uint64),
            BUF_BLURRED: ((K, H, W), np.float32), BUF_CENTRES: ((K,), CENTRE_LIST_DTYPE),
            BUF_FRAME_BGR: ((H, W, 3), np.uint8), BUF_FRAME_DEPTH: ((H, W), np.uint16),
            BUF_NORMALS: ((H, W, 4), np.float32),
        }
        shape, dt = shapes[what]
        a = np.zeros(shape, dt)
        n = self._ck(self._L.hf6d_fetch(self._h, slot, what, self._ptr(a), a.nbytes))
        assert n == a.nbytes, (n, a.nbytes)
        return a

    def texture_gather(self, slot=0):
        """Diagnostic: the slot's patches through a real CUDA texture object (the reference's own fetch), fp32 HWC."""
        _, Pp = self.counts(slot)
        ps = self.params.patch_vox
        a = np.zeros((Pp, ps, ps, 4), np.float32)
        n = self._ck(self._L.hf6d_debug_texture_gather(self._h, slot, self._ptr(a), a.nbytes))
        assert n == a.nbytes
        return a

    def inject(self, what, array, slot=0):
        a = np.ascontiguousarray(array)
        self._ck(self._L.hf6d_inject(self._h, slot, what, self._ptr(a), a.nbytes))

    def device_ptr(self, what, slot=0):
        p = C.c_void_p()
        n = C.c_size_t()
        self._ck(self._L.hf6d_device_ptr(self._h, slot, what, C.byref(p), C.byref(n)))
        return p.value, n.value

    def device_array(self, what, slot=0):
        """A __cuda_array_interface__ view of MAPS (int64 [K,H,W]) or LEAF_ORD (int32 [cap,T]) for torch.as_tensor."""
        ptr, nbytes = self.device_ptr(what, slot)
        if what == BUF_MAPS:
            return _CudaArray(ptr, (self.K, self.H, self.W), "<i8", self)
        if what == BUF_LEAF_ORD:
            return _CudaArray(ptr, (nbytes // (4 * self.T), self.T), "<i4", self)
        raise ValueError("device_array supports BUF_MAPS and BUF_LEAF_ORD")

    def pose_from_tuple(self, cx, cy, z, yaw_deg, pitch_deg, roll_deg):
        out = (C.c_float * 16)()
        self._L.hf6d_pose_from_tuple(C.byref(self.params), cx, cy, C.c_float(z), yaw_deg, pitch_deg, roll_deg, out)
        return np.array(list(out), np.float32).reshape(4, 4)
