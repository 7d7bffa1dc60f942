"""The C-ABI boundary on a machine without a GPU: libhf6d.so loads, exports every symbol include/hf6d.h declares, its
host-only format readers behave like the reference's (messages included), and every compute entry point fails loudly
instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from object_detector_6d_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "hf6d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hf6d_[a-z0-9_]+)\s*\(", text)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_every_declared_symbol_is_exported():
    L = api.load()
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/hf6d.h but not exported by libhf6d.so"
    assert sorted(api.EXPORTS) == names, set(api.EXPORTS) ^ set(names)


def test_library_is_built_for_sm_100a_and_never_links_the_oracle():
    path = api.lib_path()
    blob = open(path, "rb").read()
    assert b"sm_100a" in blob or b"sm_100" in blob
    assert b"hf6d_ref_" not in blob, "the product library must not contain the oracle"
    for f in os.listdir(os.path.join(ROOT, "object_detector_6d_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "object_detector_6d_b200", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_struct_layouts_match_the_header():
    # hf6d_params: 24 fields, one 8-byte member after 12 4-byte members -> 104 bytes with natural alignment
    assert C.sizeof(api.Params) == 104
    from oracle import oracle as O
    assert C.sizeof(O.Params) == C.sizeof(api.Params)  # parity tests hand one block to both sides
    assert api.HYP_DTYPE.itemsize == 4 * 10 + 64
    assert C.sizeof(api.ObjectOptions) == 64 + 12
    assert api.CENTRE_LIST_DTYPE.itemsize == 4 + 16 * 12
    p = api.default_params()
    assert (p.W, p.H, p.stride, p.batch_size) == (640, 480, 2, 100)
    assert (p.centers_blur_size, p.centers_nms_wsize, p.pose_blur_size, p.pose_nms_wsize) == (13, 40, 35, 35)
    assert (p.max_yaw_pitch_hypotheses, p.max_roll_hypotheses) == (7, 3)  # HFTest.h:175-183
    assert p.patch_mode == 0 and p.normals_focal == 575.0                 # HFTest.cpp:329, :356


@pytest.fixture(scope="module")
def artefacts(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("abi"))
    layers = synth.make_encoder_weights(3)
    rng = np.random.default_rng(0)
    calib = rng.random((400, 800), dtype=np.float32)
    fdir = os.path.join(d, "forest")
    stats = synth.write_forest(fdir, calib, T=3, K=6, max_depth=7, votes_per_leaf=4, seed=1)
    raw = os.path.join(d, "weights.bin")
    synth.write_weights_raw(raw, layers)
    cm = os.path.join(d, "autoencoder_iter_1.caffemodel")
    synth.write_caffemodel_v1(cm, layers)
    return dict(dir=d, forest=fdir, raw=raw, caffemodel=cm, stats=stats, layers=layers)


REFERENCE_STYLE_OPTIONS = """object_options {{
  name: "amita"
  mesh_file: "meshes/amita1_plain.ply"
  instances: 1
  nn_search_radius: 0.01
  icp_iterations: 60
  max_location_hypotheses: 12
  should_detect: true
}}
object_options {{ name: "colgate" mesh_file: "meshes/colgate.ply" max_location_hypotheses: 5 should_detect: false }}
object_options {{ name: "c" mesh_file: "c.ply" }}
object_options {{ name: "d" mesh_file: "d.ply" }}
object_options {{ name: "e" mesh_file: "e.ply" }}
object_options {{ name: "f" mesh_file: "f.ply" }}
caffe_definition: "{d}/patch_autoencoder_half.prototxt"
caffe_weights: "{w}"
forest_folder: "{f}"
num_threads: 8
stride: 2
max_depth_range_in_patch_in_m: 0.25
gpu: 0
batch_size: 100
fx: 575
fy: 575
cx: 319.5
cy: 239.5
search_single_object_instance: false
search_single_object_in_group: false
use_color_similarity: true
similarity_coeff: 10
inliers_coeff: 2.5
clutter_coeff: 1.4
location_score_coeff: 1.4
pose_score_coeff: 0.7
group_total_explain_coeff: 0.5
group_common_explain_coeff: 0.3
inliers_threshold: 0.6
clutter_threshold: 0.6
final_score_threshold: 10
cluster_eps_angle_threshold: 0.05
cluster_min_points: 5
cluster_curvature_threshold: 0.1
cluster_tolerance_near: 0.03
cluster_tolerance_far: 0.05
distance_threshold: 1.5
are_objects_segmented: false
"""


def test_options_file_as_generate_scripts_emits_it(artefacts):
    """generate_scripts.sh:541-572 + the object_options blocks of :100-140, parsed like TextFormat would."""
    path = os.path.join(artefacts["dir"], "detector_options.proto")
    with open(path, "w") as f:
        f.write(REFERENCE_STYLE_OPTIONS.format(d=artefacts["dir"], w=artefacts["caffemodel"], f=artefacts["forest"]))
    o, objs = api.parse_options(path)
    assert o.n_objects == 6 and [x["name"] for x in objs][:2] == ["amita", "colgate"]
    assert objs[1] == dict(name="colgate", should_detect=False, max_location_hypotheses=5, instances=1)
    assert objs[2]["max_location_hypotheses"] == 12 and objs[2]["should_detect"]  # proto defaults
    assert o.forest_folder.decode() == artefacts["forest"] and o.caffe_weights.decode() == artefacts["caffemodel"]
    assert (o.params.stride, o.params.batch_size, o.gpu) == (2, 100, 0)
    assert o.params.fill_random == 1  # = !are_objects_segmented (HFTest.cpp:1235)
    assert abs(o.params.cx - 319.5) < 1e-6 and abs(o.params.distance_threshold_m - 1.5) < 1e-6


def test_options_defaults_and_errors(artefacts):
    d = artefacts["dir"]

    def parse(text):
        path = os.path.join(d, "o.txt")
        with open(path, "w") as f:
            f.write(text)
        return api.parse_options(path)

    o, objs = parse(f'forest_folder: "{artefacts["forest"]}"\ncaffe_weights: "w" # comment\ncaffe_definition: "x"\n')
    assert (o.params.stride, o.gpu, o.params.batch_size, len(objs)) == (4, -1, 100, 0)  # detector_options.proto:23-27
    assert o.params.fill_random == 1
    for text, msg in (
        ('caffe_weights: "w"\n', "No forest folder specified"),            # HFTest.cpp:1166
        ('forest_folder: "f"\n', "No caffe weights model defined."),        # HFTest.cpp:1170
        ('forest_folder: "f"\ncaffe_weights: "w"\nstride: 0\n', "Stride should be more than 0"),  # HFTest.cpp:1173
        ('forest_folder: "f"\ncaffe_weights: "w"\nbogus_key: 3\n', "unknown field bogus_key"),
        ('forest_folder: "f"\ncaffe_weights: "w"\nobject_options { mesh_file: "m" }\n', "without a name"),
        ('forest_folder: "f"\ncaffe_weights: "w"\nobject_options { name: "a"\n', "unterminated"),
    ):
        with pytest.raises(api.Hf6dError) as e:
            parse(text)
        assert msg in str(e.value), (msg, str(e.value))
        assert e.value.code == -2
    with pytest.raises(api.Hf6dError) as e:
        api.parse_options(os.path.join(d, "missing.txt"))
    assert "Detector options file not found" in str(e.value)  # HFTest.cpp:1160


def test_forest_reader(artefacts):
    mi = api.inspect_forest(artefacts["forest"])
    st = artefacts["stats"]
    assert (mi.T, mi.K, mi.F, mi.patch_vox) == (3, 6, 800, 8) and abs(mi.voxel_m - 0.005) < 1e-9
    assert mi.n_leaves == sum(st["leaves"]) and mi.n_internal == sum(st["leaves"]) - 3
    assert 1 <= mi.max_depth <= 7
    with pytest.raises(api.Hf6dError) as e:
        api.inspect_forest(os.path.join(artefacts["dir"], "nope"))
    assert "forest.txt" in str(e.value)
    # truncated tree file
    bad = os.path.join(artefacts["dir"], "bad_forest")
    os.makedirs(bad, exist_ok=True)
    for n in os.listdir(artefacts["forest"]):
        data = open(os.path.join(artefacts["forest"], n), "rb").read()
        with open(os.path.join(bad, n), "wb") as f:
            f.write(data[:len(data) // 2] if n == "tree1.dat" else data)
    with pytest.raises(api.Hf6dError) as e:
        api.inspect_forest(bad)
    assert "tree1.dat" in str(e.value)


def test_weight_readers_raw_and_caffemodel(artefacts):
    assert api.inspect_weights(artefacts["raw"]) == (256, 1500, 1000, 800)
    assert api.inspect_weights(artefacts["caffemodel"]) == (256, 1500, 1000, 800)  # generate_scripts.sh:424-524
    junk = os.path.join(artefacts["dir"], "junk.bin")
    with open(junk, "wb") as f:
        f.write(b"\xff" * 100)
    with pytest.raises(api.Hf6dError):
        api.inspect_weights(junk)


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour without a CUDA device")
def test_compute_entry_points_fail_loudly_without_a_gpu(artefacts):
    """No CPU fallback: without an sm_100 device the context cannot even be created."""
    with pytest.raises(api.Hf6dError) as e:
        api.Detector(artefacts["forest"], artefacts["raw"], api.default_params(), device=0)
    assert e.value.code == -3  # HF6D_ECUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_pose_from_tuple_is_host_arithmetic():
    """hf6d_pose_from_tuple (HFTest.cpp:922-924 + MeshUtils.cpp:423-440) needs no context and no device."""
    from tests import npref
    L = api.load()
    p = api.default_params()
    out = (C.c_float * 16)()
    L.hf6d_pose_from_tuple(C.byref(p), 400, 200, C.c_float(0.83), 35, -20, 170, out)
    pose = np.array(list(out), np.float32).reshape(4, 4)
    R = npref.xtion_rotmat(np.float32(np.float32(35) / np.float32(180.0) * np.pi),
                           np.float32(np.float32(-20) / np.float32(180.0) * np.pi),
                           np.float32(np.float32(170) / np.float32(180.0) * np.pi))
    assert np.allclose(pose[:3, :3], R[:3, :3], atol=1e-6)
    assert np.allclose(pose[:3, 3], [(400 - 319.5) * 0.83 / 575, (200 - 239.5) * 0.83 / 575, 0.83], atol=1e-6)
    assert np.array_equal(pose[3], [0, 0, 0, 1])
