"""`HoughForest --test` command-line drop-in (csrc/hough_forest_main.cpp) against the reference's process interface:
flags (main.cpp:9-31), stdin pairs (HFTest.cpp:1238), `_res.txt` / `_res.png` (HFTest.cpp:1261-1311).

CPU part: flag handling, artefact validation before any device is touched, image decoding (vs cv2), no-GPU failure.
GPU part: an end-to-end run whose `_res.txt` must equal what the C ABI returns for the same frames, formatted the way
Eigen's operator<< prints a Matrix4f.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

from object_detector_6d_b200 import api, build, synth

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def cli():
    path = build.build_cli()
    assert path and os.path.exists(path)
    return path


def run(cli, args, stdin=""):
    return subprocess.run([cli] + args, input=stdin, capture_output=True, text=True, timeout=300)


def fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def eigen_format(m):
    """Eigen default IOFormat: 6 significant digits, cells right-aligned to the widest, single-space separated."""
    cells = [[f"{float(v):g}" for v in row] for row in np.asarray(m, np.float32).reshape(4, 4)]
    w = max(len(c) for r in cells for c in r)
    return "\n".join(" ".join(c.rjust(w) for c in r) for r in cells)


def test_flags_and_modes(cli):
    assert run(cli, []).returncode == 0                       # no mode flag: nothing to do (main.cpp:70-76)
    r = run(cli, ["--test"])
    assert r.returncode == 1 and "detector_options_file" in r.stderr
    r = run(cli, ["--train", "--input", "x", "--output=y", "--trees", "4"])       # main.cpp:41-51: the flag checks of --train
    assert r.returncode == 1 and "Patch Size in Voxels" in r.stderr
    r = run(cli, ["--train", "--output=y", "--patch_size_in_voxels=8", "--voxel_size_in_m=0.005"])
    assert r.returncode == 1 and "No input file specified" in r.stderr
    r = run(cli, ["--learn_transitions"])
    assert r.returncode == 2
    r = run(cli, ["--bogus_flag"])
    assert r.returncode == 1 and "unknown command line flag" in r.stderr
    assert "usage" in run(cli, ["-help"]).stdout


def test_artefacts_are_validated_before_a_device_is_selected(cli, tmp_path):
    r = run(cli, ["--test", "--detector_options_file", str(tmp_path / "missing.txt")])
    assert r.returncode == 1 and "Cannot use options file" in r.stderr
    opt = tmp_path / "opt.txt"
    synth.write_options(str(opt), str(tmp_path / "no_forest"), str(tmp_path / "w.bin"), K=2)
    r = run(cli, ["--test", f"--detector_options_file={opt}"])
    assert r.returncode == 1 and "Cannot load forest" in r.stderr
    opt.write_text(opt.read_text() + "no_such_field: 3\n")
    r = run(cli, ["--test", f"--detector_options_file={opt}"])
    assert r.returncode == 1 and "no_such_field" in r.stderr


def test_image_decoding_matches_cv2(cli, tmp_path):
    rng = np.random.default_rng(5)
    h, w = 37, 53
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    depth = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
    files = {}
    cv2.imwrite(str(tmp_path / "c.png"), bgr)                                          # RGB8, adaptive filters
    cv2.imwrite(str(tmp_path / "c9.png"), bgr, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    cv2.imwrite(str(tmp_path / "a.png"), np.dstack([bgr, grey]))                       # RGBA8
    cv2.imwrite(str(tmp_path / "g.png"), grey)                                         # grey 8 as a colour image
    cv2.imwrite(str(tmp_path / "d.png"), depth)                                        # grey 16
    cv2.imwrite(str(tmp_path / "c.ppm"), bgr)
    cv2.imwrite(str(tmp_path / "d.pgm"), depth)
    pairs = [("c.png", "d.png", bgr), ("c9.png", "d.pgm", bgr), ("a.png", "d.png", bgr), ("c.ppm", "d.png", bgr),
             ("g.png", "d.png", np.dstack([grey] * 3))]
    stdin = "".join(f"{tmp_path / a} {tmp_path / b}\n" for a, b, _ in pairs)
    r = run(cli, ["--check_inputs"], stdin)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(pairs)
    for line, (a, b, want) in zip(lines, pairs):
        # what cv::imread(rgb) / cv::imread(depth, ANYDEPTH|ANYCOLOR) give the reference (HFTest.cpp:1241-1247)
        ref_bgr = cv2.imread(str(tmp_path / a))
        ref_d = cv2.imread(str(tmp_path / b), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        assert np.array_equal(ref_bgr, want) and np.array_equal(ref_d, depth)
        assert line == (f"bgr {w}x{h} {fnv1a(ref_bgr.tobytes()):016x} depth {w}x{h} {fnv1a(ref_d.tobytes()):016x}"), (a, b)
    # unreadable / wrong files are reported and skipped, as the reference does
    (tmp_path / "junk.png").write_bytes(b"not a png")
    r = run(cli, ["--check_inputs"], f"{tmp_path / 'junk.png'} {tmp_path / 'd.png'}\n{tmp_path / 'c.png'} {tmp_path / 'c.png'}\n")
    assert r.returncode == 4 and r.stdout.count("Cannot read file") == 2


def _write_case(tmp_path, n_frames=2):
    from tests.helpers import make_case
    cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
    cs = make_case(str(tmp_path), K=2, T=2, seed=5, max_depth=10, votes_per_leaf=4, cam=cam, calib_patches=3000)
    opt = tmp_path / "detector_options.proto"
    synth.write_options(str(opt), cs["forest_dir"], cs["weights"], K=2, cam=cam, segmented=True)
    frames = []
    for i in range(n_frames):
        bgr, depth = (cs["bgr"], cs["depth"]) if i == 0 else synth.render_frame(40 + i, cam, n_objects=3)
        cv2.imwrite(str(tmp_path / f"frame{i}.png"), bgr)
        cv2.imwrite(str(tmp_path / f"frame{i}_depth.png"), depth)
        frames.append((bgr, depth))
    return cs, opt, frames


def test_no_gpu_means_no_detection(cli, tmp_path):
    """The product path fails loudly without a device (exit 3, message from the C ABI); it never falls back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    cs, opt, frames = _write_case(tmp_path, 1)
    out = tmp_path / "out"
    out.mkdir()
    r = run(cli, ["--test", f"--detector_options_file={opt}", f"--output_folder={out}"],
            f"{tmp_path / 'frame0.png'} {tmp_path / 'frame0_depth.png'}\n")
    assert r.returncode == 3 and "no CUDA device" in r.stderr
    assert not list(out.iterdir())


@pytest.mark.gpu
def test_cli_end_to_end_matches_the_c_abi(cli, tmp_path):
    cs, opt, frames = _write_case(tmp_path, 2)
    out = tmp_path / "out"
    out.mkdir()
    stdin = "".join(f"{tmp_path / f'frame{i}.png'} {tmp_path / f'frame{i}_depth.png'}\n" for i in range(len(frames)))
    stdin += f"{tmp_path / 'absent.png'} {tmp_path / 'frame0_depth.png'}\n"
    r = run(cli, ["--test", f"--detector_options_file={opt}", f"--output_dir={out}", "--stage_times"], stdin)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Detection finished. Total objects found:") == len(frames)
    assert f"Cannot read file: {tmp_path / 'absent.png'}" in r.stdout
    assert "Number of patches:" in r.stdout and "Generating Hypotheses for class: obj0" in r.stdout

    o, objs = api.parse_options(str(opt))
    det = api.Detector(options_path=str(opt), frame_size=(320, 240), device=0)
    for i, (bgr, depth) in enumerate(frames):
        hyp = det.detect(bgr, depth)
        final = ((hyp["yawpitch_score"] + hyp["roll_score"]) / np.float32(2.0)) * np.float32(o.pose_score_coeff) + \
            hyp["loc_score"] * np.float32(o.location_score_coeff)
        order = np.argsort(-final, kind="stable")
        want, seen = [], {}
        for j in order:
            c = int(hyp["cls"][j])
            if seen.get(c, 0) < objs[c]["instances"]:
                seen[c] = seen.get(c, 0) + 1
                want.append(f"{objs[c]['name']}({seen[c]}): \n{eigen_format(hyp['pose'][j])}\n\n")
        got = (out / f"frame{i}_res.txt").read_text()
        assert got == "".join(want)
        if i == 0:
            assert len(want) > 0
        img = cv2.imread(str(out / f"frame{i}_res.png"))
        assert np.array_equal(img, bgr)
    det.close()


@pytest.mark.gpu
def test_cli_writes_refined_poses_when_the_meshes_exist(cli, tmp_path):
    """With the objects' PLY files in place the CLI runs ICP + scoring + joint optimisation (HFTest.cpp:927-934, :990-994) and
    writes what hf6d_refine ranks (HFTest.cpp:1261-1303), and the overlay of MeshUtils::renderObject."""
    from tests.helpers import make_case
    cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
    cs = make_case(str(tmp_path), K=2, T=2, seed=5, max_depth=10, votes_per_leaf=4, cam=cam, calib_patches=3000)
    mesh_dir = tmp_path / "meshes"
    mesh_dir.mkdir()
    for k, (xyz, rgb) in enumerate(synth.object_models(1000, 2, spacing=0.004)):
        synth.write_ply(str(mesh_dir / f"obj{k}.ply"), xyz, rgb)
    opt = tmp_path / "detector_options.proto"
    # thresholds opened up: the forest of this case votes at random, and something should reach the output; one instance per
    # group, because with ~200 accepted random poses in one group the subsets the reference enumerates (MeshUtils.cpp:969-972)
    # are astronomically many -- hf6d_refine refuses beyond 2^20, the reference would not return
    synth.write_options(str(opt), cs["forest_dir"], cs["weights"], K=2, cam=cam, segmented=True, mesh_dir=str(mesh_dir),
                        extra="final_score_threshold: -1000\ninliers_threshold: 0\nclutter_threshold: 2\nsimilarity_coeff: 8\n"
                              "search_single_object_in_group: true")
    cv2.imwrite(str(tmp_path / "frame0.png"), cs["bgr"])
    cv2.imwrite(str(tmp_path / "frame0_depth.png"), cs["depth"])
    out = tmp_path / "out"
    out.mkdir()
    r = run(cli, ["--test", f"--detector_options_file={opt}", f"--output_folder={out}", "--stage_times"],
            f"{tmp_path / 'frame0.png'} {tmp_path / 'frame0_depth.png'}\n")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pre-ICP" not in r.stdout and "stage refine/icp" in r.stdout

    o, objs = api.parse_options(str(opt))
    det = api.Detector(options_path=str(opt), frame_size=(320, 240), device=0)
    assert abs(det.refine_params().similarity_coeff - 8.0) < 1e-6 and det.refine_params().final_score_threshold == -1000.0
    det.load_option_models()
    hyp = det.detect(cs["bgr"], cs["depth"])
    dets = det.refine(hyp)
    ranked = dets[dets["rank"] >= 0]
    ranked = ranked[np.argsort(ranked["rank"])]
    assert len(ranked) > 0
    want, seen = [], {}
    for d in ranked:
        c = int(d["cls"])
        seen[c] = seen.get(c, 0) + 1
        assert seen[c] <= objs[c]["instances"]
        want.append(f"{objs[c]['name']}({seen[c]}): \n{eigen_format(d['pose'])}\n\n")
    assert (out / "frame0_res.txt").read_text() == "".join(want)
    img = cv2.imread(str(out / "frame0_res.png"))
    changed = np.any(img != cs["bgr"], axis=2)
    assert changed.any() and np.all(img[changed][:, 1] == 255)  # renderObject with alpha = 1: green saturated, B and R untouched
    assert np.array_equal(img[..., 0], cs["bgr"][..., 0]) and np.array_equal(img[..., 2], cs["bgr"][..., 2])
    det.close()


def test_train_mode_needs_a_gpu_or_reports_the_file(cli, tmp_path):
    """--train goes to hf6d_train_forest: a missing input is reported from the C ABI; without a device it fails loudly."""
    r = run(cli, ["--train", f"--input={tmp_path / 'absent.forest'}", f"--output={tmp_path}", "--patch_size_in_voxels=8",
                  "--voxel_size_in_m=0.005"])
    assert r.returncode == 3 and "Could not open file" in r.stderr


@pytest.mark.gpu
def test_cli_trains_the_forest_the_c_abi_trains(cli, tmp_path):
    from oracle import train as T
    from tests.test_train import make_samples
    cls, dof, feat = make_samples(2000, 2, 32, seed=8)
    src = tmp_path / "patches.forest"
    T.write_patches_file(str(src), 2, cls, dof, feat)
    out = tmp_path / "cli"
    out.mkdir()
    r = run(cli, ["--train", f"--input={src}", f"--output={out}", "--trees=2", "--min_samples=20", "--tests_per_node=6",
                  "--thresholds_per_test=4", "--patch_size_in_voxels=8", "--voxel_size_in_m=0.005", "--seed=5",
                  "--threads_per_tree=8"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Tree 1 saved" in r.stdout
    ref = tmp_path / "abi"
    api.train_forest(str(ref), input_file=str(src), trees=2, min_samples=20, tests_per_node=6, thresholds_per_test=4, seed=5)
    assert (out / "forest.txt").read_text() == (ref / "forest.txt").read_text() == "2 2 32 8 0.005\n"
    for t in range(2):
        assert (out / f"tree{t}.dat").read_bytes() == (ref / f"tree{t}.dat").read_bytes()


def test_render_mode_flags(cli, tmp_path):
    r = run(cli, ["--render"])
    assert r.returncode == 1 and "--input" in r.stderr
    r = run(cli, ["--render", f"--input={tmp_path / 'absent.ply'}", f"--output={tmp_path}"])
    assert r.returncode == 3 and "not found" in r.stderr


@pytest.mark.gpu
def test_cli_renders_the_views_the_c_abi_renders(cli, tmp_path):
    """PatchGen --render (PatchGen/src/main.cpp:62-81): rgb<N>.png, depth<N>.png, pose<N>.txt per view and lighting."""
    xyz, rgb, faces = synth.object_meshes(1000, 1, 0.008)[0]
    mesh = tmp_path / "obj.ply"
    synth.write_ply_mesh(str(mesh), xyz, rgb, faces)
    out = tmp_path / "views"
    out.mkdir()
    r = run(cli, ["--render", f"--input={mesh}", f"--output={out}", "--tessel_level=1", "--inPlaceCamRot=1", "--numHeights=1",
                  "--lightings=2", "--above_z"])
    assert r.returncode == 0, r.stdout + r.stderr
    rd = api.Renderer(ply_path=str(mesh), tesselation_level=1, in_place_rotations=1, heights=1, lightings=2, above_z=1)
    n = rd.view_count()
    assert f"Total number of viewpoints: {2 * n}" in r.stdout and len(list(out.glob("rgb*.png"))) == 2 * n
    for v in (0, n - 1):
        for light in range(2):
            k = 2 * v + light
            bgr, depth = rd.render(rd.view(v), np.float32(light * 0.1))
            assert np.array_equal(cv2.imread(str(out / f"rgb{k}.png")), bgr)
            assert np.array_equal(cv2.imread(str(out / f"depth{k}.png"), cv2.IMREAD_UNCHANGED), depth)
            np.testing.assert_allclose(np.loadtxt(str(out / f"pose{k}.txt")), rd.view(v), rtol=1e-5, atol=1e-6)
    rd.close()


def test_patch_generation_flags(cli, tmp_path):
    r = run(cli, ["--genpatches", f"--output={tmp_path}"])
    assert r.returncode == 1 and "No input objects specified" in r.stderr
    r = run(cli, ["--genpatches", "--input=a,b", f"--output={tmp_path}", "--use_surface_normals"])
    assert r.returncode == 2 and "surface_normals" in r.stderr
    r = run(cli, ["--genpatches", "--binfile", "--input=a", f"--output={tmp_path}"])
    assert r.returncode == 2 and "--no_random_values" in r.stderr
    r = run(cli, ["--gentrainpatches", "--input=x", "--output=y"])
    assert r.returncode == 1 and "No caffe weights model defined" in r.stderr
    r = run(cli, ["--gentrainpatches", "--caffe_weights=w", "--output=y"])
    assert r.returncode == 1 and "No input lmdb specified" in r.stderr
    r = run(cli, ["--gentrainpatches", "--caffe_weights=w", f"--input={tmp_path / 'absent'}", "--output=y"])
    assert r.returncode == 3 and "cannot open" in r.stderr
    # no views at all: an empty but valid database and the reference's bookkeeping files (no device is touched)
    out = tmp_path / "db"
    out.mkdir()
    r = run(cli, ["--genpatches", f"--input={tmp_path / 'obj0'}, {tmp_path / 'obj1'}", f"--output={out}"])
    assert r.returncode == 0 and "using lmdb by default" in r.stdout and "Finished! Total patches: 0" in r.stdout
    assert (out / "patch_annotation_lmdb.txt").read_text() == "2\n"
    assert "Patches for obj1: 0" in (out / "patch_info_lmdb.txt").read_text()
    from oracle import patchdb as OP
    assert OP.read_lmdb(str(out)) == []
    r = run(cli, ["--genpatches", f"--input={tmp_path / 'obj0'}", f"--output={out}"])
    assert r.returncode == 1 and "Does the lmdb already exist?" in r.stderr


@pytest.mark.gpu
def test_cli_patch_database_chain(cli, tmp_path):
    """PatchGen --render -> --genpatches --lmdb -> --gentrainpatches -> HoughForest --train, every file in the reference's
    format: the database read by the independent reader, Datum bytes = the oracle's gather + normalisation of the view,
    annotation lines = the numpy restatement, training vectors = the fp32 oracle encoder of the stored patches."""
    from oracle import oracle as OR
    from oracle import patchdb as OP
    folders = []
    for o in range(2):
        xyz, rgb, faces = synth.object_meshes(1000 + o, 1, 0.008)[0]
        mesh = tmp_path / f"obj{o}.ply"
        synth.write_ply_mesh(str(mesh), xyz, rgb, faces)
        d = tmp_path / f"object{o}"
        d.mkdir()
        r = run(cli, ["--render", f"--input={mesh}", f"--output={d}", "--tessel_level=1", "--inPlaceCamRot=1", "--numHeights=1",
                      "--lightings=1", "--above_z"])
        assert r.returncode == 0, r.stdout + r.stderr
        folders.append(str(d))
    db = tmp_path / "patches"
    db.mkdir()
    r = run(cli, ["--genpatches", "--lmdb", "--input=" + ",".join(folders), f"--output={db}", "--patch_size=8", "--voxel_size=0.005",
                  "--stride=4", "--no_random_values", "--max_depth_range_in_m=0.25"])
    assert r.returncode == 0, r.stdout + r.stderr
    entries = OP.read_lmdb(str(db))
    annot = (db / "patch_annotation_lmdb.txt").read_text().split("\n")
    assert annot[0] == "2" and len(annot) == len(entries) + 2
    focal = np.float32(240.0) / np.float32(np.tan(np.float64(np.float32(45.3105) / np.float32(180.0) * np.float32(3.141592) / np.float32(2.0))))
    p = OR.default_params(W=640, H=480, stride=4, fx=float(focal), fy=float(focal), cx=319.5, cy=239.5, patch_vox=8, voxel_m=0.005,
                          max_depth_range_m=0.25, distance_threshold_m=3.0, fill_random=0, batch_size=1)
    e = 0
    per_obj = []
    for o, d in enumerate(folders):
        pid = 0
        v = 0
        while os.path.exists(f"{d}/rgb{v}.png"):
            bgr = cv2.imread(f"{d}/rgb{v}.png")
            depth = cv2.imread(f"{d}/depth{v}.png", cv2.IMREAD_UNCHANGED)
            pose = np.loadtxt(f"{d}/pose{v}.txt").astype(np.float32)
            locs = OR.scan_centres(depth, p)
            q = OR.normalise(OR.gather(bgr, depth, p, locs))
            for i in range(len(locs)):
                key, val = entries[e]
                assert key == b"%04d_%08d" % (o, pid)
                assert val == OP.datum_bytes(4, 8, 8, q[i].tobytes(), o), (o, v, i)
                x, y = locs[i]
                want = OP.annotation(640, 480, x, y, depth[y, x], pose)
                got = annot[1 + e].split(" ")
                assert got[0] == key.decode()
                np.testing.assert_allclose(np.array(got[1:], np.float64), want, rtol=2e-5, atol=2e-6)
                e += 1
                pid += 1
            v += 1
        assert v > 0
        per_obj.append(pid)
    assert e == len(entries) and min(per_obj) > 50
    info = (db / "patch_info_lmdb.txt").read_text()
    assert f"Patches for object1: {per_obj[1]}" in info and f"Total patches: {e}" in info and "Voxel size in m: 0.005" in info
    # training vectors
    layers = synth.make_encoder_weights(3)
    w = str(tmp_path / "w.bin")
    synth.write_weights_raw(w, layers)
    vec = tmp_path / "patches.forest"
    r = run(cli, ["--gentrainpatches", f"--caffe_weights={w}", "--caffe_definition=unused.prototxt", f"--input={db}", f"--output={vec}",
                  "--batch_size=1", "--gpu=0", "--encoder_mode=1"])
    assert r.returncode == 0 and f"Total patches: {e - 1}" in r.stdout, r.stdout + r.stderr
    raw = vec.read_bytes()
    rec = np.frombuffer(raw[8:], np.uint8).reshape(e - 1, 28 + 3200)
    assert np.frombuffer(raw[:8], "<i4").tolist() == [2, 800]
    assert np.array_equal(rec[:, :4].copy().view("<i4")[:, 0], [int(k[:4]) for k, _ in entries[:-1]])
    q_all = np.stack([np.frombuffer(OP.parse_datum(v)["data"], np.uint8) for _, v in entries[:-1]])
    assert np.abs(rec[:, 28:].copy().view("<f4") - OR.encode(q_all, layers)).max() < 1e-4
    dof = rec[:, 4:28].copy().view("<f4")
    np.testing.assert_array_equal(dof, np.array([[np.float32(t) for t in a.split(" ")[1:]] for a in annot[1:e]], np.float32))
    # and the forest trainer takes the file
    forest = tmp_path / "forest"
    forest.mkdir()
    r = run(cli, ["--train", f"--input={vec}", f"--output={forest}", "--trees=1", "--patch_size_in_voxels=8", "--voxel_size_in_m=0.005",
                  "--min_samples=20", "--tests_per_node=8", "--thresholds_per_test=4"])
    assert r.returncode == 0 and (forest / "tree0.dat").exists(), r.stdout + r.stderr
    assert (forest / "forest.txt").read_text().split()[:3] == ["1", "2", "800"]
