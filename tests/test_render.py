"""View renderer (PatchGen/src/render_views_tesselated_sphere_mod.cpp).

CPU: the oracle's camera geometry against known answers.  GPU: hf6d_render through the C ABI against oracle/render.py -- both
evaluate the same double expressions in the same order, so depth and colour must agree to the pixel; the camera poses agree to
float rounding (the reference mixes float and double there).  The reference's own pixels come from VTK / OpenGL and cannot be
pinned (oracle header)."""
import numpy as np
import pytest

from object_detector_6d_b200 import synth
from oracle import render as Rn


# ----------------------------------------------------------------------------------------------------------- CPU: the oracle
def test_camera_directions_of_the_tessellated_sphere():
    d0 = Rn.camera_directions(0)
    assert d0.shape == (12, 3) and np.allclose(np.linalg.norm(d0, axis=1), 1, atol=1e-6)
    d1 = Rn.camera_directions(1)
    assert d1.shape == (42, 3)                       # 12 vertices + 30 edge points
    n1 = d1 / np.linalg.norm(d1, axis=1)[:, None]
    assert np.allclose(n1[:12], d0, atol=1e-6)       # Loop subdivision moves the old vertices along their own direction
    v, f = Rn.icosahedron()
    mids = {tuple(sorted((int(t[k]), int(t[(k + 1) % 3])))) for t in f for k in range(3)}
    want = np.array([(v[a] + v[b]) / np.linalg.norm(v[a] + v[b]) for a, b in sorted(mids)])
    got = n1[12:]
    assert all(np.min(np.linalg.norm(want - g, axis=1)) < 1e-6 for g in got)  # edge points sit over the edge midpoints
    assert Rn.camera_directions(2).shape == (162, 3) and Rn.camera_directions(1, use_vertices=False).shape == (80, 3)
    up = Rn.camera_directions(1, above_z=True)
    assert (up[:, 2] >= 0).all() and 21 <= len(up) <= 30


def test_view_matrices_look_at_the_centre_of_mass():
    xyz, rgb, faces = synth.object_meshes(1000, 1, 0.01)[0]
    V = Rn.view_matrices(xyz, faces, level=0, in_place=4, heights=2, height_step=0.25, start_height=0.3)
    assert V.shape == (12 * 2 * 4, 4, 4)
    com = Rn.centre_of_mass(xyz, faces)
    radius = float((xyz.max(0) - xyz.min(0)).max()) + 0.3
    for i, m in enumerate(V):
        R = m[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
        c = m[:3, :3] @ com + m[:3, 3]                # the focal point lies straight ahead, on -z
        assert abs(c[0]) < 1e-6 and abs(c[1]) < 1e-6
        h = (i // 4) % 2
        assert abs(-c[2] - (radius + 0.25 * h)) < 1e-5
    # in-plane rotations turn the view-up by 360 / in_place degrees about the viewing direction
    a, b = V[0][:3, :3], V[1][:3, :3]
    rel = b @ a.T
    assert abs(np.degrees(np.arccos((np.trace(rel) - 1) / 2)) - 90.0) < 1e-3 and abs(abs(rel[2, 2]) - 1) < 1e-6


def test_oracle_render_contract():
    xyz, rgb, faces = synth.object_meshes(1000, 1, 0.008)[0]
    V = Rn.view_matrices(xyz, faces, level=0, in_place=1, heights=1)
    bgr, depth = Rn.render(xyz, rgb, faces, V[3], 160, 120, ambient=0.1)
    assert bgr.dtype == np.uint8 and depth.dtype == np.uint16 and bgr.shape == (120, 160, 3)
    hit = depth > 0
    assert 200 < hit.sum() < 160 * 120 / 2
    assert (bgr[~hit] == 255).all()                   # white background, depth 0 = no surface
    ys, xs = np.nonzero(hit)
    assert abs(xs.mean() - 80) < 12 and abs(ys.mean() - 60) < 12  # the object sits at the image centre
    radius = float((xyz.max(0) - xyz.min(0)).max()) + 0.3
    assert (radius - 0.2) * 1000 < depth[hit].min() and depth[hit].max() < (radius + 0.2) * 1000
    brighter, _ = Rn.render(xyz, rgb, faces, V[3], 160, 120, ambient=0.3)
    assert brighter[hit].astype(int).sum() > bgr[hit].astype(int).sum()


# ----------------------------------------------------------------------------------------------------------- GPU vs oracle
@pytest.mark.gpu
def test_gpu_views_and_pixels_match_the_oracle(tmp_path):
    from object_detector_6d_b200 import api
    meshes = synth.object_meshes(1000, 2, 0.006)
    for k, (xyz, rgb, faces) in enumerate(meshes):
        kw = dict(W=320, H=240, tesselation_level=1, in_place_rotations=3, heights=2, lightings=2)
        if k == 0:  # one mesh through the PLY reader
            path = str(tmp_path / "m.ply")
            synth.write_ply_mesh(path, xyz, rgb, faces)
            r = api.Renderer(ply_path=path, **kw)
            xyz = np.loadtxt(path, skiprows=13, max_rows=len(xyz), usecols=(0, 1, 2)).astype(np.float32)  # 6 decimals, as written
        else:
            r = api.Renderer(xyz, rgb, faces, **kw)
        V = Rn.view_matrices(xyz, faces, level=1, in_place=3, heights=2)
        assert r.view_count() == len(V) == 42 * 2 * 3
        for i in (0, 7, 100, len(V) - 1):
            np.testing.assert_allclose(r.view(i), V[i], atol=2e-6)
        for i, amb in ((5, 0.0), (130, 0.1), (251, 0.2)):
            pose = r.view(i)
            bgr, depth = r.render(pose, amb)
            b2, d2 = Rn.render(xyz, rgb, faces, pose, 320, 240, ambient=np.float64(np.float32(amb)))
            assert (depth > 0).sum() > 500
            assert np.array_equal(depth, d2), (k, i, int((depth != d2).sum()))
            assert np.array_equal(bgr, b2), (k, i, int((bgr != b2).any(2).sum()))
        r.close()


@pytest.mark.gpu
def test_gpu_render_feeds_the_detector_contract(tmp_path):
    """A rendered view is a frame the detection path accepts: white background, millimetre depth, f = 575 at 640 x 480."""
    from object_detector_6d_b200 import api
    xyz, rgb, faces = synth.object_meshes(1000, 1, 0.004)[0]
    r = api.Renderer(xyz, rgb, faces)
    assert r.view_count() == 42 * 4 * 24 and abs(480 / 2 / np.tan(np.radians(45.3105) / 2) - 575.0) < 0.01
    bgr, depth = r.render(r.view(10), 0.1)
    r.close()
    hit = depth > 0
    assert hit.sum() > 3000 and (bgr[~hit] == 255).all() and 300 < depth[hit].min() and depth[hit].max() < 800
    with pytest.raises(api.Hf6dError, match="not found"):
        api.Renderer(ply_path=str(tmp_path / "absent.ply"))
    with pytest.raises(api.Hf6dError, match="face index"):
        api.Renderer(xyz, rgb, np.array([[0, 1, len(xyz)]], np.int32))
