"""Seeded synthetic inputs for the detection path: frames, encoder weights, Hough forests, option files.

Nothing the reference needs at test time ships with it (no meshes, forest, weights or images -- SURVEY.md F4), so every
benchmark/parity input is synthesised here, to the *contracts* the reference defines:

* frames follow the PatchGen renderer's output contract (PatchGen/src/render_views_tesselated_sphere_mod.cpp:60-138):
  8-bit colour on a white background, depth as uint16 millimetres with 0 = no surface, pinhole camera f = 575 px at
  640x480, objects 0.6-1.2 m from the camera.  Procedural boxes / cylinders stand in for meshes/*.ply.
* forests are written in the reference's on-disk format (HoughForest/src/HFBase.cpp:4-38, 110-145): forest.txt plus
  pre-order treeN.dat files.  Split thresholds are drawn from a calibration batch of real encoder features, as the
  trainer does (HoughForest/src/HFTrain.cpp:365-386), so descents are balanced.
* encoder weights use the net's own fillers (generate_scripts.sh:448-456: gaussian std 1, sparse 40; zero bias).

Pure numpy; no GPU, no oracle.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

ENCODER_DIMS = (256, 1500, 1000, 800)  # generate_scripts.sh:424-524


# --------------------------------------------------------------------------------------------------------- frames
@dataclass
class Camera:
    W: int = 640
    H: int = 480
    fx: float = 575.0
    fy: float = 575.0
    cx: float = 319.5
    cy: float = 239.5

    @staticmethod
    def scaled(s: int) -> "Camera":
        """The 640x480 Xtion camera at s x resolution (s=2 is BASELINE config C4: 1280x960, f=1150)."""
        return Camera(640 * s, 480 * s, 575.0 * s, 575.0 * s, 320.0 * s - 0.5, 240.0 * s - 0.5)


def _rot_axis(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _texture(local, base, kind, freq):
    """Procedural albedo from object-frame coordinates (metres)."""
    if kind == 0:  # checker
        s = np.floor(local[..., 0] * freq) + np.floor(local[..., 1] * freq) + np.floor(local[..., 2] * freq)
        m = (np.mod(s, 2) * 0.55 + 0.45)[..., None]
    elif kind == 1:  # stripes
        m = (0.6 + 0.4 * np.sin(local[..., 1] * freq * 6.283))[..., None]
    else:  # blobs
        m = (0.65 + 0.35 * np.sin(local[..., 0] * freq * 5.1) * np.cos(local[..., 2] * freq * 4.3))[..., None]
    return base[None, :] * m


def render_frame(seed: int, cam: Camera = Camera(), n_objects: int = 6, table: bool = True):
    """Ray-cast a cluttered table-top scene.  Returns (bgr uint8 [H,W,3], depth_mm uint16 [H,W])."""
    rng = np.random.default_rng(seed)
    bgr, depth, _ = _render(rng, rng, cam, n_objects, table)
    return bgr, depth


def render_scene(frame_seed: int, object_seed: int, cam: Camera = Camera(), n_objects: int = 6, table: bool = True):
    """The same ray-caster with the object SET fixed by object_seed (shape, size, albedo, texture: the stand-ins for
    meshes/*.ply) and only the poses drawn from frame_seed -- frames of one scene family, as a detector is trained and
    tested on.  Returns (bgr, depth_mm, truth) with truth = dict(obj_id int8 [H,W] (-1: table / background),
    R [n,3,3] object -> camera rotations, centre [n,3] object centres in the camera frame [m])."""
    return _render(np.random.default_rng(object_seed), np.random.default_rng(frame_seed), cam, n_objects, table)


def _render(rng, rng_pose, cam, n_objects, table):
    H, W = cam.H, cam.W
    v, u = np.mgrid[0:H, 0:W]
    d = np.stack([(u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, np.ones((H, W))], -1).reshape(-1, 3)
    N = d.shape[0]
    tbest = np.full(N, np.inf)
    color = np.full((N, 3), 255.0)  # white background (renderer .cpp:260)
    obj_id = np.full(N, -1, np.int8)
    truth_R, truth_c = [], []
    light = np.array([0.3, -0.6, -0.74])
    light /= np.linalg.norm(light)

    # table plane: tilted away from the camera, ~0.7-1.2 m
    n_pl = np.array([0.0, -0.6, -0.8])
    n_pl /= np.linalg.norm(n_pl)
    p0 = np.array([0.0, 0.1, 0.95])
    if table:
        denom = d @ n_pl
        t = np.where(np.abs(denom) > 1e-9, (p0 @ n_pl) / denom, np.inf)
        ok = (t > 0.3) & (t < 1.495)
        hit = d * t[:, None]
        tex = 120 + 60 * (np.mod(np.floor(hit[:, 0] * 12) + np.floor(hit[:, 2] * 12), 2))
        tbest = np.where(ok, t, tbest)
        color = np.where(ok[:, None], np.stack([tex * 0.8, tex * 0.9, tex], -1), color)

    # plane frame for placing objects
    ex = np.array([1.0, 0, 0])
    ez = np.cross(ex, n_pl)
    ez /= np.linalg.norm(ez)
    for k in range(n_objects):
        kind = int(rng.integers(0, 2))  # 0 box, 1 cylinder
        size = rng.uniform(0.05, 0.2, 3)
        if n_objects == 1:
            centre = np.array([0.0, 0.0, 0.7])  # C1: one object, centred, 0.7 m
            R = _rot_axis(rng_pose.normal(size=3), rng_pose.uniform(0, 6.283))
        else:
            a, b = rng_pose.uniform(-0.33, 0.33), rng_pose.uniform(-0.22, 0.22)
            R = np.stack([ex, -n_pl, ez], 1) @ _rot_axis([0, 1, 0], rng_pose.uniform(0, 6.283))
            centre = p0 + a * ex + b * ez + n_pl * (size[1] * 0.5)
        truth_R.append(R)
        truth_c.append(centre)
        base = rng.uniform(40, 230, 3)
        tkind, freq = int(rng.integers(0, 3)), rng.uniform(15, 60)
        o = -(R.T @ centre)  # ray origin in object frame
        dl = d @ R  # ray directions in object frame
        if kind == 0:
            hs = size * 0.5
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = (-hs - o) / dl
                t2 = (hs - o) / dl
            tn = np.minimum(t1, t2).max(1)
            tf = np.maximum(t1, t2).min(1)
            ok = (tn <= tf) & (tn > 0.05)
            t = tn
            pl = o + dl * t[:, None]
            nl = np.zeros_like(pl)
            ax = np.argmax(np.abs(pl) / hs, 1)
            nl[np.arange(N), ax] = np.sign(pl[np.arange(N), ax])
        else:
            r, hh = size[0] * 0.5, size[1] * 0.5
            A = dl[:, 0] ** 2 + dl[:, 2] ** 2
            B = 2 * (o[0] * dl[:, 0] + o[2] * dl[:, 2])
            C = o[0] ** 2 + o[2] ** 2 - r * r
            disc = B * B - 4 * A * C
            with np.errstate(divide="ignore", invalid="ignore"):
                sq = np.sqrt(np.maximum(disc, 0))
                ts = (-B - sq) / (2 * A)
                ys = o[1] + dl[:, 1] * ts
                side_ok = (disc > 0) & (np.abs(ys) <= hh) & (ts > 0.05)
                tc = (-hh * np.sign(dl[:, 1]) - o[1]) / dl[:, 1]  # cap facing the ray
                pc = o + dl * tc[:, None]
                cap_ok = (pc[:, 0] ** 2 + pc[:, 2] ** 2 <= r * r) & (tc > 0.05)
            t = np.where(side_ok, ts, np.inf)
            t = np.where(cap_ok & (tc < t), tc, t)
            ok = np.isfinite(t)
            pl = o + dl * np.where(ok, t, 0)[:, None]
            is_cap = cap_ok & (t == tc)
            nl = np.stack([pl[:, 0], np.zeros(N), pl[:, 2]], 1) / r
            nl[is_cap] = np.array([0, 1.0, 0]) * -np.sign(dl[is_cap, 1])[:, None]
        ok &= t < tbest
        nw = nl @ R.T
        shade = np.clip(0.35 + 0.65 * np.maximum(nw @ (-light), 0), 0, 1)
        col = np.clip(_texture(pl, base, tkind, freq) * shade[:, None], 0, 255)
        tbest = np.where(ok, t, tbest)
        color = np.where(ok[:, None], col, color)
        obj_id = np.where(ok, np.int8(k), obj_id)

    depth = np.where(np.isfinite(tbest), np.rint(tbest * 1000.0), 0)  # z == t because d.z == 1
    depth = np.clip(depth, 0, 65535).astype(np.uint16).reshape(H, W)
    bgr = np.ascontiguousarray(np.clip(np.rint(color), 0, 255).astype(np.uint8).reshape(H, W, 3))
    truth = dict(obj_id=obj_id.reshape(H, W), R=np.array(truth_R), centre=np.array(truth_c))
    return bgr, depth, truth


# ------------------------------------------------------------------------------------------------ object models (PLY)
def object_models(object_seed: int, n_objects: int = 6, spacing: float = 0.002):
    """Surface point clouds of the procedural solids render_scene draws for `object_seed` -- the stand-ins for the reference's
    meshes/*.ply (vertices with colour, object frame, metres; README.md:42-51).  Replays the object draws of _render (kind,
    size, albedo, texture) and samples every face on a `spacing` lattice.  Returns a list of (xyz float32 [n,3], rgb uint8
    [n,3]); `R p + centre` with the frame's truth maps a model point into the camera frame."""
    rng = np.random.default_rng(object_seed)
    out = []
    for _ in range(n_objects):
        kind = int(rng.integers(0, 2))
        size = rng.uniform(0.05, 0.2, 3)
        base = rng.uniform(40, 230, 3)
        tkind, freq = int(rng.integers(0, 3)), rng.uniform(15, 60)
        pts = []
        if kind == 0:
            hs = size * 0.5
            for ax in range(3):
                a, b = [k for k in range(3) if k != ax]
                ga = np.arange(-hs[a], hs[a] + 1e-9, spacing)
                gb = np.arange(-hs[b], hs[b] + 1e-9, spacing)
                A, B = np.meshgrid(ga, gb, indexing="ij")
                for sgn in (-1.0, 1.0):
                    q = np.zeros(A.shape + (3,))
                    q[..., a], q[..., b], q[..., ax] = A, B, sgn * hs[ax]
                    pts.append(q.reshape(-1, 3))
        else:
            r, hh = size[0] * 0.5, size[1] * 0.5
            n_ang = max(8, int(np.ceil(2 * np.pi * r / spacing)))
            ang = np.arange(n_ang) * (2 * np.pi / n_ang)
            ys = np.arange(-hh, hh + 1e-9, spacing)
            A, Y = np.meshgrid(ang, ys, indexing="ij")
            pts.append(np.stack([r * np.cos(A), Y, r * np.sin(A)], -1).reshape(-1, 3))
            g = np.arange(-r, r + 1e-9, spacing)
            X, Z = np.meshgrid(g, g, indexing="ij")
            inside = X * X + Z * Z <= r * r
            for sgn in (-1.0, 1.0):
                pts.append(np.stack([X[inside], np.full(inside.sum(), sgn * hh), Z[inside]], -1))
        xyz = np.concatenate(pts).astype(np.float32)
        col = np.clip(_texture(xyz.astype(np.float64), base, tkind, freq) * 0.8, 0, 255)  # albedo x a mean shade; BGR like the frames
        out.append((xyz, np.ascontiguousarray(np.rint(col[:, ::-1]).astype(np.uint8))))
    return out


def object_meshes(object_seed: int, n_objects: int = 6, spacing: float = 0.004):
    """The solids of object_models as triangle meshes (vertices on a `spacing` lattice per face, two triangles per lattice
    cell, outward winding): what the reference's renderer reads from meshes/*.ply.  Returns [(xyz f32 [n,3], rgb u8 [n,3],
    faces i32 [m,3])]."""
    rng = np.random.default_rng(object_seed)
    out = []
    for _ in range(n_objects):
        kind = int(rng.integers(0, 2))
        size = rng.uniform(0.05, 0.2, 3)
        base = rng.uniform(40, 230, 3)
        tkind, freq = int(rng.integers(0, 3)), rng.uniform(15, 60)
        verts, faces = [], []

        def add_grid(P, flip):  # P: [na, nb, 3] lattice of one face
            na, nb = P.shape[:2]
            off = sum(len(v) for v in verts)
            verts.append(P.reshape(-1, 3))
            idx = off + np.arange(na * nb).reshape(na, nb)
            a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
            t = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
            faces.append(t[:, ::-1] if flip else t)

        if kind == 0:
            hs = size * 0.5
            for ax in range(3):
                a, b = [k for k in range(3) if k != ax]
                ga = np.linspace(-hs[a], hs[a], max(2, int(np.ceil(2 * hs[a] / spacing)) + 1))
                gb = np.linspace(-hs[b], hs[b], max(2, int(np.ceil(2 * hs[b] / spacing)) + 1))
                A, B = np.meshgrid(ga, gb, indexing="ij")
                for sgn in (-1.0, 1.0):
                    q = np.zeros(A.shape + (3,))
                    q[..., a], q[..., b], q[..., ax] = A, B, sgn * hs[ax]
                    n = np.cross(q[1, 0] - q[0, 0], q[0, 1] - q[0, 0])  # normal of the (a, b, c) winding
                    add_grid(q, flip=bool(n[ax] * sgn < 0))
        else:
            r, hh = size[0] * 0.5, size[1] * 0.5
            n_ang = max(8, int(np.ceil(2 * np.pi * r / spacing)))
            ang = np.linspace(0, 2 * np.pi, n_ang + 1)
            ys = np.linspace(-hh, hh, max(2, int(np.ceil(2 * hh / spacing)) + 1))
            A, Y = np.meshgrid(ang, ys, indexing="ij")
            side = np.stack([r * np.cos(A), Y, r * np.sin(A)], -1)
            n = np.cross(side[1, 0] - side[0, 0], side[0, 1] - side[0, 0])
            add_grid(side, flip=bool(np.dot(n, side[0, 0] * [1, 0, 1]) < 0))
            rad = np.linspace(0, r, max(2, int(np.ceil(r / spacing)) + 1))
            Rr, Aa = np.meshgrid(rad, ang, indexing="ij")
            for sgn in (-1.0, 1.0):
                cap = np.stack([Rr * np.cos(Aa), np.full(Rr.shape, sgn * hh), Rr * np.sin(Aa)], -1)
                n = np.cross(cap[1, 0] - cap[0, 0], cap[1, 1] - cap[1, 0])
                add_grid(cap, flip=bool(n[1] * sgn < 0))
        xyz = np.concatenate(verts).astype(np.float32)
        col = np.clip(_texture(xyz.astype(np.float64), base, tkind, freq) * 0.8, 0, 255)
        out.append((xyz, np.ascontiguousarray(np.rint(col[:, ::-1]).astype(np.uint8)), np.concatenate(faces).astype(np.int32)))
    return out


def write_ply_mesh(path: str, xyz: np.ndarray, rgb: np.ndarray, faces: np.ndarray) -> None:
    """ASCII PLY with coloured vertices and triangles, readable by vtkPLYReader and by MeshUtils::getPointCloudFromPLY."""
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
                "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(xyz), len(faces)))
        for (x, y, z), (r, g, b) in zip(xyz.tolist(), rgb.tolist()):
            f.write("%.6f %.6f %.6f %d %d %d 255\n" % (x, y, z, r, g, b))
        for a, b, c in faces.tolist():
            f.write("3 %d %d %d\n" % (a, b, c))


def write_ply(path: str, xyz: np.ndarray, rgb: np.ndarray) -> None:
    """ASCII PLY with `x y z r g b a` per vertex, the only layout MeshUtils::getPointCloudFromPLY reads
    (HoughForest/src/MeshUtils.cpp:68-114)."""
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
                "element face 0\nproperty list uchar int vertex_indices\nend_header\n" % len(xyz))
        for (x, y, z), (r, g, b) in zip(xyz.tolist(), rgb.tolist()):
            f.write("%.6f %.6f %.6f %d %d %d 255\n" % (x, y, z, r, g, b))


# ------------------------------------------------------------------------------------------------ encoder weights
def make_encoder_weights(seed: int, dims=ENCODER_DIMS):
    """[(W [out,in] f32, b [out] f32)] x3 with the net's own fillers: gaussian std 1, `sparse: 40`, zero bias
    (Caffe keeps a weight with probability sparse/num_output)."""
    rng = np.random.default_rng(seed)
    layers = []
    for i in range(3):
        n_in, n_out = dims[i], dims[i + 1]
        Wm = rng.standard_normal((n_out, n_in)).astype(np.float32)
        mask = rng.random((n_out, n_in)) < (40.0 / n_out)
        Wm = np.where(mask, Wm, np.float32(0)).astype(np.float32)
        layers.append((np.ascontiguousarray(Wm), np.zeros(n_out, np.float32)))
    return layers


def write_weights_raw(path: str, layers) -> None:
    """Raw container: 'HF6DW001', int32 n_layers, then per layer int32 out, int32 in, W[out][in] f32, b[out] f32."""
    with open(path, "wb") as f:
        f.write(b"HF6DW001")
        f.write(struct.pack("<i", len(layers)))
        for Wm, b in layers:
            f.write(struct.pack("<ii", Wm.shape[0], Wm.shape[1]))
            f.write(np.ascontiguousarray(Wm, np.float32).tobytes())
            f.write(np.ascontiguousarray(b, np.float32).tobytes())


def _pb_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_field(num: int, wt: int, payload: bytes) -> bytes:
    if wt == 2:
        return _pb_varint((num << 3) | 2) + _pb_varint(len(payload)) + payload
    return _pb_varint((num << 3) | wt) + payload


def write_caffemodel_v1(path: str, layers, names=("encode1", "encode2", "encode3")) -> None:
    """Minimal V1 .caffemodel: NetParameter{name=1, layers=2{name=4, type=5(INNER_PRODUCT=14), blobs=6{num=1,
    channels=2,height=3,width=4,data=5 packed}}} -- the subset the detector reads."""

    def blob(arr, shape4):
        p = b"".join(_pb_field(i + 1, 0, _pb_varint(s)) for i, s in enumerate(shape4))
        p += _pb_field(5, 2, np.ascontiguousarray(arr, "<f4").tobytes())
        return p

    net = _pb_field(1, 2, b"PATCHAutoencoder")
    for (Wm, b), nm in zip(layers, names):
        lay = _pb_field(2, 2, b"data") + _pb_field(3, 2, nm.encode()) + _pb_field(4, 2, nm.encode())
        lay += _pb_field(5, 0, _pb_varint(14))
        lay += _pb_field(6, 2, blob(Wm, (1, 1, Wm.shape[0], Wm.shape[1])))
        lay += _pb_field(6, 2, blob(b, (1, 1, 1, b.shape[0])))
        net += _pb_field(2, 2, lay)
    with open(path, "wb") as f:
        f.write(net)



# ------------------------------------------------------------------------------------------ calibration features
def calibration_features(bgr, depth, layers, n=8000, cam: Camera = Camera(), seed=0, patch_vox=8, voxel_m=0.005,
                         max_range=0.25, dist_thr=1.5):
    """Plausible encoder features for `n` patches of a frame, in plain numpy (nearest-neighbour sampling, fp32 matmul).

    Only used to draw split thresholds for synthetic forests from a realistic feature distribution; it is NOT a
    restatement of the product path (no bilinear filter, no sequential sums) and nothing checks against it."""
    rng = np.random.default_rng(seed)
    ys, xs = np.nonzero((depth > 0) & (depth < dist_thr * 1000))
    if ys.size == 0:
        return np.zeros((0, layers[-1][0].shape[0]), np.float32)
    sel = rng.choice(ys.size, size=min(n, ys.size), replace=False)
    return _patch_features(bgr, depth, ys[sel], xs[sel], layers, cam, patch_vox, voxel_m, max_range)[0]


def _patch_features(bgr, depth, ys, xs, layers, cam, patch_vox=8, voxel_m=0.005, max_range=0.25):
    """Approximate encoder features of the patches centred at (xs, ys); returns (features, ys, xs) of the patches that
    lie inside the image."""
    H, W = depth.shape
    dc = depth[ys, xs].astype(np.float32) / 1000.0
    a = (patch_vox * voxel_m / dc * cam.fx).astype(np.int64)
    ok = (xs - a // 2 >= 0) & (ys - a // 2 >= 0) & (xs - a // 2 + a - 1 < W) & (ys - a // 2 + a - 1 < H) & (a > 0)
    ys, xs, dc, a = ys[ok], xs[ok], dc[ok], a[ok]
    t = np.arange(patch_vox, dtype=np.float32)
    u = (xs - a // 2)[:, None] + (t[None, :] * (a[:, None] / patch_vox)).astype(np.int64)
    v = (ys - a // 2)[:, None] + (t[None, :] * (a[:, None] / patch_vox)).astype(np.int64)
    vv, uu = v[:, :, None], u[:, None, :]
    col = bgr[vv, uu].astype(np.float32) / 255.0  # [n,8,8,3]
    dd = depth[vv, uu].astype(np.float32) / 1000.0
    td = np.clip((dd - dc[:, None, None]) / max_range + 0.5, 0, 1)
    hole = dd <= 0
    col[hole] = 0
    td[hole] = 0
    x = np.concatenate([col.transpose(0, 3, 1, 2).reshape(len(a), -1), td.reshape(len(a), -1)], 1)  # CHW
    rgb, dch = x[:, :192], x[:, 192:]
    out = np.empty_like(x)
    for part, sl in ((rgb, slice(0, 192)), (dch, slice(192, 256))):
        m = part.mean(1, keepdims=True)
        var = ((part - m) ** 2).mean(1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.clip(part - m, -3 * var, 3 * var) / (3 * var)
        z = np.nan_to_num((z + 1) * 0.4 + 0.1, nan=0.0)
        out[:, sl] = np.floor(z * 255.0) / 255.0
    h = out.astype(np.float32)
    for Wm, b in layers:
        h = 1.0 / (1.0 + np.exp(-(h @ Wm.T + b)))
    return np.ascontiguousarray(h, np.float32), ys, xs


def euler_from_rotation(R):
    """(yaw, pitch, roll) with R = diag(1,-1,-1) Rz(yaw) Ry(pitch) Rx(roll): the convention of the forest's votes
    (HoughForest/src/HFTest.cpp:41-80) and of the pre-ICP pose (MeshUtils.cpp:29-59, 423-440)."""
    M = np.diag([1.0, -1.0, -1.0]) @ np.asarray(R, np.float64)
    return np.arctan2(M[1, 0], M[0, 0]), -np.arcsin(np.clip(M[2, 0], -1.0, 1.0)), np.arctan2(M[2, 1], M[2, 2])


def labelled_patches(bgr, depth, truth, layers, n=15000, cam: Camera = Camera(), seed=0, stride=2, dist_thr=1.5):
    """Training samples for a synthetic Hough forest from one rendered frame with ground truth: patches centred on object
    pixels of the stride grid, each with its (approximate) encoder feature vector, its class, and the 6-DoF vote a trainer
    would store for it (HoughForest/src/HFTrain.cpp:124, 198-202): the object's yaw / pitch / roll in the camera frame and the
    position of the patch in the object frame, so that  R (-x, -y, -z) + t  (HFTest.cpp:41-102) is the object centre.
    Returns (features [m,F] f32, cls [m] int, votes [m,6] f32)."""
    rng = np.random.default_rng(seed)
    oid = truth["obj_id"]
    grid = np.zeros(depth.shape, bool)
    grid[::stride, ::stride] = True
    ys, xs = np.nonzero(grid & (oid >= 0) & (depth > 0) & (depth < dist_thr * 1000))
    if ys.size == 0:
        return np.zeros((0, layers[-1][0].shape[0]), np.float32), np.zeros(0, np.int64), np.zeros((0, 6), np.float32)
    sel = rng.choice(ys.size, size=min(n, ys.size), replace=False)
    feats, ys, xs = _patch_features(bgr, depth, ys[sel], xs[sel], layers, cam)
    cls = oid[ys, xs].astype(np.int64)
    z = depth[ys, xs].astype(np.float64) / 1000.0
    t = np.stack([(xs - cam.cx) * z / cam.fx, (ys - cam.cy) * z / cam.fy, z], 1)
    votes = np.zeros((len(ys), 6), np.float32)
    for k in range(truth["R"].shape[0]):
        m = cls == k
        if not m.any():
            continue
        votes[m, :3] = euler_from_rotation(truth["R"][k])
        votes[m, 3:] = (t[m] - truth["centre"][k]) @ truth["R"][k]  # R^T (t - c): the patch in the object frame
    return feats, cls, votes

# ----------------------------------------------------------------------------------------------------- forests
def _build_tree(rng, feats, max_depth, min_samples, K, votes_per_leaf, prob_quantum=0, balanced=False):
    """Level-wise random tree over a calibration batch.  Returns dict of node arrays (index 0 = root)."""
    N, F = feats.shape
    is_leaf, mode, f1, f2, thr, left, right, depth_of = [], [], [], [], [], [], [], []

    def new_nodes(n, d):
        base = len(is_leaf)
        is_leaf.extend([True] * n)
        mode.extend([0] * n)
        f1.extend([0] * n)
        f2.extend([0] * n)
        thr.extend([0.0] * n)
        left.extend([-1] * n)
        right.extend([-1] * n)
        depth_of.extend([d] * n)
        return base

    new_nodes(1, 0)
    node_of = np.zeros(N, np.int64)
    active = np.array([0])
    for d in range(max_depth):
        if active.size == 0:
            break
        # samples per active node
        order = np.argsort(node_of, kind="stable")
        sorted_nodes = node_of[order]
        starts = np.searchsorted(sorted_nodes, active, "left")
        ends = np.searchsorted(sorted_nodes, active, "right")
        counts = ends - starts
        split = counts >= min_samples
        nodes = active[split]
        if nodes.size == 0:
            break
        s_, c_ = starts[split], counts[split]
        # A test is drawn the way the trainer draws it (HoughForest/src/HFTrain.cpp:365-386: random measure mode and
        # features, threshold = rand()/RAND_MAX * range + min over the node's samples).  The trainer keeps the best of many
        # such tests by information gain, so a test that cannot separate the node's samples (range 0: identical feature
        # values, which flat synthetic surfaces produce in bulk) is never kept: here the widest of a few candidates is taken,
        # and a node none of them can split stays a leaf, like the trainer's unsplittable nodes.
        seg_id = np.repeat(np.arange(nodes.size), c_)
        seg_smp = order[np.concatenate([np.arange(s, s + c) for s, c in zip(s_, c_)])]
        starts_rel = np.concatenate([[0], np.cumsum(c_)[:-1]])
        m = np.zeros(nodes.size, np.int64)
        a = np.zeros(nodes.size, np.int64)
        b = np.zeros(nodes.size, np.int64)
        vmin = np.zeros(nodes.size, np.float32)
        vmax = np.zeros(nodes.size, np.float32)
        for cand in range(4):
            m_c = rng.integers(0, 2, nodes.size)
            a_c = rng.integers(0, F, nodes.size)
            b_c = rng.integers(0, F, nodes.size)
            seg_val = np.where(m_c[seg_id] == 0, feats[seg_smp, a_c[seg_id]] - feats[seg_smp, b_c[seg_id]],
                               feats[seg_smp, a_c[seg_id]]).astype(np.float32)
            lo_c = np.minimum.reduceat(seg_val, starts_rel)
            hi_c = np.maximum.reduceat(seg_val, starts_rel)
            better = (hi_c - lo_c) > (vmax - vmin) if cand else np.ones(nodes.size, bool)
            m, a, b = np.where(better, m_c, m), np.where(better, a_c, a), np.where(better, b_c, b)
            vmin, vmax = np.where(better, lo_c, vmin), np.where(better, hi_c, vmax)
        if balanced:
            # the trainer keeps the best of tests_per_node x thresholds_per_test candidates by information gain
            # (HoughForest/src/main.cpp:16-18), which favours even splits: here the median of three sample values
            seg_val = np.where(m[seg_id] == 0, feats[seg_smp, a[seg_id]] - feats[seg_smp, b[seg_id]],
                               feats[seg_smp, a[seg_id]]).astype(np.float32)
            # ... placed half way to the next different value, so that no training sample sits ON a threshold
            pick = starts_rel[:, None] + (rng.random((nodes.size, 3)) * c_[:, None]).astype(np.int64)
            srt = np.sort(seg_val[pick], axis=1)
            other = np.where(srt[:, 2] > srt[:, 1], srt[:, 2], np.where(srt[:, 0] < srt[:, 1], srt[:, 0], vmax))
            other = np.where(other == srt[:, 1], vmin, other)
            th = (srt[:, 1].astype(np.float64) * 0.5 + other.astype(np.float64) * 0.5).astype(np.float32)
        else:
            th = (rng.random(nodes.size).astype(np.float32) * (vmax - vmin) + vmin).astype(np.float32)
        ok = (vmax > vmin) & (th > vmin)  # both children receive a sample
        nodes, m, a, b, th = nodes[ok], m[ok], a[ok], b[ok], th[ok]
        if nodes.size == 0:
            break
        base = new_nodes(2 * nodes.size, d + 1)
        lch = base + 2 * np.arange(nodes.size)
        for i, n in enumerate(nodes):
            is_leaf[n] = False
            mode[n], f1[n], f2[n], thr[n] = int(m[i]), int(a[i]), int(b[i]), float(th[i])
            left[n], right[n] = int(lch[i]), int(lch[i] + 1)
        # route samples
        lut = np.full(len(is_leaf), -1, np.int64)
        lut[nodes] = np.arange(nodes.size)
        idx = lut[node_of]
        sel = idx >= 0
        ii = idx[sel]
        smp = np.nonzero(sel)[0]
        val = np.where(m[ii] == 0, feats[smp, a[ii]] - feats[smp, b[ii]], feats[smp, a[ii]]).astype(np.float32)
        go_left = val < th[ii]
        node_of[smp] = np.where(go_left, lch[ii], lch[ii] + 1)
        active = np.concatenate([lch, lch + 1])
    n_nodes = len(is_leaf)
    is_leaf = np.array(is_leaf)
    # leaf payloads
    leaf_idx = np.nonzero(is_leaf)[0]
    nl = leaf_idx.size
    dom = rng.integers(0, K, nl)
    p_dom = rng.uniform(0.5, 1.0, nl).astype(np.float32)
    if prob_quantum:  # dyadic class probabilities: float vote sums are then exact whatever the summation order
        p_dom = (np.round(p_dom * prob_quantum) / prob_quantum).astype(np.float32)
    probs = np.zeros((nl, K), np.float32)
    if K > 1:
        rest = rng.random((nl, K)).astype(np.float32)
        rest[np.arange(nl), dom] = 0
        rest = rest / np.maximum(rest.sum(1, keepdims=True), 1e-9) * (1 - p_dom)[:, None]
        probs = rest
    probs[np.arange(nl), dom] = p_dom if K > 1 else 1.0
    sample_depth = float(np.mean(np.array(depth_of)[node_of])) if N else 0.0
    return dict(node_of=node_of, sample_depth=sample_depth, is_leaf=is_leaf, mode=np.array(mode, np.int32), f1=np.array(f1, np.int32), f2=np.array(f2, np.int32),
                thr=np.array(thr, np.float32), left=np.array(left, np.int64), right=np.array(right, np.int64),
                leaf_idx=leaf_idx, dom=dom, probs=probs, n_nodes=n_nodes, votes_per_leaf=votes_per_leaf)


def _random_votes(rng, n):
    v = np.empty((n, 6), np.float32)
    v[:, 0] = rng.uniform(-np.pi, np.pi, n)  # yaw
    v[:, 1] = rng.uniform(-np.pi / 2, np.pi / 2, n)  # pitch
    v[:, 2] = rng.uniform(-np.pi, np.pi, n)  # roll
    v[:, 3:] = rng.uniform(-0.1, 0.1, (n, 3))  # object-frame offset of the patch [m]
    return v


def _serialise_tree(rng, tree, K) -> bytes:
    """Pre-order, left first (HFBase.cpp:4-38).  leaf_id values are a random permutation: the trainer numbers leaves
    in hash-map order (HFTrain.cpp:167), not file order."""
    if "leaf_votes" in tree:
        return _serialise_trained_tree(rng, tree, K)
    out = bytearray()
    leaf_pos = {int(n): i for i, n in enumerate(tree["leaf_idx"])}
    ids = rng.permutation(len(leaf_pos)).astype(np.int32)
    V = tree["votes_per_leaf"]
    stack = [0]
    is_leaf, mode, f1, f2, thr = tree["is_leaf"], tree["mode"], tree["f1"], tree["f2"], tree["thr"]
    left, right = tree["left"], tree["right"]
    file_order = 0
    while stack:
        n = stack.pop()
        if is_leaf[n]:
            li = leaf_pos[n]
            out += struct.pack("<Bi", 1, int(ids[file_order]))
            file_order += 1
            out += tree["probs"][li].tobytes()
            for c in range(K):
                if c == tree["dom"][li]:
                    nv = V
                else:  # a few gated-out votes for minority classes, as a trained leaf has
                    nv = int(rng.integers(0, 3)) if tree["probs"][li, c] > 0.1 else 0
                out += struct.pack("<i", nv)
                if nv:
                    out += _random_votes(rng, nv).tobytes()
        else:
            out += struct.pack("<Biiif", 0, int(mode[n]), int(f1[n]), int(f2[n]), float(thr[n]))
            stack.append(int(right[n]))
            stack.append(int(left[n]))
    return bytes(out)


def _serialise_trained_tree(rng, tree, K) -> bytes:
    """As _serialise_tree, with the leaf payloads of write_trained_forest: probs [nl][K], leaf_votes[li][c] = [m,6]."""
    out = bytearray()
    leaf_pos = {int(n): i for i, n in enumerate(tree["leaf_idx"])}
    ids = rng.permutation(len(leaf_pos)).astype(np.int32)
    stack = [0]
    is_leaf, mode, f1, f2, thr = tree["is_leaf"], tree["mode"], tree["f1"], tree["f2"], tree["thr"]
    left, right = tree["left"], tree["right"]
    file_order = 0
    empty = np.zeros((0, 6), np.float32)
    while stack:
        n = stack.pop()
        if is_leaf[n]:
            li = leaf_pos[n]
            out += struct.pack("<Bi", 1, int(ids[file_order]))
            file_order += 1
            out += np.ascontiguousarray(tree["probs"][li], np.float32).tobytes()
            for c in range(K):
                v = tree["leaf_votes"][li].get(c, empty)
                out += struct.pack("<i", len(v))
                out += np.ascontiguousarray(v, np.float32).tobytes()
        else:
            out += struct.pack("<Biiif", 0, int(mode[n]), int(f1[n]), int(f2[n]), float(thr[n]))
            stack.append(int(right[n]))
            stack.append(int(left[n]))
    return bytes(out)


def write_trained_forest(folder: str, feats: np.ndarray, cls: np.ndarray, votes: np.ndarray, T: int = 4, K: int = 6,
                         max_depth: int = 20, min_samples: int = 2, views: int = 16, max_votes: int = 64, seed: int = 7,
                         angle_jitter_deg: float = 3.0, offset_jitter_m: float = 0.002, patch_vox: int = 8,
                         voxel_m: float = 0.005) -> dict:
    """A Hough forest with the payload a trained one has, from labelled samples (labelled_patches): random balanced trees
    as in write_forest, but every leaf holds the class distribution of the samples that reach it and, per class, the 6-DoF
    votes of those samples (HFTrain.cpp:124, 198-202 stores every training sample's).  Each sample stands for `views`
    training views of the same surface point from neighbouring camera poses -- what the reference's tessellated-sphere
    renderings provide -- i.e. `views` votes jittered by a few degrees / millimetres.  Unlike write_forest's uniformly random
    votes these are COHERENT: they pile up on the true object centres and poses, so the Hough maps have real modes (and the
    vote scatter has the hot spots a real forest produces).  Returns summary statistics."""
    os.makedirs(folder, exist_ok=True)
    feats = np.ascontiguousarray(feats, np.float32)
    F = feats.shape[1]
    rng = np.random.default_rng(seed)
    stats = dict(T=T, K=K, F=F, leaves=[], nodes=[], mean_depth=[], votes=[], trained=True)
    order_all = None
    for t in range(T):
        tree = _build_tree(rng, feats, max_depth, min_samples, K, 0, balanced=True)
        leaf_idx = tree["leaf_idx"]
        pos = np.full(tree["n_nodes"], -1, np.int64)
        pos[leaf_idx] = np.arange(leaf_idx.size)
        li_of = pos[tree["node_of"]]
        probs = np.zeros((leaf_idx.size, K), np.float32)
        np.add.at(probs, (li_of, cls), 1.0)
        tot = probs.sum(1, keepdims=True)
        probs = np.where(tot > 0, probs / np.maximum(tot, 1), 0).astype(np.float32)
        order = np.lexsort((cls, li_of))
        keys = li_of[order] * K + cls[order]
        bounds = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1], [True]]))
        leaf_votes = [dict() for _ in range(leaf_idx.size)]
        n_votes = 0
        for b0, b1 in zip(bounds[:-1], bounds[1:]):
            smp = order[b0:b1]
            li, c = int(li_of[smp[0]]), int(cls[smp[0]])
            v = np.repeat(votes[smp], views, axis=0)
            v[:, :3] += rng.normal(0, np.deg2rad(angle_jitter_deg), v[:, :3].shape).astype(np.float32)
            v[:, 3:] += rng.normal(0, offset_jitter_m, v[:, 3:].shape).astype(np.float32)
            v[:, 0] = (v[:, 0] + np.pi) % (2 * np.pi) - np.pi
            v[:, 1] = np.clip(v[:, 1], -np.pi / 2, np.pi / 2)
            v[:, 2] = (v[:, 2] + np.pi) % (2 * np.pi) - np.pi
            if len(v) > max_votes:
                v = v[rng.choice(len(v), max_votes, replace=False)]
            leaf_votes[li][c] = v.astype(np.float32)
            n_votes += len(v)
        tree["probs"] = probs
        tree["leaf_votes"] = leaf_votes
        with open(os.path.join(folder, f"tree{t}.dat"), "wb") as f:
            f.write(_serialise_tree(rng, tree, K))
        stats["leaves"].append(int(leaf_idx.size))
        stats["nodes"].append(int(tree["n_nodes"]))
        stats["mean_depth"].append(round(tree["sample_depth"], 2))
        stats["votes"].append(n_votes)
    with open(os.path.join(folder, "forest.txt"), "w") as f:
        f.write(f"{T} {K} {F} {patch_vox} {voxel_m:g}\n")
    return stats


def write_forest(folder: str, calib_features: np.ndarray, T: int = 4, K: int = 6, max_depth: int = 20,
                 votes_per_leaf: int = 16, seed: int = 7, min_samples: int = 2, patch_vox: int = 8,
                 voxel_m: float = 0.005, prob_quantum: int = 0) -> dict:
    """Write forest.txt + tree<t>.dat.  Returns summary statistics."""
    os.makedirs(folder, exist_ok=True)
    feats = np.ascontiguousarray(calib_features, np.float32)
    F = feats.shape[1]
    rng = np.random.default_rng(seed)
    stats = dict(T=T, K=K, F=F, leaves=[], nodes=[], mean_depth=[])
    for t in range(T):
        tree = _build_tree(rng, feats, max_depth, min_samples, K, votes_per_leaf, prob_quantum)
        with open(os.path.join(folder, f"tree{t}.dat"), "wb") as f:
            f.write(_serialise_tree(rng, tree, K))
        stats["leaves"].append(int(tree["leaf_idx"].size))
        stats["nodes"].append(int(tree["n_nodes"]))
        stats["mean_depth"].append(round(tree["sample_depth"], 2))
    with open(os.path.join(folder, "forest.txt"), "w") as f:
        f.write(f"{T} {K} {F} {patch_vox} {voxel_m:g}\n")  # HFTrain.cpp:1225-1231
    return stats


# ------------------------------------------------------------------------------------------------ options file
def write_options(path: str, forest_folder: str, weights_path: str, K: int, cam: Camera = Camera(), stride: int = 2,
                  segmented: bool = True, definition: str = "patch_autoencoder_half.prototxt", extra: str = "",
                  mesh_dir: str = "meshes") -> None:
    """Text-format DetectorOptions.Options as generate_scripts.sh:541-572 emits it."""
    lines = []
    for k in range(K):
        lines.append(f'object_options {{\n  name: "obj{k}"\n  mesh_file: "{mesh_dir}/obj{k}.ply"\n  instances: 1\n'
                     f"  nn_search_radius: 0.01\n  icp_iterations: 60\n  max_location_hypotheses: 12\n"
                     f"  should_detect: true\n}}")
    lines += [f'caffe_definition: "{definition}"', f'caffe_weights: "{weights_path}"',
              f'forest_folder: "{forest_folder}"', "num_threads: 8", f"stride: {stride}",
              "max_depth_range_in_patch_in_m: 0.25", "gpu: 0", "batch_size: 100", f"fx: {cam.fx:g}", f"fy: {cam.fy:g}",
              f"cx: {cam.cx:g}", f"cy: {cam.cy:g}", "distance_threshold: 1.5",
              f"are_objects_segmented: {'true' if segmented else 'false'}"]
    if extra:
        lines.append(extra)
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
