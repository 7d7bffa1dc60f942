"""CPU oracle of the view renderer: RenderViewsTesselatedSphere (PatchGen/src/render_views_tesselated_sphere_mod.cpp)
restated in numpy.  TEST INFRASTRUCTURE ONLY -- only tests/ import this; the product never does.

PARITY UNPINNED.  The reference renders through VTK / OpenGL (neither vendored nor installed): the rasteriser's fill rule, its
24-bit depth buffer, the automatic clipping range and the lighting model are the driver's.  What the reference's own code fixes
-- and what is restated line by line here -- is the camera geometry and the output contract:

  camera directions        :200-236   icosahedron, Loop-subdivided `tesselation_level` times; its vertices (use_vertices_, the
                                      default) or face centres, optionally only z >= 0 / z <= 0
  radius, heights          :186-193, :262-276   max bounding-box extent (or object_radius) + start_height + k * height_step
  view-up and in-plane rotations  :279-292, :312-322, getRotMatAroundVector :34-57
  focal point              :147-176   area-weighted centre of mass of the triangles, or 0 (render_around_0)
  projection               .h:59-61   vertical view angle 45.3105 deg -> f = H / 2 / tan(angle / 2) = 575 at 640 x 480
  outputs                  save_rendering :60-138   rgb<N>.png (white background), depth<N>.png = (int16)(depth * 1000) with
                                      0 = no surface, row 0 = top, pose<N>.txt = the 4 x 4 view transform (world -> camera,
                                      camera looking down -z, y up), surface_normals<N>.bin (surface_normals.cu at f = 575)

Choices
  V1  shading: vertex colours interpolated over the triangle (perspective-correct), times min(1, ambient + |n . v|) with n the
      triangle's geometric normal and v the direction from its centre to the camera (VTK's headlight, two-sided, diffuse 1,
      ambient = lighting * 0.1, :315).
  V2  depth is the exact perspective-correct interpolation of the vertices' camera depth (no 24-bit quantisation).
  V3  a pixel belongs to a triangle when its centre (x + 0.5, y + 0.5) lies inside or on its edges; the nearest depth wins,
      the lower triangle index on exact ties.
  V4  camera order: icosahedron vertices in the order below, then the edge points in creation order (VTK's order is its own).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def icosahedron():
    """vtkPlatonicSolidSource's solid up to vertex order: 12 vertices on the unit sphere, 20 outward triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    return v, f


def loop_subdivide(v, f):
    """One step of Loop subdivision (vtkLoopSubdivisionFilter) on a closed triangle mesh: even vertices
    (1 - n b) v + b sum(neighbours), b = (5/8 - (3/8 + cos(2 pi / n) / 4)^2) / n; edge points 3/8 (a + b) + 1/8 (c + d)."""
    nv = len(v)
    edges = {}
    opp = {}
    nbr = [set() for _ in range(nv)]
    for tri in f:
        for k in range(3):
            a, b, c = int(tri[k]), int(tri[(k + 1) % 3]), int(tri[(k + 2) % 3])
            key = (min(a, b), max(a, b))
            opp.setdefault(key, []).append(c)
            nbr[a].add(b)
            nbr[b].add(a)
    new_v = []
    for i in range(nv):
        n = len(nbr[i])
        beta = (5.0 / 8.0 - (3.0 / 8.0 + np.cos(2 * np.pi / n) / 4.0) ** 2) / n
        new_v.append((1 - n * beta) * v[i] + beta * sum(v[j] for j in sorted(nbr[i])))
    new_f = []
    for tri in f:
        mids = []
        for k in range(3):
            a, b = int(tri[k]), int(tri[(k + 1) % 3])
            key = (min(a, b), max(a, b))
            if key not in edges:
                c, d = opp[key]
                edges[key] = len(new_v)
                new_v.append(3.0 / 8.0 * (v[a] + v[b]) + 1.0 / 8.0 * (v[c] + v[d]))
            mids.append(edges[key])
        a, b, c = (int(x) for x in tri)
        ab, bc, ca = mids
        new_f += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
    return np.array(new_v), np.array(new_f, np.int64)


def camera_directions(level: int, use_vertices=True, above_z=False, below_z=False):
    v, f = icosahedron()
    for _ in range(level):
        v, f = loop_subdivide(v, f)
    pts = v if use_vertices else v[f].mean(1)
    keep = [(above_z and p[2] >= 0) or (below_z and p[2] <= 0) or (not above_z and not below_z) for p in pts]
    return pts[np.array(keep)].astype(F32)


def rot_about(vec, degrees):
    """getRotMatAroundVector (:34-57)."""
    vn = vec.astype(F32) / F32(np.linalg.norm(vec.astype(F32)))
    th = degrees / 180.0 * 3.14159265359
    c, s = np.cos(th), np.sin(th)
    x, y, z = (float(q) for q in vn)
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]], F32)


def centre_of_mass(xyz, faces):
    """:147-176: triangle centres weighted by triangle area."""
    p = xyz[faces].astype(np.float64)
    c = p.mean(1)
    a = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
    return (c * a[:, None]).sum(0) / a.sum()


def look_at(pos, focal, up):
    """vtkCamera's view transform (world -> camera): z axis from the focal point to the camera, x = up x z, y = z x x."""
    z = pos - focal
    z = z / np.linalg.norm(z)
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = x, y, z
    m[:3, 3] = -m[:3, :3] @ pos
    return m


def view_matrices(xyz, faces, level=1, in_place=24, heights=4, height_step=0.25, start_height=0.3, use_vertices=True,
                  above_z=False, below_z=False, render_around_0=False, object_radius=-1.0):
    """generateViews (:140-342): the view transforms in render order (direction, height, in-plane rotation); the reference
    renders every one `lightings` times."""
    com = np.zeros(3) if render_around_0 else centre_of_mass(xyz, faces)
    ext = xyz.max(0) - xyz.min(0)
    radius = (float(ext.max()) if object_radius < 0 else object_radius) + start_height
    out = []
    for d in camera_directions(level, use_vertices, above_z, below_z):
        for h in range(heights):
            dn = (d / F32(np.linalg.norm(d))).astype(F32)
            pos = dn.astype(np.float64) * (radius + h * height_step)
            if abs(pos[2]) > 0.00001:
                up = np.array([1, 1, (-pos[0] - pos[1]) / pos[2]], F32)
            elif abs(pos[1]) > 0.00001:
                up = np.array([1, (-pos[0] - pos[2]) / pos[1], 1], F32)
            else:
                up = np.array([0, 1, 0], F32)
            up = (up / F32(np.linalg.norm(up))).astype(F32)
            R = rot_about(dn, 360.0 / in_place)
            for _ in range(in_place):
                out.append(look_at(pos + com, com, up.astype(np.float64)))
                up = (R @ up).astype(F32)
    return np.array(out)


def render(xyz, rgb, faces, view, W=640, H=480, view_angle=45.3105, ambient=0.0):
    """One view: bgr uint8 [H, W, 3] on white, depth uint16 [H, W] millimetres (V1-V3).  Every expression is written out in the
    order csrc/render.cuh evaluates it (double, no contraction), so the two agree to the pixel."""
    f = H / 2.0 / np.tan(view_angle / 180.0 * np.pi / 2.0)
    cx, cy = W / 2.0, H / 2.0
    X, Y, Z = (xyz[:, k].astype(np.float64) for k in range(3))
    R, t = view[:3, :3].astype(np.float64), view[:3, 3].astype(np.float64)
    pc = np.stack([((X * R[k, 0] + Y * R[k, 1]) + Z * R[k, 2]) + t[k] for k in range(3)], 1)
    z = -pc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        sx = (f * pc[:, 0]) / z + cx
        sy = cy - (f * pc[:, 1]) / z
    zbuf = np.full((H, W), np.inf, np.float32)   # the z test compares floats; the lower triangle index wins ties (V3)
    depth = np.zeros((H, W))
    color = np.full((H, W, 3), 255.0)
    for ti, (a, b, c) in enumerate(faces):
        if not (z[a] > 0 and z[b] > 0 and z[c] > 0):
            continue
        x0, x1 = int(np.floor(min(sx[a], sx[b], sx[c]) - 0.5)), int(np.ceil(max(sx[a], sx[b], sx[c]) - 0.5))
        y0, y1 = int(np.floor(min(sy[a], sy[b], sy[c]) - 0.5)), int(np.ceil(max(sy[a], sy[b], sy[c]) - 0.5))
        x0, x1, y0, y1 = max(x0, 0), min(x1, W - 1), max(y0, 0), min(y1, H - 1)
        if x0 > x1 or y0 > y1:
            continue
        area = (sx[b] - sx[a]) * (sy[c] - sy[a]) - (sx[c] - sx[a]) * (sy[b] - sy[a])
        if area == 0:
            continue
        px, py = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
        w0 = ((sx[b] - px) * (sy[c] - py) - (sx[c] - px) * (sy[b] - py)) / area
        w1 = ((sx[c] - px) * (sy[a] - py) - (sx[a] - px) * (sy[c] - py)) / area
        w2 = (1.0 - w0) - w1
        inside = (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
        if not inside.any():
            continue
        q0, q1, q2 = w0 / z[a], w1 / z[b], w2 / z[c]
        d = 1.0 / ((q0 + q1) + q2)
        e1, e2 = pc[b] - pc[a], pc[c] - pc[a]
        n = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]])
        nn = np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
        cen = ((pc[a] + pc[b]) + pc[c]) / 3.0
        cl = np.sqrt((cen[0] * cen[0] + cen[1] * cen[1]) + cen[2] * cen[2])
        shade = ambient
        if nn > 0 and cl > 0:  # two-sided, as VTK lights both faces by default
            shade = min(1.0, ambient + abs(((n[0] / nn) * (-cen[0] / cl) + (n[1] / nn) * (-cen[1] / cl)) + (n[2] / nn) * (-cen[2] / cl)))
        sub_z = zbuf[y0:y1 + 1, x0:x1 + 1]
        win = inside & (d.astype(np.float32) < sub_z)
        if not win.any():
            continue
        sub_z[win] = d.astype(np.float32)[win]
        depth[y0:y1 + 1, x0:x1 + 1][win] = d[win]
        for ch in range(3):
            v = (((q0 * float(rgb[a][ch]) + q1 * float(rgb[b][ch])) + q2 * float(rgb[c][ch])) * d) * shade
            color[y0:y1 + 1, x0:x1 + 1, ch][win] = v[win]
    hit = np.isfinite(zbuf)
    dmm = np.zeros((H, W), np.uint16)
    dmm[hit] = np.clip(np.trunc(depth[hit] * 1000.0), 0, 65535).astype(np.uint16)
    bgr = np.clip(np.rint(color[..., ::-1]), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(bgr), dmm
