"""A second, independent statement of the detection path's stages in numpy / pure Python (TEST INFRASTRUCTURE).

The reference ships no golden vectors and cannot be built here, so the C oracle (oracle/hf6d_oracle.c)
is the repo's definition of "correct".  To keep a typo in that one file from silently becoming the truth, every stage is
restated here a second time, straight from the reference lines cited per function, in a different language and with a
different loop structure (vectorised over patches, sequential where the reference's float order matters).  The CPU
tests require the two statements to agree bit for bit (tolerance only for the encoder's GEMM summation order).
"""
from __future__ import annotations

import os
import struct
from collections import deque

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------ A2a scan
def scan_centres(depth, W, H, stride, ps, vox, fx, dist_thr):
    """PatchGen/src/cuda/patch_extractor.cu:372-391 -- row-major push order defines the patch index."""
    out = []
    for h in range(0, H, stride):
        row = depth[h, 0:W:stride].astype(f32)
        ws = np.arange(0, W, stride)
        ok = (row != 0) & (row / f32(1000.0) < f32(dist_thr))
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (f32(ps) * f32(vox) / (row / f32(1000.0)) * f32(fx))
        a = np.where(ok, a, 0).astype(np.int64)
        x0 = ws - a // 2
        y0 = h - a // 2
        ok &= (x0 >= 0) & (y0 >= 0) & (x0 + a - 1 < W) & (y0 + a - 1 < H)
        for w in ws[ok]:
            out.append((int(w), h))
    return np.array(out, np.int32).reshape(-1, 2)


# ------------------------------------------------------------------------------------------------ A1 + A2b gather
def _mix64(z: int) -> int:
    m = (1 << 64) - 1
    z &= m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def fill_values(seed: int, p: int):
    """The repo's counter-based stand-in for the clock64()-seeded cuRAND draws of patch_extractor.cu:236-244:
    four values k/255, k in [0, 254], keyed on (fill_seed, patch index); returns (b, g, r, d)."""
    z = _mix64(seed + 0x9E3779B97F4A7C15 * (p + 1))
    r = f32(float((z & 0xFFFF) % 255)) / f32(255.0)
    g = f32(float(((z >> 16) & 0xFFFF) % 255)) / f32(255.0)
    b = f32(float(((z >> 32) & 0xFFFF) % 255)) / f32(255.0)
    d = f32(float(((z >> 48) & 0xFFFF) % 255)) / f32(255.0)
    return b, g, r, d


def gather(bgr, depth, locs, W, H, ps, vox, fx, rng_m, fill_random=0, fill_seed=0):
    """HFTest.cpp:370-379 (texture build) + patch_extractor.cu:230-309 (one block per patch, linear texture filter:
    unnormalised coordinates, border 0, fraction in 1.8 fixed point).  Returns [P][ps][ps][4] f32, HWC."""
    tex = np.zeros((H + 2, W + 2, 4), f32)  # one texel of zero border all around
    tex[1:-1, 1:-1, :3] = bgr.astype(f32) / f32(255.0)
    tex[1:-1, 1:-1, 3] = depth.astype(f32)
    P = locs.shape[0]
    cx, cy = locs[:, 0].astype(np.int64), locs[:, 1].astype(np.int64)
    dc = tex[cy + 1, cx + 1, 3] / f32(1000.0)
    a = (f32(ps) * f32(vox) / dc * f32(fx)).astype(np.int64)
    x0 = cx - a // 2
    y0 = cy - a // 2
    step = a.astype(f32) / f32(ps)
    out = np.zeros((P, ps, ps, 4), f32)
    fills = np.zeros((P, 4), f32)
    if fill_random:
        for p in range(P):
            fills[p] = fill_values(fill_seed, p)

    def frac8(x):
        return np.floor(x * f32(256.0) + f32(0.5)) * f32(1.0 / 256.0)

    for ty in range(ps):
        v = y0.astype(f32) + f32(ty) * step
        j = np.floor(v)
        beta = frac8(v - j)
        j = j.astype(np.int64)
        for tx in range(ps):
            u = x0.astype(f32) + f32(tx) * step
            i = np.floor(u)
            alpha = frac8(u - i)
            i = i.astype(np.int64)
            ii = np.clip(i + 1, 0, W + 1)
            ii1 = np.clip(i + 2, 0, W + 1)
            jj = np.clip(j + 1, 0, H + 1)
            jj1 = np.clip(j + 2, 0, H + 1)
            w00 = (f32(1) - alpha) * (f32(1) - beta)
            w10 = alpha * (f32(1) - beta)
            w01 = (f32(1) - alpha) * beta
            w11 = alpha * beta
            S = ((w00[:, None] * tex[jj, ii] + w10[:, None] * tex[jj, ii1]) + w01[:, None] * tex[jj1, ii]) + \
                w11[:, None] * tex[jj1, ii1]
            d = S[:, 3] / f32(1000.0)
            td = np.clip((d - dc) / f32(rng_m) + f32(0.5), f32(0), f32(1))
            val = np.concatenate([S[:, :3], td[:, None]], 1)
            out[:, ty, tx, :] = np.where((d > 0)[:, None], val, fills)
    return out


# ------------------------------------------------------------------------------------------------ A2c normals variant
def surface_normals(depth, focal=575.0):
    """PatchGen/src/cuda/surface_normals.cu:11-73: central differences of the back-projected neighbours, principal
    point (W/2, H/2); zero where the pixel or one of its 4 neighbours has no depth, and on the image border."""
    H, W = depth.shape
    z = depth.astype(f32) / f32(1000.0)
    out = np.zeros((H, W, 3), f32)
    xs = np.arange(W, dtype=f32)[None, :].repeat(H, 0)
    ys = np.arange(H, dtype=f32)[:, None].repeat(W, 1)
    hw, hh, f = f32(W) / f32(2.0), f32(H) / f32(2.0), f32(focal)
    c = (slice(1, H - 1), slice(1, W - 1))
    zl, zr, zu, zd = z[1:-1, :-2], z[1:-1, 2:], z[:-2, 1:-1], z[2:, 1:-1]
    X, Y = xs[c], ys[c]
    x_left = (X - f32(1) - hw) * zl / f
    x_right = (X + f32(1) - hw) * zr / f
    x_up = (X - hw) * zu / f
    x_down = (X - hw) * zd / f
    y_left = (Y - hh) * zl / f
    y_right = (Y - hh) * zr / f
    y_up = (Y - f32(1) - hh) * zu / f
    y_down = (Y + f32(1) - hh) * zd / f
    ax, ay, az = (x_left - x_right) / f32(2), (y_left - y_right) / f32(2), (zl - zr) / f32(2)
    bx, by, bz = (x_down - x_up) / f32(2), (y_down - y_up) / f32(2), (zd - zu) / f32(2)
    nx = -(ay * bz - az * by)
    ny = -(az * bx - ax * bz)
    nz = -(ax * by - ay * bx)
    with np.errstate(divide="ignore", invalid="ignore"):
        mag = np.sqrt(nx * nx + ny * ny + nz * nz)
        n = np.stack([nx / mag, ny / mag, nz / mag], -1)
    ok = (z[c] != 0) & (zl != 0) & (zr != 0) & (zu != 0) & (zd != 0)
    out[c] = np.where(ok[..., None], n, f32(0))
    return out


def fill_values_normals(seed: int, p: int):
    """Counter-based stand-in for patch_extractor.cu:21-40: colour k/255 and a random unit normal with z >= 0."""
    golden = 0x9E3779B97F4A7C15
    z = _mix64(seed + golden * (p + 1))
    r = f32(float((z & 0xFFFF) % 255)) / f32(255.0)
    g = f32(float(((z >> 16) & 0xFFFF) % 255)) / f32(255.0)
    b = f32(float(((z >> 32) & 0xFFFF) % 255)) / f32(255.0)
    n = (f32(0), f32(0), f32(1))
    for _ in range(4):
        z2 = _mix64(z + golden)
        z3 = _mix64(z2 + golden)
        z = z3
        xr = f32(float((z2 & 0xFFFFFFFF) % 100000)) - f32(50000.0)
        yr = f32(float((z2 >> 32) % 100000)) - f32(50000.0)
        zr = f32(float((z3 & 0xFFFFFFFF) % 50000))
        norm = np.sqrt(f32(f32(xr * xr + yr * yr) + zr * zr))
        if norm != 0:
            n = (xr / norm, yr / norm, zr / norm)
            break
    return (b, g, r) + n


def gather_normals(bgr, depth, nrm, locs, W, H, ps, vox, focal, fill_random=0, fill_seed=0):
    """patch_extractor.cu:12-111 on the 7-channel texture of HFTest.cpp:333-346.  Returns [P][ps][ps][6] f32, HWC."""
    tex = np.zeros((H + 2, W + 2, 7), f32)
    tex[1:-1, 1:-1, :3] = bgr.astype(f32) / f32(255.0)
    tex[1:-1, 1:-1, 3] = depth.astype(f32)
    tex[1:-1, 1:-1, 4:] = nrm
    P = locs.shape[0]
    cx, cy = locs[:, 0].astype(np.int64), locs[:, 1].astype(np.int64)
    dc = tex[cy + 1, cx + 1, 3] / f32(1000.0)
    a = (f32(ps) * f32(vox) / dc * f32(focal)).astype(np.int64)
    x0, y0 = cx - a // 2, cy - a // 2
    step = a.astype(f32) / f32(ps)
    out = np.zeros((P, ps, ps, 6), f32)
    fills = np.zeros((P, 6), f32)
    if fill_random:
        for p in range(P):
            fills[p] = fill_values_normals(fill_seed, p)

    def frac8(x):
        return np.floor(x * f32(256.0) + f32(0.5)) * f32(1.0 / 256.0)

    for ty in range(ps):
        v = y0.astype(f32) + f32(ty) * step
        j = np.floor(v)
        beta = frac8(v - j)
        j = j.astype(np.int64)
        for tx in range(ps):
            u = x0.astype(f32) + f32(tx) * step
            i = np.floor(u)
            alpha = frac8(u - i)
            i = i.astype(np.int64)
            ii, ii1 = np.clip(i + 1, 0, W + 1), np.clip(i + 2, 0, W + 1)
            jj, jj1 = np.clip(j + 1, 0, H + 1), np.clip(j + 2, 0, H + 1)
            w00, w10 = (f32(1) - alpha) * (f32(1) - beta), alpha * (f32(1) - beta)
            w01, w11 = (f32(1) - alpha) * beta, alpha * beta
            S = ((w00[:, None] * tex[jj, ii] + w10[:, None] * tex[jj, ii1]) + w01[:, None] * tex[jj1, ii]) + \
                w11[:, None] * tex[jj1, ii1]
            d = S[:, 3] / f32(1000.0)
            x, y, z = S[:, 4], S[:, 5], S[:, 6]
            norm = np.sqrt((x * x + y * y) + z * z)
            with np.errstate(divide="ignore", invalid="ignore"):
                val = np.stack([S[:, 0], S[:, 1], S[:, 2], x / norm, y / norm, z / norm], 1)
            inside = (d > 0) & (norm > 0)
            out[:, ty, tx, :] = np.where(inside[:, None], val, fills)
    return out


def quantise_normals(patches):
    """HFTest.cpp:443-470: HWC -> CHW; colour (uchar)(v*255.0f); normals (uchar)((v/2.0 + 0.5f)*255.0f) in double."""
    P, ps = patches.shape[0], patches.shape[1]
    buf = np.ascontiguousarray(patches.transpose(0, 3, 1, 2)).reshape(P, 6, ps * ps)
    q = np.zeros((P, 6, ps * ps), np.uint8)
    col = buf[:, :3] * f32(255.0)
    q[:, :3] = (np.where(np.isnan(col), f32(0), col).astype(np.int64) & 0xFF).astype(np.uint8)
    nr = (buf[:, 3:].astype(np.float64) / 2.0 + 0.5) * 255.0
    q[:, 3:] = (np.where(np.isnan(nr), 0.0, nr).astype(np.int64) & 0xFF).astype(np.uint8)
    return q.reshape(P, 6 * ps * ps)


# ------------------------------------------------------------------------------------------------ A3 normalise
def normalise(patches):
    """HFTest.cpp:500-570: HWC -> CHW, sequential sums (means in float, variance terms in double), variance (never sqrt'ed), clip, scale, truncate.
    (unsigned char)(NaN) == 0 on x86."""
    P, ps = patches.shape[0], patches.shape[1]
    buf = np.ascontiguousarray(patches.transpose(0, 3, 1, 2)).reshape(P, 4 * ps * ps).astype(f32)
    n3, n1 = 3 * ps * ps, ps * ps
    mean_rgb = np.zeros(P, f32)
    for j in range(n3):
        mean_rgb = mean_rgb + buf[:, j] / f32(n3)
    mean_d = np.zeros(P, f32)
    for j in range(n3, n3 + n1):
        mean_d = mean_d + buf[:, j] / f32(n1)
    # `std += pow(x - mean, 2) / N`: pow(double, double) on the reference's toolchain -- the float deviation squared in
    # double, divided by (double)N, added to (double)std, narrowed back to float for every element
    f64 = np.float64
    var_rgb = np.zeros(P, f32)
    for j in range(n3):
        dlt = (buf[:, j] - mean_rgb).astype(f64)
        var_rgb = (var_rgb.astype(f64) + (dlt * dlt) / f64(n3)).astype(f32)
    var_d = np.zeros(P, f32)
    for j in range(n3, n3 + n1):
        dlt = (buf[:, j] - mean_d).astype(f64)
        var_d = (var_d.astype(f64) + (dlt * dlt) / f64(n1)).astype(f32)
    q = np.zeros((P, 4 * ps * ps), np.uint8)
    with np.errstate(divide="ignore", invalid="ignore"):
        for sl, m, var in ((slice(0, n3), mean_rgb, var_rgb), (slice(n3, n3 + n1), mean_d, var_d)):
            lim = f32(3.0) * var
            x = buf[:, sl] - m[:, None]
            x = np.where(x > lim[:, None], lim[:, None], x)
            x = np.where(x < -lim[:, None], -lim[:, None], x)
            x = x / lim[:, None]
            x = (x + f32(1.0)) * f32(0.4) + f32(0.1)
            y = x * f32(255.0)
            q[:, sl] = (np.where(np.isnan(y), f32(0), y).astype(np.int64) & 0xFF).astype(np.uint8)
    return q


# ------------------------------------------------------------------------------------------------ A4 encoder
def encode(q, layers):
    """generate_scripts.sh:424-524: three InnerProduct + Sigmoid layers on q/255 (HFTest.cpp:566, 585-596)."""
    h = q.astype(f32) / f32(255.0)
    for Wm, b in layers:
        z = h.astype(np.float64) @ Wm.T.astype(np.float64) + b.astype(np.float64)
        h = (1.0 / (1.0 + np.exp(-z))).astype(f32)
    return h


# ------------------------------------------------------------------------------------------------ A5 forest files
class Node:
    __slots__ = ("leaf", "leaf_id", "mode", "f1", "f2", "thr", "left", "right", "class_prob", "votes", "ordinal")


def read_tree(path: str, K: int):
    """HFBase.cpp:58-108: recursive pre-order; bool = 1 byte; leaves carry K floats + per class (int n, n x 6 floats)."""
    data = open(path, "rb").read()
    pos = 0
    leaves = []

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, data, pos)
        pos += struct.calcsize("<" + fmt)
        return v

    def node():
        n = Node()
        (n.leaf,) = rd("B")
        if n.leaf:
            (n.leaf_id,) = rd("i")
            n.class_prob = np.array(rd(f"{K}f"), f32)
            n.votes = []
            for _ in range(K):
                (m,) = rd("i")
                n.votes.append(np.array(rd(f"{6 * m}f"), f32).reshape(m, 6))
            n.ordinal = len(leaves)
            leaves.append(n)
        else:
            n.mode, n.f1, n.f2 = rd("iii")
            (n.thr,) = rd("f")
            n.thr = f32(n.thr)
            n.left = node()
            n.right = node()
        return n

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    root = node()
    assert pos == len(data), "trailing bytes in tree file"
    return root, leaves


def read_forest(folder: str):
    T, K, F, ps, vox = open(os.path.join(folder, "forest.txt")).read().split()
    T, K, F, ps, vox = int(T), int(K), int(F), int(ps), float(vox)
    trees = [read_tree(os.path.join(folder, f"tree{t}.dat"), K) for t in range(T)]
    return dict(T=T, K=K, F=F, ps=ps, vox=vox, trees=trees)


# ------------------------------------------------------------------------------------------------ A6 traversal
def traverse(forest, feats):
    """HFTest.cpp:144-163: val = f[f1]-f[f2] (mode 0) or f[f1] (mode 1); val < thr -> left, else (incl. NaN) right."""
    P = feats.shape[0]
    out = np.zeros((P, forest["T"]), np.int32)
    for t, (root, _) in enumerate(forest["trees"]):
        stack = [(root, np.arange(P))]
        while stack:
            n, idx = stack.pop()
            if idx.size == 0:
                continue
            if n.leaf:
                out[idx, t] = n.ordinal
                continue
            val = feats[idx, n.f1] - feats[idx, n.f2] if n.mode == 0 else feats[idx, n.f1]
            go_left = val < n.thr
            stack.append((n.left, idx[go_left]))
            stack.append((n.right, idx[~go_left]))
    return out


# ------------------------------------------------------------------------------------------------ A7 votes
def _matmul4(A, B):
    """4x4 * 4x4 in f32, k accumulated left to right (Eigen's coefficient-wise product order)."""
    C = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            acc = A[i, 0] * B[0, j]
            for k in range(1, 4):
                acc = f32(acc + A[i, k] * B[k, j])
            C[i, j] = acc
    return C


def xtion_rotmat(yaw, pitch, roll):
    """HFTest.cpp:45-80: cos/sin through <math.h> double overloads narrowed to float; corr * (Rz * Ry * Rx)."""
    import math
    cy_, sy_ = f32(math.cos(float(yaw))), f32(math.sin(float(yaw)))
    cp, sp = f32(math.cos(float(pitch))), f32(math.sin(float(pitch)))
    cr, sr = f32(math.cos(float(roll))), f32(math.sin(float(roll)))
    Rz = np.array([[cy_, -sy_, 0, 0], [sy_, cy_, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], f32)
    Ry = np.array([[cp, 0, sp, 0], [0, 1, 0, 0], [-sp, 0, cp, 0], [0, 0, 0, 1]], f32)
    Rx = np.array([[1, 0, 0, 0], [0, cr, -sr, 0], [0, sr, cr, 0], [0, 0, 0, 1]], f32)
    corr = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], f32)
    return _matmul4(corr, _matmul4(_matmul4(Rz, Ry), Rx))


def _f2i(x):
    """C float -> int conversion (truncation); NaN / overflow -> INT_MIN as cvttss2si does."""
    bad = ~np.isfinite(x) | (x >= f32(2147483648.0)) | (x < f32(-2147483648.0))
    return np.where(bad, np.int64(-2147483648), np.where(bad, 0, x).astype(np.int64))


def cast_votes(forest, leaf_ord, locs, depth, W, H, fx, fy, cx, cy, should_detect=None, want_entries=False):
    """HFTest.cpp:166-217 with get_obj_center_vote_from_6dof (:41-102) and Point3DToImage (:21-37).
    Q16 integer weights (the repo's choice C8).  Returns maps [K][H][W] uint64 (and the back-map entries)."""
    K, T = forest["K"], forest["T"]
    maps = np.zeros((K, H, W), np.uint64)
    entries = [[] for _ in range(K)]
    P = leaf_ord.shape[0]
    px, py = locs[:P, 0].astype(np.int64), locs[:P, 1].astype(np.int64)
    z = depth[py, px].astype(f32) / f32(1000.0)
    tx = (px.astype(f32) - f32(cx)) * z / f32(fx)
    ty = (py.astype(f32) - f32(cy)) * z / f32(fy)
    for t in range(T):
        _, leaves = forest["trees"][t]
        order = np.argsort(leaf_ord[:, t], kind="stable")
        sorted_ord = leaf_ord[order, t]
        bounds = np.searchsorted(sorted_ord, np.arange(len(leaves) + 1))
        for lo in range(len(leaves)):
            idx = order[bounds[lo]:bounds[lo + 1]]
            if idx.size == 0:
                continue
            leaf = leaves[lo]
            for c in range(K):
                if should_detect is not None and not should_detect[c]:
                    continue
                if not (leaf.class_prob[c] >= f32(0.5)):
                    continue
                w = np.uint64(int(f32(leaf.class_prob[c]) * f32(65536.0) + f32(0.5)))
                for vote in leaf.votes[c]:
                    R = xtion_rotmat(vote[0], vote[1], vote[2])
                    vx, vy, vz = -vote[3], -vote[4], -vote[5]
                    c3 = []
                    for r, tt in ((0, tx[idx]), (1, ty[idx]), (2, z[idx])):
                        acc = f32(f32(R[r, 0] * vx) + R[r, 1] * vy)
                        acc = f32(acc + R[r, 2] * vz)
                        c3.append(acc + tt * f32(1.0))
                    with np.errstate(divide="ignore", invalid="ignore"):
                        uu = _f2i(c3[0] / c3[2] * f32(fx) + f32(cx) + f32(0.5))
                        vv = _f2i(c3[1] / c3[2] * f32(fy) + f32(cy) + f32(0.5))
                    zero = c3[2] == 0
                    uu = np.where(zero, 0, uu)
                    vv = np.where(zero, 0, vv)
                    inb = (uu >= 0) & (uu < W) & (vv >= 0) & (vv < H)
                    np.add.at(maps[c], (vv[inb], uu[inb]), w)
                    if want_entries:
                        entries[c].extend(zip(uu.tolist(), vv.tolist(), [(t, lo)] * idx.size))
    return (maps, entries) if want_entries else maps


# ------------------------------------------------------------------------------------------------ A9 NMS
def nms(img, wx, wy):
    """HFTest.cpp:219-268, literally: two monotonic deques; emits (value, x, y); sorted by value descending (stable)."""
    rows, cols = img.shape
    res = [[None] * (cols - wx + 1) for _ in range(rows)]
    for i in range(rows):
        q = deque()
        for j in range(cols):
            if q and q[0][2] == j - wx:
                q.popleft()
            val = img[i, j]
            while q and q[-1][0] < val:
                q.pop()
            q.append((val, i, j))
            if j >= wx - 1:
                res[i][j - wx + 1] = (q[0][1], q[0][2])
    out = []
    for j in range(cols - wx + 1):
        q = deque()
        for i in range(rows - wy + 1):
            if q and q[0][1] == i - wy:
                q.popleft()
            r, c = res[i][j]
            val = img[r, c]
            while q and q[-1][0] < val:
                q.pop()
            q.append((val, r, c))
            if i >= wy - 1:
                ccx = j + wx // 2
                ccy = (i - wy + 1) + wy // 2
                if q[0][0] != 0 and q[0][1] == ccy and q[0][2] == ccx:
                    out.append((q[0][0], q[0][2], q[0][1]))
    out.sort(key=lambda t: -t[0])
    return out


def blur_reference(acc_q16, k):
    """cv::blur(map, map, Size(k,k)) on the float map the Q16 accumulator stands for (HFTest.cpp:702): normalised box
    filter, BORDER_REFLECT_101, evaluated in float64 and rounded once."""
    m = acc_q16.astype(np.float64) / 65536.0
    pad = k // 2
    mp = np.pad(m, pad, mode="reflect")
    cs = np.cumsum(np.cumsum(mp, 0), 1)
    cs = np.pad(cs, ((1, 0), (1, 0)))
    H, W = m.shape
    s = cs[k:k + H, k:k + W] - cs[0:H, k:k + W] - cs[k:k + H, 0:W] + cs[0:H, 0:W]
    return (s * (1.0 / (k * k))).astype(f32)


# ------------------------------------------------------------------------------------------------ A10 - A12 pose seeking
def _d2i(x: float) -> int:
    """C double -> int conversion (truncation toward zero)."""
    import math
    return int(math.trunc(x)) if math.isfinite(x) and abs(x) < 2147483648.0 else -2147483648


def seek_poses(forest, entries, maps, depth, W, H, fx, fy, cx, cy, *, centers_blur=15, centers_nms=40, pose_blur=35,
               pose_nms=35, max_loc=12, max_yaw_pitch=7, max_roll=3, min_loc_ratio=0.5, min_yp_ratio=0.5, should_detect=None):
    """HFTest.cpp:694-925, literally, per class: blur + NMS of the centre map, then for every kept centre the window walk
    over the back-map (`entries`, from cast_votes(want_entries=True): one (u, v, (tree, leaf)) per CAST VOTE, so a leaf that
    cast n votes on a pixel is listed n times there and all n of its votes are walked every time, :763-766), z histogram
    with the window pixel's depth standing in for the patch centre (:766-775), yaw/pitch map with its +-360 wrap copies
    (:778-791), z mode (:803-812), yaw/pitch peaks kept in [180, 540]^2 (:817-836), roll histogram from the leaves whose
    (yaw, pitch) bins lie in the blur box of the peak (:852-872), roll peaks at least 7 degrees apart (:903-921).
    Choices shared with the oracle (its header, C8 / C9 / C10): Q16 integer weights, window pixels outside the image carry
    no depth, bins outside a histogram are dropped.  Returns [(cls, cx, cy, z, yaw, pitch, roll, loc_score, yp_score,
    roll_score)] in the reference's emission order."""
    import math
    K = forest["K"]
    NB, ZB = 720, 300
    z_bin_size = f32(0.01)
    out = []
    for c in range(K):
        if should_detect is not None and not should_detect[c]:
            continue
        centres = nms(blur_reference(maps[c], centers_blur), centers_nms, centers_nms)
        back = {}
        for (u, v, tl) in entries[c]:
            back.setdefault((u, v), []).append(tl)
        n_loc = min(max_loc[c] if hasattr(max_loc, "__len__") else max_loc, len(centres))
        half = centers_nms // 2
        for k in range(n_loc):
            loc_score, ctr_x, ctr_y = centres[k]
            if f32(loc_score) / f32(centres[0][0]) < f32(min_loc_ratio):
                continue
            zacc = np.zeros(ZB, np.uint64)
            ypacc = np.zeros((NB, NB), np.uint64)
            roll_map = {}
            for row in range(ctr_y - half, ctr_y + half):
                for col in range(ctr_x - half, ctr_x + half):
                    if (col, row) not in back:
                        continue
                    inside = 0 <= row < H and 0 <= col < W
                    dpix = int(depth[row, col]) if inside else 0
                    for (t, lo) in back[(col, row)]:
                        leaf = forest["trees"][t][1][lo]
                        w = np.uint64(int(f32(leaf.class_prob[c]) * f32(65536.0) + f32(0.5)))
                        for vote in leaf.votes[c]:
                            if dpix != 0:
                                R = xtion_rotmat(vote[0], vote[1], vote[2])
                                zpix = f32(dpix) / f32(1000.0)
                                acc = f32(f32(R[2, 0] * -vote[3]) + R[2, 1] * -vote[4])
                                acc = f32(acc + R[2, 2] * -vote[5])
                                zc = f32(acc + zpix * f32(1.0))
                                zb = int(_f2i(np.array([zc / z_bin_size], f32))[0])
                                if 0 <= zb < ZB:
                                    zacc[zb] += w
                            yaw = _d2i(float(vote[0]) / math.pi * 180.0)
                            pitch = _d2i(float(vote[1]) / math.pi * 180.0)
                            sy = -1 if yaw < 0 else 1
                            sp = -1 if pitch < 0 else 1
                            for k1 in range(2):
                                for k2 in range(2):
                                    cy_ = yaw - sy * k1 * 360 + 360
                                    cp_ = pitch - sp * k2 * 360 + 360
                                    if 0 <= cy_ < NB and 0 <= cp_ < NB:
                                        ypacc[cy_, cp_] += w
                                    if k1 == 0 and k2 == 0:
                                        roll_map.setdefault((cy_, cp_), []).append((t, lo))
            zf = (zacc.astype(np.float64) / 65536.0).astype(f32).reshape(ZB, 1)
            zh = nms(zf, 1, 20)
            if not zh:
                continue
            mode_z = f32(zh[0][2]) * z_bin_size
            yph = [h for h in nms(blur_reference(ypacc, pose_blur), pose_nms, pose_nms)
                   if not (h[1] < 180 or h[1] > 540 or h[2] < 180 or h[2] > 540)]
            bh = pose_blur // 2
            for h2 in range(min(max_yaw_pitch, len(yph))):
                yp_score = f32(yph[h2][0]) / f32(yph[0][0])
                if yp_score < f32(min_yp_ratio):
                    break
                Pp, Yp = yph[h2][1], yph[h2][2]  # x = column = pitch bin, y = row = yaw bin
                racc = np.zeros(NB, np.uint64)
                for row in range(Yp - bh, Yp + bh):
                    for col in range(Pp - bh, Pp + bh):
                        for (t, lo) in roll_map.get((row, col), ()):
                            leaf = forest["trees"][t][1][lo]
                            w = np.uint64(int(f32(leaf.class_prob[c]) * f32(65536.0) + f32(0.5)))
                            for vote in leaf.votes[c]:
                                r = _d2i(float(f32(vote[2]) * f32(180.0)) / math.pi)
                                b0, b1 = r + 360, (r + 720 if r < 0 else r)
                                if 0 <= b0 < NB:
                                    racc[b0] += w
                                if 0 <= b1 < NB:
                                    racc[b1] += w
                # cv::blur(Size(1, k)) on the 720 x 1 column
                m = racc.astype(np.float64) / 65536.0
                mp = np.pad(m, pose_blur // 2, mode="reflect")
                cs = np.concatenate([[0.0], np.cumsum(mp)])
                rf = ((cs[pose_blur:pose_blur + NB] - cs[0:NB]) * (1.0 / pose_blur)).astype(f32).reshape(NB, 1)
                rh = [h for h in nms(rf, 1, pose_nms) if not (h[2] < 180 or h[2] > 540)]
                n_roll, prev = 0, f32(3.402823466e+38)
                for h in rh:
                    if n_roll >= max_roll:
                        break
                    roll_score = f32(h[0]) / f32(rh[0][0])
                    ry = h[2]
                    a = float(f32(prev / f32(180.0))) * math.pi
                    b = float(f32(f32(ry) / f32(180.0))) * math.pi
                    dot = f32(math.cos(a) * math.cos(b) + math.sin(a) * math.sin(b))
                    if n_roll == 0 or math.acos(float(dot)) / math.pi * float(f32(180.0)) > 7:
                        out.append((c, ctr_x, ctr_y, float(mode_z), Yp - 360, Pp - 360, ry - 360, float(f32(loc_score)),
                                    float(yp_score), float(roll_score)))
                        prev = f32(ry)
                        n_roll += 1
    return out
