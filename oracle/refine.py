"""CPU oracle of the step after the Hough stage: ICP refinement, hypothesis scoring, joint optimisation.
TEST INFRASTRUCTURE ONLY -- only tests/ and bench.py's CPU legs import this; the product never does.

PARITY UNPINNED.  The reference delegates this step's arithmetic to PCL 1.7 (VoxelGrid, NormalEstimation, KdTreeFLANN,
IterativeClosestPoint + DefaultConvergenceCriteria), which is not vendored under /root/reference and is not installed here.
What follows restates (a) the reference's own code, line by line, and (b) PCL's published algorithms at the reference's call
sites, in numpy / scipy (cKDTree stands in for FLANN: the same neighbour SETS).  Every place a choice had to be made is
marked R1..R9; comparisons against the CUDA path are tolerance comparisons (tests/test_refine.py states the tolerances).

  reference                                                       here
  MeshUtils::getPointCloudFromPLY      MeshUtils.cpp:68-137       read_ply / ObjectModel
  MeshUtils::insertObjectFromPLY       MeshUtils.h:213-247        ObjectModel (VoxelGrid 5 mm + normals, NaN rows dropped)
  pcl::VoxelGrid<PointXYZRGB>          (PCL voxel_grid.hpp)        voxel_grid
  MeshUtils::get_normals_not_nan       MeshUtils.cpp:196-232      normals_not_nan (radius 0.03, viewpoint = origin)
  MeshUtils::extractEuclideanClustersSmooth  MeshUtils.cpp:245-337 smooth_clusters (the literal seed-queue walk)
  MeshUtils::setScene                  MeshUtils.cpp:340-420      Scene
  MeshUtils::icp                       MeshUtils.cpp:423-464      icp (pcl::IterativeClosestPoint restated)
  MeshUtils::evaluate_hypothesis       MeshUtils.cpp:629-793      evaluate_hypothesis
  MeshUtils::optimize_hypotheses*      MeshUtils.cpp:800-1168     optimize_hypotheses
  HFTest::test_image tail / DetectObjects  HFTest.cpp:927-994, 1261-1303   refine_frame / select_instances

Choices
  R1  radius searches return the points with squared distance STRICTLY below r^2 (FLANN's RadiusResultSet), nearest first.
  R2  VoxelGrid centroids are the exact means rounded to float (PCL sums floats in sorted-index order); colours are the
      truncated means of the 8-bit channels, as PCL's int conversion of the float mean.
  R3  normals: centred covariance of the neighbours (PCL accumulates the raw second moments in float in one pass, which is
      summation-order noise at the 1e-3 rad level), smallest eigenvector, flipped towards the viewpoint (0,0,0);
      curvature = lambda_0 / trace.  Fewer than 3 neighbours (the point itself included) -> NaN -> the point is dropped.
  R4  the reference's scene cloud holds a point at (0,0,0) for every invalid pixel (value-initialised pcl::PointXYZRGB,
      MeshUtils.cpp:346-364); they collapse into one voxel whose normal is NaN, so they never survive setScene.  Kept.
  R5  cluster walk: `fabs(acos(dot)) < eps` with dot evaluated in float and widened, NaN (dot > 1) compares false.
  R6  scene_clusters_ / scene_indices_to_cluster_ are never cleared by the reference (they grow across frames); here every
      frame starts clean, which is what the first frame of a run sees.
  R7  evaluate_hypothesis reads scene_depth_.at(row, col) without a bounds check; a model point that projects outside the
      image counts as "no scene depth" (visible).
  R8  ICP: transformation from the correspondences by the closed-form least-squares rotation (SVD with the determinant
      correction, PCL TransformationEstimationSVD); convergence exactly as DefaultConvergenceCriteria with the values
      IterativeClosestPoint installs for transformation_epsilon = 0 and euclidean_fitness_epsilon = -max:
      iterations >= max, or an exactly-identity increment, or |mse - previous mse| < 1e-12.  Fewer than 3 correspondences
      -> not converged -> the pose stays the Hough pose.
  R9  ties: std::sort / OpenMP leave the order of equal scores open; here equal scores keep their original order and the
      first best solution wins.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np
from scipy.spatial import cKDTree

F32 = np.float32
CORR = np.diag([1.0, -1.0, -1.0, 1.0]).astype(F32)  # vtk camera -> xtion frame, MeshUtils.cpp:430-436


@dataclass
class RefineParams:
    """MeshUtils' members (MeshUtils.h:115-156) with the values DetectObjects installs from the options
    (HFTest.cpp:1203-1232; defaults detector_options.proto:33-66)."""
    fx: float = 575.0
    fy: float = 575.0
    cx: float = 319.5
    cy: float = 239.5
    distance_threshold: float = 1.5
    scene_leaf: float = 0.005
    object_leaf: float = 0.005
    normals_radius: float = 0.03            # MeshUtils.cpp:198
    nn_search_radius: float = 0.01          # member default, MeshUtils.h:127 (also the divisor of the depth score)
    occlusion_threshold: float = 0.02
    similarity_coeff: float = 10.0
    inliers_coeff: float = 2.5
    clutter_coeff: float = 1.4
    location_score_coeff: float = 1.0
    pose_score_coeff: float = 0.7
    group_total_explain_coeff: float = 0.5
    group_common_explain_coeff: float = 0.3
    inliers_threshold: float = 0.6
    clutter_threshold: float = 0.6
    final_score_threshold: float = 10.0
    cluster_eps_angle: float = 0.05
    cluster_min_points: int = 5
    cluster_curvature: float = 0.1
    cluster_tolerance_near: float = 0.03
    cluster_tolerance_far: float = 0.05
    use_color_similarity: bool = True
    use_normal_similarity: bool = True
    single_object_instance: bool = False
    single_object_in_group: bool = False
    default_icp_iterations: int = 60


# ------------------------------------------------------------------------------------------------------- PLY, VoxelGrid
def read_ply(path: str):
    """MeshUtils.cpp:68-137: ASCII PLY, `x y z r g b a` per vertex.  Returns xyz float32 [n,3], rgb uint8 [n,3],
    max_center_length (largest distance from the mean vertex to a corner of the bounding box)."""
    with open(path) as f:
        tok = f.read().split()
    i, n = 0, 0
    while i < len(tok):
        if tok[i] == "element" and tok[i + 1] == "vertex":
            n = int(tok[i + 2])
        if tok[i] == "end_header":
            i += 1
            break
        i += 1
    a = np.array(tok[i:i + 7 * n], dtype=np.float64).reshape(n, 7)
    xyz = a[:, :3].astype(F32)
    rgb = a[:, 3:6].astype(np.int64).astype(np.uint8)
    return xyz, rgb, max_center_length(xyz)


def max_center_length(xyz):
    n = F32(len(xyz))
    mean = np.zeros(3, F32)
    for k in range(3):  # `mean_x += x / (float)nVertex`, sequential float adds
        mean[k] = np.add.accumulate((xyz[:, k] / n).astype(F32), dtype=F32)[-1] if len(xyz) else 0
    lo, hi = xyz.min(0), xyz.max(0)
    best = F32(0)
    for cz, cy_, cx_ in itertools.product((0, 1), (0, 1), (0, 1)):
        p = np.array([hi[0] if cx_ else lo[0], hi[1] if cy_ else lo[1], hi[2] if cz else lo[2]], F32)
        d = F32(np.sqrt(np.sum((mean - p).astype(F32) ** 2, dtype=F32)))
        best = max(best, d)
    return float(best)


def voxel_grid(xyz, rgb, leaf):
    """pcl::VoxelGrid<PointXYZRGB>::applyFilter, downsample_all_data = true: one point per occupied leaf-sized voxel, the
    centroid of the points in it, in ascending voxel-index order (i + j*dx + k*dx*dy).  Returns xyz, rgb, voxel ijk."""
    inv = F32(1.0) / F32(leaf)
    ijk = np.floor(xyz.astype(F32) * inv).astype(np.int64)
    lo = ijk.min(0)
    d = ijk.max(0) - lo + 1
    key = (ijk[:, 0] - lo[0]) + (ijk[:, 1] - lo[1]) * d[0] + (ijk[:, 2] - lo[2]) * d[0] * d[1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    start = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    cnt = np.diff(np.r_[start, len(ks)])
    sx = np.add.reduceat(xyz[order].astype(np.float64), start, axis=0)
    out_xyz = (sx / cnt[:, None]).astype(F32)  # R2
    sc = np.add.reduceat(rgb[order].astype(np.float64), start, axis=0)
    out_rgb = np.floor(sc / cnt[:, None]).astype(np.uint8)
    return out_xyz, out_rgb, ijk[order][start]


# ------------------------------------------------------------------------------------------------------- normals
def _radius_lists(tree: cKDTree, pts, r):
    """R1: neighbours with d^2 < r^2, nearest first (ties by index)."""
    cand = tree.query_ball_point(pts.astype(np.float64), r * (1 + 1e-9))
    data = tree.data
    out = []
    r2 = np.float32(r) * np.float32(r)
    for i, c in enumerate(cand):
        c = np.asarray(sorted(c), np.int64)
        d2 = np.sum((data[c].astype(F32) - pts[i].astype(F32)) ** 2, axis=1, dtype=F32) if len(c) else np.zeros(0, F32)
        keep = d2 < r2
        c, d2 = c[keep], d2[keep]
        o = np.argsort(d2, kind="stable")
        out.append((c[o], d2[o]))
    return out


def estimate_normals(xyz, radius):
    """pcl::NormalEstimation with setRadiusSearch(radius), viewpoint (0,0,0) (R3).  Returns normals float32 [n,3] (NaN rows
    where fewer than 3 neighbours) and curvature [n]."""
    n = len(xyz)
    nrm = np.full((n, 3), np.nan, F32)
    curv = np.full(n, np.nan, F32)
    if n == 0:
        return nrm, curv
    tree = cKDTree(xyz.astype(np.float64))
    for i, (idx, _) in enumerate(_radius_lists(tree, xyz, radius)):
        if len(idx) < 3:
            continue
        q = xyz[idx].astype(np.float64)
        c = q.mean(0)
        cov = (q - c).T @ (q - c) / len(idx)
        w, v = np.linalg.eigh(cov)
        nv = v[:, 0]
        if np.dot(-xyz[i].astype(np.float64), nv) < 0:  # flipNormalTowardsViewpoint
            nv = -nv
        tr = cov[0, 0] + cov[1, 1] + cov[2, 2]
        nrm[i] = nv.astype(F32)
        curv[i] = F32(abs(w[0] / tr)) if tr != 0 else F32(0)
    return nrm, curv


def normals_not_nan(xyz, rgb, radius):
    """MeshUtils.cpp:196-232: normals, then both arrays compacted to the rows with finite normals."""
    nrm, curv = estimate_normals(xyz, radius)
    ok = np.isfinite(nrm).all(1)
    return xyz[ok], rgb[ok], nrm[ok], curv[ok], ok


# ------------------------------------------------------------------------------------------------------- objects, scene
class ObjectModel:
    """MeshUtils::insertObjectFromPLY (MeshUtils.h:213-247)."""

    def __init__(self, xyz, rgb, p: RefineParams, nn_search_radius: float = -1.0, icp_iterations: int = -1, name: str = ""):
        self.full_xyz, self.full_rgb = xyz.astype(F32), rgb.astype(np.uint8)
        self.max_center_length = max_center_length(self.full_xyz)
        vx, vc, _ = voxel_grid(self.full_xyz, self.full_rgb, p.object_leaf)
        self.xyz, self.rgb, self.normals, self.curvature, _ = normals_not_nan(vx, vc, p.normals_radius)
        self.nn_search_radius = nn_search_radius
        self.icp_iterations = icp_iterations
        self.name = name

    @staticmethod
    def from_ply(path, p, **kw):
        xyz, rgb, _ = read_ply(path)
        return ObjectModel(xyz, rgb, p, **kw)


def smooth_clusters(xyz, nrm, curv, p: RefineParams):
    """MeshUtils.cpp:245-337, literally.  Returns cluster id per point (-1: none) and the list of cluster sizes."""
    n = len(xyz)
    tree = cKDTree(xyz.astype(np.float64))
    processed = np.zeros(n, bool)
    label = np.full(n, -1, np.int32)
    sizes = []
    r_near = _radius_lists(tree, xyz, p.cluster_tolerance_near)
    far = np.flatnonzero(xyz[:, 2] > F32(1.3))
    r_far = dict(zip(far.tolist(), _radius_lists(tree, xyz[far], p.cluster_tolerance_far))) if len(far) else {}
    thr = F32(p.cluster_curvature)
    for i in range(n):
        if processed[i]:
            continue
        queue = [i]
        processed[i] = True
        k = 0
        while k < len(queue):
            s = queue[k]
            k += 1
            if curv[s] > thr:
                continue
            idx, _ = r_far[s] if s in r_far else r_near[s]
            if len(idx) == 0:
                continue
            for j in idx[1:]:  # nn_indices[0] is the seed itself
                if processed[j] or curv[j] > thr:
                    continue
                dot = F32(F32(F32(nrm[s, 0] * nrm[j, 0]) + F32(nrm[s, 1] * nrm[j, 1])) + F32(nrm[s, 2] * nrm[j, 2]))
                with np.errstate(invalid="ignore"):
                    ok = abs(np.arccos(np.float64(dot))) < np.float64(F32(p.cluster_eps_angle))  # float member widened; R5: NaN -> False
                if ok:
                    processed[j] = True
                    queue.append(int(j))
        if len(queue) >= p.cluster_min_points:
            label[np.asarray(queue)] = len(sizes)
            sizes.append(len(queue))
    return label, np.asarray(sizes, np.int32)


class Scene:
    """MeshUtils::setScene (MeshUtils.cpp:340-420): organised cloud -> VoxelGrid -> normals (NaN rows dropped) -> smooth
    clusters.  bgr uint8 [H,W,3], depth uint16 [H,W] millimetres."""

    def __init__(self, bgr, depth, p: RefineParams):
        H, W = depth.shape
        self.depth = depth
        self.shape = (H, W)
        thr = F32(p.distance_threshold) * F32(1000)  # `distance_threshold *= 1000`
        valid = (depth != 0) & (depth.astype(F32) < thr)
        row, col = np.nonzero(valid)
        z = depth[row, col].astype(F32) / F32(1000.0)
        x = ((col.astype(F32) - F32(p.cx)) * z / F32(p.fx)).astype(F32)
        y = ((row.astype(F32) - F32(p.cy)) * z / F32(p.fy)).astype(F32)
        xyz = np.stack([x, y, z], 1).astype(F32)
        rgb = bgr[row, col][:, ::-1]
        n_invalid = H * W - len(row)
        if n_invalid:  # R4
            xyz = np.concatenate([xyz, np.zeros((n_invalid, 3), F32)])
            rgb = np.concatenate([rgb, np.zeros((n_invalid, 3), np.uint8)])
        vx, vc, vijk = voxel_grid(xyz, rgb, p.scene_leaf)
        self.n_voxels = len(vx)
        self.xyz, self.rgb, self.normals, self.curvature, ok = normals_not_nan(vx, vc, p.normals_radius)
        self.voxel_ijk = vijk[ok]
        self.tree = cKDTree(self.xyz.astype(np.float64))
        self.cluster, self.cluster_sizes = smooth_clusters(self.xyz, self.normals, self.curvature, p)


# ------------------------------------------------------------------------------------------------------- ICP
def rotmat_from_ypr(yaw, pitch, roll):
    """MeshUtils::get_rotmat_from_yaw_pitch_roll (MeshUtils.cpp:30-60): Rz * Ry * Rx, float cos/sin."""
    cy_, sy = F32(np.cos(np.float64(F32(yaw)))), F32(np.sin(np.float64(F32(yaw))))
    cp, sp = F32(np.cos(np.float64(F32(pitch)))), F32(np.sin(np.float64(F32(pitch))))
    cr, sr = F32(np.cos(np.float64(F32(roll)))), F32(np.sin(np.float64(F32(roll))))
    Rz = np.array([[cy_, -sy, 0, 0], [sy, cy_, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], F32)
    Ry = np.array([[cp, 0, sp, 0], [0, 1, 0, 0], [-sp, 0, cp, 0], [0, 0, 0, 1]], F32)
    Rx = np.array([[1, 0, 0, 0], [0, cr, -sr, 0], [0, sr, cr, 0], [0, 0, 0, 1]], F32)
    return (Rz @ Ry @ Rx).astype(F32)


def initial_pose(p: RefineParams, row, col, z, yaw, pitch, roll):
    """Head of MeshUtils::icp (MeshUtils.cpp:423-440)."""
    z = F32(z)
    x = (F32(col) - F32(p.cx)) * z / F32(p.fx)
    y = (F32(row) - F32(p.cy)) * z / F32(p.fy)
    m = (CORR @ rotmat_from_ypr(yaw, pitch, roll)).astype(F32)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def _transform(xyz, m):
    return (xyz.astype(F32) @ m[:3, :3].T.astype(F32) + m[:3, 3].astype(F32)).astype(F32)


def rigid_from_pairs(src, tgt):
    """Least-squares rigid transform src -> tgt (PCL TransformationEstimationSVD), in double."""
    src = src.astype(np.float64)
    tgt = tgt.astype(np.float64)
    cs, ct = src.mean(0), tgt.mean(0)
    Hm = (src - cs).T @ (tgt - ct)
    u, _, vt = np.linalg.svd(Hm)
    v = vt.T
    if np.linalg.det(u) * np.linalg.det(v) < 0:
        v[:, 2] *= -1
    R = v @ u.T
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = ct - R @ cs
    return T


def icp(scene: Scene, model: ObjectModel, p: RefineParams, pose0):
    """MeshUtils::icp (MeshUtils.cpp:442-462) around pcl::IterativeClosestPoint (R8).  Returns (pose, converged, iterations)."""
    max_dist = model.nn_search_radius if model.nn_search_radius != -1.0 else 0.0  # operator[] on a missing key: 0
    iters = model.icp_iterations if model.icp_iterations != -1 else p.default_icp_iterations
    src = _transform(model.xyz, pose0).astype(np.float64)
    final = np.eye(4)
    prev_mse = np.finfo(np.float32).max
    max_d2 = F32(max_dist) * F32(max_dist)
    n_it, converged = 0, False
    if len(scene.xyz) == 0 or len(src) == 0:
        return pose0.copy(), False, 0
    while True:
        d, j = scene.tree.query(src, k=1)
        d2 = (d * d).astype(F32)
        keep = ~(d2 > max_d2)
        if keep.sum() < 3:
            converged = False
            break
        T = rigid_from_pairs(src[keep], scene.xyz[j[keep]])
        src = src @ T[:3, :3].T + T[:3, 3]
        final = T @ final
        n_it += 1
        if n_it >= iters:
            converged = True
            break
        cos_angle = 0.5 * (T[0, 0] + T[1, 1] + T[2, 2] - 1)
        tr2 = float(T[0, 3] ** 2 + T[1, 3] ** 2 + T[2, 3] ** 2)
        if cos_angle >= 1.0 and tr2 <= 0.0:
            converged = True
            break
        mse = float(np.mean(d2[keep].astype(np.float64)))
        if abs(mse - prev_mse) < 1e-12:
            converged = True
            break
        prev_mse = mse
    pose = (final @ pose0.astype(np.float64)).astype(F32) if converged else pose0.copy()
    return pose, converged, n_it


# ------------------------------------------------------------------------------------------------------- scoring
@dataclass
class Evaluation:
    accepted: bool = False
    similarity_score: float = 0.0
    inliers_ratio: float = 0.0
    clutter_score: float = 0.0
    location_score: float = 0.0
    pose_score: float = 0.0
    final_score: float = 0.0
    visible: int = 0
    inliers: int = 0
    explained: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))


def evaluate_hypothesis(scene: Scene, model: ObjectModel, p: RefineParams, pose, location_score, pose_score) -> Evaluation:
    """MeshUtils::evaluate_hypothesis (MeshUtils.cpp:629-793)."""
    ev = Evaluation(location_score=float(location_score), pose_score=float(pose_score))
    ev.explained = np.zeros(len(scene.xyz), bool)
    if pose[2, 3] > F32(1.5):
        return ev
    pts = _transform(model.xyz, pose)
    pts, rgb, nrm, _, _ = normals_not_nan(pts, model.rgb, p.normals_radius)
    H, W = scene.shape
    with np.errstate(divide="ignore", invalid="ignore"):
        rowf = pts[:, 1] * F32(p.fy) / pts[:, 2] + F32(p.cy)
        colf = pts[:, 0] * F32(p.fx) / pts[:, 2] + F32(p.cx)
    ok = np.isfinite(rowf) & np.isfinite(colf) & (np.abs(rowf) < 2e9) & (np.abs(colf) < 2e9)
    row = np.where(ok, np.trunc(np.where(ok, rowf, 0)), -1).astype(np.int64)
    col = np.where(ok, np.trunc(np.where(ok, colf, 0)), -1).astype(np.int64)
    inside = ok & (row >= 0) & (row < H) & (col >= 0) & (col < W)
    sd = np.zeros(len(pts), F32)
    sd[inside] = scene.depth[row[inside], col[inside]].astype(F32) / F32(1000.0)  # R7
    visible = np.flatnonzero((sd == 0) | (pts[:, 2] < sd + F32(p.occlusion_threshold)))
    ev.visible = len(visible)
    radius = model.nn_search_radius if model.nn_search_radius != -1.0 else p.nn_search_radius
    n_cl = len(scene.cluster_sizes)
    scene_cl = np.zeros(n_cl, np.int64)
    model_cl = np.zeros(n_cl, np.int64)
    inliers = not_in_cluster = 0
    sim = F32(0)
    lists = _radius_lists(scene.tree, pts[visible], radius) if len(visible) and len(scene.xyz) else []
    for vi, (idx, d2) in zip(visible, lists):
        if len(idx) == 0:
            continue
        depth_score = F32(1.0) - d2 / F32(p.nn_search_radius)  # squared distance over the GLOBAL radius, as written
        if p.use_normal_similarity:
            sn = scene.normals[idx]
            ns = (sn[:, 0] * nrm[vi, 0] + sn[:, 1] * nrm[vi, 1] + sn[:, 2] * nrm[vi, 2]).astype(F32) / F32(2.0) + F32(0.5)
        else:
            ns = np.ones(len(idx), F32)
        sc = scene.rgb[idx].astype(F32)
        oc = rgb[vi].astype(F32)
        cdiff = np.maximum(np.maximum(np.abs(sc[:, 0] - oc[0]), np.abs(sc[:, 1] - oc[1])), np.abs(sc[:, 2] - oc[2]))
        cs = (1.0 - cdiff.astype(np.float64) / 255.0).astype(F32)
        score = ((ns + depth_score + cs) / F32(3.0) if p.use_color_similarity else (ns + depth_score) / F32(2.0)).astype(F32)
        best, best_id = F32(0), -1
        for k in range(len(idx)):
            if score[k] > best:
                best, best_id = score[k], int(idx[k])
        new = idx[~ev.explained[idx]]
        ev.explained[new] = True
        cl = scene.cluster[new]
        np.add.at(scene_cl, cl[cl >= 0], 1)
        sim = F32(sim + best)
        if best_id >= 0 and scene.cluster[best_id] >= 0:
            model_cl[scene.cluster[best_id]] += 1
        else:
            not_in_cluster += 1
        inliers += 1
    ev.inliers = inliers
    with np.errstate(divide="ignore", invalid="ignore"):
        ev.similarity_score = float(F32(sim) / F32(inliers))
        if inliers - not_in_cluster <= 0:
            ev.clutter_score = 1.0
        else:
            c = F32(0)
            for k in range(n_cl):
                if scene_cl[k] != 0:
                    non = int(scene.cluster_sizes[k]) - int(scene_cl[k])
                    c = F32(c + F32(F32(non) / F32(scene.cluster_sizes[k])) * F32(F32(model_cl[k]) / F32(inliers - not_in_cluster)))
            ev.clutter_score = float(c)
        ev.inliers_ratio = float(F32(inliers) / F32(len(visible)))
    fs = (F32(ev.similarity_score) * F32(p.similarity_coeff) + F32(ev.inliers_ratio) * F32(p.inliers_coeff)
          - F32(ev.clutter_score) * F32(p.clutter_coeff) + F32(ev.pose_score) * F32(p.pose_score_coeff)
          + F32(ev.location_score) * F32(p.location_score_coeff))
    ev.final_score = float(fs)
    if ev.clutter_score > p.clutter_threshold or ev.inliers_ratio < p.inliers_threshold:
        ev.accepted = False
    else:
        ev.accepted = bool(fs > F32(p.final_score_threshold))
    return ev


# ------------------------------------------------------------------------------------------------------- optimisation
def _next_solution(sol, group, excl, single_in_group):
    """MeshUtils::get_next_solution_vector (MeshUtils.cpp:800-861)."""
    n = len(sol)
    if single_in_group:
        i = 0
        while i < n and not sol[i]:
            i += 1
        if i == n:
            sol[0] = True
            return True
        if i == n - 1:
            return False
        sol[i] = False
        sol[i + 1] = True
        return True
    pos = 0
    while True:
        found = False
        for i in range(pos, n):
            if sol[i]:
                sol[i] = False
            else:
                sol[i] = True
                pos = i
                found = True
                break
        if not found:
            return False
        valid = True
        for i in range(pos + 1, n):
            if sol[i] and (group[pos], group[i]) in excl:
                valid = False
                break
        if valid:
            return True
        sol[pos] = False
        pos += 1


def optimize_hypotheses(hyps, models, p: RefineParams, max_solutions: int = 1 << 20):
    """MeshUtils::optimize_hypotheses (MeshUtils.cpp:864-1168).  hyps: list of (cls, pose, Evaluation), already sorted by
    final score (HFTest.cpp:990).  Returns the indices of the chosen hypotheses, in the reference's order."""
    n = len(hyps)
    if p.single_object_instance:  # optimize_hypotheses_single, MeshUtils.cpp:1086-1155
        if n == 0:
            return []
        best, best_i = F32(0), -1
        for i, (_, _, ev) in enumerate(hyps):
            fs = (F32(ev.similarity_score) * F32(p.similarity_coeff) + F32(ev.inliers_ratio) * F32(p.inliers_coeff)
                  - F32(ev.clutter_score) * F32(p.clutter_coeff) + F32(ev.pose_score) * F32(p.pose_score_coeff))
            ev.final_score = float(fs)
            if fs > best:
                best, best_i = fs, i
        return [best_i]
    excl = set()
    for i in range(n):
        for j in range(i + 1, n):
            ci, cj = hyps[i][1][:3, 3].astype(F32), hyps[j][1][:3, 3].astype(F32)
            dist = F32(np.sqrt(np.sum((ci - cj) ** 2, dtype=F32)))
            if dist < F32(models[hyps[i][0]].max_center_length) + F32(models[hyps[j][0]].max_center_length):
                ei, ej = hyps[i][2].explained, hyps[j][2].explained
                common = int(np.count_nonzero(ei & ej))
                ti, tj = int(np.count_nonzero(ei)), int(np.count_nonzero(ej))
                with np.errstate(divide="ignore", invalid="ignore"):
                    if F32(common) / F32(ti) > F32(0.4) or F32(common) / F32(tj) > F32(0.4):
                        excl.add((i, j))
                        excl.add((j, i))
    visited = [False] * n
    groups = []
    for i in range(n):
        if visited[i]:
            continue
        visited[i] = True
        q = [i]
        groups.append([i])
        while q:
            cur = q.pop(0)
            for j in range(n):
                if not visited[j] and (cur, j) in excl:
                    groups[-1].append(j)
                    q.append(j)
                    visited[j] = True
    result = []
    for g in groups:
        member = np.stack([hyps[i][2].explained for i in g])  # [len(g), S]
        occupied = member.any(0)
        total_by_group = int(np.count_nonzero(occupied))
        sol = [False] * len(g)
        best_sol, best_score = [False] * len(g), F32(0)
        count = 0
        while _next_solution(sol, g, excl, p.single_object_in_group):
            count += 1
            if count > max_solutions:
                raise RuntimeError("solution space too large")
            cur = [s for s in range(len(g)) if sol[s]]
            k = member[cur].sum(0)
            total = int(np.count_nonzero(k > 0))
            common = int(np.sum(np.maximum(k - 1, 0)))
            avg = dict(clutter=F32(0), inliers=F32(0), sim=F32(0), loc=F32(0), pose=F32(0))
            m = F32(len(cur))
            for s in cur:
                ev = hyps[g[s]][2]
                avg["clutter"] = F32(avg["clutter"] + F32(ev.clutter_score) / m)
                avg["inliers"] = F32(avg["inliers"] + F32(ev.inliers_ratio) / m)
                avg["sim"] = F32(avg["sim"] + F32(ev.similarity_score) / m)
                avg["loc"] = F32(avg["loc"] + F32(ev.location_score) / m)
                avg["pose"] = F32(avg["pose"] + F32(ev.pose_score) / m)
            with np.errstate(divide="ignore", invalid="ignore"):
                total_ratio = F32(total) / F32(total_by_group)
                common_ratio = F32(common) / F32(total_by_group)
            tr_reg = F32(0) if p.single_object_in_group else F32(p.group_total_explain_coeff)
            cr_reg = F32(0) if p.single_object_in_group else F32(p.group_common_explain_coeff)
            fs = (avg["sim"] * F32(p.similarity_coeff) + avg["inliers"] * F32(p.inliers_coeff) - avg["clutter"] * F32(p.clutter_coeff)
                  + avg["pose"] * F32(p.pose_score_coeff) + avg["loc"] * F32(p.location_score_coeff)
                  + total_ratio * tr_reg - common_ratio * cr_reg)
            if fs > best_score:  # R9: first best wins
                best_score, best_sol = fs, list(sol)
        result.extend(g[s] for s in range(len(g)) if best_sol[s])
    return result


def select_instances(chosen, instances):
    """DetectObjects' output loop (HFTest.cpp:1261-1303): by final score, at most instances[cls] per object.
    chosen: list of (cls, pose, Evaluation)."""
    order = sorted(range(len(chosen)), key=lambda i: -chosen[i][2].final_score)  # stable: R9
    count = {}
    out = []
    for i in order:
        c = chosen[i][0]
        if count.get(c, 0) < instances[c]:
            count[c] = count.get(c, 0) + 1
            out.append(i)
    return out


def refine_frame(scene: Scene, models, p: RefineParams, hyps):
    """Tail of HFTest::test_image (HFTest.cpp:922-994) for the hypothesis tuples of the Hough stage.  hyps: structured array
    with cls, loc_score, yawpitch_score, roll_score and the pre-ICP pose (oracle.HYP_DTYPE).
    Returns dict(poses [n,4,4], converged [n], evals [n], accepted indices, chosen indices (into the accepted list, sorted by
    final score))."""
    poses, conv, evals = [], [], []
    for h in hyps:
        m = models[int(h["cls"])]
        pose0 = np.asarray(h["pose"], F32).reshape(4, 4)  # the head of MeshUtils::icp, already part of the tuple (initial_pose)
        pose, ok, _ = icp(scene, m, p, pose0)
        ev = evaluate_hypothesis(scene, m, p, pose, h["loc_score"], (F32(h["yawpitch_score"]) + F32(h["roll_score"])) / F32(2.0))
        poses.append(pose)
        conv.append(ok)
        evals.append(ev)
    acc = [i for i, e in enumerate(evals) if e.accepted]
    acc.sort(key=lambda i: -evals[i].final_score)  # std::sort(hcomparator), R9
    triples = [(int(hyps[i]["cls"]), poses[i], evals[i]) for i in acc]
    chosen = optimize_hypotheses(triples, models, p)
    return dict(poses=np.array(poses, F32).reshape(-1, 4, 4), converged=np.array(conv, bool), evals=evals, accepted=acc,
                chosen=chosen)
