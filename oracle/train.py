"""CPU oracle of Hough-forest training: HFTrain (HoughForest/src/HFTrain.cpp:14-1265) restated in numpy.
TEST INFRASTRUCTURE ONLY -- only tests/ import this; the product never does.

PARITY: the reference draws every random number from libc rand(), seeded with time(NULL) and called from OpenMP threads
(HFTrain.cpp:244-246, 386, 1091, 1208): its forests are not reproducible even by itself, and it ships no trained forest.  What
can be pinned is the ALGORITHM given the draws; so the draws come from a counter-based generator (rng_* below, the same
function in csrc/train.cuh) and everything else follows the reference line by line:

  HFTrain::getTrainSet                 HFTrain.cpp:14-67      read_patches_file
  suffle_training_set + the 2/3 rule   :72-85, :1142-1145     shuffle
  get_random_features                  :231-262               tests per node: tests_per_node x (mode, f1, f2), each repeated
                                                              thresholds_per_test times
  get_min_max_count_samples            :267-362               value range of every test over the node's samples
  get_random_thresholds                :365-392               threshold = u * (max - min) + min
  optimize_level's choice of objective :1089-1094             level < 4: classification, else one of three at random
  find_classification_split            :399-523               sum of child entropies weighted by child size
  find_regression_location_split       :531-757               within-child scatter of the (x, y, z) offsets
  find_regression_pose_split           :762-994               within-child scatter of (cos, sin) of yaw, pitch, roll
  apply_tests_to_train_samples         :999-1047              val < threshold -> left
  train_tree's leaf rule               :1159-1176             a child with <= min_samples samples is a leaf
  make_leafs                           :87-207                class_prob[c] = n_c / sum_i (N_c / N_i) n_i ; votes per class
  saveTreeNode / forest.txt            HFBase.cpp:4-38, HFTrain.cpp:1225-1231

Choices
  T1  random numbers: rng_u64(seed, tree, level, node, test, draw) (splitmix64 finaliser chain), never rand().
  T2  the regression objectives are evaluated as  sum |v|^2 - |sum v|^2 / n  per child in double (the reference makes two float
      passes whose sums depend on the OpenMP schedule); ties and near-ties between tests may therefore resolve differently.
  T3  among tests with equal objective the first wins (the reference: `<` in test order -- the same).
  T4  leaf ids count up in the order the leaves are written (pre-order); the reference numbers them in hash-map order.
  T5  the votes of a leaf are stored in ascending sample order (the reference: OpenMP merge order).
"""
from __future__ import annotations

import os
import struct

import numpy as np

F32 = np.float32
M64 = (1 << 64) - 1


def mix64(z: int) -> int:
    z &= M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rng_u64(seed: int, tree: int, level: int, node: int, test: int, draw: int) -> int:
    """T1.  level -1 = the shuffle (node = position), node 0xFFFFFFFF = the level's objective."""
    h = mix64((seed ^ (0x9E3779B97F4A7C15 * (tree + 1))) & M64)
    h = mix64((h + (level & 0xFFFFFFFF)) & M64)
    h = mix64((h + (node & 0xFFFFFFFF)) & M64)
    return mix64((h + ((test << 3) | draw)) & M64)


DRAW_MODE, DRAW_F1, DRAW_F2, DRAW_THR = 0, 1, 2, 3
LEVEL_SHUFFLE, NODE_OBJECTIVE = -1, 0xFFFFFFFF


def read_patches_file(path: str):
    """HFTrain::getTrainSet (HFTrain.cpp:14-67): int32 K, int32 F, then records {int32 class, 6 x f32 dof, F x f32}."""
    raw = np.fromfile(path, np.uint8)
    K, F = struct.unpack("<ii", raw[:8].tobytes())
    rec = 4 + 24 + 4 * F
    n = (len(raw) - 8) // rec
    body = raw[8:8 + n * rec].reshape(n, rec)
    cls = body[:, :4].copy().view("<i4").reshape(n)
    dof = body[:, 4:28].copy().view("<f4").reshape(n, 6)
    feat = body[:, 28:].copy().view("<f4").reshape(n, F)
    return K, F, cls.astype(np.int32), dof.astype(F32), feat.astype(F32)


def write_patches_file(path: str, K: int, cls, dof, feat):
    n, F = feat.shape
    rec = np.zeros((n, 4 + 24 + 4 * F), np.uint8)
    rec[:, :4] = np.ascontiguousarray(cls, "<i4").view(np.uint8).reshape(n, 4)
    rec[:, 4:28] = np.ascontiguousarray(dof, "<f4").view(np.uint8).reshape(n, 24)
    rec[:, 28:] = np.ascontiguousarray(feat, "<f4").view(np.uint8).reshape(n, 4 * F)
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", K, F))
        f.write(rec.tobytes())


def shuffle(n: int, seed: int, tree: int):
    """suffle_training_set (HFTrain.cpp:72-85) on an index array; the first int(2/3 * n) entries train the tree (:1143)."""
    order = np.arange(n, dtype=np.int64)
    n_train = int(F32(2.0) / F32(3.0) * F32(n))
    for i in range(n_train):
        k = rng_u64(seed, tree, LEVEL_SHUFFLE, i, 0, 0) % (n - i) + i
        order[i], order[k] = order[k], order[i]
    return order, n_train


class Node:
    __slots__ = ("leaf", "test", "left", "right", "samples", "leaf_id", "class_prob", "votes")

    def __init__(self):
        self.leaf, self.test, self.left, self.right, self.samples = False, None, None, None, None
        self.leaf_id, self.class_prob, self.votes = -1, None, None


def _values(feat, mode, f1, f2):
    """[n_samples, n_tests] test values (HFTrain.cpp:312-316): mode 0: f[f1] - f[f2], mode 1: f[f1]."""
    a = feat[:, f1]
    return np.where(mode[None, :] == 0, (a - feat[:, f2]).astype(F32), a)


def train_tree(cls, dof, feat, K, tree, seed, min_samples=30, tests_per_node=30, thresholds_per_test=10):
    """HFTrain::train_tree (HFTrain.cpp:1135-1195).  Returns the root Node."""
    n, F = feat.shape
    order, n_train = shuffle(n, seed, tree)
    train = np.sort(order[:n_train])  # T5: sample order inside a node is ascending index
    spc = np.bincount(cls[train], minlength=K).astype(np.int64)  # samples_per_class of the training subset
    root = Node()
    root.samples = train
    level_nodes, level = [root], 0
    NT = tests_per_node * thresholds_per_test
    pose_v = np.stack([np.cos(dof[:, 0].astype(np.float64)), np.sin(dof[:, 0].astype(np.float64)),
                       np.cos(dof[:, 1].astype(np.float64)), np.sin(dof[:, 1].astype(np.float64)),
                       np.cos(dof[:, 2].astype(np.float64)), np.sin(dof[:, 2].astype(np.float64))], 1)
    loc_v = dof[:, 3:6].astype(np.float64)
    while level_nodes:
        method = 0 if level < 4 else rng_u64(seed, tree, level, NODE_OBJECTIVE, 0, 0) % 3
        for ni, node in enumerate(level_nodes):
            s = node.samples
            mode = np.zeros(NT, np.int32)
            f1 = np.zeros(NT, np.int64)
            f2 = np.zeros(NT, np.int64)
            for t in range(tests_per_node):
                mm = rng_u64(seed, tree, level, ni, t, DRAW_MODE) % 2
                a = rng_u64(seed, tree, level, ni, t, DRAW_F1) % F
                b = rng_u64(seed, tree, level, ni, t, DRAW_F2) % F
                mode[t * thresholds_per_test:(t + 1) * thresholds_per_test] = mm
                f1[t * thresholds_per_test:(t + 1) * thresholds_per_test] = a
                f2[t * thresholds_per_test:(t + 1) * thresholds_per_test] = b
            vals = _values(feat[s], mode, f1, f2)
            lo, hi = vals.min(0), vals.max(0)
            u = np.array([F32(rng_u64(seed, tree, level, ni, t, DRAW_THR) >> 40) / F32(16777216.0) for t in range(NT)], F32)
            thr = (u * (hi - lo).astype(F32) + lo).astype(F32)
            left = vals < thr[None, :]
            nl = left.sum(0)
            nr = len(s) - nl
            ok = (nl > 0) & (nr > 0)
            obj = np.full(NT, np.inf)
            if method == 0:
                c = cls[s]
                for t in np.flatnonzero(ok):
                    cl = np.bincount(c[left[:, t]], minlength=K)
                    cr = np.bincount(c, minlength=K) - cl
                    el = er = F32(0)
                    for k in range(K):
                        p = F32(cl[k]) / F32(nl[t])
                        if p != 0:
                            el = F32(np.float64(el) - np.float64(p) * np.log(np.float64(p)))
                        p = F32(cr[k]) / F32(nr[t])
                        if p != 0:
                            er = F32(np.float64(er) - np.float64(p) * np.log(np.float64(p)))
                    obj[t] = F32(F32(el * F32(nl[t])) + F32(er * F32(nr[t])))
            else:
                v = loc_v[s] if method == 1 else pose_v[s]
                # |(cos, sin) of three angles|^2 is 3 for every sample: the pose scatter needs no sum of squares
                tot, tot2 = v.sum(0), (np.sum(v * v) if method == 1 else 3.0 * len(s))
                for t in np.flatnonzero(ok):
                    vl = v[left[:, t]]
                    sl, sl2 = vl.sum(0), (np.sum(vl * vl) if method == 1 else 3.0 * nl[t])
                    sr, sr2 = tot - sl, tot2 - sl2
                    obj[t] = F32((sl2 - np.dot(sl, sl) / nl[t]) + (sr2 - np.dot(sr, sr) / nr[t]))  # T2
            if not ok.any():
                node.leaf = True
                continue
            best = int(np.argmin(obj))  # T3: first minimum
            node.test = (int(mode[best]), int(f1[best]), int(f2[best]), F32(thr[best]))
            node.left, node.right = Node(), Node()
            node.left.samples, node.right.samples = s[left[:, best]], s[~left[:, best]]
        nxt = []
        for node in level_nodes:
            if node.leaf:
                continue
            for ch in (node.left, node.right):
                if len(ch.samples) > min_samples:
                    nxt.append(ch)
                else:
                    ch.leaf = True
        level_nodes, level = nxt, level + 1
    # make_leafs (HFTrain.cpp:87-207) + T4 leaf ids in pre-order
    counter = [0]

    def finish(node):
        if node.leaf:
            node.leaf_id = counter[0]
            counter[0] += 1
            s = node.samples
            cnt = np.bincount(cls[s], minlength=K)
            prob = np.zeros(K, F32)
            with np.errstate(divide="ignore", invalid="ignore"):
                for c in range(K):
                    norm = F32(0)
                    for i in range(K):
                        norm = F32(norm + F32(F32(spc[c]) / F32(spc[i])) * F32(cnt[i]))
                    prob[c] = F32(cnt[c]) / norm
            node.class_prob = prob
            node.votes = [dof[s[cls[s] == c]] for c in range(K)]
        else:
            finish(node.left)
            finish(node.right)
    finish(root)
    return root


def serialise(node: Node, K: int) -> bytes:
    """HFBase::saveTreeNode (HFBase.cpp:4-38)."""
    out = bytearray()

    def rec(nd):
        out.extend(struct.pack("<B", 1 if nd.leaf else 0))
        if nd.leaf:
            out.extend(struct.pack("<i", nd.leaf_id))
            out.extend(np.asarray(nd.class_prob, "<f4").tobytes())
            for c in range(K):
                v = np.asarray(nd.votes[c], "<f4").reshape(-1, 6)
                out.extend(struct.pack("<i", len(v)))
                out.extend(v.tobytes())
        else:
            mm, a, b, thr = nd.test
            out.extend(struct.pack("<iiif", mm, a, b, float(thr)))
            rec(nd.left)
            rec(nd.right)
    rec(node)
    return bytes(out)


def train_forest(out_dir, cls, dof, feat, K, trees=3, seed=1, start_tree_no=0, patch_size_in_voxels=8, voxel_size_in_m=0.005,
                 **kw):
    """HFTrain::train (HFTrain.cpp:1199-1265): forest.txt + tree<N>.dat."""
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "forest.txt"), "w") as f:
        f.write(f"{trees} {K} {feat.shape[1]} {patch_size_in_voxels} {voxel_size_in_m:g}\n")
    roots = []
    for t in range(start_tree_no, start_tree_no + trees):
        root = train_tree(cls, dof, feat, K, t, seed, **kw)
        with open(os.path.join(out_dir, f"tree{t}.dat"), "wb") as f:
            f.write(serialise(root, K))
        roots.append(root)
    return roots
