"""Forest training on the GPU at the size of the bench workload's training set.

  python tools/profile_train.py [--samples 160000] [--features 800] [--classes 6] [--trees 2]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detector_6d_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=160000)
    ap.add_argument("--features", type=int, default=800)
    ap.add_argument("--classes", type=int, default=6)
    ap.add_argument("--trees", type=int, default=2)
    ap.add_argument("--min_samples", type=int, default=30)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    n, K, F = a.samples, a.classes, a.features
    cls = rng.integers(0, K, n).astype(np.int32)
    dof = np.zeros((n, 6), np.float32)
    dof[:, :3] = rng.uniform(-np.pi, np.pi, (n, 3))
    dof[:, 3:] = rng.uniform(-0.1, 0.1, (n, 3))
    feat = rng.uniform(0, 1, (n, F)).astype(np.float32)
    feat[:, :64] += cls[:, None] * 0.2
    feat[:, 64:128] += dof[:, 3:4] * 3.0
    feat[:, 128:192] += np.cos(dof[:, 0:1]) * 0.3
    with tempfile.TemporaryDirectory() as d:
        for rep in range(2):
            t0 = time.time()
            st = api.train_forest(d, cls, dof, feat, K=K, trees=a.trees, min_samples=a.min_samples, seed=1 + rep)
            wall = time.time() - t0
            evals = st.training_samples * 301.0 * 3 * st.max_depth  # (min/max + statistics + apply) x tests x levels, upper bound
            print(f"run {rep}: {a.trees} trees x {st.training_samples} samples x {F} features, 30 x 10 tests per node: "
                  f"{st.nodes} nodes, {st.leaves} leaves, depth {st.max_depth}; GPU {st.train_ms:.1f} ms "
                  f"({st.train_ms / a.trees:.1f} ms per tree), wall {wall:.2f} s (H2D of {feat.nbytes / 1e6:.0f} MB included); "
                  f"<= {evals * a.trees / (st.train_ms * 1e-3) / 1e9:.1f} G test evaluations/s")


if __name__ == "__main__":
    main()
