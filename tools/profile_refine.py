"""Stage REFINE on the bench workload (configs[1]): detect -> hypotheses -> ICP + scoring + joint optimisation, timed.

  python tools/profile_refine.py [--frames 3] [--icp-iterations 60]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from object_detector_6d_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--icp-iterations", type=int, default=60)
    ap.add_argument("--radius", type=float, default=0.015)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.make_workload(d, 2)
        det = api.Detector(forest_dir, wpath, api.default_params(fill_random=1, fill_seed=1), device=0, n_slots=1)
        t0 = time.time()
        for k, (x, c) in enumerate(synth.object_models(bench.OBJECT_SEED, bench.K_CLASSES)):
            det.set_object_model(k, x, c, a.radius, a.icp_iterations)
        print("models loaded in %.2f s" % (time.time() - t0))
        for i in range(a.frames):
            hyp = det.detect(frames[i % 2][0], frames[i % 2][1])
            t0 = time.time()
            dets = det.refine(hyp)
            wall = (time.time() - t0) * 1e3
            ms = det.refine_ms()
            print(f"frame {i}: {len(hyp)} hypotheses, {int(dets['icp_converged'].sum())} ICP converged, "
                  f"{int(dets['accepted'].sum())} accepted, {int(dets['selected'].sum())} selected, "
                  f"{int((dets['rank'] >= 0).sum())} written; wall {wall:.2f} ms; stage ms "
                  + ", ".join(f"{k} {v:.3f}" for k, v in ms.items())
                  + f"; scene points {len(det.refine_fetch(api.RBUF_SCENE_POINTS))}, clusters {len(det.refine_fetch(api.RBUF_CLUSTER_SIZES))}")
            sel = dets[dets["rank"] >= 0]
            for r in sel[np.argsort(sel["rank"])]:
                print(f"   rank {r['rank']} class {r['cls']} final {r['final_score']:.3f} inliers {r['inliers_ratio']:.3f} "
                      f"clutter {r['clutter']:.3f} icp its {r['icp_iterations']}")
        det.close()


if __name__ == "__main__":
    main()
