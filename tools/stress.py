"""Long-run determinism check: the same frames through the pipelined path thousands of times, every result compared with
the first one (a race in a barrier protocol shows up as a changed hypothesis list, a trap, or a hang -- run under `timeout`).

  python tools/stress.py [--frames 4000] [--slots 4] [--distinct 3]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from object_detector_6d_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4000)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--distinct", type=int, default=3)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.make_workload(d, a.distinct)
        det = api.Detector(forest_dir, wpath, api.default_params(fill_random=1, fill_seed=1), device=0, n_slots=a.slots)
        ref = [det.detect(f[0], f[1]).copy() for f in frames]
        feat_ref = []
        for f in frames:
            det.upload(0, f[0], f[1])
            det.run(0)
            det.sync(0)
            feat_ref.append((det.fetch(api.BUF_FEATURES).copy(), det.fetch(api.BUF_LEAF_ORD).copy()))
        t0 = time.perf_counter()
        bad = 0
        pending = []
        for i in range(a.frames):
            j = i % a.distinct
            if len(pending) == a.slots:
                tk, jj = pending.pop(0)
                h = det.wait(tk)
                if len(h) != len(ref[jj]) or any(not np.array_equal(h[n], ref[jj][n]) for n in h.dtype.names):
                    bad += 1
            pending.append((det.submit(frames[j][0], frames[j][1]), j))
        for tk, jj in pending:
            h = det.wait(tk)
            if len(h) != len(ref[jj]) or any(not np.array_equal(h[n], ref[jj][n]) for n in h.dtype.names):
                bad += 1
        dt = time.perf_counter() - t0
        # intermediates once more at the end
        for j, f in enumerate(frames):
            det.upload(0, f[0], f[1])
            det.run(0)
            det.sync(0)
            if not (np.array_equal(det.fetch(api.BUF_FEATURES), feat_ref[j][0]) and np.array_equal(det.fetch(api.BUF_LEAF_ORD), feat_ref[j][1])):
                bad += 1
        det.close()
        print(f"stress: {a.frames} frames, {a.slots} slots, {a.frames / dt:.0f} frames/s (host buffers are pageable here), mismatches: {bad}")
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
