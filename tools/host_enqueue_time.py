"""How long does the HOST need to enqueue one frame (all launches + event records of hf6d_run)?  If that is close to the
device time per frame, the multi-stream throughput is bound by the submitting thread, not by the GPU.

  python tools/host_enqueue_time.py [--slots 4] [--frames 256]
"""
import argparse
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from object_detector_6d_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--frames", type=int, default=256)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.make_workload(d, 2)
        det = api.Detector(forest_dir, wpath, api.default_params(fill_random=1, fill_seed=1), device=0, n_slots=a.slots)
        for s in range(a.slots):
            det.upload(s, frames[s % 2][0], frames[s % 2][1])
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(a.frames):
                det.run(i % a.slots)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"rep {rep}: enqueue {1e6 * (t1 - t0) / a.frames:.1f} us/frame, total {1e6 * (t2 - t0) / a.frames:.1f} us/frame "
                  f"({a.frames / (t2 - t0):.0f} frames/s), {det.launch_count(0)} launches/frame")
        det.close()


if __name__ == "__main__":
    main()
