"""Small driver for ncu: the bench workload (configs[1]), a few device-resident frames through every stage.

  python tools/profile_frame.py [--frames 3] [--distinct 2]
"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from object_detector_6d_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--distinct", type=int, default=2)
    ap.add_argument("--trees", type=int, default=bench.T_TREES)
    ap.add_argument("--lib", default=None, help="load this build of libhf6d.so instead (timing experiments)")
    a = ap.parse_args()
    if a.lib:
        from object_detector_6d_b200 import build as _b
        _b.build_lib = lambda *x, **k: os.path.abspath(a.lib)
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.make_workload(d, a.distinct, T=a.trees)
        det = api.Detector(forest_dir, wpath, api.default_params(fill_random=1, fill_seed=1), device=0, n_slots=1)
        for i in range(a.frames):
            det.upload(0, frames[i % a.distinct][0], frames[i % a.distinct][1])
            det.run(0)
            det.sync(0)
        print("stage ms", dict(zip(api.STAGE_NAMES, np.round(det.stage_ms(0), 4))), "launches", det.launch_count(0))
        print("encoder layer ms", np.round(det.encoder_layer_ms(0), 4))
        print("hyps", len(det.collect(0)), "counts", det.counts(0))
        det.close()


if __name__ == "__main__":
    main()
