"""Marginal cost of every stage of the hot path with frames in flight: frames/s of the stage prefixes SCAN..X on the bench's
4-slot context (the same workload and streams as bench.py's `value`).  The difference between two prefixes is what a stage
costs once its kernels overlap other frames' kernels -- the number the serial per-stage table cannot give.

  python tools/pipelined_marginal.py [--config c2] [--slots 4] [--reps 3]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from object_detector_6d_b200 import api

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    distinct = cfg["distinct"]
    names = ["scan", "gather", "encode", "traverse", "vote", "centres", "pose"]
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.config_workload(cfg, d)
        p = bench.params_for(cfg, api)
        det = api.Detector(forest_dir, wpath, p, device=0, n_slots=args.slots)
        bgr_all = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
        dep_all = torch.from_numpy(np.stack([f[1] for f in frames]).view(np.int16)).cuda()
        main_s = torch.cuda.Stream()
        streams = [torch.cuda.Stream() for _ in range(args.slots)]
        for s in range(args.slots):
            det.set_stream(s, streams[s].cuda_stream)

        def batch(last):
            for i in range(bench.BATCH):
                s = i % args.slots
                j = i % distinct
                det.bind_frame(s, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
                det.run(s, api.STAGE_SCAN, last)

        out = {}
        prev = 0.0
        for last in range(api.STAGE_COUNT):
            batch(api.STAGE_POSE)  # every buffer valid
            batch(last)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            ev = torch.cuda.Event()
            ev.record(main_s)
            for st in streams:
                st.wait_event(ev)
            for _ in range(args.reps):
                batch(last)
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                main_s.wait_event(ev)
            e1.record(main_s)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (args.reps * bench.BATCH)
            out[names[last]] = {"prefix_ms_per_frame": round(ms, 4), "marginal_ms": round(ms - prev, 4)}
            prev = ms
        print(json.dumps({"config": args.config, "slots": args.slots, "pipelined_prefixes": out}))


if __name__ == "__main__":
    main()
