"""Feature storage (and tuning switch) A/B on the bench's workload: frames/s with frames in flight (the device-resident `value` of bench.py) and
the serial encode / traverse times, for fp32 rows (storage 0) and fp16 rows (HF6D_FEATURES=fp16), over several slot counts.

  python tools/sweep_feat16.py [--config c2] [--slots 4,6] [--reps 3]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def measure(api, torch, cfg, workload, slots, reps):
    frames, layers, forest_dir, wpath, stats = workload
    distinct = cfg["distinct"]
    p = bench.params_for(cfg, api)
    det = api.Detector(forest_dir, wpath, p, device=0, n_slots=slots)
    try:
        bgr_all = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
        dep_all = torch.from_numpy(np.stack([f[1] for f in frames]).view(np.int16)).cuda()
        main_s = torch.cuda.Stream()
        streams = [torch.cuda.Stream() for _ in range(slots)]
        for s in range(slots):
            det.set_stream(s, streams[s].cuda_stream)

        def batch():
            for i in range(bench.BATCH):
                s = i % slots
                j = i % distinct
                det.bind_frame(s, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
                det.run(s)

        batch()
        batch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        ev = torch.cuda.Event()
        ev.record(main_s)
        for st in streams:
            st.wait_event(ev)
        for _ in range(reps):
            batch()
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main_s.wait_event(ev)
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * bench.BATCH)
        # serial pass: one frame at a time on slot 0
        enc, trv, pose = [], [], []
        for j in range(distinct):
            det.bind_frame(0, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
            det.run(0)
            det.sync(0)
            enc.append(det.encoder_layer_ms(0))
            sm = det.stage_ms(0)
            trv.append(sm[api.STAGE_TRAVERSE])
            pose.append(sm[api.STAGE_POSE])
        hyp = det.collect(0)
        return dict(ms_per_frame=round(ms, 4), frames_per_s=round(1000.0 / ms, 1),
                    encoder_layer_ms=[round(float(x), 4) for x in np.mean(np.array(enc), 0)],
                    traverse_ms=round(float(np.mean(trv)), 4), pose_ms=round(float(np.mean(pose)), 4),
                    hypotheses_last_frame=int(len(hyp)))
    finally:
        det.close()


def main():
    import torch
    from object_detector_6d_b200 import api

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--slots", default="4,6")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--features", default="fp32,fp16")
    ap.add_argument("--env", default="", help="';'-separated NAME=VALUE settings to sweep as well (tuning switches read at create)")
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    out = {"config": args.config, "runs": []}
    with tempfile.TemporaryDirectory() as d:
        workload = bench.config_workload(cfg, d)
        for slots in [int(x) for x in args.slots.split(",")]:
            for storage in args.features.split(","):
                for setting in (args.env.split(";") if args.env else [""]):
                    os.environ["HF6D_FEATURES"] = storage
                    if setting:
                        name, value = setting.split("=", 1)
                        os.environ[name] = value
                    r = measure(api, torch, cfg, workload, slots, args.reps)
                    if setting:
                        os.environ.pop(name, None)
                    r.update(slots=slots, features=storage, env=setting)
                    out["runs"].append(r)
                    print(json.dumps(r), flush=True)
    os.environ.pop("HF6D_FEATURES", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
