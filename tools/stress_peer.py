"""Long-run check of the peer exchange (tree-sharded mode over NVLink peer memory): thousands of frames through all slots,
every slot's result compared with the first result for that frame.  Launch with torchrun, one rank per GPU, under `timeout`:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/stress_peer.py --frames 3000
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import bench  # noqa: E402
from object_detector_6d_b200 import api, sharded  # noqa: E402


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(a[n], b[n]) for n in a.dtype.names)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3000)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--distinct", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = bench.make_workload(d, a.distinct)
        p = api.default_params(fill_random=1, fill_seed=1)
        sd = sharded.TreeShardedDetector(forest_dir, wpath, p, device=local, n_slots=a.slots, exchange="peer")
        ref = [sd.detect(f[0], f[1], slot=0, gather=False).copy() for f in frames]
        t0 = time.perf_counter()
        bad = 0
        which = [None] * a.slots
        for i in range(a.frames):
            s, j = i % a.slots, (i * 7 + i // 5) % a.distinct
            if which[s] is not None and not same(sd.det.collect(s), ref[which[s]]):
                bad += 1
            sd.det.upload(s, frames[j][0], frames[j][1])
            sd.run(s)
            which[s] = j
        for s in range(a.slots):
            if which[s] is not None and not same(sd.det.collect(s), ref[which[s]]):
                bad += 1
        dt = time.perf_counter() - t0
        t = torch.tensor([bad + (1 if sd.det.peer_timed_out() else 0)], device="cuda")
        dist.all_reduce(t)
        sd.close()
        if rank == 0:
            print(f"peer stress: {a.frames} frames, {world} ranks, {a.slots} slots, {a.frames / dt:.0f} frames/s, "
                  f"{len(ref[0])} hypotheses on rank 0, mismatches over all ranks: {int(t.item())}")
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
