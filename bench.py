#!/usr/bin/env python
"""bench.py -- HoughForest test-time detection path on B200 (BASELINE.json metric: frames/s and patch-tree
traversals/s per 640x480 RGB-D frame).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libhf6d.so through its C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores (CPU oracle)

A "step" is one pass of the hot path (scan -> gather -> encode -> traverse -> vote -> centres -> pose) over one batch
of BATCH synthetic frames (BASELINE.json configs[1]: 6-object forest, T=4, depth ~20, 16 votes per leaf, 64 cluttered
640x480 frames).  One JSON line on stdout (rank 0).

Multi-GPU (torchrun, one rank per GPU): frames are independent (the reference's frame loop carries no state,
HFTest.cpp:1238), so ranks take their own batch -- weak scaling, no data-path collective -- and additionally the
tree-sharded mode the north star names (trees t % N == rank, vote maps summed with NCCL before mode seeking) is run and
reported under "tree_sharded".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# one hardware queue per stream (default 8): frame slots on aliased queues would serialise, and a slot waiting for a
# peer's flag (tree-sharded mode) must never hold up another slot's kernels.  Read by the driver at CUDA initialisation.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from object_detector_6d_b200 import synth  # noqa: E402

BATCH = 64            # frames per step (configs[1])
DISTINCT_FRAMES = 8   # rendered once (the numpy ray-caster takes ~2 s per frame); the batch cycles through them
K_CLASSES, T_TREES, MAX_DEPTH, VOTES = 6, 4, 20, 16
ENC_FLOP_PER_PATCH = 2 * (256 * 1500 + 1500 * 1000 + 1000 * 800)
METRIC = "frames/s (640x480 RGB-D, HoughForest --test hot path)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return dict(hbm=float(pk["hbm_gbs"]), tf_burst=float(pk["bf16_tflops"]),
                    tf_sustained=float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


OBJECT_SEED = 1000    # the six procedural objects that stand in for meshes/*.ply (shape, size, albedo, texture)
TRAIN_FRAMES = 4      # frames whose object patches (every pixel) train the forest: same objects, their own poses
VIEWS, MIN_SAMPLES = 4, 8   # training views a labelled patch stands for / node size that stops splitting (-> ~16 votes per leaf)
WORKLOAD_VERSION = 4  # bump when synth or the recipe below changes (invalidates the on-disk cache)


def _build_workload(out: str, n_frames: int, seed0: int, T: int, forest: str):
    layers = synth.make_encoder_weights(3)
    forest_dir = os.path.join(out, "forest")
    if forest == "random":
        frames = [synth.render_frame(seed0 + i) for i in range(n_frames)]
        calib = synth.calibration_features(frames[0][0], frames[0][1], layers, n=30000)
        stats = synth.write_forest(forest_dir, calib, T=T, K=K_CLASSES, max_depth=MAX_DEPTH, votes_per_leaf=VOTES, seed=7)
    else:
        scenes = [synth.render_scene(seed0 + i, OBJECT_SEED) for i in range(max(n_frames, TRAIN_FRAMES))]
        frames = [(b, d) for b, d, _ in scenes[:n_frames]]
        lab = [synth.labelled_patches(b, d, tr, layers, n=80000, seed=i, stride=1) for i, (b, d, tr) in enumerate(scenes[:TRAIN_FRAMES])]
        feats, cls, votes = (np.concatenate([x[j] for x in lab]) for j in range(3))
        stats = synth.write_trained_forest(forest_dir, feats, cls, votes, T=T, K=K_CLASSES, max_depth=MAX_DEPTH,
                                           min_samples=MIN_SAMPLES, views=VIEWS, seed=7)
        stats["training_samples"] = int(len(cls))
    stats["forest"] = forest
    synth.write_weights_raw(os.path.join(out, "weights.bin"), layers)
    np.savez(os.path.join(out, "frames.npz"), bgr=np.stack([f[0] for f in frames]), depth=np.stack([f[1] for f in frames]))
    with open(os.path.join(out, "stats.json"), "w") as f:
        json.dump(stats, f)


def make_workload(tmpdir: str, n_frames: int, seed0: int = 1, T: int = T_TREES, forest: str = "trained"):
    """Frames + encoder weights + forest on disk.  forest = "trained": leaves hold the class distributions and the 6-DoF
    votes of labelled object patches (synth.write_trained_forest) -- coherent votes, real Hough modes, the vote hot spots
    a trained forest produces; "random": uniformly random votes, 16 per leaf (round 1's workload: no modes, maximal
    scatter; kept for the stage-level parity tests and for continuity).

    Building it takes ~40 s of numpy, so it is cached under the system temp directory, keyed by the recipe: the ranks of a
    torchrun launch (and successive bench / test processes on one box) share one copy -- whoever creates the directory
    builds, the others wait for its `done` marker.  `tmpdir` is unused when the cache can be used."""
    key = f"hf6d_workload_v{WORKLOAD_VERSION}_{forest}_n{n_frames}_s{seed0}_T{T}"
    root = os.path.join(tempfile.gettempdir(), key)
    done = os.path.join(root, "done")
    try:
        os.makedirs(root)
        owner = True
    except FileExistsError:
        owner = False
    if owner:
        try:
            _build_workload(root, n_frames, seed0, T, forest)
            open(done, "w").close()
        except BaseException:
            import shutil
            shutil.rmtree(root, ignore_errors=True)
            raise
    else:
        t0 = time.time()
        while not os.path.exists(done):
            if not os.path.isdir(root) or time.time() - t0 > 900:  # the builder failed or died: build privately
                root = os.path.join(tmpdir, key)
                os.makedirs(root, exist_ok=True)
                _build_workload(root, n_frames, seed0, T, forest)
                break
            time.sleep(0.5)
    z = np.load(os.path.join(root, "frames.npz"))
    frames = [(np.ascontiguousarray(z["bgr"][i]), np.ascontiguousarray(z["depth"][i])) for i in range(n_frames)]
    with open(os.path.join(root, "stats.json")) as f:
        stats = json.load(f)
    return frames, synth.make_encoder_weights(3), os.path.join(root, "forest"), os.path.join(root, "weights.bin"), stats


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU algorithm for the path (its own binary cannot be built here: SURVEY.md F5), i.e. the oracle
    port, OpenMP over all host cores.  One step = one frame of the batch (bounded sample)."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = O.set_threads()  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, _, stats = make_workload(d, DISTINCT_FRAMES, forest=args.forest)
        forest = O.Forest(forest_dir)
        p = O.default_params(fill_random=1, fill_seed=1)
        n_trav = 0
        for i in range(args.warmup):
            O.detect(forest, frames[i % len(frames)][0], frames[i % len(frames)][1], p, layers)
        t0 = time.perf_counter()
        stage = np.zeros(6)
        for i in range(args.steps):
            _, (P, Pp), st = O.detect(forest, frames[i % len(frames)][0], frames[i % len(frames)][1], p, layers)
            n_trav += Pp * forest.T
            stage += st
        dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(stats, sample="1 frame per step"),
        "traversals_per_s": n_trav / dt,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frames of the batch, one per step, all stages, OpenMP on {cores} threads",
                         "stage_s_per_frame": {k: float(v / args.steps) for k, v in
                                               zip(("gather", "normalise", "encode", "traverse", "vote+modes", "total"), stage)}},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(stats, **extra):
    cfg = {"workload": "configs[1]: 6-object forest, batch of 64 synthetic cluttered 640x480 RGB-D frames",
           "frame": "640x480", "stride": 2, "classes": K_CLASSES, "trees": T_TREES, "mean_leaf_depth":
           float(np.mean(stats["mean_depth"])), "forest": stats.get("forest"), "leaves": int(sum(stats["leaves"])),
           "batch_frames": BATCH, "distinct_frames": DISTINCT_FRAMES, "fill": "random (are_objects_segmented: false)",
           "l2": "per-frame intermediates (0.9 GB) exceed the 126 MB L2; frames cycle through 8 distinct inputs"}
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from object_detector_6d_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the libhf6d path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = load_peaks()

    with tempfile.TemporaryDirectory() as d:
        frames, layers, forest_dir, wpath, stats = make_workload(d, DISTINCT_FRAMES, forest=args.forest)
        p = api.default_params(fill_random=1, fill_seed=1)
        n_slots = args.slots
        det = api.Detector(forest_dir, wpath, p, device=local_rank, n_slots=n_slots)

        # ---- device-resident inputs
        bgr_all = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
        dep_all = torch.from_numpy(np.stack([f[1] for f in frames]).view(np.int16)).cuda()
        main = torch.cuda.Stream()
        streams = [torch.cuda.Stream() for _ in range(n_slots)]
        for s in range(n_slots):
            det.set_stream(s, streams[s].cuda_stream)

        def step_resident():
            # frames are independent: slot s (its own stream and workspace) takes frames s, s + n_slots, ...; the small
            # kernels of one frame overlap the encoder of another
            launches = 0
            for i in range(BATCH):
                s = i % n_slots
                j = i % DISTINCT_FRAMES
                det.bind_frame(s, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
                det.run(s)
                launches += det.launch_count(s)
            return launches

        def fork():
            ev = torch.cuda.Event()
            ev.record(main)
            for st in streams:
                st.wait_event(ev)

        def join():
            for st in streams:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)

        for _ in range(args.warmup):
            step_resident()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        fork()
        launches = 0
        for _ in range(args.steps):
            launches += step_resident()
        join()
        e1.record(main)
        main.synchronize()
        torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        counts = [det.counts(s) for s in range(n_slots)]
        # per-stage times: a serial pass (one frame at a time, nothing else on the GPU) so a stage's events bracket only
        # its kernels.  It runs on a ONE-slot context -- the library's latency configuration, whose encoder kernels are the
        # ones that are fastest alone (one more ring stage than the pipelined context's, see HF6D_ENC_CONFIGS)
        det1 = api.Detector(forest_dir, wpath, p, device=local_rank, n_slots=1)
        st_acc, enc_acc, n_ser = np.zeros(api.STAGE_COUNT), np.zeros(3), 0
        for rep in range(2):
            for j in range(DISTINCT_FRAMES):
                det1.bind_frame(0, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
                det1.run(0)
                det1.sync(0)
                if rep:
                    st_acc += det1.stage_ms(0)
                    enc_acc += det1.encoder_layer_ms(0)
                    n_ser += 1
        stage_ms, enc_ms = st_acc / n_ser, enc_acc / n_ser
        det1.bind_frame(0, None, None)
        det1.close()
        # patches per frame: exact, from the scan of every distinct frame
        Pp_frames, votes_frames = [], []
        for j in range(DISTINCT_FRAMES):
            det.bind_frame(0, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
            det.run(0, api.STAGE_SCAN, api.STAGE_TRAVERSE)
            Pp_frames.append(det.counts(0)[1])
            votes_frames.append(det.count_cast_votes(0))
        Pp_mean = float(np.mean([Pp_frames[i % DISTINCT_FRAMES] for i in range(BATCH)]))
        votes_mean = float(np.mean([votes_frames[i % DISTINCT_FRAMES] for i in range(BATCH)]))
        for s in range(n_slots):
            det.bind_frame(s, None, None)
            det.set_stream(s, None)
        if world > 1:
            t = torch.tensor([ms_total], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())

        # ---- end to end through the public API: pinned host frames -> hf6d_submit / hf6d_wait -> host hypotheses
        pin_b = [api.PinnedArray((480, 640, 3), np.uint8) for _ in range(DISTINCT_FRAMES)]
        pin_d = [api.PinnedArray((480, 640), np.uint16) for _ in range(DISTINCT_FRAMES)]
        for j in range(DISTINCT_FRAMES):
            pin_b[j].array[...] = frames[j][0]
            pin_d[j].array[...] = frames[j][1]

        def step_e2e():
            tickets, nh = [], 0
            for i in range(BATCH):
                j = i % DISTINCT_FRAMES
                if len(tickets) == n_slots:
                    nh += len(det.wait(tickets.pop(0)))
                tickets.append(det.submit(pin_b[j].array, pin_d[j].array))
            while tickets:
                nh += len(det.wait(tickets.pop(0)))
            return nh

        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_hyp = 0
        for _ in range(args.steps):
            n_hyp += step_e2e()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        clocks = sampler.stop() if sampler else None
        for a in pin_b + pin_d:
            a.free()
        d2h = BATCH * det.result_bytes()
        h2d = BATCH * (640 * 480 * 5)

        # ---- CPU baseline (rank 0, N == 1 only): the oracle port on a bounded sample, all host cores
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            cpu_threads = O.set_threads()
            forest = O.Forest(forest_dir)
            po = O.default_params(fill_random=1, fill_seed=1)
            O.detect(forest, frames[0][0], frames[0][1], po, layers)  # warm
            n_s = 2
            t0 = time.perf_counter()
            for i in range(n_s):
                O.detect(forest, frames[i][0], frames[i][1], po, layers)
            dt = time.perf_counter() - t0
            cpu = {"value": n_s / dt, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                   "sample": f"{n_s} frames of the batch, all stages, OpenMP on all host cores"}
        det.close()

        def build_line(tree):
            """The JSON line from everything measured so far (also called by the watchdog of the tree-sharded arm)."""
            frames_total = BATCH * args.steps * world
            sec = ms_total * 1e-3
            fps = frames_total / sec
            ms_frame = ms_total / (BATCH * args.steps)
            # stage rooflines from SURVEY.md §8(d)'s algorithmic work per frame
            Pp = Pp_mean
            votes_cast = votes_mean  # counted from the leaf tables of the frames (hf6d_count_cast_votes)
            alg = {
                "scan": (640 * 480 * 2 + Pp * 8, "hbm"),
                "gather": (640 * 480 * 5 + Pp * 512, "hbm"),                 # frame once + bf16 A operand [P'][256]
                "encode": (Pp * ENC_FLOP_PER_PATCH, "tensor"),
                "traverse": (Pp * 800 * 4 + Pp * T_TREES * 4, "hbm"),
                "vote": (Pp * T_TREES * 4 + votes_cast * 12 + K_CLASSES * 640 * 480 * 8, "hbm"),
                "centres": (K_CLASSES * 640 * 480 * (8 + 4), "hbm"),
                "pose": (2 * (Pp * T_TREES * 4 + votes_cast * 12), "hbm"),
            }
            stages = {}
            for name, ms in zip(api.STAGE_NAMES, stage_ms):
                work, bound = alg[name]
                if ms <= 0:
                    continue
                if bound == "tensor":
                    ach = work / (ms * 1e-3) / 1e12
                    stages[name] = {"ms": float(ms), "bound": bound, "achieved": ach, "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"]}
                else:
                    ach = work / (ms * 1e-3) / 1e9
                    stages[name] = {"ms": float(ms), "bound": bound, "achieved": ach, "unit": "GB/s", "frac": ach / peaks["hbm"]}
            # dominant kernel: encoder layer 2 (K=1536 -> N=1024 padded; algorithmic 1500 x 1000), timed alone in the serial
            # pass with the SM clock at its maximum -> the burst bf16 peak is the denominator (the sustained figure is for a
            # kernel inside a long power-limited tensor step; this path spends ~1/3 of a frame on the tensor pipe)
            l2_flop = 2.0 * Pp * 1500 * 1000
            l2_ach = l2_flop / (enc_ms[1] * 1e-3) / 1e12 if enc_ms[1] > 0 else 0.0
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    traffic = json.load(f).get("encoder_layer_2", {}).get("dram_bytes_per_launch")
            except Exception:
                pass
            roofline = {"kernel": "encoder_layer_kernel<256,false,6,1,2,64,8> (layer 2: 1500->1000, CTA pairs, tcgen05 cta_group::2; one-slot context)",
                        "bound": "tensor", "achieved": l2_ach,
                        "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": l2_ach / peaks["tf_burst"],
                        "traffic": traffic, "peak_source": peaks["source"] + " (burst bf16: kernel timed alone, SM clock at max)",
                        "frac_of_sustained_peak": l2_ach / peaks["tf_sustained"],
                        "frac_of_nominal_dense_peak": l2_ach / 2250.0,
                        "note": "achieved counts ALGORITHMIC flops (1500 x 1000 per patch; the kernel multiplies the padded 1536 x 1024); "
                                "the measured peak is a cuBLAS bf16 GEMM on this pool's B200s, so a fraction near or above 1 means "
                                "the kernel matches the library's throughput, not that it exceeds the hardware (nominal 2250 TFLOP/s)",
                        "encoder_layer_ms": [float(x) for x in enc_ms],
                        "encoder_stage_tflops": float(Pp * ENC_FLOP_PER_PATCH / (sum(enc_ms) * 1e-3) / 1e12) if sum(enc_ms) > 0 else 0.0}
            e2e_fps = frames_total / e2e_s
            line = {
                "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(stats, patches_per_frame=Pp, votes_cast_per_frame=votes_cast,
                                          parallelism=f"frames x{world}" if world > 1 else "1 GPU"),
                "ms_per_frame": ms_frame, "traversals_per_s": fps * Pp * T_TREES,
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "timing": "wall clock around hf6d_submit/hf6d_wait with pinned host frames, device sync both sides",
                        "hypotheses_per_frame": n_hyp / (BATCH * args.steps)},
                "gpu_launches": launches,
                "roofline": roofline, "stages": stages, "stages_note": "serial pass: one frame at a time on a one-slot context (5/6/6-stage encoder "
                "rings), sum = %.3f ms/frame; `value` and `e2e` run %d frames in flight on separate streams of a %d-slot context, whose "
                "encoder kernels trade one ring stage (4/5/5) for room beside them: other frames' CTAs co-reside, +3-5 %% frames/s"
                % (float(np.sum(stage_ms)), n_slots, n_slots),
                "cpu_baseline": cpu, "clocks": clocks,
            }
            if tree is not None:
                line["tree_sharded"] = tree
            return line

        # ---- the north star's multi-GPU mode: trees sharded over the ranks, ONE exchange step per frame (NCCL
        # all-reduce SUM of the Q16 vote maps + MAX of the leaf table), every rank works on the same frames
        tree = None
        if world > 1:
            import threading
            from object_detector_6d_b200 import sharded
            tree_modes = {}
            tree_done = threading.Event()

            def watchdog():  # a rank that dies inside the exchange would leave the others waiting on its flags for ever
                if not tree_done.wait(timeout=args.tree_timeout):
                    if rank == 0:
                        print(json.dumps(build_line({"unavailable": "the tree-sharded arm did not finish within %d s" % args.tree_timeout,
                                                     "exchanges": tree_modes})), flush=True)
                    os._exit(0)
            threading.Thread(target=watchdog, daemon=True).start()
            for mode in ("peer", "nccl"):
                try:
                    sd = sharded.TreeShardedDetector(forest_dir, wpath, p, device=local_rank, n_slots=n_slots, exchange=mode)
                except Exception as e:  # e.g. no P2P path between the GPUs: the NCCL exchange still runs
                    tree_modes[mode] = {"unavailable": str(e).splitlines()[0][:200]}
                    continue
                for s in range(n_slots):
                    sd.det.bind_frame(s, bgr_all[0].data_ptr(), dep_all[0].data_ptr())

                def step_tree():
                    for i in range(BATCH):
                        s = i % n_slots
                        j = i % DISTINCT_FRAMES
                        sd.det.bind_frame(s, bgr_all[j].data_ptr(), dep_all[j].data_ptr())
                        sd.run(s)

                main_t = torch.cuda.Stream()
                for _ in range(args.warmup):
                    step_tree()
                torch.cuda.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record(main_t)
                for st_ in sd.streams:
                    st_.wait_event(t0e)
                for _ in range(args.steps):
                    step_tree()
                for st_ in sd.streams:
                    ev = torch.cuda.Event()
                    ev.record(st_)
                    main_t.wait_event(ev)
                t1e.record(main_t)
                main_t.synchronize()
                t = torch.tensor([t0e.elapsed_time(t1e)], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_tree = float(t.item())
                tree_modes[mode] = {"frames_per_s": BATCH * args.steps / (ms_tree * 1e-3), "ms_per_frame": ms_tree / (BATCH * args.steps),
                                    "kernel_launches_per_frame": sd.launches_per_frame()}
                trees_per_rank, classes_per_rank = len(sd.trees), len(sd.classes)
                for s in range(n_slots):
                    sd.det.bind_frame(s, None, None)
                sd.close()
            tree_done.set()
            best = max((m for m in tree_modes if "frames_per_s" in tree_modes[m]), key=lambda m: tree_modes[m]["frames_per_s"])
            tree = dict(tree_modes[best])
            tree.update({
                "exchange": best, "exchanges": tree_modes,
                "trees_per_rank": trees_per_rank, "classes_per_rank": classes_per_rank,
                "exchange_bytes_per_frame": {"nccl": int(K_CLASSES * 640 * 480 * 8 + counts[0][1] * T_TREES * 4),
                                             "peer": int((world - 1) * classes_per_rank * 640 * 480 * 8
                                                         + counts[0][1] * (T_TREES - trees_per_rank) * 4)},
                "scaling": "strong (same frames on every rank, trees t % N == rank)",
                "note": "scan/gather/encode are replicated (every rank needs all features); traverse+vote are sharded by tree, "
                        "centres+pose by class after the exchange; %d frames in flight.  exchange 'peer': the blur's row pass and the "
                        "pose stage read the other ranks' vote maps / leaf tables in place over NVLink (CUDA IPC, flags in peer "
                        "memory, no collective); 'nccl': all-reduce SUM of the maps + MAX of the leaf table" % n_slots})

        line = build_line(tree)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slots", type=int, default=4, help="frames in flight (one stream + workspace each)")
    ap.add_argument("--forest", default="trained", choices=["trained", "random"],
                    help="synthetic forest: leaf payloads from labelled patches (coherent votes) or uniformly random votes")
    ap.add_argument("--tree-timeout", type=int, default=240, help="seconds the tree-sharded arm (N > 1) may take")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
