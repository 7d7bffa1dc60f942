"""One stream of frames sharded across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Two splits of a frame's work (TreeShardedDetector(split=...)):

  "trees"    the north star's: rank r owns trees {t : t % N == r} (below).  Scan, gather and encode are REPLICATED (every
             rank needs every patch's features), so the gain stops at ~1.4x however many GPUs there are.
  "patches"  rank r owns a contiguous share of the frame's patches (128-patch row blocks, in the reference's patch order --
             its batches of 100 already run in an OpenMP loop, HFTest.cpp:612): gather, encode, traverse (all trees; the
             forest is a few MB) and vote shard with them, only the scan (0.02 ms) is replicated.  Same exchange, same
             class-sharded mode seeking afterwards.  This is the split that scales.

The rest of this text describes the tree split; the patch split differs only in what a rank's partial maps / leaf table hold.

The reference's voting loop runs trees outermost (HoughForest/src/HFTest.cpp:177) and trees interact only through the
per-class vote maps and the centre->leaf back-map, so rank r of N owns trees {t : t % N == r}:

  every rank      : scan -> gather -> encode the frame (replicated: no traffic, and all ranks need all P' x F features)
  rank r          : traverse + vote its own trees into its own Q16 maps          hf6d_run(SCAN .. VOTE)
  ONE exchange    : all-reduce(SUM) of the maps   [K][H][W] uint64 (as int64)    14.7 MB at 6 x 480 x 640
                    all-reduce(MAX) of the leaf table [cap][T] int32             (entries of foreign trees are -1)
  rank r          : centres + pose mode seeking on the summed maps / merged table hf6d_run(CENTRES .. POSE)
                    for ITS classes (k % N == r, hf6d_set_class_shard): the centres of different classes are
                    independent units of work, so the stage that follows the exchange shards too
  (optional)      : the ranks' hypothesis lists, concatenated in class order, are the unsharded list (gather())

Vote weights are Q16 integers, so the summed maps -- and everything downstream -- are bit-identical to the single-GPU
result whatever N is (the reference's float maps depend on the OpenMP schedule, HFTest.cpp:645-654).

The exchange runs on the slot's own CUDA stream (NCCL is stream-ordered after the vote kernel; no host sync in between),
and every frame slot has its own stream, so the exchange of frame i overlaps the encoder of frame i+1.
`exchange()` itself is backend-agnostic: the CPU test-suite drives it with gloo at world_size 2.

exchange="peer" (GPUs of one NVLink / NVSwitch box) drops the two all-reduces: the ranks map each other's maps and leaf
tables once (CUDA IPC; torch.distributed only carries the handles), and the first kernels after the exchange point read
them in place -- the blur's row pass sums the ranks' maps of THIS rank's classes while it loads them (half of a
reduce-scatter's traffic, no all-gather at all), the pose stage reads a tree's leaf ordinals from the rank that traversed
it -- with flags in peer memory as the only synchronisation (include/hf6d.h, hf6d_peer_attach).
"""
from __future__ import annotations

import numpy as np

from . import api


def owned_trees(rank: int, world: int, T: int):
    """Trees of rank `rank` (the rule hf6d_set_tree_shard implements on the device side)."""
    return [t for t in range(T) if t % world == rank]


def owned_patches(rank: int, world: int, Pp: int):
    """[lo, hi) of the frame's Pp processed patches that rank `rank` gathers, encodes, traverses and votes (the rule
    hf6d_set_patch_shard implements on the device side: cuts on multiples of 128 patches)."""
    mb = (Pp + 127) // 128
    return min(Pp, mb * rank // world * 128), min(Pp, mb * (rank + 1) // world * 128)


def exchange(maps, leaf_table, group=None):
    """The path's one exchange step: in-place SUM of the vote maps and MAX of the leaf table across the group.

    maps: int64 tensor (Q16 sums; uint64 bit pattern, sums stay far below 2^63), leaf_table: int32 tensor with -1 for
    trees a rank does not own.  Works on CUDA tensors (nccl) and CPU tensors (gloo)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    h1 = dist.all_reduce(maps, op=dist.ReduceOp.SUM, group=group, async_op=True)
    h2 = dist.all_reduce(leaf_table, op=dist.ReduceOp.MAX, group=group, async_op=True)
    h1.wait()
    h2.wait()


class TreeShardedDetector:
    """This rank's libhf6d context plus the exchange.  Construct it on every rank after init_process_group."""

    def __init__(self, forest_dir, weights_path, params=None, device=0, n_slots=2, group=None, shard_classes=True,
                 exchange="nccl", split="trees"):
        import torch
        import torch.distributed as dist
        if split not in ("trees", "patches"):
            raise ValueError("split must be 'trees' or 'patches'")
        self.torch = torch
        self.group = group
        self.split = split
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.det = api.Detector(forest_dir, weights_path, params, device=device, n_slots=n_slots)
        if split == "patches":
            self.det.set_patch_shard(self.rank, self.world)
            self.det.set_peer_split(1)
        else:
            self.det.set_tree_shard(self.rank, self.world)
        self.shard_classes = bool(shard_classes) and self.world > 1
        if self.shard_classes:
            self.det.set_class_shard(self.rank, self.world)
        self.exchange_mode = exchange if self.world > 1 else "none"
        if self.exchange_mode == "peer":
            if not self.shard_classes:
                raise ValueError("the peer exchange shards the classes as well")
            # all ranks attach or none does: a rank left out would never write the flags the others wait for
            err = None
            try:
                blob = self.det.peer_export()
            except Exception as e:  # noqa: BLE001
                blob, err = b"", e
            blobs = [None] * self.world
            dist.all_gather_object(blobs, blob, group=group)
            if err is None and all(blobs):
                try:
                    self.det.peer_attach(self.rank, self.world, blobs)
                except Exception as e:  # noqa: BLE001
                    err = e
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None and all(blobs), group=group)  # also: everyone has mapped everything
            if not all(oks):
                self.det.peer_detach()
                self.det.close()
                raise RuntimeError(f"peer exchange unavailable on rank(s) {[r for r, ok in enumerate(oks) if not ok]}: {err}")
        self.device = device
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_slots)]
        self.stream = self.streams[0]
        self._views = []
        for s in range(n_slots):
            self.det.set_stream(s, self.streams[s].cuda_stream)
            maps = torch.as_tensor(self.det.device_array(api.BUF_MAPS, s), device=f"cuda:{device}")
            leaf = torch.as_tensor(self.det.device_array(api.BUF_LEAF_ORD, s), device=f"cuda:{device}")
            self._views.append((maps, leaf))

    @property
    def trees(self):
        return list(range(self.det.T)) if self.split == "patches" else owned_trees(self.rank, self.world, self.det.T)

    @property
    def classes(self):
        """Classes whose centres / poses this rank seeks."""
        K = self.det.K
        return [k for k in range(K) if not self.shard_classes or k % self.world == self.rank]

    def run(self, slot=0):
        """Launch one frame (already uploaded / bound on `slot`) asynchronously on the slot's stream."""
        torch = self.torch
        if self.exchange_mode == "peer":  # nothing to do here: the kernels read peer memory, flags order the ranks
            self.det.run(slot, api.STAGE_SCAN, api.STAGE_POSE)
            self._launches = self.det.launch_count(slot)
            return
        with torch.cuda.stream(self.streams[slot]):
            self.det.run(slot, api.STAGE_SCAN, api.STAGE_VOTE)
            n = self.det.launch_count(slot)
            maps, leaf = self._views[slot]
            exchange(maps, leaf, self.group)
            self.det.run(slot, api.STAGE_CENTRES, api.STAGE_POSE)
            self._launches = n + self.det.launch_count(slot)

    def detect(self, bgr, depth, slot=0, gather=True):
        """One frame.  With gather (default) every rank returns the complete hypothesis list."""
        self.det.upload(slot, bgr, depth)
        self.run(slot)
        mine = self.det.collect(slot)
        return self.gather(mine) if gather and self.shard_classes else mine

    def gather(self, mine):
        """All ranks' hypothesis arrays, merged in class order (= the order of the unsharded list)."""
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, mine, group=self.group)
        allh = np.concatenate(parts) if parts else mine
        return allh[np.argsort(allh["cls"], kind="stable")]

    def launches_per_frame(self) -> int:
        """Kernels of this repo launched by the last run() (the two NCCL all-reduces are not counted)."""
        return getattr(self, "_launches", 0)

    def close(self):
        if self.exchange_mode == "peer":
            import torch.distributed as dist
            self.torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still read
            self.det.peer_detach()
            dist.barrier(group=self.group)
        for s in range(self.det.n_slots):
            self.det.set_stream(s, None)
        self._views = []
        self.det.close()


def merge_numpy(parts_maps, parts_leaf):
    """Host-side statement of the exchange (used by tests): sum of maps, max of leaf tables."""
    maps = np.zeros_like(parts_maps[0])
    leaf = np.full_like(parts_leaf[0], -1)
    for m, lf in zip(parts_maps, parts_leaf):
        maps += m
        leaf = np.maximum(leaf, lf)
    return maps, leaf
