"""Tree sharding: host-side logic on CPU with gloo at world_size 2, and the NCCL path on 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_owned_trees_partition():
    from object_detector_6d_b200.sharded import owned_trees
    for T in (1, 4, 7, 80):
        for world in (1, 2, 4, 8):
            parts = [owned_trees(r, world, T) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(T))
            assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1


def test_owned_patches_partition():
    from object_detector_6d_b200.sharded import owned_patches
    for Pp in (0, 100, 128, 129, 70900, 290600):
        for world in (1, 2, 3, 4, 8):
            parts = [owned_patches(r, world, Pp) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == Pp
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 128 == 0 for lo, _ in parts)


def _gloo_worker(rank, world, port, tmpdir, out_q, split="trees"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from object_detector_6d_b200 import sharded, synth
    from oracle import oracle as O
    from tests.helpers import make_case
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cam = synth.Camera(320, 240, 287.5, 287.5, 159.5, 119.5)
        # every rank builds the same seeded case in its own directory
        cs = make_case(os.path.join(tmpdir, f"r{rank}"), K=2, T=3, seed=5, max_depth=8, votes_per_leaf=3, cam=cam,
                       calib_patches=2000)
        p = cs["params"]
        forest = O.Forest(cs["forest_dir"])
        locs = O.scan_centres(cs["depth"], p)
        Pp = (len(locs) // p.batch_size) * p.batch_size
        feats = O.encode(O.normalise(O.gather(cs["bgr"], cs["depth"], p, locs[:Pp])), cs["layers"])
        _, ords = O.traverse(forest, feats)
        # this rank's shard: foreign trees / patches are -1, exactly what the device leaves in its leaf table
        part = np.full_like(ords, -1)
        if split == "patches":
            lo, hi = sharded.owned_patches(rank, world, Pp)
            part[lo:hi] = ords[lo:hi]
        else:
            mine = sharded.owned_trees(rank, world, forest.T)
            part[:, mine] = ords[:, mine]
        maps_part, _ = O.vote(forest, part, locs[:Pp], cs["depth"], p)
        t_maps = torch.from_numpy(maps_part.view(np.int64).copy())
        t_leaf = torch.from_numpy(part.copy())
        sharded.exchange(t_maps, t_leaf)
        maps_full, _ = O.vote(forest, ords, locs[:Pp], cs["depth"], p)
        ok_maps = np.array_equal(t_maps.numpy().view(np.uint64), maps_full)
        ok_leaf = np.array_equal(t_leaf.numpy(), ords)
        hyp_full = O.hypotheses(forest, ords, locs[:Pp], cs["depth"], p)
        hyp_merged = O.hypotheses(forest, t_leaf.numpy(), locs[:Pp], cs["depth"], p)
        ok_hyp = len(hyp_full) > 0 and len(hyp_full) == len(hyp_merged) and all(
            np.array_equal(hyp_full[n], hyp_merged[n]) for n in hyp_full.dtype.names)
        out_q.put((rank, ok_maps, ok_leaf, ok_hyp, int((maps_part != maps_full).sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split", ["trees", "patches"])
def test_exchange_with_gloo_world_size_2(tmp_path, split):
    """Each rank votes only its own trees / patches (oracle arithmetic); after the exchange both hold the full maps / table."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, str(tmp_path), q, split)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, ok_maps, ok_leaf, ok_hyp, n_diff_before in sorted(res):
        assert n_diff_before > 0, "a shard alone must not already equal the full maps"
        assert ok_maps and ok_leaf and ok_hyp, (rank, ok_maps, ok_leaf, ok_hyp)


# ------------------------------------------------------------------------------------------------ GPU, NCCL
def _nccl_worker(rank, world, port, tmpdir, out_q, mode="nccl", split="trees"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from object_detector_6d_b200 import api, sharded
    from tests.helpers import make_case, to_api_params
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        cs = make_case(os.path.join(tmpdir, f"r{rank}"), K=3, T=5, seed=2, max_depth=12, votes_per_leaf=6)
        p = to_api_params(cs["params"])
        single = api.Detector(cs["forest_dir"], cs["weights"], p, device=rank)
        hyp_single = single.detect(cs["bgr"], cs["depth"])
        maps_single = single.fetch(api.BUF_MAPS)
        leaf_single = single.fetch(api.BUF_LEAF_ORD)
        single.close()
        sd = sharded.TreeShardedDetector(cs["forest_dir"], cs["weights"], p, device=rank, n_slots=2, exchange=mode, split=split)
        hyp = sd.detect(cs["bgr"], cs["depth"], slot=1)
        if mode == "peer":  # several frames through both slots: the flags must keep the ranks in step
            for i in range(6):
                again = sd.detect(cs["bgr"], cs["depth"], slot=i % 2)
                assert len(again) == len(hyp) and all(np.array_equal(again[n], hyp[n]) for n in hyp.dtype.names)
            assert not sd.det.peer_timed_out()
        maps = sd.det.fetch(api.BUF_MAPS, slot=1)
        leaf = sd.det.fetch(api.BUF_LEAF_ORD, slot=1)
        n_launch = sd.launches_per_frame()
        sd.close()
        same = len(hyp) == len(hyp_single) and len(hyp) > 0 and all(
            np.array_equal(hyp[n], hyp_single[n]) for n in hyp.dtype.names)
        if mode == "peer":  # the local buffers stay partial: the sums exist only inside the kernels that read the peers
            if split == "patches":
                lo, hi = sharded.owned_patches(rank, world, leaf_single.shape[0])
                ok_leaf = bool(np.array_equal(leaf[lo:hi], leaf_single[lo:hi]))
            else:
                mine = sharded.owned_trees(rank, world, leaf_single.shape[1])
                ok_leaf = bool(np.array_equal(leaf[:, mine], leaf_single[:, mine]))
            out_q.put((rank, True, ok_leaf, same, n_launch))
        else:
            out_q.put((rank, bool(np.array_equal(maps, maps_single)), bool(np.array_equal(leaf, leaf_single)), same,
                       n_launch))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("split", ["trees", "patches"])
def test_tree_sharded_nccl_equals_single_gpu(tmp_path, split):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, str(tmp_path), q, "nccl", split)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=900) for _ in procs]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, ok_maps, ok_leaf, ok_hyp, n_launch in sorted(res):
        assert ok_maps and ok_leaf and ok_hyp, (rank, ok_maps, ok_leaf, ok_hyp)
        assert n_launch >= 18


@pytest.mark.gpu
@pytest.mark.parametrize("split", ["trees", "patches"])
def test_tree_sharded_peer_exchange_equals_single_gpu(tmp_path, split):
    """The same, with the exchange done by the kernels themselves over peer memory (no NCCL collective on the data path)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, str(tmp_path), q, "peer", split)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=900) for _ in procs]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, ok_maps, ok_leaf, ok_hyp, n_launch in sorted(res):
        assert ok_leaf and ok_hyp, (rank, ok_leaf, ok_hyp)
        assert n_launch >= 18
