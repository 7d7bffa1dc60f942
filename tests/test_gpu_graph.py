"""GPU tests of the whole-frame CUDA graph (HF6D_GRAPH=1): a frame replayed from the graph of an earlier frame must be the frame
the eager launches give, bit for bit -- for new frame contents in the same buffers (every kernel reads its sizes from device
memory), across configuration changes (the graph is re-captured), with stage-isolated runs in between, and pipelined."""
import numpy as np
import pytest

from tests.helpers import make_case, to_api_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("graph"))
    cs = make_case(d, K=3, T=4, seed=1, max_depth=14, votes_per_leaf=8)
    # a second frame of the same size: the first one shifted and with a hole, so that patch count and votes differ
    bgr2 = np.roll(cs["bgr"], 37, axis=1).copy()
    depth2 = np.roll(cs["depth"], 37, axis=1).copy()
    depth2[100:160, 200:300] = 0
    cs["frames"] = [(cs["bgr"], cs["depth"]), (bgr2, depth2)]
    return cs


def _same(a, b):
    return len(a) == len(b) and all(np.array_equal(a[n], b[n]) for n in a.dtype.names)


def _detector(case, monkeypatch, graph, n_slots=2):
    from object_detector_6d_b200 import api
    if graph:
        monkeypatch.setenv("HF6D_GRAPH", "1")
    else:
        monkeypatch.delenv("HF6D_GRAPH", raising=False)
    return api.Detector(case["forest_dir"], case["weights"], to_api_params(case["params"]), device=0, n_slots=n_slots)


def test_replayed_frames_equal_eager_frames(case, monkeypatch):
    from object_detector_6d_b200 import api
    eager = _detector(case, monkeypatch, False)
    graph = _detector(case, monkeypatch, True)
    try:
        ref = []
        for bgr, depth in case["frames"]:
            hyp = eager.detect(bgr, depth)
            ref.append((hyp, eager.fetch(api.BUF_LEAF_ORD), eager.fetch(api.BUF_MAPS), eager.counts(0), eager.launch_count(0)))
        assert len(ref[0][0]) > 0 and ref[0][3] != ref[1][3]  # the two frames differ in patch count
        # frame 0 three times (eager, captured + replayed, replayed), then frame 1 replayed from frame 0's graph, then back
        for j in (0, 0, 0, 1, 1, 0):
            bgr, depth = case["frames"][j]
            hyp = graph.detect(bgr, depth)
            assert _same(hyp, ref[j][0])
            assert np.array_equal(graph.fetch(api.BUF_LEAF_ORD), ref[j][1])
            assert np.array_equal(graph.fetch(api.BUF_MAPS), ref[j][2])
            assert graph.counts(0) == ref[j][3] and graph.launch_count(0) == ref[j][4]
    finally:
        eager.close()
        graph.close()


def test_configuration_changes_and_partial_runs_invalidate_the_graph(case, monkeypatch):
    from object_detector_6d_b200 import api
    eager = _detector(case, monkeypatch, False)
    graph = _detector(case, monkeypatch, True)
    bgr, depth = case["frames"][0]
    try:
        for _ in range(3):
            h_all = graph.detect(bgr, depth)
        assert _same(h_all, eager.detect(bgr, depth))
        # one class switched off, fewer centres for another: both contexts alike
        for det in (eager, graph):
            det.set_objects(should_detect=[0, 1, 1], max_loc=[12, 2, 12])
        h_ref = eager.detect(bgr, depth)
        assert not _same(h_ref, h_all)
        for _ in range(3):
            assert _same(graph.detect(bgr, depth), h_ref)
        # a stage-isolated run and an injected leaf table between whole frames
        graph.upload(0, bgr, depth)
        graph.run(0, api.STAGE_SCAN, api.STAGE_TRAVERSE)
        leaf = graph.fetch(api.BUF_LEAF_ORD)
        graph.inject(api.BUF_LEAF_ORD, np.zeros_like(leaf))
        graph.run(0, api.STAGE_VOTE, api.STAGE_POSE)
        graph.sync(0)
        for _ in range(3):
            assert _same(graph.detect(bgr, depth), h_ref)
        # encoder mode and feature storage
        for det in (eager, graph):
            det.set_encoder_mode(2)
            det.set_feature_storage(0)
        h2 = eager.detect(bgr, depth)
        for _ in range(3):
            assert _same(graph.detect(bgr, depth), h2)
    finally:
        eager.close()
        graph.close()


def test_pipelined_submit_wait_with_graphs(case, monkeypatch):
    eager = _detector(case, monkeypatch, False, n_slots=1)
    graph = _detector(case, monkeypatch, True, n_slots=3)
    try:
        ref = [eager.detect(b, d) for b, d in case["frames"]]
        order = [0, 1, 0, 0, 1, 1, 0, 1, 0, 1, 1, 0]
        tickets = []
        got = []
        for j in order:
            if len(tickets) == 3:
                got.append(graph.wait(tickets.pop(0)))
            tickets.append(graph.submit(*case["frames"][j]))
        while tickets:
            got.append(graph.wait(tickets.pop(0)))
        assert len(got) == len(order)
        for j, h in zip(order, got):
            assert _same(h, ref[j])
    finally:
        eager.close()
        graph.close()
