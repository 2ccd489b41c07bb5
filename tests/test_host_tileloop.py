"""CPU: host-side tile-loop logic (sam_road_b200/inferencer.py) against the oracle's line-faithful
restatement of inferencer.py / graph_extraction.py, plus a world_size-2 gloo run of the sharding +
all-gather exchange used at N > 1."""
import os

import numpy as np
import pytest
import torch

from oracle import samroad_oracle as O
from sam_road_b200 import inferencer as I

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_tile_grid_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "tileloop.npz"))
    for tag, args in {"c2_cityscale_16": (0, 2048, 64, 512, 16), "c4_cityscale_8": (0, 2048, 64, 512, 8),
                      "c3_spacenet_16": (0, 400, 0, 256, 16), "spacenet_4": (0, 400, 0, 256, 4)}.items():
        mine = np.array([[x0, y0, x1, y1] for _, (x0, y0), (x1, y1) in I.get_patch_info_one_img(*args)])
        assert np.array_equal(mine, g[f"tiles_{tag}"]), tag


def test_keypoint_extraction_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "tileloop.npz"))
    cfg = dict(ITSC_THRESHOLD=0.3, ROAD_THRESHOLD=0.4, ITSC_NMS_RADIUS=8, ROAD_NMS_RADIUS=16)
    pts = I.extract_graph_points(g["kp_mask"], g["road_mask"], cfg)
    assert np.array_equal(pts, g["graph_points"])
    empty = I.extract_graph_points(np.zeros((64, 64), np.uint8), np.zeros((64, 64), np.uint8), cfg)
    assert empty.shape == (0, 2)


def test_pair_queries_and_edge_aggregation_match_oracle_loop():
    rng = np.random.RandomState(0)
    gp = np.unique(rng.randint(0, 400, size=(300, 2)), axis=0).astype(np.int64)
    tiles = I.get_patch_info_one_img(0, 400, 0, 256, 4)
    K, R = 16, 64.0
    mine = [I.build_pair_queries(gp, t, K, R) for t in tiles]
    ref = [O.build_pair_queries(gp, t, K, R) for t in tiles]
    for a, b in zip(mine, ref):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # prefix-valid masks (neighbours sorted by distance, misses at the end): SURVEY.md §8a a17
    for _, _, _, valid in mine:
        assert np.all(valid[:, :-1] >= valid[:, 1:])
    # edge aggregation: reference triple loop (inferencer.py:210-229) vs the vectorised version
    nmax = max(q[1].shape[0] for q in mine)
    scores = [rng.rand(nmax, K).astype(np.float32) for _ in tiles]
    from collections import defaultdict
    es, ec = defaultdict(float), defaultdict(float)
    for ti, (idx, pts, pairs, valid) in enumerate(mine):
        for si in range(pts.shape[0]):
            for pi in range(K):
                if not valid[si, pi]:
                    continue
                s, t = pairs[si, pi]
                es[(idx[s], idx[t])] += scores[ti][si, pi]
                ec[(idx[s], idx[t])] += 1.0
    for thr in (0.3, 0.5, 0.7):
        ref_edges = np.array([e for e, s in es.items() if s / ec[e] > thr]).reshape(-1, 2)
        got = I.aggregate_edges([q[2] for q in mine], [q[3] for q in mine], [q[0] for q in mine],
                                scores, thr)
        assert np.array_equal(got, ref_edges)      # same edges, same (first-occurrence) order
    assert I.aggregate_edges([], [], [], [], 0.5).shape == (0, 2)


def test_sat2graph_format():
    nodes = np.array([[10.2, 20.7], [30.0, 40.0], [5.0, 5.0]])
    edges = np.array([[0, 1], [1, 2]])
    g = I.convert_to_sat2graph_format(nodes, edges)
    assert g[(10, 21)] == [(30, 40)] and set(g[(30, 40)]) == {(10, 21), (5, 5)}


def _gloo_worker(rank, world, port, n_tiles, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi, per = I._shard(n_tiles, rank, world)
        full = torch.arange(n_tiles * 6, dtype=torch.float32).view(n_tiles, 3, 2) * 0.5   # "scores"
        gathered = torch.zeros((per * world, 3, 2))
        mine = gathered[rank * per: rank * per + per]
        mine[: hi - lo] = full[lo:hi]
        dist.all_gather_into_tensor(gathered, mine.clone())
        ok = torch.equal(gathered[:n_tiles], full) and bool((gathered[n_tiles:] == 0).all())
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(bool(t.item() == 1.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [16, 7])
def test_shard_and_allgather_world2_gloo(n_tiles):
    """N>1 exchange step on CPU: contiguous block sharding + one all_gather_into_tensor reproduces
    the tile-ordered tensor of the single-rank run (ragged last shard included)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_tiles
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_tiles, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
    covered = []
    for r in range(2):
        lo, hi, _ = I._shard(n_tiles, r, 2)
        covered += list(range(lo, hi))
    assert covered == list(range(n_tiles))


def test_shard_partitions_every_case():
    """_shard: contiguous blocks in rank order that partition range(n) for any (n, world), equal
    `per_rank` on every rank (the all-gather needs equal shard shapes), empty tail shards allowed."""
    for world in range(1, 10):
        for n in range(0, 41):
            covered, pers = [], set()
            for r in range(world):
                lo, hi, per = I._shard(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                covered += list(range(lo, hi))
                pers.add(per)
            assert covered == list(range(n)) and len(pers) == 1
            assert pers.pop() * world >= n
