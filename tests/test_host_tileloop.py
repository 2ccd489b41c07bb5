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
    """The oracle's restatement of extract_graph_points against the fixture the unmodified reference
    functions produced (tools/make_golden.py).  The fixture was written on an AVX-512 host: np.argsort's
    order of equal priorities in the third NMS pass is CPU-dependent (oracle.visiting_order), so on other
    CPUs only the NMS invariants are checked."""
    g = np.load(os.path.join(GOLD, "tileloop.npz"))
    pts = O.extract_graph_points(g["kp_mask"], g["road_mask"], 0.3, 0.4, 8, 16)
    from numpy._core._multiarray_umath import __cpu_features__ as feat
    if feat.get("AVX512_SKX", False):
        assert np.array_equal(pts, g["graph_points"])
    _check_nms_invariants(pts, g["kp_mask"], g["road_mask"], 0.3 * 255, 0.4 * 255, 16)
    empty = O.extract_graph_points(np.zeros((64, 64), np.uint8), np.zeros((64, 64), np.uint8), 0.3, 0.4, 8, 16)
    assert empty.shape == (0, 2)


def _check_nms_invariants(pts, kp, road, thr0, thr1, radius):
    """What every valid tie order must satisfy after the merged pass: survivors are candidates, pairwise
    farther apart than the radius, and every candidate lies within the radius of a survivor."""
    import scipy.spatial
    assert pts.shape[0] > 0
    assert np.all((kp[pts[:, 1], pts[:, 0]] > thr0) | (road[pts[:, 1], pts[:, 0]] > thr1))
    tree = scipy.spatial.KDTree(pts)
    assert len(tree.query_pairs(r=radius)) == 0
    cand = np.column_stack(np.where((kp > thr0) | (road > thr1)))[:, ::-1]
    d, _ = tree.query(cand, k=1)
    assert d.max() <= radius


def test_nms_tie_orders_and_shortcut():
    """tie_order="stable" is a valid execution of the reference (same invariants, same count class),
    and the all-immune shortcut of the oracle equals the literal loop."""
    rng = np.random.RandomState(1)
    kp = (rng.rand(160, 200) * 255).astype(np.uint8)
    road = (rng.rand(160, 200) * 255).astype(np.uint8)
    for tie in ("numpy", "stable"):
        pts = O.extract_graph_points(kp, road, 0.97, 0.9, 4, 8, tie)
        _check_nms_invariants(pts, kp, road, 0.97 * 255, 0.9 * 255, 8)
        p, sc = np.column_stack(np.where(road > 230))[:, ::-1], road[road > 230]
        assert np.array_equal(O.nms_points(p, sc, 8, tie, shortcut=True), O.nms_points(p, sc, 8, tie, shortcut=False))
    # a threshold below 1/255 admits score 1, which is NOT immune: mixed immune / mortal pass
    low = rng.randint(0, 4, size=(50, 50)).astype(np.uint8)
    p, sc = np.column_stack(np.where(low > 0.5))[:, ::-1], low[low > 0.5]
    out = O.nms_points(p, sc, 3, "stable")
    assert (low[out[:, 1], out[:, 0]] >= 2).sum() == (low >= 2).sum()     # every immune point survives
    assert out.shape[0] < p.shape[0]
    # stable order == argsort(kind="stable")[::-1]
    s = rng.randint(0, 5, size=1000).astype(np.uint8)
    o = O.visiting_order(s, "stable")
    assert np.all(np.diff(s[o].astype(int)) <= 0)
    same = np.diff(s[o].astype(int)) == 0
    assert np.all(np.diff(o)[same] < 0)


def test_pair_queries_index_ties_vs_scipy():
    """knn_by_index pins scipy's order of equidistant neighbours to ascending index; wherever the 17
    nearest distances of a query are all distinct the two must agree exactly."""
    import scipy.spatial
    rng = np.random.RandomState(0)
    gp = np.unique(rng.randint(0, 400, size=(300, 2)), axis=0).astype(np.int64)
    tiles = I.get_patch_info_one_img(0, 400, 0, 256, 4)
    K, R = 16, 64.0
    n_rows = n_tied = 0
    for t in tiles:
        a = O.build_pair_queries(gp, t, K, R, "scipy")
        b = O.build_pair_queries(gp, t, K, R, "index")
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
        pts = a[1]
        d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        for i in range(pts.shape[0]):
            near = np.sort(d2[i][d2[i] < R * R])[: K + 2]
            tied = np.any(np.diff(near) == 0)
            n_rows += 1
            n_tied += int(tied)
            if not tied:
                assert np.array_equal(a[2][i], b[2][i])
            else:   # same distances slot by slot, whatever the order inside a tie group
                da = d2[i][a[2][i, :, 1]][a[3][i]]
                db = d2[i][b[2][i, :, 1]][b[3][i]]
                assert np.array_equal(da, db)
        assert np.all(a[3][:, :-1] >= a[3][:, 1:])     # prefix-valid
    assert n_rows > 500 and 0 < n_tied < n_rows
    # exclusive upper bound, missing slots, border points (inclusive box)
    gp2 = np.array([[0, 0], [64, 0], [0, 63], [256, 256], [255, 200]], dtype=np.int64)
    idx, pts, pairs, valid = O.build_pair_queries(gp2, (0, (0, 0), (256, 256)), K, R, "index")
    assert list(idx) == [0, 1, 2, 3, 4]
    assert list(pairs[0, :2, 1]) == [2, 0] and list(valid[0, :2]) == [True, False]   # (64,0) is NOT < 64 away


def test_batch_plan_layout():
    for n, bs, world in ((256, 64, 1), (256, 64, 8), (64, 64, 4), (7, 4, 2), (5, 64, 8)):
        plan = I.batch_plan(n, bs, world)
        covered = [t for (_, b0, nb) in plan for t in range(b0, b0 + nb)]
        assert covered == list(range(n))
        for r, b0, nb in plan:
            lo, hi, _ = I._shard(n, r, world)
            assert lo <= b0 and b0 + nb <= hi and 0 < nb <= bs


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
