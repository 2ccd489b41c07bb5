"""Tile-loop driver: drop-in for the reference's `inferencer.infer_one_img` / CLI (inferencer.py:61-349).

    pred_nodes, pred_edges, keypoint_mask, road_mask = infer_one_img(net, img, config)

Same signature, argument meaning and return values as the reference.  What changes is
where the work happens:
  * the uint8 scene is uploaded once; tiles are cropped on the device and fed to the encoder as uint8
    (the reference converts every crop to float32 on the CPU and copies it synchronously,
    inferencer.py:52-58,94);
  * mask fusion (inferencer.py:79-110) is one kernel that adds the tiles in tile-list order, so the
    uint8 masks are bit-identical to the reference's accumulation for identical scores;
  * with torch.distributed initialised, tiles are sharded over ranks in contiguous blocks, the
    per-tile mask scores are exchanged with ONE all-gather (and the topology scores with another),
    and every rank fuses in global tile order -> identical masks / graph on all ranks and at any
    world size.
Keypoint extraction and kNN pair construction stay on the CPU with the reference's semantics
(graph_extraction.py:130-139, graph_utils.py:572-591, inferencer.py:126-197); the rtree box query is
an inclusive numpy box test with ascending indices; the edge aggregation (inferencer.py:206-230) is
vectorised but keeps the reference's float32 accumulation order.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .model import _cfg_get

TileInfo = Tuple[int, Tuple[int, int], Tuple[int, int]]


# --------------------------------------------------------------------------------------------------
# host helpers (CPU, numpy)
# --------------------------------------------------------------------------------------------------
def get_patch_info_one_img(image_index: int, image_size: int, sample_margin: int, patch_size: int,
                           patches_per_edge: int) -> List[TileInfo]:
    """Tile grid: round(linspace(margin, size-(P+margin), n)), x outer / y inner (dataset.py:56-67)."""
    lo, hi = sample_margin, image_size - (patch_size + sample_margin)
    origins = [round(v) for v in np.linspace(start=lo, stop=hi, num=patches_per_edge)]
    return [(image_index, (x, y), (x + patch_size, y + patch_size)) for x in origins for y in origins]


def nms_points(points: np.ndarray, scores: np.ndarray, radius: float) -> np.ndarray:
    """Greedy radius NMS in descending score order; scores > 1 are never suppressed
    (graph_utils.py:572-591)."""
    import scipy.spatial
    order = np.argsort(scores)[::-1]
    pts, sc = points[order, :], scores[order]
    if pts.shape[0] == 0:
        return pts
    kept = np.ones(order.shape[0], dtype=bool)
    tree = scipy.spatial.KDTree(pts)
    for i in range(pts.shape[0]):
        if not kept[i]:
            continue
        nbr = tree.query_ball_point(pts[i], r=radius)
        kept[nbr] = sc[nbr] > 1.0
        kept[i] = True
    return pts[kept]


def extract_graph_points(keypoint_mask: np.ndarray, road_mask: np.ndarray, config) -> np.ndarray:
    """Threshold + 3x NMS, intersections prioritised (graph_extraction.py:24-28,130-139) -> [N,2] xy."""
    def candidates(mask, thr):
        sel = mask > thr
        rc = np.column_stack(np.where(sel))
        return rc[:, ::-1], mask[sel]
    p0, s0 = candidates(keypoint_mask, _cfg_get(config, "ITSC_THRESHOLD") * 255)
    k0 = nms_points(p0, s0, _cfg_get(config, "ITSC_NMS_RADIUS"))
    p1, s1 = candidates(road_mask, _cfg_get(config, "ROAD_THRESHOLD") * 255)
    k1 = nms_points(p1, s1, _cfg_get(config, "ROAD_NMS_RADIUS"))
    pts = np.concatenate([k0, k1], axis=0)
    pri = np.concatenate([np.ones(k0.shape[0]), np.zeros(k1.shape[0])], axis=0)
    return nms_points(pts, pri, _cfg_get(config, "ROAD_NMS_RADIUS"))


def build_pair_queries(graph_points: np.ndarray, tile: TileInfo, max_nbr: int, radius: float):
    """Pair queries of one tile (inferencer.py:148-176): points inside the tile box (inclusive),
    kNN (k+1, drop self) within `radius`, prefix-valid mask, invalid slots point back at the source."""
    import scipy.spatial
    _, (x0, y0), (x1, y1) = tile
    gx, gy = graph_points[:, 0], graph_points[:, 1]
    idx = np.nonzero((gx >= x0) & (gx <= x1) & (gy >= y0) & (gy <= y1))[0]
    n = idx.shape[0]
    pts = graph_points[idx, :] - np.array([[x0, y0]], dtype=graph_points.dtype)
    if n == 0:
        return idx, pts, np.zeros((0, max_nbr, 2), dtype=np.int64), np.zeros((0, max_nbr), dtype=bool)
    _, knn = scipy.spatial.KDTree(pts).query(pts, k=max_nbr + 1, distance_upper_bound=radius)
    knn = knn.reshape(n, -1)[:, 1:]
    src = np.tile(np.arange(n)[:, None], (1, max_nbr))
    valid = knn < n
    tgt = np.where(valid, knn, src)
    return idx, pts, np.stack([src, tgt], axis=-1), valid


def aggregate_edges(all_pairs: Sequence[np.ndarray], all_valid: Sequence[np.ndarray],
                    all_idx: Sequence[np.ndarray], all_scores: Sequence[np.ndarray],
                    threshold: float) -> np.ndarray:
    """Edge aggregation of inferencer.py:206-230, vectorised: per directed (src,tgt) the scores of all
    valid slots are summed in float32 in (tile, sample, pair) order -- the reference's loop order and
    NumPy-2 scalar arithmetic -- then averaged and thresholded; edges keep first-occurrence order."""
    srcs, tgts, vals = [], [], []
    for pairs, valid, idx, scores in zip(all_pairs, all_valid, all_idx, all_scores):
        if pairs.shape[0] == 0:
            continue
        v = valid.reshape(-1)
        p = pairs.reshape(-1, 2)[v]
        srcs.append(idx[p[:, 0]])
        tgts.append(idx[p[:, 1]])
        vals.append(scores[: pairs.shape[0]].reshape(-1)[v].astype(np.float32))
    if not srcs:
        return np.zeros((0, 2), dtype=np.int64)
    src, tgt, val = np.concatenate(srcs), np.concatenate(tgts), np.concatenate(vals)
    val = np.where(np.isnan(val), np.float32(-100.0), val)        # inferencer.py:206
    assert np.all((val >= 0.0) & (val <= 1.0)), "topology score outside [0,1]"   # inferencer.py:219
    key = src.astype(np.int64) * (int(max(src.max(), tgt.max())) + 1) + tgt.astype(np.int64)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    sums = np.zeros(uniq.shape[0], dtype=np.float32)
    np.add.at(sums, inv, val)                                      # sequential float32 adds, in order
    counts = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float32)
    keep = (sums / counts) > np.float32(threshold)
    order = np.argsort(first[keep], kind="stable")                 # dict insertion order
    sel = np.nonzero(keep)[0][order]
    return np.stack([src[first[sel]], tgt[first[sel]]], axis=1)


# --------------------------------------------------------------------------------------------------
# the driver
# --------------------------------------------------------------------------------------------------
def _shard(n_items: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous block of items owned by `rank`; returns (begin, end, per_rank)."""
    per = (n_items + world - 1) // world
    b = min(n_items, rank * per)
    return b, min(n_items, b + per), per


def fuse_masks_device(scores: torch.Tensor, tiles: Sequence[TileInfo], H: int, W: int):
    """scores [n_tiles,P,P,2] fp32 (device, tile-list order) -> uint8 keypoint / road masks [H,W]."""
    dev = scores.device
    P = scores.shape[1]
    x0 = torch.tensor([t[1][0] for t in tiles], dtype=torch.int32, device=dev)
    y0 = torch.tensor([t[1][1] for t in tiles], dtype=torch.int32, device=dev)
    kp = torch.empty((H, W), dtype=torch.uint8, device=dev)
    road = torch.empty((H, W), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().samroad_fuse_masks(scores.data_ptr(), len(tiles), P, x0.data_ptr(),
                                                  y0.data_ptr(), H, W, kp.data_ptr(), road.data_ptr(),
                                                  _lib.current_stream_ptr()), "samroad_fuse_masks")
    return kp, road


def infer_one_img(net, img: np.ndarray, config, device: Optional[torch.device] = None,
                  group=None, timings: Optional[dict] = None, shard: bool = True):
    """Whole-scene inference (inferencer.py:61-234).

    img: uint8 [H,W,3] RGB.  Returns (pred_nodes [N,2] (r,c), pred_edges [E,2], fused_keypoint_mask
    uint8 [H,W], fused_road_mask uint8 [H,W]) -- identical on every rank when run distributed.
    `shard=False` makes a rank process the whole scene alone even if torch.distributed is up."""
    import torch.distributed as dist
    distributed = shard and dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    if device is None:
        device = next(net.parameters()).device
    device = torch.device(device)
    t_start = time.perf_counter()

    H, W = int(img.shape[0]), int(img.shape[1])
    P = int(_cfg_get(config, "PATCH_SIZE"))
    bs = int(_cfg_get(config, "INFER_BATCH_SIZE"))
    tiles = get_patch_info_one_img(0, H, int(_cfg_get(config, "SAMPLE_MARGIN")), P,
                                   int(_cfg_get(config, "INFER_PATCHES_PER_EDGE")))
    n_tiles = len(tiles)
    lo, hi, per = _shard(n_tiles, rank, world)
    my_tiles = tiles[lo:hi]

    # ---- pass 1: masks + image features of the tiles this rank owns -------------------------------
    img_d = torch.as_tensor(np.ascontiguousarray(img), device=device)           # one H2D of the scene
    scores_all = torch.zeros((per * world, P, P, 2), dtype=torch.float32, device=device)
    my_scores = scores_all[rank * per: rank * per + per]
    feats: List[torch.Tensor] = []
    for b0 in range(0, len(my_tiles), bs):
        batch = my_tiles[b0:b0 + bs]
        rgb = torch.stack([img_d[y0:y1, x0:x1, :] for _, (x0, y0), (x1, y1) in batch], 0)
        sc, ft = net.infer_masks_and_img_features(rgb)
        my_scores[b0:b0 + len(batch)].copy_(sc)
        feats.append(ft)
    if world > 1:   # the exchange step: one all-gather of per-tile mask scores
        dist.all_gather_into_tensor(scores_all, my_scores.clone(), group=group)
    kp_d, road_d = fuse_masks_device(scores_all[:n_tiles], tiles, H, W)
    kp_mask, road_mask = kp_d.cpu().numpy(), road_d.cpu().numpy()
    t_pass1 = time.perf_counter()

    # ---- keypoints (CPU, every rank: deterministic) -------------------------------------------------
    graph_points = extract_graph_points(kp_mask, road_mask, config)
    if graph_points.shape[0] == 0:
        return graph_points, np.zeros((0, 2), dtype=np.int32), kp_mask, road_mask
    t_points = time.perf_counter()

    # ---- pass 2: TopoNet on the stored features ---------------------------------------------------------
    K = int(_cfg_get(config, "MAX_NEIGHBOR_QUERIES"))
    R = float(_cfg_get(config, "NEIGHBOR_RADIUS"))
    queries = [build_pair_queries(graph_points, t, K, R) for t in tiles]   # cheap, needed by rank 0
    nmax = max((q[1].shape[0] for q in queries), default=0)
    topo_all = torch.zeros((per * world, max(nmax, 1), K), dtype=torch.float32, device=device)
    my_topo = topo_all[rank * per: rank * per + per]
    if nmax > 0:
        def pad(a):
            return np.pad(a, [(0, nmax - a.shape[0])] + [(0, 0)] * (a.ndim - 1))
        for bi, b0 in enumerate(range(0, len(my_tiles), bs)):
            q = queries[lo + b0: min(lo + b0 + bs, hi)]     # this rank's tiles of the batch only
            if max(x[1].shape[0] for x in q) == 0:       # inferencer.py:188-189
                continue
            pts = torch.as_tensor(np.stack([pad(x[1]) for x in q]), device=device)
            prs = torch.as_tensor(np.stack([pad(x[2]) for x in q]), device=device)
            val = torch.as_tensor(np.stack([pad(x[3]) for x in q]), device=device)
            ts = net.infer_toponet(feats[bi], pts, prs, val)
            my_topo[b0:b0 + len(q)].copy_(ts.squeeze(-1))
    if world > 1:
        dist.all_gather_into_tensor(topo_all, my_topo.clone(), group=group)
    topo_np = topo_all[:n_tiles].cpu().numpy()
    pred_edges = aggregate_edges([q[2] for q in queries], [q[3] for q in queries],
                                 [q[0] for q in queries], list(topo_np),
                                 float(_cfg_get(config, "TOPO_THRESHOLD")))
    pred_nodes = graph_points[:, ::-1]   # to (r, c), inferencer.py:230
    if timings is not None:
        t_end = time.perf_counter()
        timings.update(pass1_s=t_pass1 - t_start, keypoints_s=t_points - t_pass1,
                       pass2_s=t_end - t_points, total_s=t_end - t_start, n_tiles=n_tiles,
                       n_points=int(graph_points.shape[0]))
    return pred_nodes, pred_edges, kp_mask, road_mask


# --------------------------------------------------------------------------------------------------
# CLI (inferencer.py:24-35, 239-349): same flags, same output tree
# --------------------------------------------------------------------------------------------------
class _Cfg(dict):
    """addict.Dict-like: attribute access, missing keys read as an empty (falsy) _Cfg (utils.py:6-9)."""

    def __getattr__(self, k):
        return self[k] if k in self else _Cfg()


def load_config(path: str) -> _Cfg:
    import yaml
    with open(path) as f:
        return _Cfg(yaml.safe_load(f))


def convert_to_sat2graph_format(nodes: np.ndarray, edges: np.ndarray) -> dict:
    """{(r,c): [(r,c) neighbours]} with reverse edges added (graph_utils.py:383-405)."""
    int_nodes = [(round(float(x)), round(float(y))) for x, y in nodes]
    adj: List[list] = [[] for _ in int_nodes]
    for a, b in list(edges) + [e[::-1] for e in edges]:
        if int(b) not in adj[int(a)]:
            adj[int(a)].append(int(b))
    return {int_nodes[i]: [int_nodes[j] for j in nbrs] for i, nbrs in enumerate(adj)}


def main(argv=None):
    import pickle
    from argparse import ArgumentParser

    import cv2
    from .model import SAMRoad
    ap = ArgumentParser()
    ap.add_argument("--checkpoint", default=None, help="checkpoint of the model to test.")
    ap.add_argument("--config", default=None, help="model config.")
    ap.add_argument("--output_dir", default=None, help="Name of the output dir under ./save/")
    ap.add_argument("--device", default="cuda", help="device to use")
    args = ap.parse_args(argv)
    config = load_config(args.config)
    device = torch.device(args.device)
    net = SAMRoad(config)
    ckpt = torch.load(args.checkpoint, map_location="cpu")
    print(f"##### Loading Trained CKPT {args.checkpoint} #####")
    net.load_state_dict(ckpt["state_dict"], strict=True)
    net.eval().to(device)

    if config.DATASET == "cityscale":
        test_ids = [x for x in range(180) if x % 10 == 9 or x % 20 == 8]       # dataset.py:21-41
        rgb_pattern = "./cityscale/20cities/region_{}_sat.png"
    elif config.DATASET == "spacenet":
        import json
        test_ids = json.load(open("./spacenet/data_split.json"))["test"]        # dataset.py:44-53
        rgb_pattern = "./spacenet/RGB_1.0_meter/{}__rgb.png"
    else:
        raise SystemExit(f"config.DATASET must be 'cityscale' or 'spacenet', got {config.DATASET!r}")
    out_dir = f"./save/{args.output_dir}" if args.output_dir else \
        "./save/infer_" + time.strftime("%Y%m%d_%H%M%S")
    for sub in ("mask", "viz", "graph"):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    total = 0.0
    for img_id in test_ids:
        print(f"Processing {img_id}")
        img = cv2.cvtColor(cv2.imread(rgb_pattern.format(img_id)), cv2.COLOR_BGR2RGB)
        t0 = time.time()
        nodes, edges, itsc_mask, road_mask = infer_one_img(net, img, config, device=device)
        total += time.time() - t0
        cv2.imwrite(os.path.join(out_dir, "mask", f"{img_id}_road.png"), road_mask)
        cv2.imwrite(os.path.join(out_dir, "mask", f"{img_id}_itsc.png"), itsc_mask)
        viz = cv2.cvtColor(img.copy(), cv2.COLOR_RGB2BGR)
        for a, b in edges:
            cv2.line(viz, (int(nodes[a][1]), int(nodes[a][0])), (int(nodes[b][1]), int(nodes[b][0])),
                     (15, 160, 253), 4)
        for r, c in nodes:
            cv2.circle(viz, (int(c), int(r)), 4, (0, 255, 255), -1)
        cv2.imwrite(os.path.join(out_dir, "viz", f"{img_id}.png"), viz)
        if config.DATASET == "spacenet":    # r, c -> sat2graph convention (inferencer.py:332-334)
            nodes = np.stack([400 - nodes[:, 0], nodes[:, 1]], axis=1)
        with open(os.path.join(out_dir, "graph", f"{img_id}.p"), "wb") as f:
            pickle.dump(convert_to_sat2graph_format(nodes, np.asarray(edges).reshape(-1, 2)), f)
        print(f"Done for {img_id}.")
    msg = f"Inference completed for {args.config} in {total} seconds."
    print(msg)
    with open(os.path.join(out_dir, "inference_time.txt"), "w") as f:
        f.write(msg)


if __name__ == "__main__":
    main()
