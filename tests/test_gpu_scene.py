"""Scene-level parity of sam_road_b200.inferencer.infer_one_img against the oracle's restatement of
inferencer.py:61-234 (SURVEY.md §4 level iii): uint8 masks equal within 1 LSB; the graph built from
identical masks has the same nodes, and edges differ only where the mean topology score lies within
2e-3 of TOPO_THRESHOLD."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import samroad_oracle as O  # noqa: E402
from sam_road_b200 import SAMRoad, synth  # noqa: E402
from sam_road_b200.inferencer import fuse_masks_device, get_patch_info_one_img, infer_one_img, infer_scenes  # noqa: E402

DEV = "cuda:0"


def _scene_cfg(patch, per_edge, margin):
    return dict(SAM_VERSION="vit_b", PATCH_SIZE=patch, USE_SAM_DECODER=False, ENCODER_LORA=False,
                TOPONET_VERSION="normal", NO_SAM=False, INFER_BATCH_SIZE=6, SAMPLE_MARGIN=margin,
                INFER_PATCHES_PER_EDGE=per_edge, ITSC_THRESHOLD=0.56, ROAD_THRESHOLD=0.50,
                TOPO_THRESHOLD=0.5, ITSC_NMS_RADIUS=8, ROAD_NMS_RADIUS=16, NEIGHBOR_RADIUS=64,
                MAX_NEIGHBOR_QUERIES=16)


def test_fuse_masks_bit_exact():
    """The fusion kernel adds tiles in tile-list order like inferencer.py:99-104: bit-exact uint8."""
    tiles = get_patch_info_one_img(0, 400, 0, 256, 4)
    g = torch.Generator().manual_seed(0)
    scores = torch.rand((len(tiles), 256, 256, 2), generator=g)
    kp, road = fuse_masks_device(scores.to(DEV), tiles, 400, 400)
    okp, oroad = O.fuse_masks(list(scores.numpy()), tiles, 400, 400)
    assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(road.cpu().numpy(), oroad)
    tiles_m = get_patch_info_one_img(0, 700, 64, 256, 3)     # uncovered border -> 0 (NaN cast)
    sc = torch.rand((len(tiles_m), 256, 256, 2), generator=g)
    kp, road = fuse_masks_device(sc.to(DEV), tiles_m, 700, 700)
    okp, oroad = O.fuse_masks(list(sc.numpy()), tiles_m, 700, 700)
    assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(road.cpu().numpy(), oroad)
    assert kp[:64].sum().item() == 0


def _calibrated_scene(cfg, img, seed, gain, cand_frac=(0.004, 0.05)):
    """Model + thresholds for a synthetic scene.  Random weights give noise-like masks, so the two
    thresholds are set from OUR fused masks' quantiles (the masks do not depend on them) to get a
    realistic candidate density: ~0.4 % intersection and ~5 % road pixels."""
    sd = synth.make_state_dict(cfg, seed=seed, logit_gain=gain)
    net = SAMRoad(cfg)
    net.load_state_dict(sd, strict=True)
    net.eval().to(DEV)
    probe = dict(cfg, ITSC_THRESHOLD=2.0, ROAD_THRESHOLD=2.0)           # nothing passes: masks only
    nodes, edges, kp, road = infer_one_img(net, img, probe, device=torch.device(DEV))
    assert nodes.shape == (0, 2) and edges.shape == (0, 2)
    cfg = dict(cfg, ITSC_THRESHOLD=float(np.quantile(kp, 1 - cand_frac[0])) / 255,
               ROAD_THRESHOLD=float(np.quantile(road, 1 - cand_frac[1])) / 255)
    return net, sd, cfg, (kp, road)


def _check_scene(cfg, img, seed, gain, cand_frac, tie, min_points, require_edges=True):
    net, sd, cfg, masks0 = _calibrated_scene(cfg, img, seed, gain, cand_frac)
    timings = {}
    nodes, edges, kp, road = infer_one_img(net, img, cfg, device=torch.device(DEV), timings=timings,
                                           nms_tie_order=tie)
    assert np.array_equal(kp, masks0[0]) and np.array_equal(road, masks0[1])       # deterministic
    assert kp.dtype == np.uint8 and kp.shape == img.shape[:2] and nodes.dtype == np.int64
    spec = O.ModelSpec.from_config(cfg)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    # (1) masks against the oracle model + reference fusion
    o_nodes, o_edges, o_kp, o_road = O.infer_one_img(sd_dev, spec, img, cfg, tie_order=tie, knn_ties="index")
    assert np.abs(kp.astype(int) - o_kp.astype(int)).max() <= 1
    assert np.abs(road.astype(int) - o_road.astype(int)).max() <= 1
    # (2) the graph stage, exactly: continue the reference loop from OUR masks with OUR topology scores
    tiles = get_patch_info_one_img(0, img.shape[0], cfg["SAMPLE_MARGIN"], cfg["PATCH_SIZE"],
                                   cfg["INFER_PATCHES_PER_EDGE"])
    bs = cfg["INFER_BATCH_SIZE"]
    img_d = torch.as_tensor(img).to(DEV)
    my_feats = []
    for b0 in range(0, len(tiles), bs):
        rgb = torch.stack([img_d[y0:y1, x0:x1, :] for _, (x0, y0), (x1, y1) in tiles[b0:b0 + bs]], 0)
        my_feats.append(net.infer_masks_and_img_features(rgb)[1])

    def cuda_topo(bi, _feat, pts, prs, val):
        return net.infer_toponet(my_feats[bi], pts, prs, val)

    ex = O.infer_one_img(sd_dev, spec, img, cfg, masks_override=(kp, road), tie_order=tie, knn_ties="index",
                         topo_fn=cuda_topo)
    assert np.array_equal(nodes, ex[0]), (nodes.shape, ex[0].shape)
    assert nodes.shape[0] >= min_points, nodes.shape
    assert np.array_equal(edges, ex[1]), (edges.shape, ex[1].shape)         # same edges, same order
    # (3) and with the oracle TopoNet: edges may differ only at the threshold
    o2 = O.infer_one_img(sd_dev, spec, img, cfg, masks_override=(kp, road), tie_order=tie, knn_ties="index",
                         return_edge_scores=True)
    mine = {tuple(e) for e in edges.astype(np.int64).tolist()}
    ref = {tuple(int(v) for v in e) for e in o2[1].tolist()}
    for e in mine ^ ref:
        assert abs(o2[4][e] - cfg["TOPO_THRESHOLD"]) < 2e-3, (e, o2[4][e])
    assert len(ref) > 0 or not require_edges
    return timings, cfg


def test_infer_one_img_scene_parity():
    """Small scene (4x4 tiles of 256 on 400^2, batches of 6: ragged last batch), both tie orders."""
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(400, 400, 3)).astype(np.uint8)
    for tie in ("numpy", "stable"):
        _check_scene(_scene_cfg(256, 4, 0), img, seed=7, gain=6.0, cand_frac=(0.01, 0.08), tie=tie, min_points=10)


@pytest.mark.parametrize("name,size,patch,per_edge,margin,cand", [
    ("c3_spacenet_16x16", 400, 256, 16, 0, (0.004, 0.05)),       # 256 tiles of 256^2 on a 400^2 scene
    ("c4_cityscale_8x8", 2048, 512, 8, 64, (0.004, 0.05)),       # 64 tiles of 512^2, one batch
    ("c2_cityscale_16x16", 2048, 512, 16, 64, (0.004, 0.05)),    # 256 tiles of 512^2, four batches of 64
])
def test_infer_one_img_baseline_grids(name, size, patch, per_edge, margin, cand, report_dir):
    """The BASELINE scene grids with INFER_BATCH_SIZE = 64."""
    import json
    rng = np.random.RandomState(size + per_edge)
    img = rng.randint(0, 256, size=(size, size, 3)).astype(np.uint8)
    cfg = dict(_scene_cfg(patch, per_edge, margin), INFER_BATCH_SIZE=64)
    timings, cfg = _check_scene(cfg, img, seed=11, gain=6.0, cand_frac=cand, tie="numpy", min_points=50)
    path = os.path.join(report_dir, "scene_parity.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old[name] = {k: v for k, v in timings.items()}
    json.dump(old, open(path, "w"), indent=1, sort_keys=True, default=str)


def test_scene_with_dense_candidates():
    """Thresholds that let 30 % / 60 % of the pixels through (what the reference's default thresholds do on an
    untrained model): every candidate has hundreds of candidates inside its NMS disc."""
    rng = np.random.RandomState(8)
    img = rng.randint(0, 256, size=(400, 400, 3)).astype(np.uint8)
    for tie in ("numpy", "stable"):
        timings, _ = _check_scene(dict(_scene_cfg(256, 4, 0), INFER_BATCH_SIZE=16), img, seed=9, gain=6.0,
                                  cand_frac=(0.3, 0.6), tie=tie, min_points=100)
        assert sum(timings["graph_stats"]["candidates"]) > 100000


def test_scene_with_empty_tiles_and_no_keypoints():
    """C4 grid with thresholds so high that only a handful of keypoints survive: most tiles hold no
    point (zero-row queries inside a non-empty batch), and with thresholds nothing passes the early
    return of inferencer.py:123-124 is taken."""
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, size=(2048, 2048, 3)).astype(np.uint8)
    cfg = dict(_scene_cfg(512, 8, 64), INFER_BATCH_SIZE=64)
    timings, cfg2 = _check_scene(cfg, img, seed=13, gain=6.0, cand_frac=(2e-6, 8e-6), tie="numpy", min_points=4,
                                 require_edges=False)
    assert timings["n_points"] < 64        # fewer points than tiles: empty tiles are certain


def test_infer_scenes_pipeline_equals_sequential():
    """The multi-scene driver (images pulled by a loader thread while the GPU works) returns exactly what
    infer_one_img returns scene by scene, in order -- scenes of different content AND different size."""
    cfg = dict(_scene_cfg(256, 4, 0), INFER_BATCH_SIZE=16, ITSC_THRESHOLD=0.53, ROAD_THRESHOLD=0.5)
    sd = synth.make_state_dict(cfg, seed=7, logit_gain=6.0)
    net = SAMRoad(cfg)
    net.load_state_dict(sd, strict=True)
    net.eval().to(DEV)
    rng = np.random.RandomState(9)
    imgs = [rng.randint(0, 256, size=(sz, sz, 3)).astype(np.uint8) for sz in (400, 400, 512, 400, 300)]
    seq = [infer_one_img(net, im, cfg, device=torch.device(DEV)) for im in imgs]

    def gen():
        for im in imgs:
            yield im
    pipe = list(infer_scenes(net, gen(), cfg, device=torch.device(DEV)))
    assert len(pipe) == len(seq)
    for a, b in zip(seq, pipe):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert sum(r[0].shape[0] for r in seq) > 50
    with pytest.raises(RuntimeError):          # loader errors surface in the consumer
        def bad():
            yield imgs[0]
            raise RuntimeError("decode failed")
        list(infer_scenes(net, bad(), cfg, device=torch.device(DEV)))
