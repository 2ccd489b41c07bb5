"""Scene-level parity of sam_road_b200.inferencer.infer_one_img against the oracle's restatement of
inferencer.py:61-234 (SURVEY.md §4 level iii): uint8 masks equal within 1 LSB; the graph built from
identical masks has the same nodes, and edges differ only where the mean topology score lies within
2e-3 of TOPO_THRESHOLD."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import samroad_oracle as O  # noqa: E402
from sam_road_b200 import SAMRoad, synth  # noqa: E402
from sam_road_b200.inferencer import fuse_masks_device, get_patch_info_one_img, infer_one_img  # noqa: E402

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


def test_infer_one_img_scene_parity():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = _scene_cfg(256, 4, 0)
    sd = synth.make_state_dict(cfg, seed=7, logit_gain=6.0)
    net = SAMRoad(cfg)
    net.load_state_dict(sd, strict=True)
    net.eval().to(DEV)
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(400, 400, 3)).astype(np.uint8)
    timings = {}
    nodes, edges, kp, road = infer_one_img(net, img, cfg, device=torch.device(DEV), timings=timings)
    assert kp.dtype == np.uint8 and kp.shape == (400, 400) and timings["n_tiles"] == 16

    spec = O.ModelSpec.from_config(cfg)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    o_nodes, o_edges, o_kp, o_road = O.infer_one_img(sd_dev, spec, img, cfg)
    assert np.abs(kp.astype(int) - o_kp.astype(int)).max() <= 1
    assert np.abs(road.astype(int) - o_road.astype(int)).max() <= 1
    # continue the oracle from OUR masks: identical keypoints, then compare the graphs
    o2 = O.infer_one_img(sd_dev, spec, img, cfg, masks_override=(kp, road), return_edge_scores=True)
    assert np.array_equal(nodes, o2[0])
    assert nodes.shape[0] > 10, "test scene should produce keypoints"
    mine = {tuple(e) for e in edges.tolist()}
    ref = {tuple(e) for e in o2[1].tolist()}
    for e in mine ^ ref:
        assert abs(o2[4][e] - cfg["TOPO_THRESHOLD"]) < 2e-3, (e, o2[4][e])
    assert len(ref) > 0
