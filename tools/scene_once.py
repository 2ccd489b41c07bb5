"""One C2-grid scene through infer_one_img (for `ncu --metrics gpu__time_duration.sum` launch lists of the
graph stage): python tools/scene_once.py [numpy|stable]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam_road_b200 import SAMRoad, synth  # noqa: E402
from sam_road_b200.inferencer import infer_one_img  # noqa: E402

tie = sys.argv[1] if len(sys.argv) > 1 else "numpy"
cfg = dict(SAM_VERSION="vit_b", PATCH_SIZE=512, USE_SAM_DECODER=False, ENCODER_LORA=False, TOPONET_VERSION="normal",
           NO_SAM=False, INFER_BATCH_SIZE=64, SAMPLE_MARGIN=64, INFER_PATCHES_PER_EDGE=16, TOPO_THRESHOLD=0.5,
           ITSC_NMS_RADIUS=8, ROAD_NMS_RADIUS=16, NEIGHBOR_RADIUS=64, MAX_NEIGHBOR_QUERIES=16, ITSC_THRESHOLD=2.0,
           ROAD_THRESHOLD=2.0)
dev = torch.device("cuda:0")
net = SAMRoad(cfg)
net.load_state_dict(synth.make_state_dict(cfg, seed=0, logit_gain=6.0), strict=True)
net.eval().to(dev)
img = np.random.RandomState(17).randint(0, 256, size=(2048, 2048, 3)).astype(np.uint8)
_, _, kp, road = infer_one_img(net, img, cfg, device=dev)
cfg.update(ITSC_THRESHOLD=float(np.quantile(kp, 0.996)) / 255, ROAD_THRESHOLD=float(np.quantile(road, 0.95)) / 255)
tm = {}
torch.cuda.nvtx.range_push("scene")
nodes, edges, kp, road = infer_one_img(net, img, cfg, device=dev, nms_tie_order=tie, timings=tm)
torch.cuda.nvtx.range_pop()
print(tie, nodes.shape, edges.shape, {k: round(1e3 * v, 2) for k, v in tm.items() if k.endswith("_s")})
