"""Multi-GPU check (run under torchrun, one rank per GPU): infer_one_img sharded over W ranks with the
two NCCL all-gathers returns, on every rank, exactly what a single rank computes alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_scene_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sam_road_b200 import SAMRoad, synth  # noqa: E402
from sam_road_b200.inferencer import infer_one_img  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(SAM_VERSION="vit_b", PATCH_SIZE=256, USE_SAM_DECODER=False, ENCODER_LORA=False,
               TOPONET_VERSION="normal", NO_SAM=False, INFER_BATCH_SIZE=5, SAMPLE_MARGIN=0,
               INFER_PATCHES_PER_EDGE=5, ITSC_THRESHOLD=0.56, ROAD_THRESHOLD=0.50, TOPO_THRESHOLD=0.5,
               ITSC_NMS_RADIUS=8, ROAD_NMS_RADIUS=16, NEIGHBOR_RADIUS=64, MAX_NEIGHBOR_QUERIES=16)
    net = SAMRoad(cfg)
    net.load_state_dict(synth.make_state_dict(cfg, seed=7, logit_gain=6.0), strict=True)
    net.eval().to(dev)
    img = np.random.RandomState(3).randint(0, 256, size=(400, 400, 3)).astype(np.uint8)
    t_sharded, t_alone = {}, {}
    sharded = infer_one_img(net, img, cfg, device=dev, timings=t_sharded)
    alone = infer_one_img(net, img, cfg, device=dev, shard=False, timings=t_alone)
    ok = all(np.array_equal(a, b) for a, b in zip(sharded, alone))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world={dist.get_world_size()} tiles={t_sharded['n_tiles']} nodes={sharded[0].shape[0]} "
              f"edges={sharded[1].shape[0]} identical_on_all_ranks={bool(flag.item() == 1.0)} "
              f"sharded_s={t_sharded['total_s']:.3f} alone_s={t_alone['total_s']:.3f}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
