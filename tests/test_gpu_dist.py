"""N > 1 scene path on real GPUs (NCCL): `infer_one_img` sharded over world_size ranks must return,
on EVERY rank, exactly what a single GPU returns -- same uint8 masks, same nodes, same edges in the
same order (contiguous tile shards, one all-gather of mask scores, one all-reduce of the
disjointly written topology scores, fusion / aggregation in global tile order).
Skipped on boxes with a single GPU; the CPU suite covers the sharding + exchange logic with gloo."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from sam_road_b200 import SAMRoad, synth
    from sam_road_b200.inferencer import infer_one_img
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = dict(SAM_VERSION="vit_b", PATCH_SIZE=256, USE_SAM_DECODER=False, ENCODER_LORA=False,
                   TOPONET_VERSION="normal", NO_SAM=False, INFER_BATCH_SIZE=16, SAMPLE_MARGIN=0,
                   INFER_PATCHES_PER_EDGE=7, ITSC_THRESHOLD=2.0, ROAD_THRESHOLD=2.0, TOPO_THRESHOLD=0.5,
                   ITSC_NMS_RADIUS=8, ROAD_NMS_RADIUS=16, NEIGHBOR_RADIUS=64, MAX_NEIGHBOR_QUERIES=16)
        net = SAMRoad(cfg)
        net.load_state_dict(synth.make_state_dict(cfg, seed=7, logit_gain=6.0), strict=True)
        net.eval().to(dev)
        img = np.random.RandomState(3).randint(0, 256, size=(400, 400, 3)).astype(np.uint8)
        # thresholds from the masks (identical on every rank: same weights, same scene)
        _, _, kp, road = infer_one_img(net, img, cfg, device=dev, shard=False)
        cfg.update(ITSC_THRESHOLD=float(np.quantile(kp, 0.99)) / 255, ROAD_THRESHOLD=float(np.quantile(road, 0.93)) / 255)
        single = infer_one_img(net, img, cfg, device=dev, shard=False)          # whole scene on this GPU
        sharded = infer_one_img(net, img, cfg, device=dev, shard=True)          # 49 tiles over `world` ranks (ragged)
        ok = all(np.array_equal(a, b) for a, b in zip(single, sharded))
        ok = ok and single[0].shape[0] > 20 and single[1].shape[0] > 5
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        # every rank must hold the same graph: compare checksums across ranks
        chk = torch.tensor([float(sharded[0].sum()), float(sharded[1].sum()), float(sharded[2].astype(np.int64).sum()),
                            float(sharded[3].astype(np.int64).sum())], dtype=torch.float64, device=dev)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        if rank == 0:
            with open(os.path.join(out_dir, "result.txt"), "w") as f:
                f.write(f"{int(flag.item())} {int(same)} {single[0].shape[0]} {single[1].shape[0]}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_scene_equals_single_gpu(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 29600 + (os.getpid() % 300) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    ok, same, n_nodes, n_edges = open(os.path.join(str(tmp_path), "result.txt")).read().split()
    assert ok == "1" and same == "1", (ok, same, n_nodes, n_edges)
