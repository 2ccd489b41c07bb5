"""Generate tests/golden/*.npz from the UNMODIFIED reference (only runs where /root/reference exists).

The reference has no golden vectors for the hot path, so the reference itself is the
source of truth: this script imports its modules as they lie under /root/reference (tools/ref_loader
stubs lightning/torchmetrics/matplotlib), runs them on seeded synthetic weights/inputs and stores
small slices + checksums of the outputs.  It also checks the oracle restatement against the same
runs and refuses to write fixtures if they disagree.  `tests/test_oracle_golden.py` re-checks the
oracle against these files on any machine (no reference needed).

    python tools/make_golden.py
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import samroad_oracle as O  # noqa: E402
from sam_road_b200 import synth  # noqa: E402
from tools import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ORACLE_TOL = 2e-5


def _cfg(patch, version="vit_b", topo="normal", lora=0, samdec=False):
    return dict(SAM_VERSION=version, PATCH_SIZE=patch, USE_SAM_DECODER=samdec, ENCODER_LORA=lora > 0,
                LORA_RANK=lora, TOPONET_VERSION=topo, NO_SAM=False, FOCAL_LOSS=False)


def _stats(t: torch.Tensor):
    d = t.double()
    return np.array([d.sum().item(), d.abs().sum().item(), (d * d).sum().item()])


def model_fixture(name, cfg, seed, n_points):
    torch.manual_seed(0)
    net = ref_loader.load_reference_model(cfg)
    sd = synth.make_state_dict(cfg, seed=seed)
    net.load_state_dict(sd, strict=True)
    P = cfg["PATCH_SIZE"]
    rgb = synth.make_tiles(1, P, seed=seed + 21, dtype=torch.float32)
    pts, prs, val = synth.make_topo_inputs(1, P, n_points, seed=seed + 22)
    with torch.no_grad():
        ref = net(rgb, pts, prs, val)                              # reference SAMRoad.forward
        ref_scores, ref_feat = net.infer_masks_and_img_features(rgb)
        ref_ts = net.infer_toponet(ref_feat, pts, prs, val)
        spec = O.ModelSpec.from_config(cfg)
        ora = O.forward(sd, spec, rgb, pts, prs, val)
        _, ora_feat = O.infer_masks_and_img_features(sd, spec, rgb)
    errs = [(a - b).abs().max().item() for a, b in zip(ref, ora)]
    errs.append((ref_feat - ora_feat).abs().max().item())
    assert max(errs) < ORACLE_TOL, f"{name}: oracle disagrees with the reference: {errs}"
    assert torch.equal(ref[1], ref_scores) and torch.equal(ref[3], ref_ts)
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        seed=seed, n_points=n_points, patch=P,
        mask_logits_sub=ref[0][0, ::8, ::8, :].numpy(), mask_scores_sub=ref[1][0, ::8, ::8, :].numpy(),
        feat_sub=ref_feat[0, ::8, ::2, ::2].numpy(),
        topo_logits=ref[2].numpy(), topo_scores=ref[3].numpy(),
        mask_logits_stats=_stats(ref[0]), feat_stats=_stats(ref_feat), topo_logits_stats=_stats(ref[2]),
        oracle_vs_reference_maxabs=np.array(errs))
    print(f"{name}: oracle-vs-reference max-abs {max(errs):.2e}  "
          f"(mask logits in [{ref[0].min():.3f}, {ref[0].max():.3f}])")


def toponet_fixture():
    out = {}
    for topo in ("normal", "no_offset", "no_transformer", "no_tgt_features"):
        cfg = _cfg(256, topo=topo)
        torch.manual_seed(0)
        net = ref_loader.load_reference_model(cfg)
        sd = synth.make_state_dict(cfg, seed=5)
        net.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(77)
        feat = torch.randn(2, 256, 16, 16, generator=g)
        pts, prs, val = synth.make_topo_inputs(2, 256, 20, seed=9)
        val[1, 3] = False          # an all-invalid row (flipped to valid, model.py:128-130)
        pts[0, 0] = torch.tensor([256, 256])   # border point (inclusive box query)
        with torch.no_grad():
            ref = net.infer_toponet(feat, pts, prs, val)
            ref_f = net.infer_toponet(feat, pts.float() + 0.25, prs, val)
            spec = O.ModelSpec.from_config(cfg)
            ora = O.infer_toponet(sd, spec, feat, pts, prs, val)
            ora_f = O.infer_toponet(sd, spec, feat, pts.float() + 0.25, prs, val)
        err = max((ref - ora).abs().max().item(), (ref_f - ora_f).abs().max().item())
        assert err < ORACLE_TOL, f"toponet {topo}: oracle disagrees with the reference: {err}"
        out[f"{topo}_int"] = ref.numpy()
        out[f"{topo}_float"] = ref_f.numpy()
        print(f"toponet[{topo}]: oracle-vs-reference max-abs {err:.2e}")
    np.savez_compressed(os.path.join(OUT, "toponet_versions.npz"), **out)


def _extract_functions(path, names):
    """exec selected top-level function definitions of a reference file (it cannot be imported as a
    whole here: rtree / igraph / shapely / tcod are absent) -- the source is read where it lies."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "numpy": np}
    import scipy
    import scipy.spatial
    ns["scipy"] = scipy
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def tileloop_fixture():
    ref = ref_loader.REF_ROOT
    ds = _extract_functions(os.path.join(ref, "dataset.py"), {"get_patch_info_one_img"})
    gu = _extract_functions(os.path.join(ref, "graph_utils.py"), {"nms_points"})
    ge = _extract_functions(os.path.join(ref, "graph_extraction.py"),
                            {"get_points_and_scores_from_mask", "extract_graph_points"})
    ge["nms_points"] = gu["nms_points"]
    out = {}
    for tag, args in {"c2_cityscale_16": (0, 2048, 64, 512, 16), "c4_cityscale_8": (0, 2048, 64, 512, 8),
                      "c3_spacenet_16": (0, 400, 0, 256, 16), "spacenet_4": (0, 400, 0, 256, 4)}.items():
        info = ds["get_patch_info_one_img"](*args)
        mine = O.get_patch_info_one_img(*args)
        assert info == mine, tag
        out[f"tiles_{tag}"] = np.array([[x0, y0, x1, y1] for _, (x0, y0), (x1, y1) in info])
    # keypoint extraction on a seeded synthetic pair of masks
    rng = np.random.RandomState(5)
    kp = np.zeros((400, 400), np.uint8)
    road = np.zeros((400, 400), np.uint8)
    for _ in range(60):
        y, x = rng.randint(5, 395, 2)
        kp[y - 2:y + 3, x - 2:x + 3] = rng.randint(60, 255)
    for _ in range(25):
        y = rng.randint(5, 395)
        road[y - 1:y + 2, rng.randint(0, 200):rng.randint(200, 400)] = rng.randint(90, 255)

    class Cfg:
        ITSC_THRESHOLD, ROAD_THRESHOLD, ITSC_NMS_RADIUS, ROAD_NMS_RADIUS = 0.3, 0.4, 8, 16
    pts_ref = ge["extract_graph_points"](kp, road, Cfg)
    pts_mine = O.extract_graph_points(kp, road, 0.3, 0.4, 8, 16)
    assert np.array_equal(pts_ref, pts_mine)
    out.update(kp_mask=kp, road_mask=road, graph_points=pts_ref)
    np.savez_compressed(os.path.join(OUT, "tileloop.npz"), **out)
    print(f"tileloop: {len(pts_ref)} graph points, tile grids identical")


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-new" in sys.argv:      # fixtures added in round 2 (the round-1 files stay byte-identical)
        model_fixture("vitl_256", _cfg(256, version="vit_l"), seed=6, n_points=16)
        model_fixture("vitb_1024", _cfg(1024), seed=7, n_points=40)
        return
    model_fixture("vitb_256", _cfg(256), seed=0, n_points=24)
    model_fixture("vitb_512", _cfg(512), seed=1, n_points=40)
    model_fixture("vitb_256_lora4", _cfg(256, lora=4), seed=2, n_points=16)
    model_fixture("vitb_256_samdec", _cfg(256, samdec=True), seed=3, n_points=16)
    model_fixture("vith_256", _cfg(256, version="vit_h"), seed=4, n_points=16)
    model_fixture("vitl_256", _cfg(256, version="vit_l"), seed=6, n_points=16)
    model_fixture("vitb_1024", _cfg(1024), seed=7, n_points=40)
    toponet_fixture()
    tileloop_fixture()


if __name__ == "__main__":
    main()
