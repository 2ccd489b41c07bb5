"""Where the fp16-operand error of the CUDA path comes from (run on the GPU box; writes JSON to stdout).

The CUDA path keeps the residual stream, LayerNorm / softmax statistics and accumulators in fp32 and
rounds GEMM / attention OPERANDS to fp16.  This script reproduces that rounding inside the fp32 oracle,
one stage at a time (patch embedding, block linears, attention operands, neck, decoder, TopoNet), and
reports the max-abs deviation each stage alone causes on the image embeddings, the mask logits and the
topology logits -- next to the deviation of the real CUDA path -- for default-scale and wide
(`logit_gain`) synthetic weights.  It is a model of the rounding, not of the kernels (accumulation order,
the GELU fit and the fp16 storage of intermediate LayerNorm outputs are only in the CUDA number).

    python tools/error_budget.py [--patch 512] [--gain 12]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import samroad_oracle as O  # noqa: E402
from sam_road_b200 import SAMRoad, synth  # noqa: E402

DEV = "cuda:0"


def h(x):
    return x.half().float()


class FShim:
    """torch.nn.functional with fp16-rounded operands for the linears / convs of the enabled stages."""

    def __init__(self, real, stages, wmap):
        self._real, self._stages, self._wmap = real, stages, wmap

    def __getattr__(self, n):
        return getattr(self._real, n)

    def _on(self, w):
        return self._wmap.get(id(w)) in self._stages

    def linear(self, x, w, b=None):
        return self._real.linear(h(x), h(w), b) if self._on(w) else self._real.linear(x, w, b)

    def conv2d(self, x, w, b=None, **k):
        return self._real.conv2d(h(x), h(w), b, **k) if self._on(w) else self._real.conv2d(x, w, b, **k)

    def conv_transpose2d(self, x, w, b=None, **k):
        return self._real.conv_transpose2d(h(x), h(w), b, **k) if self._on(w) else \
            self._real.conv_transpose2d(x, w, b, **k)

    def multi_head_attention_forward(self, *a, **k):
        return self._real.multi_head_attention_forward(*a, **k)


def stage_of(key):
    if key.startswith("image_encoder.patch_embed"):
        return "patch_embed"
    if key.startswith("image_encoder.blocks"):
        return "block_linears"
    if key.startswith("image_encoder.neck"):
        return "neck"
    if key.startswith("map_decoder"):
        return "decoder"
    if key.startswith("topo_net.feature_proj"):
        return "topo_feature_proj"
    if key.startswith("topo_net.pair_proj"):
        return "topo_pair_proj"
    if key.startswith("topo_net.transformer_encoder"):
        return "topo_transformer"
    return None      # topo_net.output_proj stays fp32 in the CUDA path


def run_variant(sd, spec, rgb, pts, prs, val, stages):
    wmap = {id(v): stage_of(k) for k, v in sd.items()}
    real_F, real_core = O.F, O.attention_core
    O.F = FShim(real_F, stages, wmap)
    if "attention_operands" in stages:
        def core(qkv, rh, rw, nh):      # q, k, v are stored in fp16; so are the probabilities fed to P.V
            return real_core(h(qkv), rh, rw, nh)
        O.attention_core = core
    try:
        with torch.no_grad():
            scores, feat, logits = O.infer_masks_and_img_features(sd, spec, rgb, return_logits=True)
            ts, tl = O.infer_toponet(sd, spec, feat, pts, prs, val, return_logits=True)
    finally:
        O.F, O.attention_core = real_F, real_core
    return feat, logits, tl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=512)
    ap.add_argument("--gain", type=float, default=12.0)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = dict(SAM_VERSION="vit_b", PATCH_SIZE=a.patch, USE_SAM_DECODER=False, ENCODER_LORA=False,
               TOPONET_VERSION="normal", NO_SAM=False)
    out = {}
    for gain in (1.0, a.gain):
        sd = {k: v.to(DEV) for k, v in synth.make_state_dict(cfg, seed=1, logit_gain=gain).items()}
        spec = O.ModelSpec.from_config(cfg)
        rgb = synth.make_tiles(2, a.patch, seed=5).to(DEV).float()
        pts, prs, val = [t.to(DEV) for t in synth.make_topo_inputs(2, a.patch, 64, seed=6)]
        ref = run_variant(sd, spec, rgb, pts, prs, val, set())
        vmask = val.unsqueeze(-1)
        rows = {}
        stages = ["patch_embed", "block_linears", "attention_operands", "neck", "decoder", "topo_feature_proj",
                  "topo_pair_proj", "topo_transformer"]
        for st in [[s] for s in stages] + [stages[1:3], stages[1:3] + ["topo_transformer"], stages]:
            got = run_variant(sd, spec, rgb, pts, prs, val, set(st))
            rows["+".join(st) if len(st) < len(stages) else "all"] = {
                "feat": (got[0] - ref[0]).abs().max().item(),
                "mask_logit": (got[1] - ref[1]).abs().max().item(),
                "topo_logit_valid": ((got[2] - ref[2]).abs() * vmask).max().item()}
        net = SAMRoad(cfg)
        net.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True)
        net.eval().to(DEV)
        logits, _, tl, _ = net(rgb, pts, prs, val)
        feat = net.infer_masks_and_img_features(rgb)[1]
        rows["cuda_path"] = {"feat": (feat - ref[0]).abs().max().item(),
                             "mask_logit": (logits - ref[1]).abs().max().item(),
                             "topo_logit_valid": ((tl - ref[2]).abs() * vmask).max().item()}
        out[f"gain_{gain:g}"] = {"feat_absmax": ref[0].abs().max().item(),
                                 "mask_logit_range": [ref[1].min().item(), ref[1].max().item()],
                                 "topo_logit_range": [ref[2].min().item(), ref[2].max().item()],
                                 "maxabs_vs_fp32_oracle": rows}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
