"""CPU oracle for the USE_SAM_DECODER: True mask head -- TEST INFRASTRUCTURE ONLY (see samroad_oracle.py).

Restates, functionally and driven by the reference state_dict, what sam_road runs when the SAM mask
decoder is enabled (model.py:260-282, 426-443, 471-488): the null-prompt PromptEncoder
(prompt_encoder.py:128-168, dense PE :62-71,171-205), MaskDecoder.predict_masks
(mask_decoder.py:112-149) with the TwoWayTransformer (transformer.py:62-106, 151-182, 185-240), and
the x4 bilinear upsampling of the two low-res masks (model.py:482-487).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def dense_pe(sd: Dict[str, Tensor], h: int, w: int) -> Tensor:
    """PromptEncoder.get_dense_pe (prompt_encoder.py:62-71,185-205) -> [1, 256, h, w]."""
    G = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones((h, w), dtype=torch.float32, device=G.device)
    y = (grid.cumsum(dim=0) - 0.5) / h
    x = (grid.cumsum(dim=1) - 0.5) / w
    c = 2 * torch.stack([x, y], dim=-1) - 1
    c = 2 * np.pi * (c @ G)
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1).unsqueeze(0)


def _attn(sd, p: str, q: Tensor, k: Tensor, v: Tensor, heads: int = 8) -> Tensor:
    """transformer.py:185-240: projections (possibly down-sampled internal dim), scaled dot-product
    attention with the scale applied after QK^T, out_proj.  Batch dims broadcast (token batch 1 vs B)."""
    q = F.linear(q, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(k, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(v, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])

    def split(x):
        b, n, c = x.shape
        return x.reshape(b, n, heads, c // heads).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    att = (q @ k.permute(0, 1, 3, 2)) / math.sqrt(q.shape[-1])
    out = torch.softmax(att, dim=-1) @ v
    b, hds, n, c = out.shape
    out = out.transpose(1, 2).reshape(b, n, hds * c)
    return F.linear(out, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def two_way_transformer(sd, src: Tensor, pos: Tensor, tokens: Tensor):
    """TwoWayTransformer.forward (transformer.py:62-106) with depth 2; layer 0 skips the PE on its
    self-attention and REPLACES the queries (transformer.py:155-161)."""
    t = "mask_decoder.transformer."
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries, query_pe = tokens, tokens
    for i in range(2):
        p = f"{t}layers.{i}."
        if i == 0:
            queries = _attn(sd, p + "self_attn.", queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + _attn(sd, p + "self_attn.", q, q, queries)
        queries = _ln(queries, sd, p + "norm1.")
        q, k = queries + query_pe, keys + key_pe
        queries = _ln(queries + _attn(sd, p + "cross_attn_token_to_image.", q, k, keys), sd, p + "norm2.")
        mlp = F.linear(F.relu(F.linear(queries, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])),
                       sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
        queries = _ln(queries + mlp, sd, p + "norm3.")
        q, k = queries + query_pe, keys + key_pe
        keys = _ln(keys + _attn(sd, p + "cross_attn_image_to_token.", k, q, queries), sd, p + "norm4.")
    q, k = queries + query_pe, keys + key_pe
    queries = _ln(queries + _attn(sd, t + "final_attn_token_to_image.", q, k, keys), sd,
                  t + "norm_final_attn.")
    return queries, keys


def sam_low_res_masks(feat: Tensor, sd) -> Tensor:
    """MaskDecoder.forward(multimask_output=True) on null prompts -> [B, 2, 4s, 4s]
    (mask_decoder.py:71-149; sparse prompts are empty, dense prompt = no_mask_embed broadcast,
    prompt_encoder.py:164-166)."""
    from .samroad_oracle import layer_norm_2d
    B, C, h, w = feat.shape
    tokens = torch.cat([sd["mask_decoder.iou_token.weight"], sd["mask_decoder.mask_tokens.weight"]], 0)
    tokens = tokens.unsqueeze(0)                                         # [1, 4, 256]
    src = feat + sd["prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1)
    hs, keys = two_way_transformer(sd, src, dense_pe(sd, h, w), tokens)
    mask_tokens_out = hs[:, 1:4, :]
    up = keys.transpose(1, 2).reshape(B, C, h, w)
    u = "mask_decoder.output_upscaling."
    up = F.conv_transpose2d(up, sd[u + "0.weight"], sd[u + "0.bias"], stride=2)
    up = F.gelu(layer_norm_2d(up, sd[u + "1.weight"], sd[u + "1.bias"]))
    up = F.gelu(F.conv_transpose2d(up, sd[u + "3.weight"], sd[u + "3.bias"], stride=2))
    hyper = []
    for i in range(3):
        x = mask_tokens_out[:, i, :]
        m = f"mask_decoder.output_hypernetworks_mlps.{i}.layers."
        x = F.relu(F.linear(x, sd[m + "0.weight"], sd[m + "0.bias"]))
        x = F.relu(F.linear(x, sd[m + "1.weight"], sd[m + "1.bias"]))
        hyper.append(F.linear(x, sd[m + "2.weight"], sd[m + "2.bias"]))
    hyper = torch.stack(hyper, dim=1)                                    # [B, 3, 32]
    b, c, hh, ww = up.shape
    masks = (hyper @ up.view(b, c, hh * ww)).view(b, -1, hh, ww)
    return masks[:, 1:, :, :]                                            # multimask_output=True


def sam_mask_logits(feat: Tensor, sd, spec) -> Tensor:
    """mask logits [B, 2, P, P]: low-res masks upsampled x4, bilinear, align_corners=False
    (model.py:482-487)."""
    low = sam_low_res_masks(feat, sd)
    return F.interpolate(low, (spec.patch_size, spec.patch_size), mode="bilinear", align_corners=False)
