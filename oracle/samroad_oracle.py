"""CPU oracle for the sam_road tiled-inference hot path -- TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch restatement (functional, driven directly by the reference's state_dict) of the
algorithm the reference runs on this path.  It exists to check the CUDA implementation; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.  The product path (`sam_road_b200/`) never does.

Parity status: the reference holds NO golden vectors or tests for this path, so the
oracle is pinned against the reference itself: `tools/make_golden.py` imports the unmodified
reference modules from /root/reference (with stub modules for lightning/torchmetrics/matplotlib),
checks this file against them (max-abs ~1e-6, see tests/golden/README.md) and commits small golden
fixtures that `tests/test_oracle_golden.py` re-checks everywhere.

Every function cites the reference lines it restates (paths relative to the reference root; `sam/`
is the vendored segment-anything fork).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

PIXEL_MEAN = (123.675, 116.28, 103.53)   # model.py:229
PIXEL_STD = (58.395, 57.12, 57.375)      # model.py:230

_VIT = {  # model.py:198-218
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_h": dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31)),
}


@dataclass
class ModelSpec:
    """What SAMRoad.__init__ derives from the YAML config (model.py:193-300)."""
    patch_size: int = 512
    sam_version: str = "vit_b"
    use_sam_decoder: bool = False
    toponet_version: str = "normal"
    lora_rank: int = 0
    window_size: int = 14                      # model.py:256
    embed_dim: int = field(init=False)
    depth: int = field(init=False)
    num_heads: int = field(init=False)
    global_attn_indexes: Tuple[int, ...] = field(init=False)

    def __post_init__(self):
        v = _VIT[self.sam_version]
        self.embed_dim, self.depth, self.num_heads = v["embed_dim"], v["depth"], v["num_heads"]
        self.global_attn_indexes = tuple(v["global_attn_indexes"])

    @property
    def grid(self) -> int:
        return self.patch_size // 16

    @classmethod
    def from_config(cls, config) -> "ModelSpec":
        def g(key, default=None):
            v = config.get(key, default) if hasattr(config, "get") else getattr(config, key, default)
            return default if (v is None or (isinstance(v, dict) and not v)) else v
        return cls(patch_size=int(g("PATCH_SIZE")), sam_version=g("SAM_VERSION", "vit_b"),
                   use_sam_decoder=bool(g("USE_SAM_DECODER", False)),
                   toponet_version=g("TOPONET_VERSION", "normal") or "normal",
                   lora_rank=int(g("LORA_RANK", 0)) if g("ENCODER_LORA", False) else 0)


# --------------------------------------------------------------------------------------------------
# image encoder  (sam/segment_anything/modeling/image_encoder.py)
# --------------------------------------------------------------------------------------------------
def layer_norm_2d(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm over channels of NCHW, biased variance, eps inside the sqrt (common.py:31-43)."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def rel_pos_gather(q_len: int, k_len: int, table: Tensor) -> Tensor:
    """R[q, k] = table[q - k + (k_len - 1)] for q_len == k_len and an exactly sized table
    (image_encoder.py:292-322; the interpolation branch :306-313 is never taken on this path because
    tables are created at the exact size, image_encoder.py:158,221-222)."""
    assert q_len == k_len and table.shape[0] == 2 * k_len - 1
    q = torch.arange(q_len).view(-1, 1)
    k = torch.arange(k_len).view(1, -1)
    return table[(q - k + (k_len - 1)).long()]          # [q_len, k_len, hd]


def attention(x: Tensor, sd: StateDict, prefix: str, num_heads: int) -> Tensor:
    """Multi-head attention with decomposed rel-pos on a [B, H, W, D] token grid
    (image_encoder.py:224-240 and 325-361).  The rel-pos term uses the UNscaled q."""
    B, H, W, D = x.shape
    hd = D // num_heads
    w = sd[prefix + "qkv.weight"]
    if prefix + "qkv.linear_a_q.weight" in sd:           # _LoRA_qkv, model.py:179-186
        qkv = F.linear(x, w, sd[prefix + "qkv.bias"])
        new_q = F.linear(F.linear(x, sd[prefix + "qkv.linear_a_q.weight"]),
                         sd[prefix + "qkv.linear_b_q.weight"])
        new_v = F.linear(F.linear(x, sd[prefix + "qkv.linear_a_v.weight"]),
                         sd[prefix + "qkv.linear_b_v.weight"])
        qkv = qkv.clone()
        qkv[..., :D] += new_q
        qkv[..., -D:] += new_v
    else:
        qkv = F.linear(x, w, sd[prefix + "qkv.bias"])
    out = attention_core(qkv, sd[prefix + "rel_pos_h"], sd[prefix + "rel_pos_w"], num_heads)
    return F.linear(out, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def attention_core(qkv: Tensor, rel_pos_h: Tensor, rel_pos_w: Tensor, num_heads: int) -> Tensor:
    """softmax((q*scale) k^T + rel_h + rel_w) v on a [B, H, W, 3D] qkv grid -> [B, H, W, D]
    (image_encoder.py:227-237, 347-361)."""
    B, H, W, D3 = qkv.shape
    D = D3 // 3
    hd = D // num_heads
    qkv = qkv.reshape(B, H * W, 3, num_heads, hd).permute(2, 0, 3, 1, 4)   # [3, B, h, HW, hd]
    q, k, v = qkv.reshape(3, B * num_heads, H * W, hd).unbind(0)
    scores = (q * hd ** -0.5) @ k.transpose(-2, -1)                        # :231
    Rh = rel_pos_gather(H, H, rel_pos_h)
    Rw = rel_pos_gather(W, W, rel_pos_w)
    q_grid = q.reshape(B * num_heads, H, W, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", q_grid, Rh)                      # :354
    rel_w = torch.einsum("bhwc,wkc->bhwk", q_grid, Rw)                      # :355
    scores = (scores.view(-1, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :])
    probs = scores.view(-1, H * W, H * W).softmax(dim=-1)
    return (probs @ v).view(B, num_heads, H, W, hd).permute(0, 2, 3, 1, 4).reshape(B, H, W, D)


def window_split(x: Tensor, win: int) -> Tuple[Tensor, Tuple[int, int]]:
    """Zero-pad bottom/right to a multiple of `win` and cut into windows (image_encoder.py:243-264)."""
    B, H, W, C = x.shape
    ph, pw = (win - H % win) % win, (win - W % win) % win
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // win, win, Wp // win, win, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, win, win, C), (Hp, Wp)


def window_merge(wins: Tensor, win: int, padded: Tuple[int, int], hw: Tuple[int, int]) -> Tensor:
    """Inverse of window_split + crop (image_encoder.py:267-289)."""
    Hp, Wp = padded
    H, W = hw
    B = wins.shape[0] // ((Hp // win) * (Wp // win))
    x = wins.view(B, Hp // win, Wp // win, win, win, -1).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, Hp, Wp, -1)[:, :H, :W, :]


def encoder_block(x: Tensor, sd: StateDict, i: int, spec: ModelSpec) -> Tensor:
    """One ViTDet block (image_encoder.py:166-182); LN eps 1e-6 (model.py:250)."""
    p = f"image_encoder.blocks.{i}."
    D = x.shape[-1]
    win = 0 if i in spec.global_attn_indexes else spec.window_size
    y = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    if win > 0:
        H, W = y.shape[1], y.shape[2]
        y, padded = window_split(y, win)       # padding happens AFTER norm1 (:168-172)
        y = attention(y, sd, p + "attn.", spec.num_heads)
        y = window_merge(y, win, padded, (H, W))
    else:
        y = attention(y, sd, p + "attn.", spec.num_heads)
    x = x + y
    z = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    z = F.linear(z, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])
    z = F.gelu(z)                              # exact erf GELU (common.py:18)
    z = F.linear(z, sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
    return x + z


def image_encoder(x_nchw: Tensor, sd: StateDict, spec: ModelSpec,
                  taps: Optional[dict] = None) -> Tensor:
    """ImageEncoderViT.forward (image_encoder.py:106-116): patch embed (:387-395), + pos_embed,
    blocks, neck (:88-104).  Returns [B, 256, s, s]."""
    x = F.conv2d(x_nchw, sd["image_encoder.patch_embed.proj.weight"],
                 sd["image_encoder.patch_embed.proj.bias"], stride=16).permute(0, 2, 3, 1)
    x = x + sd["image_encoder.pos_embed"]
    if taps is not None:
        taps["tokens0"] = x
    for i in range(spec.depth):
        x = encoder_block(x, sd, i, spec)
        if taps is not None:
            taps[f"block{i}"] = x
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd["image_encoder.neck.0.weight"])
    x = layer_norm_2d(x, sd["image_encoder.neck.1.weight"], sd["image_encoder.neck.1.bias"])
    x = F.conv2d(x, sd["image_encoder.neck.2.weight"], padding=1)
    x = layer_norm_2d(x, sd["image_encoder.neck.3.weight"], sd["image_encoder.neck.3.bias"])
    return x


def map_decoder(feat: Tensor, sd: StateDict) -> Tensor:
    """Naive decoder: 4x ConvTranspose2d(k2,s2) with LN2d+GELU / GELU / GELU (model.py:286-295)."""
    x = F.conv_transpose2d(feat, sd["map_decoder.0.weight"], sd["map_decoder.0.bias"], stride=2)
    x = F.gelu(layer_norm_2d(x, sd["map_decoder.1.weight"], sd["map_decoder.1.bias"]))
    x = F.gelu(F.conv_transpose2d(x, sd["map_decoder.3.weight"], sd["map_decoder.3.bias"], stride=2))
    x = F.gelu(F.conv_transpose2d(x, sd["map_decoder.5.weight"], sd["map_decoder.5.bias"], stride=2))
    return F.conv_transpose2d(x, sd["map_decoder.7.weight"], sd["map_decoder.7.bias"], stride=2)


def normalize_rgb(rgb: Tensor) -> Tensor:
    """[B,H,W,3] 0..255 -> normalised NCHW (model.py:465-467)."""
    x = rgb.permute(0, 3, 1, 2)
    mean = torch.tensor(PIXEL_MEAN, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    std = torch.tensor(PIXEL_STD, dtype=x.dtype, device=x.device).view(-1, 1, 1)
    return (x - mean) / std


def mask_head(feat: Tensor, sd: StateDict, spec: ModelSpec) -> Tensor:
    """mask logits [B, 2, P, P] from image embeddings (model.py:471-491)."""
    if spec.use_sam_decoder:
        from . import sam_decoder_oracle   # SAM TwoWayTransformer decoder path (model.py:260-282)
        return sam_decoder_oracle.sam_mask_logits(feat, sd, spec)
    return map_decoder(feat, sd)


def infer_masks_and_img_features(sd: StateDict, spec: ModelSpec, rgb: Tensor,
                                 return_logits: bool = False):
    """SAMRoad.infer_masks_and_img_features (model.py:459-495)."""
    feat = image_encoder(normalize_rgb(rgb.to(torch.float32)), sd, spec)
    logits = mask_head(feat, sd, spec)
    scores = torch.sigmoid(logits).permute(0, 2, 3, 1)
    if return_logits:
        return scores, feat, logits.permute(0, 2, 3, 1)
    return scores, feat


# --------------------------------------------------------------------------------------------------
# TopoNet  (model.py:29-148)
# --------------------------------------------------------------------------------------------------
def bilinear_sample(feat: Tensor, points: Tensor, patch_size: int) -> Tensor:
    """BilinearSampler.forward (model.py:34-58): grid_sample, bilinear, align_corners=False, zeros."""
    grid = (points / patch_size) * 2.0 - 1.0
    out = F.grid_sample(feat, grid.unsqueeze(2).to(feat.dtype), mode="bilinear", align_corners=False)
    return out.squeeze(-1).permute(0, 2, 1)


def _encoder_layer_masked(x: Tensor, keep: Tensor, sd: StateDict, p: str, heads: int = 4) -> Tensor:
    """Post-norm TransformerEncoderLayer(d=128, ff=128, relu, eps 1e-5) with a key-padding mask
    (model.py:74-85; torch/nn/modules/transformer.py).  x: [R, L, 128], keep: [R, L] bool."""
    R, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.view(R, L, 3, heads, hd).permute(2, 0, 3, 1, 4)          # [R, h, L, hd]
    att = (q / math.sqrt(hd)) @ k.transpose(-2, -1)
    att = att.masked_fill(~keep[:, None, None, :], float("-inf")).softmax(dim=-1)
    y = (att @ v).permute(0, 2, 1, 3).reshape(R, L, D)
    y = F.linear(y, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + y, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    y = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(x + y, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)


def toponet(points: Tensor, point_features: Tensor, pairs: Tensor, pairs_valid: Tensor,
            sd: StateDict, version: str = "normal") -> Tuple[Tensor, Tensor]:
    """TopoNet.forward (model.py:88-148) with the eval-mode semantics of torch's nested-tensor fast
    path: masked keys do not take part in attention and masked slots leave the
    encoder as zeros, so their logit is output_proj.bias.  Rows with no valid pair are flipped to
    all-valid first (model.py:128-130)."""
    pf = F.relu(F.linear(point_features, sd["topo_net.feature_proj.weight"],
                         sd["topo_net.feature_proj.bias"]))
    B, Ns, Np, _ = pairs.shape
    flat = pairs.reshape(B, -1, 2).long()
    bidx = torch.arange(B, device=pairs.device).view(-1, 1).expand(-1, Ns * Np)
    src_f, tgt_f = pf[bidx, flat[:, :, 0]], pf[bidx, flat[:, :, 1]]
    pts = points.to(pf.dtype)
    offset = pts[bidx, flat[:, :, 1]] - pts[bidx, flat[:, :, 0]]
    if version == "no_offset":                 # model.py:113-116 ('no_tgt_features' is overwritten)
        offset = torch.zeros_like(offset)
    x = torch.cat([src_f, tgt_f, offset], dim=2)
    x = F.relu(F.linear(x, sd["topo_net.pair_proj.weight"], sd["topo_net.pair_proj.bias"]))
    x = x.view(B * Ns, Np, -1)
    keep = pairs_valid.reshape(B * Ns, Np).bool()
    keep = keep | (keep.sum(dim=-1, keepdim=True) == 0)
    if version != "no_transformer":
        for l in range(3):
            x = _encoder_layer_masked(x, keep, sd, f"topo_net.transformer_encoder.layers.{l}.")
        x = x * keep.unsqueeze(-1)             # fast path: to_padded_tensor(0.0)
    logits = F.linear(x.view(B, Ns, Np, -1), sd["topo_net.output_proj.weight"],
                      sd["topo_net.output_proj.bias"])
    return logits, torch.sigmoid(logits)


def infer_toponet(sd: StateDict, spec: ModelSpec, image_embeddings: Tensor, graph_points: Tensor,
                  pairs: Tensor, valid: Tensor, return_logits: bool = False):
    """SAMRoad.infer_toponet (model.py:498-508)."""
    pf = bilinear_sample(image_embeddings, graph_points, spec.patch_size)
    logits, scores = toponet(graph_points, pf, pairs, valid, sd, spec.toponet_version)
    return (scores, logits) if return_logits else scores


def forward(sd: StateDict, spec: ModelSpec, rgb: Tensor, graph_points: Tensor, pairs: Tensor,
            valid: Tensor):
    """SAMRoad.forward (model.py:414-457) -> (mask_logits, mask_scores, topo_logits, topo_scores)."""
    scores, feat, logits = infer_masks_and_img_features(sd, spec, rgb, return_logits=True)
    t_scores, t_logits = infer_toponet(sd, spec, feat, graph_points, pairs, valid, True)
    return logits, scores, t_logits, t_scores


# --------------------------------------------------------------------------------------------------
# tile loop  (inferencer.py:61-234, dataset.py:56-67, graph_extraction.py:24-28,130-139,
#             graph_utils.py:572-591)
# --------------------------------------------------------------------------------------------------
def get_patch_info_one_img(image_index: int, image_size: int, sample_margin: int, patch_size: int,
                           patches_per_edge: int):
    """dataset.py:56-67: round(linspace) origins, x outer / y inner."""
    lo, hi = sample_margin, image_size - (patch_size + sample_margin)
    origins = [round(v) for v in np.linspace(start=lo, stop=hi, num=patches_per_edge)]
    return [(image_index, (x, y), (x + patch_size, y + patch_size)) for x in origins for y in origins]


def visiting_order(scores: np.ndarray, tie_order: str = "numpy") -> np.ndarray:
    """`np.argsort(scores)[::-1]` (graph_utils.py:574).  NumPy's default sort is unstable: the order of
    EQUAL scores is an implementation detail (introsort for uint8, AVX-512 / AVX2 sorting networks for
    float64 -- different CPUs give different permutations).  tie_order="numpy" is the reference verbatim
    on this host; tie_order="stable" pins the ties: argsort(kind="stable")[::-1] (descending score,
    equal scores in descending index) -- the order the device-only sort of csrc/graph.cu produces."""
    if tie_order == "numpy":
        return np.argsort(scores)[::-1]
    assert tie_order == "stable", tie_order
    return np.argsort(scores, kind="stable")[::-1]


def nms_points(points: np.ndarray, scores: np.ndarray, radius: float, tie_order: str = "numpy",
               shortcut: bool = True) -> np.ndarray:
    """Greedy radius NMS in descending score order; score > 1 is never suppressed
    (graph_utils.py:572-591).  `shortcut`: when EVERY score is > 1.0 (always the case for mask scores
    above a threshold >= 1/255) the loop cannot suppress anything and returns the sorted points; the
    literal loop (shortcut=False) is kept and compared in tests/test_host_tileloop.py."""
    import scipy.spatial
    order = visiting_order(scores, tie_order)
    pts, sc = points[order, :], scores[order]
    kept = np.ones(order.shape[0], dtype=bool)
    if pts.shape[0] == 0:
        return pts
    if shortcut and bool(np.all(np.greater(sc, 1.0))):
        return pts
    tree = scipy.spatial.KDTree(pts)
    for i, p in enumerate(pts):
        if not kept[i]:
            continue
        nbr = tree.query_ball_point(p, r=radius)
        kept[nbr] = np.greater(sc[nbr], 1.0)
        kept[i] = True
    return pts[kept]


def extract_graph_points(keypoint_mask: np.ndarray, road_mask: np.ndarray, itsc_thr: float,
                         road_thr: float, itsc_radius: float, road_radius: float,
                         tie_order: str = "numpy") -> np.ndarray:
    """graph_extraction.py:24-28,130-139 (thresholds are given in 0..1 and scaled by 255)."""
    def cand(mask, thr):
        rc = np.column_stack(np.where(mask > thr))
        return rc[:, ::-1], mask[mask > thr]
    p0, s0 = cand(keypoint_mask, itsc_thr * 255)
    k0 = nms_points(p0, s0, itsc_radius, tie_order)
    p1, s1 = cand(road_mask, road_thr * 255)
    k1 = nms_points(p1, s1, road_radius, tie_order)
    allp = np.concatenate([k0, k1], axis=0)
    alls = np.concatenate([np.ones(k0.shape[0]), np.zeros(k1.shape[0])], axis=0)
    return nms_points(allp, alls, road_radius, tie_order)


def knn_by_index(pts: np.ndarray, k: int, radius: float) -> np.ndarray:
    """`KDTree(pts).query(pts, k=k+1, distance_upper_bound=radius)[1][:, 1:]` (inferencer.py:159-163)
    with the one thing scipy leaves open pinned: neighbours at EQUAL distance come in ascending index
    (cKDTree returns them in traversal order).  Strictly closer than `radius` (scipy's upper bound is
    exclusive), ascending distance, missing slots = n.  Needs distinct points (true after NMS: the
    nearest neighbour the reference drops is then the point itself)."""
    n = pts.shape[0]
    out = np.full((n, k), n, dtype=np.int64)
    if n == 0:
        return out
    p = pts.astype(np.int64)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.iinfo(np.int64).max)
    lim = float(radius) * float(radius)
    for i in range(n):
        cand = np.nonzero(d2[i] < lim)[0]
        cand = cand[np.lexsort((cand, d2[i, cand]))][:k]
        out[i, :cand.shape[0]] = cand
    return out


def build_pair_queries(graph_points: np.ndarray, tile, max_nbr: int, radius: float, ties: str = "scipy"):
    """Per-tile pair queries (inferencer.py:148-176).  The rtree box query (inclusive bounds,
    inferencer.py:150) is restated as a numpy box test with ascending indices.  ties="scipy" is the
    reference verbatim (cKDTree's order of equidistant neighbours); ties="index" pins it (knn_by_index)."""
    import scipy.spatial
    _, (x0, y0), (x1, y1) = tile
    gx, gy = graph_points[:, 0], graph_points[:, 1]
    idx = np.nonzero((gx >= x0) & (gx <= x1) & (gy >= y0) & (gy <= y1))[0]
    n = idx.shape[0]
    pts = graph_points[idx, :] - np.array([[x0, y0]], dtype=graph_points.dtype)
    if n == 0:
        return idx, pts, np.zeros((0, max_nbr, 2), dtype=np.int64), np.zeros((0, max_nbr), bool)
    if ties == "scipy":
        tree = scipy.spatial.KDTree(pts)
        _, knn = tree.query(pts, k=max_nbr + 1, distance_upper_bound=radius)
        knn = knn[:, 1:]
    else:
        assert ties == "index", ties
        knn = knn_by_index(pts, max_nbr, radius)
    src = np.tile(np.arange(n)[:, None], (1, max_nbr))
    valid = knn < n
    tgt = np.where(valid, knn, src)
    return idx, pts, np.stack([src, tgt], axis=-1), valid


def fuse_masks(tile_scores: Sequence[np.ndarray], tiles, H: int, W: int):
    """inferencer.py:79-110: accumulate in tile order, divide by coverage, *255, truncate to uint8."""
    kp = torch.zeros((H, W), dtype=torch.float32)
    road = torch.zeros((H, W), dtype=torch.float32)
    cnt = torch.zeros((H, W), dtype=torch.float32)
    for sc, (_, (x0, y0), (x1, y1)) in zip(tile_scores, tiles):
        sc = torch.as_tensor(sc)
        kp[y0:y1, x0:x1] += sc[:, :, 0]
        road[y0:y1, x0:x1] += sc[:, :, 1]
        cnt[y0:y1, x0:x1] += 1.0
    kp = torch.nan_to_num(kp / cnt, nan=0.0)
    road = torch.nan_to_num(road / cnt, nan=0.0)
    return (kp * 255).to(torch.uint8).numpy(), (road * 255).to(torch.uint8).numpy()


def infer_one_img(sd: StateDict, spec: ModelSpec, img: np.ndarray, config, masks_override=None,
                  return_edge_scores: bool = False, tie_order: str = "numpy", knn_ties: str = "scipy",
                  topo_fn=None) -> tuple:
    """Whole-scene driver (inferencer.py:61-234) on top of the oracle model.  `masks_override`
    (kp_mask, road_mask) lets a test continue from given fused masks; `return_edge_scores` also
    returns {(src,tgt): mean score} before thresholding.  `tie_order` / `knn_ties` select how the two
    third-party tie orders are resolved (see visiting_order / build_pair_queries); `topo_fn(batch_index,
    feats, pts, prs, val) -> scores` replaces the oracle TopoNet (tests feed the CUDA scores through
    the reference aggregation loop to check it bit for bit)."""
    from collections import defaultdict
    H = img.shape[0]
    bs = int(config["INFER_BATCH_SIZE"])
    tiles = get_patch_info_one_img(0, H, int(config["SAMPLE_MARGIN"]), int(config["PATCH_SIZE"]),
                                   int(config["INFER_PATCHES_PER_EDGE"]))
    scores_all, feats = [], []
    dev0 = next(iter(sd.values())).device
    skip_model = masks_override is not None and topo_fn is not None     # nothing of the oracle model is needed
    for b0 in range(0, len(tiles), bs):
        batch = tiles[b0:b0 + bs]
        if skip_model:
            feats.append(None)
            continue
        rgb = torch.stack([torch.tensor(img[y0:y1, x0:x1, :], dtype=torch.float32)
                           for _, (x0, y0), (x1, y1) in batch], 0)
        with torch.no_grad():
            sc, ft = infer_masks_and_img_features(sd, spec, rgb.to(dev0))
        feats.append(ft)
        scores_all.extend([s.cpu().numpy() for s in sc])
    if masks_override is not None:
        kp_mask, road_mask = masks_override
    else:
        kp_mask, road_mask = fuse_masks(scores_all, tiles, img.shape[0], img.shape[1])
    gp = extract_graph_points(kp_mask, road_mask, float(config["ITSC_THRESHOLD"]),
                              float(config["ROAD_THRESHOLD"]), float(config["ITSC_NMS_RADIUS"]),
                              float(config["ROAD_NMS_RADIUS"]), tie_order)
    if gp.shape[0] == 0:
        return gp, np.zeros((0, 2), dtype=np.int32), kp_mask, road_mask
    edge_sum, edge_cnt = defaultdict(float), defaultdict(float)
    K, R = int(config["MAX_NEIGHBOR_QUERIES"]), float(config["NEIGHBOR_RADIUS"])
    for bi, b0 in enumerate(range(0, len(tiles), bs)):
        batch = tiles[b0:b0 + bs]
        q = [build_pair_queries(gp, t, K, R, knn_ties) for t in batch]
        nmax = max(x[1].shape[0] for x in q)
        if nmax == 0:
            continue
        pad = lambda a: np.pad(a, [(0, nmax - a.shape[0])] + [(0, 0)] * (a.ndim - 1))
        dev = dev0
        pts = torch.tensor(np.stack([pad(x[1]) for x in q])).to(dev)
        prs = torch.tensor(np.stack([pad(x[2]) for x in q])).to(dev)
        val = torch.tensor(np.stack([pad(x[3]) for x in q])).to(dev)
        with torch.no_grad():
            ts = topo_fn(bi, feats[bi], pts, prs, val) if topo_fn is not None else \
                infer_toponet(sd, spec, feats[bi], pts, prs, val)
        ts = torch.where(torch.isnan(ts), -100.0, ts).squeeze(-1).cpu().numpy()
        for ti in range(len(batch)):
            idx = q[ti][0]
            for si in range(q[ti][1].shape[0]):
                for pi in range(K):
                    if not q[ti][3][si, pi]:
                        continue
                    s_p, t_p = q[ti][2][si, pi]
                    key = (int(idx[s_p]), int(idx[t_p]))
                    edge_sum[key] += ts[ti, si, pi]
                    edge_cnt[key] += 1.0
    thr = float(config["TOPO_THRESHOLD"])
    edges = [e for e, s in edge_sum.items() if s / edge_cnt[e] > thr]
    out = (gp[:, ::-1], np.array(edges).reshape(-1, 2), kp_mask, road_mask)
    if return_edge_scores:
        return out + ({e: s / edge_cnt[e] for e, s in edge_sum.items()},)
    return out
