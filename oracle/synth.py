"""Deterministic synthetic weights and inputs for the sam_road hot path -- TEST INFRASTRUCTURE.

No SAM / sam_road checkpoint exists offline (SURVEY.md §7 "Hard parts"), so parity and throughput
are measured on seeded random weights with the reference's exact state_dict key set and shapes
(SURVEY.md §8b).  pos_embed and rel_pos tables are randomised (the reference zero-initialises them,
image_encoder.py:68-70,221-222, which would hide rel-pos bugs).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .samroad_oracle import ModelSpec


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, gen, name, out_f, in_f, bias=True):
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(gen, (out_f, in_f), bound)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), bound)


def _norm(sd, gen, name, n):
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=gen)
    sd[name + ".bias"] = 0.05 * torch.randn(n, generator=gen)


def make_state_dict(spec: ModelSpec, seed: int = 0, logit_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Reference-layout state_dict (fp32, CPU).  `logit_gain` scales the last decoder layer and
    TopoNet's output_proj so logits span a wider range than random init gives (+-0.3)."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D, s, hd = spec.embed_dim, spec.grid, spec.embed_dim // spec.num_heads
    e = "image_encoder."
    bound = 1.0 / math.sqrt(3 * 256)
    sd[e + "pos_embed"] = 0.02 * torch.randn(1, s, s, D, generator=gen)
    sd[e + "patch_embed.proj.weight"] = _uniform(gen, (D, 3, 16, 16), bound)
    sd[e + "patch_embed.proj.bias"] = _uniform(gen, (D,), bound)
    for i in range(spec.depth):
        p = f"{e}blocks.{i}."
        _norm(sd, gen, p + "norm1", D)
        rows = 2 * s - 1 if i in spec.global_attn_indexes else 2 * spec.window_size - 1
        sd[p + "attn.rel_pos_h"] = 0.02 * torch.randn(rows, hd, generator=gen)
        sd[p + "attn.rel_pos_w"] = 0.02 * torch.randn(rows, hd, generator=gen)
        _linear(sd, gen, p + "attn.qkv", 3 * D, D)
        if spec.lora_rank > 0:
            r = spec.lora_rank
            for nm in ("q", "v"):
                sd[p + f"attn.qkv.linear_a_{nm}.weight"] = _uniform(gen, (r, D), 1.0 / math.sqrt(D))
                sd[p + f"attn.qkv.linear_b_{nm}.weight"] = _uniform(gen, (D, r), 0.3 / math.sqrt(r))
        _linear(sd, gen, p + "attn.proj", D, D)
        _norm(sd, gen, p + "norm2", D)
        _linear(sd, gen, p + "mlp.lin1", 4 * D, D)
        _linear(sd, gen, p + "mlp.lin2", D, 4 * D)
    sd[e + "neck.0.weight"] = _uniform(gen, (256, D, 1, 1), 1.0 / math.sqrt(D))
    _norm(sd, gen, e + "neck.1", 256)
    sd[e + "neck.2.weight"] = _uniform(gen, (256, 256, 3, 3), 1.0 / math.sqrt(256 * 9))
    _norm(sd, gen, e + "neck.3", 256)

    if spec.use_sam_decoder:
        from .sam_decoder_oracle import add_sam_decoder_weights
        add_sam_decoder_weights(sd, gen, spec, logit_gain)
    else:
        for idx, (cin, cout) in zip((0, 3, 5, 7), ((256, 128), (128, 64), (64, 32), (32, 2))):
            b = 1.0 / math.sqrt(cout * 4)       # torch's fan_in for ConvTranspose2d weights
            g = logit_gain if idx == 7 else 1.0
            sd[f"map_decoder.{idx}.weight"] = _uniform(gen, (cin, cout, 2, 2), b) * g
            sd[f"map_decoder.{idx}.bias"] = _uniform(gen, (cout,), b) * g
        _norm(sd, gen, "map_decoder.1", 128)

    t = "topo_net."
    _linear(sd, gen, t + "feature_proj", 128, 256)
    _linear(sd, gen, t + "pair_proj", 128, 258)
    if spec.toponet_version != "no_transformer":
        for l in range(3):
            p = f"{t}transformer_encoder.layers.{l}."
            sd[p + "self_attn.in_proj_weight"] = _uniform(gen, (384, 128), math.sqrt(6.0 / (384 + 128)))
            sd[p + "self_attn.in_proj_bias"] = 0.02 * torch.randn(384, generator=gen)
            _linear(sd, gen, p + "self_attn.out_proj", 128, 128)
            _linear(sd, gen, p + "linear1", 128, 128)
            _linear(sd, gen, p + "linear2", 128, 128)
            _norm(sd, gen, p + "norm1", 128)
            _norm(sd, gen, p + "norm2", 128)
    _linear(sd, gen, t + "output_proj", 1, 128)
    sd[t + "output_proj.weight"] *= logit_gain
    sd[t + "output_proj.bias"] *= logit_gain
    return sd


def make_tiles(batch: int, patch_size: int, seed: int = 0, dtype=torch.uint8) -> torch.Tensor:
    """Uniform random RGB tiles [B,P,P,3] (SURVEY.md §8d synthetic inputs)."""
    gen = torch.Generator().manual_seed(1000 + seed)
    t = torch.randint(0, 256, (batch, patch_size, patch_size, 3), generator=gen, dtype=torch.uint8)
    return t if dtype == torch.uint8 else t.to(dtype)


def make_topo_inputs(batch: int, patch_size: int, n_points: int, seed: int = 0,
                     max_nbr: int = 16, radius: float = 64.0, ragged: bool = True):
    """Keypoints on a jittered lattice per tile, kNN pair queries built exactly like
    inferencer.py:156-176 (prefix-valid masks), padded to the batch maximum (inferencer.py:179-185).
    Returns int64 points [B,N,2] (x,y), int64 pairs [B,N,K,2], bool valid [B,N,K]."""
    import scipy.spatial
    rng = np.random.RandomState(2000 + seed)
    side = int(math.ceil(math.sqrt(n_points)))
    step = patch_size / side
    pts_l, pairs_l, valid_l = [], [], []
    for b in range(batch):
        n = n_points if not ragged else max(1, n_points - (b * 7) % max(1, n_points // 3))
        gy, gx = np.divmod(rng.permutation(side * side)[:n], side)
        jit = rng.uniform(-0.35, 0.35, size=(n, 2)) * step
        xy = np.stack([(gx + 0.5) * step, (gy + 0.5) * step], 1) + jit
        xy = np.clip(np.round(xy), 0, patch_size).astype(np.int64)
        tree = scipy.spatial.KDTree(xy)
        k = min(max_nbr + 1, max(2, n))
        _, knn = tree.query(xy, k=k, distance_upper_bound=radius)
        knn = knn.reshape(n, -1)[:, 1:]
        if knn.shape[1] < max_nbr:
            knn = np.pad(knn, [(0, 0), (0, max_nbr - knn.shape[1])], constant_values=n)
        src = np.tile(np.arange(n)[:, None], (1, max_nbr))
        valid = knn < n
        tgt = np.where(valid, knn, src)
        pts_l.append(xy)
        pairs_l.append(np.stack([src, tgt], -1))
        valid_l.append(valid)
    nmax = max(p.shape[0] for p in pts_l)
    pad = lambda a: np.pad(a, [(0, nmax - a.shape[0])] + [(0, 0)] * (a.ndim - 1))
    return (torch.tensor(np.stack([pad(p) for p in pts_l])),
            torch.tensor(np.stack([pad(p) for p in pairs_l])),
            torch.tensor(np.stack([pad(v) for v in valid_l])))
