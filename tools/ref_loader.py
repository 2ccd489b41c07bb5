"""Import the UNMODIFIED reference (htcr/sam_road) from /root/reference for oracle validation.

Only usable where /root/reference exists (the build container); nothing under tests -m gpu, smoke()
or bench.py depends on it.  The reference's model.py needs `lightning`, `torchmetrics` and
`matplotlib`, which are absent here, so three stub modules are injected into sys.modules
 -- the reference's own source files are imported as they lie, never copied.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("SAMROAD_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


def _install_stubs():
    import torch.nn as nn
    if "lightning" not in sys.modules:
        lightning = types.ModuleType("lightning")
        pl = types.ModuleType("lightning.pytorch")

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        lightning.pytorch = pl
        sys.modules["lightning"] = lightning
        sys.modules["lightning.pytorch"] = pl
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tmc = types.ModuleType("torchmetrics.classification")

        class _Metric(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        tmc.BinaryJaccardIndex = _Metric
        tmc.F1Score = _Metric
        tmc.BinaryPrecisionRecallCurve = _Metric
        tm.classification = tmc
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.classification"] = tmc
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    os.environ.setdefault("WANDB_MODE", "disabled")


class AttrDict(dict):
    """addict.Dict-like config: missing keys read as an empty (falsy) AttrDict (utils.py:6-9)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return AttrDict()

    def __setattr__(self, k, v):
        self[k] = v


def load_reference_model(config: dict):
    """Instantiate the reference SAMRoad (eval mode, CPU) with a synthetic SAM_CKPT_PATH that only
    holds a correctly sized pos_embed (model.py:367-396 loads it unconditionally)."""
    import contextlib
    import io

    import torch
    assert available(), f"reference not found under {REF_ROOT}"
    _install_stubs()
    for p in (REF_ROOT, os.path.join(REF_ROOT, "sam")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import model as ref_model  # noqa: the reference's model.py

    cfg = AttrDict(config)
    from oracle.samroad_oracle import ModelSpec
    spec = ModelSpec.from_config(cfg)
    with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
        torch.save({"image_encoder.pos_embed": torch.zeros(1, spec.grid, spec.grid, spec.embed_dim)}, f)
        cfg["SAM_CKPT_PATH"] = f.name
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            net = ref_model.SAMRoad(cfg)
    finally:
        os.unlink(cfg["SAM_CKPT_PATH"])
    net.eval()
    return net
