"""Build and run the UNMODIFIED reference DETRPose model (whole-model parity tests, `model_e2e` bench leg).

The reference's python sources are taken from ``baseline/_ref`` (vendored copy, see ``baseline/vendor.py``)
or, in the build container, from ``/root/reference``.  Its package ``__init__`` files import libraries this
image does not have (omegaconf, pycocotools, xtcocotools, iopath, calflops), so the
modules are imported through empty stand-in packages (only ``__path__`` set) plus two stub modules; every
class and function then runs exactly as the reference wrote it.  Models are built by direct constructor
calls with the values of ``configs/detrpose/include/detrpose_hgnetv2.py:29-100`` and the per-size
overrides (``detrpose_hgnetv2_{n,s,m,l,x}.py``), because ``LazyConfig`` itself needs the real omegaconf.

This is test / measurement infrastructure: nothing under ``detrpose_b200/`` imports it.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace
from typing import List, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")

__all__ = ["reference_root", "available", "load_reference", "MODEL_CONFIGS", "build_model", "build_criterion",
           "build_postprocessor", "synthetic_targets", "install_kernels", "uninstall_kernels"]


def reference_root() -> Optional[str]:
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "src", "models", "detrpose", "transformer.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


_ref = None


def load_reference() -> SimpleNamespace:
    """Import the reference's model modules (once) and return them in a namespace."""
    global _ref
    if _ref is not None:
        return _ref
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found: run `python baseline/vendor.py` in the build container")
    src = os.path.join(root, "src")

    def fake_pkg(name, path):
        m = sys.modules.get(name)
        if m is None or not hasattr(m, "__path__"):
            m = types.ModuleType(name)
            sys.modules[name] = m
        m.__path__ = [path]
        return m

    fake_pkg("src", src)
    fake_pkg("src.models", src + "/models")
    fake_pkg("src.models.detrpose", src + "/models/detrpose")
    fake_pkg("src.misc", src + "/misc")
    fake_pkg("src.nn", src + "/nn")
    fake_pkg("src.nn.backbone", src + "/nn/backbone")
    fake_pkg("src.data", src + "/data")
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.OmegaConf = type("OmegaConf", (), {})
        oc.DictConfig = dict
        sys.modules["omegaconf"] = oc
    if "src.data.dataloader" not in sys.modules:
        # misc/dist_utils.py:12 only needs the name; the real module pulls pycocotools / xtcocotools
        dl = types.ModuleType("src.data.dataloader")
        dl.DataLoader = torch.utils.data.DataLoader
        sys.modules["src.data.dataloader"] = dl

    import src.models.detrpose.ms_deform_attn as msda          # noqa: E402
    import src.models.detrpose.transformer as transformer      # noqa: E402
    import src.models.detrpose.hybrid_encoder as hybrid_encoder  # noqa: E402
    import src.models.detrpose.detrpose as detrpose            # noqa: E402
    import src.models.detrpose.criterion as criterion          # noqa: E402
    import src.models.detrpose.matcher as matcher              # noqa: E402
    import src.models.detrpose.postprocesses as postprocesses  # noqa: E402
    import src.models.detrpose.dn_component as dn_component    # noqa: E402
    import src.nn.backbone.hgnetv2 as hgnetv2                  # noqa: E402
    _ref = SimpleNamespace(root=root, msda=msda, transformer=transformer, hybrid_encoder=hybrid_encoder,
                           detrpose=detrpose, criterion=criterion, matcher=matcher,
                           postprocesses=postprocesses, dn_component=dn_component, hgnetv2=hgnetv2)
    return _ref


# configs/detrpose/include/detrpose_hgnetv2.py:29-100 + detrpose_hgnetv2_{n,s,m,l,x}.py overrides
_BASE = dict(
    backbone=dict(name="B4", use_lab=False, return_idx=[1, 2, 3], freeze_stem_only=True, freeze_at=-1,
                  freeze_norm=True, pretrained=False),
    encoder=dict(in_channels=[512, 1024, 2048], feat_strides=[8, 16, 32], n_levels=3, hidden_dim=256, nhead=8,
                 dim_feedforward=1024, dropout=0.0, enc_act="gelu", expansion=1.0, depth_mult=1.0, act="silu",
                 temperatureH=20, temperatureW=20, eval_spatial_size=(640, 640)),
    transformer=dict(hidden_dim=256, dropout=0.0, nhead=8, num_queries=60, dim_feedforward=1024,
                     num_decoder_layers=6, normalize_before=False, return_intermediate_dec=True,
                     activation="relu", num_feature_levels=3, dec_n_points=4, learnable_tgt_init=True,
                     two_stage_type="standard", num_body_points=17, aux_loss=True, num_classes=2,
                     dec_pred_class_embed_share=False, dec_pred_pose_embed_share=False,
                     two_stage_class_embed_share=False, two_stage_bbox_embed_share=False, cls_no_bias=False,
                     feat_strides=[8, 16, 32], eval_spatial_size=(640, 640), reg_max=32, reg_scale=4,
                     energy_decrease_weight=0.0),
)
_OVERRIDES = {
    "n": dict(backbone=dict(name="B0", use_lab=True, return_idx=[2, 3]),
              encoder=dict(in_channels=[512, 1024], feat_strides=[16, 32], n_levels=2, use_encoder_idx=[1],
                           depth_mult=0.5, expansion=0.34, hidden_dim=128, dim_feedforward=512),
              transformer=dict(num_decoder_layers=3, num_feature_levels=2, dim_feedforward=512,
                               feat_strides=[16, 32], hidden_dim=128, dec_n_points=6, use_kan=False, kan_grid=3)),
    "s": dict(backbone=dict(name="B0", use_lab=True),
              encoder=dict(in_channels=[256, 512, 1024], depth_mult=0.34, expansion=0.5),
              transformer=dict(num_decoder_layers=3)),
    "m": dict(backbone=dict(name="B2", use_lab=True),
              encoder=dict(in_channels=[384, 768, 1536], depth_mult=0.67),
              transformer=dict(num_decoder_layers=4)),
    "l": dict(),
    "x": dict(backbone=dict(name="B5"), encoder=dict(hidden_dim=384, dim_feedforward=2048),
              transformer=dict(hidden_dim=384, reg_scale=8)),
}


def _merged(size: str) -> dict:
    cfg = {k: dict(v) for k, v in _BASE.items()}
    for part, upd in _OVERRIDES[size].items():
        cfg[part].update(upd)
    return cfg


MODEL_CONFIGS = {s: _merged(s) for s in _OVERRIDES}


def build_model(size: str, seed: int = 0, quiet: bool = True):
    """DETRPose-{n,s,m,l,x} with random-init weights (``pretrained=False``), reference classes throughout."""
    ref = load_reference()
    cfg = MODEL_CONFIGS[size]
    torch.manual_seed(seed)
    out = open(os.devnull, "w") if quiet else sys.stdout
    old = sys.stdout
    sys.stdout = out                      # the reference constructors print parameter counts
    try:
        model = ref.detrpose.DETRPose(
            backbone=ref.hgnetv2.HGNetv2(**cfg["backbone"]),
            encoder=ref.hybrid_encoder.HybridEncoder(**cfg["encoder"]),
            transformer=ref.transformer.Transformer(**cfg["transformer"]))
    finally:
        sys.stdout = old
        if quiet:
            out.close()
    return model


def build_criterion():
    """configs/detrpose/include/detrpose_hgnetv2.py:85-98."""
    ref = load_reference()
    matcher = ref.matcher.HungarianMatcher(cost_class=2.0, cost_keypoints=10.0, cost_oks=4.0, focal_alpha=0.25)
    return ref.criterion.Criterion(num_classes=2,
                                   weight_dict={"loss_vfl": 2.0, "loss_keypoints": 10.0, "loss_oks": 4.0},
                                   focal_alpha=0.25, losses=["vfl", "keypoints"], matcher=matcher,
                                   num_body_points=17)


def build_postprocessor():
    ref = load_reference()
    return ref.postprocesses.PostProcess(num_select=60, num_body_points=17)


def synthetic_targets(batch: int, device, seed: int = 0, max_persons: int = 8) -> List[dict]:
    """Per-image targets in the format of src/data/transforms.py:287-311 / coco.py:127-140:
    k ~ U{1..max_persons} persons, labels = 1, boxes cxcywh in-image, keypoints (k, 51) = 34 normalised xy
    inside the box followed by 17 visibilities ~ Bernoulli(0.7), area = w*h*0.53."""
    g = torch.Generator().manual_seed(1000 + seed)
    targets = []
    for _ in range(batch):
        k = int(torch.randint(1, max_persons + 1, (1,), generator=g))
        wh = 0.1 + 0.4 * torch.rand(k, 2, generator=g)
        cxcy = wh / 2 + (1 - wh) * torch.rand(k, 2, generator=g)
        xy = cxcy[:, None, :] - wh[:, None, :] / 2 + wh[:, None, :] * torch.rand(k, 17, 2, generator=g)
        vis = (torch.rand(k, 17, generator=g) < 0.7).float()
        kpts = torch.cat([(xy * vis[..., None]).reshape(k, 34), vis], dim=1)
        targets.append({
            "labels": torch.ones(k, dtype=torch.int64, device=device),
            "boxes": torch.cat([cxcy, wh], dim=1).to(device),
            "keypoints": kpts.to(device),
            "area": (wh[:, 0] * wh[:, 1] * 0.53 * 640 * 640).to(device),
            "iscrowd": torch.zeros(k, dtype=torch.int64, device=device),
        })
    return targets


def install_kernels(*, core: bool = True, gate: bool = True, lqe: bool = True,
                    value_producer: bool = True, fused_forward: bool = True) -> None:
    """Drop this package's kernels into the loaded reference modules (class / module-global patches, the
    model objects themselves stay untouched): core (rows a1-a3), fused prologue (f1), value hand-over (f2), Gate (f3), LQE (f4)."""
    import detrpose_b200 as dp
    ref = load_reference()
    if core:
        dp.patch.install(ref.msda)
    if fused_forward:
        dp.patch.install_forward(ref.msda)
    if value_producer:
        dp.patch.install_value_producer(ref.transformer)
    if gate:
        dp.gate.install_gate(ref.transformer)
    if lqe:
        dp.lqe.install_lqe(ref.transformer)


def uninstall_kernels() -> None:
    import detrpose_b200 as dp
    ref = load_reference()
    dp.patch.uninstall(ref.msda)
    dp.patch.uninstall_forward(ref.msda)
    dp.patch.uninstall_value_producer(ref.transformer)
    dp.gate.uninstall_gate(ref.transformer)
    dp.lqe.uninstall_lqe(ref.transformer)
