"""`build_sam2` / `build_sam2_video_predictor` with the reference's signatures (build_sam.py:15-89).

The reference composes its model tree with Hydra; this package carries a ~40-line instantiator for
the same YAML schema (`_target_` + kwargs, `++model.a.b=value` overrides), so no Hydra/OmegaConf is
needed.  Config names resolve to `medsam2_b200/configs/<name>.yaml`; a path to the reference's own
YAML also works (`sam2_train.*` targets are mapped onto this package).  Checkpoints use the
reference layout: `torch.load(path)["model"]`, loaded strictly.
"""
import importlib
import logging
import os
import re

import torch
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
_FLOAT_RE = re.compile(r"^[-+]?\d+(\.\d*)?[eE][-+]?\d+$")
_IMAGE_OVERRIDES = [
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_via_stability=true",
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_delta=0.05",
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_thresh=0.98",
]
_VIDEO_OVERRIDES = _IMAGE_OVERRIDES + [
    "++model.binarize_mask_from_pts_for_mem_enc=true",
    "++model.fill_hole_area=8",
]


def _resolve_config(config_file):
    if os.path.isfile(config_file):
        return config_file
    name = os.path.basename(config_file)
    if not name.endswith(".yaml"):
        name += ".yaml"
    path = os.path.join(_HERE, "configs", name)
    if not os.path.isfile(path):
        raise FileNotFoundError(f"config {config_file!r} not found (looked for {path})")
    return path


def _apply_override(tree, ov):
    m = re.match(r"^\+{0,2}([\w.]+)=(.*)$", ov)
    if not m:
        raise ValueError(f"cannot parse override {ov!r}")
    keys, val = m.group(1).split("."), yaml.safe_load(m.group(2))
    node = tree
    for k in keys[:-1]:
        if node.get(k) is None:
            node[k] = {}
        node = node[k]
    node[keys[-1]] = val


def _target(name):
    if name.startswith("sam2_train."):
        name = "medsam2_b200." + name[len("sam2_train."):]
    mod, cls = name.rsplit(".", 1)
    return getattr(importlib.import_module(mod), cls)


def instantiate(node):
    if isinstance(node, dict):
        kwargs = {k: instantiate(v) for k, v in node.items() if k != "_target_"}
        return _target(node["_target_"])(**kwargs) if "_target_" in node else kwargs
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if isinstance(node, str) and _FLOAT_RE.match(node):      # PyYAML reads `1e-6` as a string
        return float(node)
    return node


def _build(config_file, ckpt_path, device, mode, overrides):
    tree = yaml.safe_load(open(_resolve_config(config_file)))
    for ov in overrides:
        _apply_override(tree, ov)
    model = instantiate(tree["model"])
    _load_checkpoint(model, ckpt_path)
    model = model.to(device)
    if mode == "eval":
        model.eval()
    return model


def build_sam2(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=[],
               apply_postprocessing=True):
    overrides = list(hydra_overrides_extra)
    if apply_postprocessing:
        overrides += _IMAGE_OVERRIDES
    return _build(config_file, ckpt_path, device, mode, overrides)


def build_sam2_video_predictor(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=[],
                               apply_postprocessing=True):
    overrides = ["++model._target_=medsam2_b200.sam2_video_predictor.SAM2VideoPredictor"]
    overrides += list(hydra_overrides_extra)
    if apply_postprocessing:
        overrides += _VIDEO_OVERRIDES
    return _build(config_file, ckpt_path, device, mode, overrides)


def _load_checkpoint(model, ckpt_path):
    if ckpt_path is None:
        return
    sd = torch.load(ckpt_path, map_location="cpu")["model"]
    missing, unexpected = model.load_state_dict(sd)
    if missing:
        logging.error(missing)
        raise RuntimeError()
    if unexpected:
        logging.error(unexpected)
        raise RuntimeError()
    logging.info("Loaded checkpoint sucessfully")
