"""Model registry / loaders — the reference's plug point (viewformer/models/__init__.py:6-82,
viewformer/utils/torch.py:9-17, viewformer/utils/tensorflow.py:20-63).

``AutoModel.from_config`` / ``AutoModelTH.from_config`` return the viewformer_b200 classes for
``cfg.model in {'vqgan','migt'}``; ``load_model(dir_or_ckpt)`` reads ``config.json`` next to the checkpoint
and ingests a torch-pickled ``state_dict`` (Lightning ``.ckpt`` layout: ``ckpt['state_dict']``).
"""
import json
import os

import torch

from .config import load_config, ModelNotFoundError
from .vqgan import VQGAN
from .migt import MIGT

_REPOSITORY = {"vqgan": VQGAN, "migt": MIGT}


class AutoModel:
    @staticmethod
    def from_config(config, **kwargs):
        config = load_config(config)
        if config.model not in _REPOSITORY:
            raise ModelNotFoundError(f"Model {config.model} is not supported")
        return _REPOSITORY[config.model](config, **kwargs)


AutoModelTH = AutoModel   # the torch/TF split of the reference collapses: one CUDA implementation serves both


def load_model(checkpoint, restore_weights=True, precision="bf16", **config_overrides):
    """checkpoint: directory holding config.json (+ a .ckpt/.pt file) or the checkpoint file itself."""
    ckpt_file, tf_prefix = None, None
    if os.path.isdir(checkpoint):
        model_dir = checkpoint
        for f in sorted(os.listdir(checkpoint)):
            if f.endswith((".ckpt", ".pt", ".pth")):
                ckpt_file = os.path.join(checkpoint, f)
            elif f.endswith(".index"):                       # TF2 checkpoint: <name>.index + <name>.data-*  (utils/tensorflow.py:26-31)
                tf_prefix = os.path.join(checkpoint, f[: -len(".index")])
    elif os.path.exists(checkpoint + ".index"):              # '<dir>/model' as the reference passes it to load_weights
        model_dir, tf_prefix = os.path.dirname(checkpoint), checkpoint
    else:
        model_dir, ckpt_file = os.path.dirname(checkpoint), checkpoint
    with open(os.path.join(model_dir, "config.json")) as f:
        cfg = json.load(f)
    cfg.update(config_overrides)
    model = AutoModel.from_config(cfg, precision=precision)
    if restore_weights and tf_prefix is not None and ckpt_file is None:
        model.load_weights(tf_prefix)
        return model
    if restore_weights:
        if ckpt_file is None:
            raise FileNotFoundError(f"no checkpoint file next to {model_dir}/config.json")
        data = torch.load(ckpt_file, map_location="cpu")
        model.load_state_dict(data.get("state_dict", data))
    return model
