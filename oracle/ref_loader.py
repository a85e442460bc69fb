"""TEST INFRASTRUCTURE — import the *real* reference torch VQGAN from /root/reference.

Container only (the GPU box has no /root/reference).  Three absent third-party modules are stubbed
(aparse.Literal, pytorch_lightning.LightningModule, lpips.LPIPS -> zeros); the reference sources are
loaded in place by file path — nothing is copied.  Recipe from SURVEY.md §8(c).
"""
import importlib.util
import os
import sys
import types
import typing

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VIEWFORMER_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "viewformer", "models", "vqgan_th.py"))


class _ZeroLPIPS(nn.Module):
    def __init__(self, **kw):
        super().__init__()

    def forward(self, a, b):
        return torch.zeros(a.shape[0], 1, 1, 1)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_cache = {}


def load_reference_modules():
    """Returns (config_module, utils_th_module, vqgan_th_module) of the real reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)
    if "aparse" not in sys.modules:
        ap = types.ModuleType("aparse")
        ap.Literal = typing.Literal
        sys.modules["aparse"] = ap
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = pl
    if "lpips" not in sys.modules:
        lp = types.ModuleType("lpips")
        lp.LPIPS = _ZeroLPIPS
        sys.modules["lpips"] = lp
    for pkg in ("viewformer", "viewformer.utils", "viewformer.models"):
        if pkg not in sys.modules:
            p = types.ModuleType(pkg)
            p.__path__ = []
            sys.modules[pkg] = p
    R = os.path.join(REFERENCE_ROOT, "viewformer")
    _load("viewformer.utils.schedules", R + "/utils/schedules.py")
    cfg = _load("viewformer.models.config", R + "/models/config.py")
    uth = _load("viewformer.models.utils_th", R + "/models/utils_th.py")
    vq = _load("viewformer.models.vqgan_th", R + "/models/vqgan_th.py")
    _cache["mods"] = (cfg, uth, vq)
    return _cache["mods"]


def build_reference_vqgan(state_dict=None, **cfg_overrides):
    """Instantiate the real reference VQGAN (eval mode — see SURVEY.md §8 a6 quirk)."""
    cfg_mod, _, vq = load_reference_modules()
    cfg = cfg_mod.VQGANConfig(**cfg_overrides)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = vq.VQGAN(cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model.eval()


def load_reference_registry(extra_paths=()):
    """Executes the REAL ``viewformer/models/__init__.py`` (AutoModel / AutoModelTH / load_config and the two override tables)
    as the package ``viewformer.models``.  ``extra_paths`` are appended to the package ``__path__`` (how a maintainer's copied
    shim modules become importable as ``viewformer.models.<shim>``)."""
    load_reference_modules()
    R = os.path.join(REFERENCE_ROOT, "viewformer", "models")
    spec = importlib.util.spec_from_file_location("viewformer.models", os.path.join(R, "__init__.py"),
                                                  submodule_search_locations=[R] + list(extra_paths))
    m = importlib.util.module_from_spec(spec)
    sys.modules["viewformer.models"] = m
    spec.loader.exec_module(m)
    return m
