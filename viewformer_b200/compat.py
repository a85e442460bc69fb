"""Drop-in classes for the reference's two model registries + the patch that installs them.

The reference resolves models with ``importlib.import_module(f'.{package}', 'viewformer.models')``
(viewformer/models/__init__.py:15-59): a replacement has to be importable as a SUBMODULE of
``viewformer.models`` and be named in the ``_TH_REPOSITORY`` / ``_TF_REPOSITORY`` override tables (:6-8).
Two ways to get there, both tested against the real registry in tests/test_registry_boundary.py:

  * ``viewformer_b200.compat.install()`` — at run time, no file of the reference is touched: registers the
    shim modules ``viewformer.models.{vqgan_b200_th, vqgan_b200, migt_b200}`` in ``sys.modules`` and fills
    the two tables;
  * copy ``integration/viewformer/models/*.py`` (three 1-line shims) next to the reference's models and add
    the three table entries shown in INTEGRATION.md.

Layouts: the torch flavour is NCHW (``VQGAN_TH`` = ``viewformer_b200.VQGAN`` unchanged); the TF flavour's callers
pass NHWC (evaluate/evaluate_transformer.py:106-109,127; models/vqgan.py:284-301), so ``VQGAN_TF`` exposes
``encode`` / ``decode_code`` / ``__call__`` in NHWC with the Keras ``training=`` keyword.
"""
import sys
import types

import torch

from . import _lib as L
from .vqgan import VQGAN
from .migt import MIGT
from .ops import linear

VQGAN_TH = VQGAN


class VQGAN_TF(VQGAN):
    """TF-twin surface (viewformer/models/vqgan.py:284-301): NHWC tensors, ``training`` keyword."""

    def encode(self, input, training=False):
        was = self.training
        self.training = bool(training)
        try:
            return self.encode_nhwc(input)
        finally:
            self.training = was

    def decode_code(self, code, training=False):
        return self.decode_code_nhwc(code)

    def decode(self, quant, training=False):
        self._need_weights()
        q = self._in(quant)
        n, hh, ww, c = q.shape
        z = linear(self.exact, q.reshape(-1, c), self._w["post_quant_conv"], torch.float32).reshape(n, hh, ww, -1)
        return self._decoder(z)

    def call(self, input, training=False):
        quant, diff, idx = self.encode(input, training=training)
        return self.decode(quant), diff, quant, idx

    __call__ = call

    def embed_code(self, embed_id):
        ids = self._in(embed_id, torch.int64)
        n, hh, ww = ids.shape
        return L.gather_rows(self._w["q"]["et"], ids.reshape(-1)).reshape(n, hh, ww, -1)


MIGT_TF = MIGT          # the transformer exists in the TF flavour only; its call surface has no image layout

_SHIMS = {
    "vqgan_b200_th": ("VQGAN", VQGAN_TH),
    "vqgan_b200": ("VQGAN", VQGAN_TF),
    "migt_b200": ("MIGT", MIGT_TF),
}
TH_REPOSITORY = {"vqgan": ("vqgan_b200_th", "VQGAN")}
TF_REPOSITORY = {"vqgan": ("vqgan_b200", "VQGAN"), "migt": ("migt_b200", "MIGT")}


def install(models_package=None):
    """Make ``viewformer.models.AutoModelTH`` / ``AutoModel`` return viewformer_b200 classes.  ``models_package`` defaults to
    the imported ``viewformer.models``.  Idempotent; returns the package."""
    if models_package is None:
        import importlib
        models_package = importlib.import_module("viewformer.models")
    pkg_name = models_package.__name__
    for mod_name, (cls_name, cls) in _SHIMS.items():
        full = f"{pkg_name}.{mod_name}"
        m = sys.modules.get(full)
        if m is None:
            m = types.ModuleType(full)
            m.__package__ = pkg_name
            sys.modules[full] = m
        setattr(m, cls_name, cls)
        setattr(models_package, mod_name, m)
    models_package._TH_REPOSITORY.update(TH_REPOSITORY)
    models_package._TF_REPOSITORY.update(TF_REPOSITORY)
    return models_package
