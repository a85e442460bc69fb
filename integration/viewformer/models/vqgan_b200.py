# shim: copy to <reference>/viewformer/models/ — TF-flavour (NHWC) codebook served by viewformer_b200
from viewformer_b200.compat import VQGAN_TF as VQGAN  # noqa: F401
