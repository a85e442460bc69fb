# shim: copy to <reference>/viewformer/models/ — torch-flavour (NCHW) codebook served by viewformer_b200
from viewformer_b200.compat import VQGAN_TH as VQGAN  # noqa: F401
