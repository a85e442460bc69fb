# shim: copy to <reference>/viewformer/models/ — context-view transformer served by viewformer_b200
from viewformer_b200.compat import MIGT_TF as MIGT  # noqa: F401
