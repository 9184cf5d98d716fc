"""atoma-infer B200 attention path: hand-written sm_100a kernels behind the reference's C ABI.

`atoma_infer_b200.ops` mirrors the reference's Rust crate `csrc`; `atoma_infer_b200.flash_attention`
mirrors the operator `models::FlashAttention`.  Importing the package does not load CUDA; the shared
library is loaded (and required — there is no fallback) on first use.
"""
from . import _lib  # noqa: F401
from .ops import AtomaError  # noqa: F401

__all__ = ["ops", "flash_attention", "tp", "AtomaError"]
