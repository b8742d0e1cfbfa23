"""Global numeric mode of the conv path.

"fp32": fp32 activations/weights, fp32 FMA kernels (PN_IMPL_SIMT) — the 1e-3 parity mode.
"bf16": bf16 activations/weights, fp32 accumulation in TMEM on tcgen05 (PN_IMPL_TCGEN05) — the fast mode.
"bf16x3": fp32 activations, every conv on tcgen05 with split-bf16 operands (x = hi + lo, three bf16 products per fp32
          product, fp32 accumulation): the tensor-core path that meets the 1e-3 parity tolerance (~1e-5 per layer).
          Inference only; the layers run through the gather conv (no padded dense layout, no window plans).
"""
import os
import torch

from ._lib import PN_IMPL_SIMT, PN_IMPL_TCGEN05

_precision = "bf16"


def set_precision(p):
    global _precision
    if p not in ("fp32", "bf16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    _precision = p


def get_precision():
    return _precision


def act_dtype():
    return torch.bfloat16 if _precision == "bf16" else torch.float32


def conv_impl():
    return PN_IMPL_SIMT if _precision == "fp32" else PN_IMPL_TCGEN05


_overlap_rulebooks = os.environ.get("PN_OVERLAP_RULEBOOKS", "1") != "0"


def overlap_rulebooks():
    """build the strided-level rulebooks on a side stream, concurrently with the first stage's convs"""
    return _overlap_rulebooks


_deconv_gemm = os.environ.get("PN_DECONV_GEMM", "1") != "0"


def deconv_gemm():
    """ConvTranspose2d(k=2,s=2) as one GEMM with a scattering epilogue (layers.dense_deconv2x2) instead of a 4-tap
    gather conv; PN_DECONV_GEMM=0 restores the latter (A/B measurements, parity tests)."""
    return _deconv_gemm


def set_deconv_gemm(on):
    global _deconv_gemm
    _deconv_gemm = bool(on)
