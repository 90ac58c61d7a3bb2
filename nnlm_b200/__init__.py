"""nnlm_b200 — B200-native alternating non-negative least squares behind the NNLM API.

Host-side mirror of the reference's R front-end for the one hot path this package accelerates:
    nnmf()  <- R/nnmf.R:135-225  (arguments, defaults, returned W/H/loss objects)
    nnlm()  <- R/nnlm.R:70-145
    predict() <- R/nnmf_methods.R:22-48 (predict.nnmf)
Both call the C ABI of include/nnlm_b200.h (the boundary the R `.Call` shim binds) which drives
hand-written sm_100a CUDA kernels. No CPU fallback: without the built library or a CUDA device, calls raise.
"""
from .api import Nnmf, Nnlm, nnmf, nnlm, predict, nnlm_update, mse_mkl, get_method_code, reformat_input, na_mask, cross
from .session import Session
from . import shard
from . import _capi

__all__ = ["nnmf", "nnlm", "predict", "nnlm_update", "mse_mkl", "get_method_code", "reformat_input", "na_mask", "cross", "Nnmf", "Nnlm",
           "Session", "shard", "_capi"]
