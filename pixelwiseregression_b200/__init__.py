"""pixelwiseregression_b200 — the B200 (sm_100a) hot path of PixelwiseRegression.

SFR target builder (`sfr.build_sfr`), differentiable decoder (`ops.fused_decoder`,
drop-in `model.PixelwiseRegression`) and decoder backward fused with the stage
loss (`ops.fused_decoder_loss`, `PixelwiseRegression.forward_loss`), all backed by
hand-written CUDA kernels in libpwr_b200.so behind the C ABI of include/pwr.h.
Importing the package does not need a GPU; calling into it does, and there is
no CPU fallback.
"""
from . import _lib, synth  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
