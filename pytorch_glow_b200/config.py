"""Process-wide knobs of the engine (read at call time, never by the kernels)."""
from . import _C

# Arithmetic type of the coupling-network GEMM operands:
#   "bf16": tcgen05/TMEM kernels, bf16 operands, fp32 accumulate (the production path);
#   "fp32": CUDA-core fp32 kernels, used for the strict 1e-4 parity runs;
#   "auto": bf16 whenever the layer shape fits the tensor-core tiling, else fp32.
conv_dtype = "auto"

# FlowModel.encode / decode keep the flow state pixel-major between the NCHW tensors of the API
# (rows_path.py) whenever the model fits those kernels; False forces the per-layer NCHW kernels.
use_rows_path = True

# Activation recompute for training (the exact inverse network/model.py:119-154 is what makes it
# possible in a flow).  False: every FlowStep keeps a1 / h1 / h2 / their ReLU masks / the coupling's (shift, scale) for
# its backward pass (~2.4 KB per pixel at 12 channels: ~100 MB per 64x64 image).  True: a step keeps only its input
# and output rows (96 B per pixel; the output is the next step's input anyway) and its backward pass re-runs the fused
# coupling-net forward on z1 -- which the coupling leaves untouched in the output -- to rebuild the rest, bit for bit:
# ~20x less activation memory for one more fused forward per step (~ +20 % step time).  GLOWK_RECOMPUTE=1 sets it.
import os as _os
recompute_activations = _os.environ.get("GLOWK_RECOMPUTE", "0") == "1"


def resolve_conv_dtype(hidden_channels, override=None):
    mode = override or conv_dtype
    if mode == "fp32":
        return _C.F32
    fits = hidden_channels % 64 == 0
    if mode == "bf16":
        if not fits:
            raise ValueError("conv_dtype='bf16' needs hidden_channels %% 64 == 0 (got %d)" % hidden_channels)
        if not _C.has_tcgen05():
            raise _C.GlowkError("conv_dtype='bf16' needs an sm_100 device with the tcgen05 GEMM built")
        return _C.BF16
    if mode != "auto":
        raise ValueError("conv_dtype must be 'auto', 'fp32' or 'bf16'")
    return _C.BF16 if (fits and _C.has_tcgen05()) else _C.F32
