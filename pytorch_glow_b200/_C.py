"""ctypes binding of libglowk.so (the C ABI declared in include/glowk.h).

There is no CPU fallback: every op that reaches `lib()` without the compiled
extension, or `check_cuda()` with a non-CUDA tensor, raises.  The library is
built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLOWK_LIB") or os.path.join(_HERE, "libglowk.so")     # GLOWK_LIB: A/B builds (tools/)

F32, BF16 = 0, 1
EPI_STORE, EPI_ACTNORM_RELU, EPI_ACTNORM, EPI_ZEROS, EPI_RELU_BWD = 0, 1, 2, 3, 4

_c = ctypes
_p, _i64, _i32, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES).  Mirrors include/glowk.h 1:1;
# tests/test_abi.py checks that every symbol declared in the header is exported and listed here.
SIGNATURES = {
    "glowk_last_error": [],
    "glowk_version": [],
    "glowk_has_tcgen05": [],
    "glowk_debug_gemm_trace": [_p],
    "glowk_actnorm": [_p, _p, _p, _p, _f32, _i64, _i64, _i64, _i32, _p],
    "glowk_actnorm_init": [_p, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _p, _p, _p],
    "glowk_invconv_prepare": [_p, _i64, _p, _p, _p],
    "glowk_invconv_prepare_batched": [_p, _i64, _i64, _p, _p, _p],
    "glowk_invconv_lu_assemble": [_p, _p, _p, _p, _p, _i64, _p, _p, _p, _p],
    "glowk_invconv_lu_grads": [_p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p],
    "glowk_actnorm_init_ex": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _i32, _i32, _p, _p, _p],
    "glowk_actnorm_mix": [_p, _p, _p, _p, _p, _p, _f32, _i64, _i64, _i64, _i32, _p],
    "glowk_squeeze2d": [_p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "glowk_im2col": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i64, _p],
    "glowk_im2col_rows": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i64, _p],
    "glowk_im2col_rows_ones": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i64, _i64, _p],
    "glowk_rows_to_nchw": [_p, _i32, _i64, _p, _i64, _i64, _i64, _p],
    "glowk_tapsum_to_nchw": [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "glowk_pack_conv_weight": [_p, _i64, _i64, _i32, _i32, _p, _i32, _i64, _i64, _p],
    "glowk_gemm": [_p, _i64, _p, _i64, _i32, _i64, _i64, _i64, _i32, _p, _p, _f32, _p, _i64, _p, _p,
                   _p, _i32, _i64, _p],
    "glowk_gemm_ex": [_p, _i64, _p, _i64, _i32, _i64, _i64, _i64, _i32, _p, _p, _f32, _p, _i64, _p, _p,
                      _p, _i32, _i64, _i32, _i32, _p],
    "glowk_gemm_wgrad": [_p, _i64, _p, _i64, _i32, _i64, _i64, _i64, _p, _i64, _p],
    "glowk_debug_cnet_trace": [_p],
    "glowk_debug_cnet_timeline": [_p],
    "glowk_cnet_fused_supported": [_i32, _i64, _i64, _i64],
    "glowk_cnet_forward": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _f32, _p, _p, _f32,
                           _p, _i64, _p, _p, _i64, _p],
    "glowk_cnet_forward_masked": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _f32, _p, _p, _f32, _p, _i64, _p, _p, _i64, _p, _p, _p],
    "glowk_cnet_forward_implicit": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64,
                                    _i64, _i64, _i64, _p, _p, _f32, _p, _p, _f32, _p, _i64, _p, _p, _i64, _p],
    "glowk_cnet_relu_mask_bytes": [_i64],
    "glowk_cnet_forward_implicit_masked": [_p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _p, _p, _f32, _p, _p, _f32, _p, _i64, _p, _p, _i64, _p, _p, _p],
    "glowk_cnet_backward_implicit_masked": [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _p, _f32, _p, _f32, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _p, _p, _p],
    "glowk_cnet_backward": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _f32, _p, _f32, _p, _p,
                            _p, _p, _i64, _p, _i64, _p, _p, _p],
    "glowk_cnet_backward_implicit": [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64,
                                     _i64, _p, _f32, _p, _f32, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _p],
    "glowk_coupling_nblk": [_i64],
    "glowk_coupling": [_p, _i64, _p, _p, _f32, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "glowk_logdet_finish": [_p, _p, _p, _i64, _f32, _p, _p, _i64, _i64, _f32, _i64, _p],
    "glowk_gaussian_logp": [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p],
    "glowk_split2d_sample": [_p, _i64, _p, _p, _p, _i64, _i64, _i64, _p],
    "glowk_coupling_bwd": [_p, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _p],
    "glowk_split2d_bwd": [_p, _p, _i64, _p, _i64, _p, _p, _f32, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _p],
    "glowk_actnorm_mix_bwd": [_p, _p, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "glowk_logdet_param_grad": [_p, _i64, _i64, _f32, _p, _i64, _p, _p, _p],
    "glowk_unpack_weight_grad": [_p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p],
    "glowk_optim_workspace_floats": [],
    "glowk_optim_clip_norm": [_p, _i64, _f32, _f32, _p, _p],
    "glowk_optim_adam": [_p, _p, _p, _p, _i64, _p, _p, _f32, _f32, _f32, _f32, _i64, _p],
    "glowk_optim_adamax": [_p, _p, _p, _p, _i64, _p, _p, _f32, _f32, _f32, _f32, _i64, _p],
    "glowk_optim_schedule": [_p, _p, _f32, _i64, _f32, _f32, _f32, _p],
    "glowk_rows_max_channels": [],
    "glowk_rows_actnorm_mix": [_p, _p, _p, _p, _p, _p, _f32, _i64, _i64, _i32, _p],
    "glowk_rows_coupling_nblk": [_i64, _i64],
    "glowk_rows_coupling": [_p, _i64, _p, _p, _f32, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _f32,
                            _p, _f32, _p, _p, _p],
    "glowk_rows_coupling_rev_mix": [_p, _i64, _p, _p, _f32, _p, _p, _p, _p, _p, _p, _f32, _i64, _i64, _i64, _i64, _i32, _p],
    "glowk_rows_coupling_bwd": [_p, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _p],
    "glowk_rows_actnorm_mix_bwd": [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _i64, _i64, _i64,
                                   _i64, _p, _p, _p],
    "glowk_rows_actnorm_mix_bwd_ex": [_p, _p, _p, _i32, _i64, _i64, _p, _p, _p, _p, _f32, _p, _p, _p, _p, _i64, _i64,
                                      _i64, _i64, _p, _p, _p],
    "glowk_rows_max_channels_wide": [],
    "glowk_rows_actnorm_bwd": [_p, _p, _p, _p, _f32, _p, _p, _p, _i64, _i64, _p],
    "glowk_rows_gaussian_logp": [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p],
    "glowk_rows_split2d_sample": [_p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _p],
    "glowk_rows_split2d_bwd": [_p, _p, _i64, _p, _p, _f32, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _p],
    "glowk_rows_tapsum": [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "glowk_rows_squeeze_add": [_p, _p, _i32, _i64, _p, _i32, _i64, _i64, _i64, _i64, _i64, _i32, _p],
    "glowk_nll_head": [_p, _p, _f32, _f32, _i64, _i64, _p, _p, _p, _p],
    "glowk_nll_head_bwd": [_p, _p, _p, _p, _f32, _i64, _i64, _p, _p, _p],
    "glowk_rows_squeeze": [_p, _i32, _i64, _p, _i32, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _p],
    "glowk_pack_conv_weights_batched": [_p, _i64, _i64, _i32, _p],
    "glowk_unpack_weight_grads_batched": [_p, _i64, _i64, _p],
    "glowk_conv_actnorm_finish_batched": [_p, _i64, _i64, _p],
}
_RESTYPES = {"glowk_last_error": _c.c_char_p, "glowk_coupling_nblk": _i64, "glowk_optim_workspace_floats": _i64,
             "glowk_rows_coupling_nblk": _i64, "glowk_cnet_relu_mask_bytes": _i64}

_lib = None
launch_count = 0  # kernels-launching C-ABI calls made by this process (bench.py's gpu_launches claim)


class GlowkError(RuntimeError):
    pass


def lib():
    """Load libglowk.so (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GlowkError(
                "libglowk.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'`.  pytorch_glow_b200 has no CPU / eager fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _c.c_int)
        _lib = L
    return _lib


def check_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: `call` launches on torch's current stream of that device
    (one process per GPU sets it once; under torch.nn.DataParallel each replica thread has its own)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise GlowkError("pytorch_glow_b200 runs on CUDA tensors only (got a %s tensor); "
                             "there is no CPU fallback" % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise GlowkError("tensor on cuda:%d but the current device is cuda:%d: wrap the call in "
                             "torch.cuda.device(tensor.device) (kernels launch on the current device's stream)"
                             % (t.device.index, cur))


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "glowk expects contiguous tensors"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an int-returning entry point on torch's current stream; raise on failure."""
    global launch_count
    L = lib()
    rc = getattr(L, name)(*args, stream())
    if rc != 0:
        msg = L.glowk_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError("%s: %s" % (name, msg))
        raise GlowkError("%s failed (code %d): %s" % (name, rc, msg))
    launch_count += 1


def has_tcgen05():
    return bool(lib().glowk_has_tcgen05())


def coupling_nblk(hw):
    return int(lib().glowk_coupling_nblk(hw))
