"""ctypes binding of ``liblamp_b200.so`` (the C ABI declared in ``include/lamp_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m lamp_b200.build``.  There is NO
fallback: if the shared object is missing, or a kernel reports an error, the caller gets an exception.
torch is used only to obtain device pointers and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LAMP_B200_LIB') or os.path.join(_HERE, 'liblamp_b200.so')  # env: A/B builds of the library

PREC_FP32 = 0   # 3-term split-bf16 tensor-core products, fp32-grade results
PREC_BF16 = 1   # plain bf16 operands

_lib = None

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

class LampSplitJob(C.Structure):
    """``LampSplitJob`` of include/lamp_b200.h (one matrix of a ``lamp_split_planes_multi`` launch)."""
    _fields_ = [('src', C.c_void_p), ('hi', C.c_void_p), ('lo', C.c_void_p), ('rows', C.c_int), ('cols', C.c_int),
                ('ld', C.c_int64), ('ldp', C.c_int64), ('transpose', C.c_int)]


_SIGNATURES = {
    'lamp_version': ([], _i),
    'lamp_last_error': ([], C.c_char_p),
    'lamp_device_check': ([], _i),
    'lamp_sm_count': ([], _i),
    'lamp_set_tuning': ([_i, _i], _i),
    'lamp_split_planes': ([_vp, _i64, _i, _i64, _vp, _vp, _i64, _vp], _i),
    'lamp_split_planes_multi': ([_vp, _i, _vp], _i),
    'lamp_gemm_planes': ([_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp, _i64, _i, _vp, _i64, _vp, _vp,
                          _i64, _vp, _vp], _i),
    'lamp_gemm_planes_drop': ([_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _vp, _f, C.c_uint64, _vp, _vp, _i64, _i, _vp,
                               _i64, _vp], _i),
    'lamp_gemm_planes_pres': ([_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i, _vp, _i64, _vp,
                               _vp, _i64, _vp, _vp], _i),
    'lamp_gemm_ln_planes': ([_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _i64, _i, _vp, _vp, _f, _vp, _i64,
                             _vp, _vp, _i64, _vp], _i),
    'lamp_gemm_stats_parts': ([_i], _i),
    'lamp_gemm_planes_dln': ([_vp, _vp, _i64, _vp, _i, _f, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i64,
                              _vp, _vp], _i),
    'lamp_gemm_planes_rstats': ([_vp, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i,
                                 _f, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp], _i),
    'lamp_ln_apply': ([_vp, _vp, _vp, _i, _vp, _vp, _f, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    'lamp_diag_proj_ln': ([_vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _i64, _i, _i, _vp, _vp], _i),
    'lamp_attn_core_planes': ([_vp, _vp, _i64, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _i64,
                               _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _i),
    'lamp_pack_mask_bits': ([_vp, _i64, _i64, _i64, _i64, _i, _i, _vp, _vp], _i),
    'lamp_attn_core_planes_mbits': ([_vp, _vp, _i64, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _i64,
                                     _i64, _vp, _vp, _i64, _vp, _i64, _vp], _i),
    'lamp_attn_core_bwd_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'lamp_attn_core_bwd': ([_vp] * 10 + [_i, _i, _i, _i, _f, _f, _vp, _sz, _vp], _i),
    'lamp_layernorm_bwd': ([_vp, _vp, _vp, _f, _i64, _i, _vp, _vp, _vp, _vp], _i),
    'lamp_layernorm_bwd_drop': ([_vp, _vp, _vp, _f, _i64, _i, _vp, _vp, _vp, _f, C.c_uint64, _vp, _vp, _vp, _vp], _i),
    'lamp_gemm_tn_acc': ([_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i, _i, _vp, _vp, _vp], _i),
    'lamp_diag_proj_bwd': ([_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp], _i),
    'lamp_attn_core_planes_train': ([_vp, _vp, _i64, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _i64,
                                     _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _f, C.c_uint64, _vp, _vp], _i),
    'lamp_attn_core_planes_train_mbits': ([_vp, _vp, _i64, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp,
                                           _i64, _i64, _vp, _vp, _i64, _vp, _vp, _f, C.c_uint64, _vp, _vp], _i),
    'lamp_attn_bwd_planes_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'lamp_attn_bwd_planes': ([_vp, _vp, _i64, _i, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp,
                              _i64, _i, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _i64, _i64, _i64,
                              C.c_uint64, _vp, _vp, _sz, _vp], _i),
    'lamp_dropout_add': ([_vp, _vp, _i64, _i, _i, _f, C.c_uint64, _vp, _vp, _vp], _i),
    'lamp_dropout_split': ([_vp, _i64, _i, _f, C.c_uint64, _vp, _vp, _vp, _vp], _i),
    'lamp_relu_mask_planes': ([_vp, _vp, _vp, _i64, _vp], _i),
    'lamp_gold_binary': ([_vp, _i64, _i, _i, _i, _vp, _vp], _i),
    'lamp_bce_logits_workspace_bytes': ([], _sz),
    'lamp_bce_logits': ([_vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp], _i),
    'lamp_layernorm': ([_vp, _vp, _i, _vp, _vp, _f, _i64, _i, _vp, _vp, _vp, _vp, _vp], _i),
    'lamp_embed': ([_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    'lamp_embed_bwd': ([_vp, _vp, _vp, _i64, _i, _i64, _i64, _vp, _vp, _vp], _i),
    'lamp_gather_rows': ([_vp, _vp, _i64, _i, _vp, _vp], _i),
    'lamp_zero_guard_rows': ([_vp, _vp, _i64, _i, _vp, _i64, _i, _vp], _i),
    'lamp_diag_proj': ([_vp, _vp, _vp, _i64, _i, _i, _vp, _vp], _i),
    'lamp_sdpa_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'lamp_sdpa_fwd': ([_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _sz, _vp], _i),
    'lamp_sdpa_fwd_train': ([_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _f, C.c_uint64,
                             _vp, _vp, _sz, _vp], _i),
    'lamp_mha_workspace_bytes': ([_i, _i, _i, _i, _i, _i, _i, _i], _sz),
    'lamp_mha_fwd': ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i,
                      _i, _f, _vp, _sz, _vp], _i),
    'lamp_ffn_workspace_bytes': ([_i64, _i, _i], _sz),
    'lamp_ffn_fwd': ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _vp, _sz, _vp], _i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the native library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: the CUDA extension is not built (run `python -c "import __graft_entry__ as g; '
                f'g.build()"` at the repo root).  lamp_b200 has no CPU or eager fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
        # LAMP_TUNE="key=value,key=value": process-wide tuning knobs (include/lamp_b200.h) for experiments
        for kv in filter(None, os.environ.get('LAMP_TUNE', '').split(',')):
            k, v = kv.split('=')
            check(handle.lamp_set_tuning(int(k), int(v)), f'LAMP_TUNE {kv}')
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lamp_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  ``torch.cuda.current_stream()`` builds a Python
    Stream object and re-validates the device on every call (~15 us -- a third of the host time of a training step with
    ~350 launches); the raw accessor is a plain C call."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('lamp_b200 kernels need CUDA tensors: this package has no CPU path '
                               '(use the reference implementation or oracle/ for CPU runs)')
