"""ctypes binding of libregtr_b200.so (the C ABI declared in include/regtr_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails,
`RegtrLibError` is raised.  torch is imported first so that its libcudart is the one in the process
(the library links the CUDA runtime only -- no cuBLAS or other compute library).
"""
from __future__ import annotations

import ctypes
import os
import re

import torch  # noqa: F401  (loads libcudart first)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libregtr_b200.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'regtr_b200.h')


class RegtrLibError(RuntimeError):
    pass


_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float
_Z = _c.c_size_t

# name -> (restype, argtypes); mirrors include/regtr_b200.h one to one
SIGNATURES = {
    'regtr_version': (_I, []),
    'regtr_build_info': (_c.c_char_p, []),
    'regtr_grid_subsample_ws_bytes': (_Z, [_I, _I]),
    'regtr_grid_subsample_state_bytes': (_Z, [_I]),
    'regtr_grid_subsample': (_I, [_P, _P, _I, _I, _F, _P, _I, _P, _P, _P, _Z, _P, _Z, _P]),
    'regtr_grid_subsample_sorted_ws_bytes': (_Z, [_I]),
    'regtr_grid_subsample_sorted': (_I, [_P, _P, _I, _I, _F, _P, _I, _P, _P, _P, _Z, _P]),
    'regtr_cellgrid_bytes': (_Z, [_I]),
    'regtr_cellgrid_ws_bytes': (_Z, [_I]),
    'regtr_cellgrid_state_bytes': (_Z, [_I]),
    'regtr_cellgrid_build': (_I, [_P, _P, _I, _I, _F, _P, _P, _P, _P, _Z, _P, _Z, _P]),
    'regtr_ball_query': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P]),
    'regtr_kpconv_ws_bytes': (_Z, [_I, _I, _I]),
    'regtr_kpconv_fwd_ws_bytes': (_Z, [_I, _I, _I, _I]),
    'regtr_kpconv_fwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _F, _P, _P, _Z, _P]),
    'regtr_kpconv_aggregate': (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _F, _P, _P, _I, _P]),
    'regtr_max_pool': (_I, [_P, _P, _I, _I, _P, _I, _I, _P, _P]),
    'regtr_instnorm_ws_bytes': (_Z, [_I, _I, _I]),
    'regtr_instnorm_counter_bytes': (_Z, [_I, _I]),
    'regtr_instnorm_act': (_I, [_P, _P, _I, _I, _I, _F, _P, _F, _P, _P, _P, _Z, _P, _P]),
    'regtr_instnorm_apply': (_I, [_P, _P, _I, _I, _I, _P, _P, _F, _P, _P, _P]),
    'regtr_instnorm_part_bytes': (_Z, [_I, _I]),
    'regtr_gemm_tf32x3_instats': (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _F, _P, _P, _P, _Z, _P]),
    'regtr_split_tf32': (_I, [_P, _c.c_longlong, _P, _P, _P]),
    'regtr_gemm_ws_bytes': (_Z, [_I, _I, _I]),
    'regtr_gemm_tf32x3': (_I, [_P, _I, _P, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P, _Z, _P]),
    'regtr_pos_embed_sine': (_I, [_P, _I, _P, _I, _I, _F, _P, _P]),
    'regtr_layernorm_pos': (_I, [_P, _P, _P, _P, _I, _P, _I, _F, _P, _P, _P]),
    'regtr_attention_plan': (_I, [_P, _I, _P, _P]),
    'regtr_corr_decode_fwd': (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    'regtr_mha_varlen_fwd': (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _F, _P]),
    'regtr_gemm_tf32x3_qkv_bf16': (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P]),
    'regtr_mha_bf16_tc_fwd': (_I, [_P, _I, _P, _I, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    'regtr_gemm_tf32x3_qkv_split': (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _F, _P, _I, _P, _I, _P, _P]),
    'regtr_mha_tf32_tc_fwd': (_I, [_P, _I, _P, _I, _I, _P, _I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _P]),
    'regtr_kabsch_fwd': (_I, [_P, _P, _P, _P, _I, _P, _P]),
    'regtr_pose_from_corr': (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    'regtr_status_clear': (_I, [_P, _P]),
}

_ERR = {-1: 'REGTR_ERR_ARG (rejected argument)', -2: 'REGTR_ERR_WORKSPACE (workspace too small)',
        -3: 'REGTR_ERR_UNSUPPORTED (shape outside the hot path)'}

_lib = None


def header_symbols():
    """Every function name declared in include/regtr_b200.h."""
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(regtr_[a-z0-9_]+)\s*\(', text)))


def load():
    """Load (building first if the .so is absent and nvcc exists).  Raises RegtrLibError."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:
            raise RegtrLibError(f'{LIB_PATH} is missing and could not be built: {exc}') from exc
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise RegtrLibError(f'cannot load {LIB_PATH}: {exc}') from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise RegtrLibError(f'{LIB_PATH} does not export {name}') from exc
        fn.restype = res
        fn.argtypes = args
    if lib.regtr_version() != 1:
        raise RegtrLibError('ABI version mismatch')
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc <= -1000:
        raise RegtrLibError(f'{what}: CUDA error {-rc - 1000} at launch')
    raise RegtrLibError(f'{what}: {_ERR.get(rc, rc)}')
