"""ctypes binding of ``libswem_b200.so`` (C ABI declared in ``include/swem_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C swem_b200/csrc``.  There is
no fallback: if it is missing, or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libswem_b200.so')

PATH_AUTO, PATH_GENERIC, PATH_FUSED = 0, 1, 2

EXPORTS = (
    'swem_abi_version', 'swem_decode_tail_masks', 'swem_last_error', 'swem_device_check', 'swem_last_launch_count', 'swem_total_launch_count',
    'swem_em_workspace_bytes', 'swem_em_forward', 'swem_em_fused_supported',
    'swem_readout_workspace_bytes', 'swem_readout_forward', 'swem_readout_fused_supported',
    'swem_em_masks', 'swem_decode_tail', 'swem_set_profile_buffer',
    'swem_em_backward_workspace_bytes', 'swem_em_backward', 'swem_readout_backward_workspace_bytes', 'swem_readout_backward', 'swem_upsample_add', 'swem_bias_add_act', 'swem_glu_gate', 'swem_maxpool3x3s2', 'swem_tf32_split', 'swem_tf32_split_bf16', 'swem_bf16_widen_add', 'swem_stem_input', 'swem_resblock_tail_pred',
    'swem_cbam_channel_gate', 'swem_cbam_spatial_pool', 'swem_cbam_apply',
    'swem_fusion_weight_bytes', 'swem_fusion_prepare_weights', 'swem_fusion_workspace_bytes', 'swem_fusion_conv_glu',
)


class SwemDims(C.Structure):
    _fields_ = [('B', C.c_int32), ('N', C.c_int32), ('Ck', C.c_int32), ('Cv', C.c_int32),
                ('HW', C.c_int32), ('L', C.c_int32), ('n_iters', C.c_int32), ('n_banks', C.c_int32),
                ('topl', C.c_int32), ('tau', C.c_float)]


class SwemEmArgs(C.Structure):
    _fields_ = [('dims', SwemDims),
                ('x', C.c_void_p), ('v', C.c_void_p), ('masks', C.c_void_p),
                ('kappa_prior', C.c_void_p), ('nu_prior', C.c_void_p), ('zita_prior', C.c_void_p),
                ('kappa', C.c_void_p), ('nu', C.c_void_p), ('zita', C.c_void_p),
                ('z_last', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
                ('path', C.c_int32), ('v_pixel_major', C.c_int32),
                ('image_workspace', C.c_void_p), ('image_bank', C.c_int32), ('image_n_banks', C.c_int32)]


class SwemEmBwdArgs(C.Structure):
    _fields_ = [('dims', SwemDims),
                ('z_last', C.c_void_p), ('zita_prior', C.c_void_p), ('zita', C.c_void_p), ('grad_nu', C.c_void_p),
                ('grad_v', C.c_void_p), ('grad_nu_prior', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class SwemReadArgs(C.Structure):
    _fields_ = [('dims', SwemDims),
                ('qk', C.c_void_p),
                ('kappa', C.c_void_p * 2), ('nu', C.c_void_p * 2),
                ('out', C.c_void_p),
                ('out_channels', C.c_int32), ('mem_channel', C.c_int32), ('s_channel', C.c_int32),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
                ('path', C.c_int32), ('out_pixel_major', C.c_int32), ('bank_images_valid', C.c_int32),
                ('mkm_kernels', C.c_int32), ('mkm_sigma', C.c_float), ('mkm_width', C.c_int32), ('drop_mask', C.c_void_p)]


class SwemReadBwdArgs(C.Structure):
    _fields_ = [('dims', SwemDims),
                ('qk', C.c_void_p), ('kappa', C.c_void_p * 2), ('nu', C.c_void_p * 2), ('grad_out', C.c_void_p),
                ('out_channels', C.c_int32), ('mem_channel', C.c_int32), ('s_channel', C.c_int32),
                ('grad_qk', C.c_void_p), ('grad_nu', C.c_void_p * 2),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t), ('drop_mask', C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: the SWEM hot path is CUDA-only (no CPU fallback). '
            'Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C swem_b200/csrc`.')
    lib = C.CDLL(LIB_PATH)
    lib.swem_abi_version.restype = C.c_int
    lib.swem_last_error.restype = C.c_char_p
    lib.swem_device_check.argtypes = [C.c_int]
    lib.swem_last_launch_count.restype = C.c_int
    lib.swem_total_launch_count.restype = C.c_longlong
    lib.swem_em_workspace_bytes.argtypes = [C.POINTER(SwemDims), C.c_int32]
    lib.swem_em_workspace_bytes.restype = C.c_size_t
    lib.swem_em_forward.argtypes = [C.POINTER(SwemEmArgs), C.c_void_p]
    lib.swem_em_fused_supported.argtypes = [C.POINTER(SwemDims)]
    lib.swem_em_backward_workspace_bytes.argtypes = [C.POINTER(SwemDims)]
    lib.swem_em_backward_workspace_bytes.restype = C.c_size_t
    lib.swem_em_backward.argtypes = [C.POINTER(SwemEmBwdArgs), C.c_void_p]
    lib.swem_readout_backward_workspace_bytes.argtypes = [C.POINTER(SwemDims)]
    lib.swem_readout_backward_workspace_bytes.restype = C.c_size_t
    lib.swem_readout_backward.argtypes = [C.POINTER(SwemReadBwdArgs), C.c_void_p]
    lib.swem_readout_workspace_bytes.argtypes = [C.POINTER(SwemDims), C.c_int32]
    lib.swem_readout_workspace_bytes.restype = C.c_size_t
    lib.swem_readout_forward.argtypes = [C.POINTER(SwemReadArgs), C.c_void_p]
    lib.swem_readout_fused_supported.argtypes = [C.POINTER(SwemDims)]
    lib.swem_em_masks.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.swem_decode_tail.argtypes = [C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p] * 4
    lib.swem_decode_tail_masks.argtypes = [C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p] * 6
    lib.swem_set_profile_buffer.argtypes = [C.c_void_p, C.c_size_t]
    lib.swem_upsample_add.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p] * 3
    lib.swem_maxpool3x3s2.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2
    lib.swem_tf32_split.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.swem_tf32_split_bf16.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.swem_bf16_widen_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.swem_stem_input.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 6 + [C.c_void_p] * 2
    lib.swem_resblock_tail_pred.argtypes = [C.c_void_p] * 4 + [C.c_float] + [C.c_int32] * 4 + [C.c_void_p] * 2
    lib.swem_cbam_channel_gate.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 3
    lib.swem_cbam_spatial_pool.argtypes = [C.c_void_p] * 2 + [C.c_int32, C.c_int64, C.c_int32] + [C.c_void_p] * 2
    lib.swem_cbam_apply.argtypes = [C.c_void_p] * 3 + [C.c_int32, C.c_int64, C.c_int32] + [C.c_void_p] * 2
    lib.swem_bias_add_act.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.swem_glu_gate.argtypes = [C.c_void_p] * 3 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.swem_fusion_weight_bytes.restype = C.c_size_t
    lib.swem_fusion_weight_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.swem_fusion_prepare_weights.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    lib.swem_fusion_workspace_bytes.restype = C.c_size_t
    lib.swem_fusion_workspace_bytes.argtypes = [C.c_int32] * 4
    lib.swem_fusion_conv_glu.argtypes = ([C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 +
                                         [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p])
    if lib.swem_abi_version() != 3:
        raise RuntimeError(f'libswem_b200.so ABI version {lib.swem_abi_version()} != 3')
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().swem_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (status {rc}): {msg}')
