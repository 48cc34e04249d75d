"""ctypes binding of libpcfd_sm100.so (the C ABI in include/pcfd.h).

The library is the ONLY implementation of the hot path: if it is missing, or a call returns a
non-zero status, this module raises -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PCFD_LIB_PATH') or os.path.join(HERE, 'libpcfd_sm100.so')   # override: experiments with variant builds

ACT_NONE, ACT_SILU, ACT_TANH = 0, 1, 2
ACT_CODES = {None: ACT_NONE, 'none': ACT_NONE, 'silu': ACT_SILU, 'tanh': ACT_TANH}
LOSS_KINDS = {'manufactured': 0, 'fixed': 1, 'variable': 2}
LAP_MODES = {'reference': 0, 'true': 1}
LOSS_OUT_FLOATS = 48
ENGINE_FFMA, ENGINE_THIN, ENGINE_TCGEN05 = 0, 1, 2   # answers of pcfd_jet_linear_engine

ERRORS = {1: 'bad argument (shape / null pointer / unsupported channel count)', 2: 'misaligned pointer',
          3: 'workspace too small', 4: 'device is not sm_100 (no fallback path exists)'}


class PcfdError(RuntimeError):
    pass


class InTrans(C.Structure):
    _fields_ = [('act', C.c_int32), ('act_cols', C.c_int32), ('escale', C.c_void_p), ('ldescale', C.c_int32),
                ('drop_p', C.c_float), ('seed_dev', C.c_void_p), ('salt', C.c_uint32), ('reserved', C.c_uint32)]


class DpPeers(C.Structure):
    _fields_ = [('grad', C.c_void_p * 16), ('param', C.c_void_p * 16), ('flags', C.c_void_p * 16), ('grad_mc', C.c_void_p),
                ('param_mc', C.c_void_p), ('rank', C.c_int32), ('world', C.c_int32)]


class ResidualParams(C.Structure):
    _fields_ = [('dims', C.c_int32), ('loss_kind', C.c_int32), ('lap_mode', C.c_int32), ('enable_data_loss', C.c_int32),
                ('nu', C.c_float), ('d', C.c_float), ('f', C.c_float),
                ('c_std', C.c_float * 3), ('u_std', C.c_float * 3), ('u_mean', C.c_float * 3),
                ('p_std', C.c_float), ('p_mean', C.c_float),
                ('d_min', C.c_float * 3), ('d_range', C.c_float * 3), ('f_min', C.c_float * 3), ('f_range', C.c_float * 3),
                ('col_u', C.c_int32 * 3), ('col_p', C.c_int32), ('col_zone', C.c_int32),
                ('col_d', C.c_int32 * 3), ('col_f', C.c_int32 * 3), ('weights', C.c_float * 16)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_IT = C.POINTER(InTrans)

# name -> (restype, argtypes); mirrors include/pcfd.h declaration by declaration
SIGNATURES = {
    'pcfd_abi_version': (C.c_int, []),
    'pcfd_device_arch': (C.c_int, [C.POINTER(C.c_int)]),
    'pcfd_jet_linear_fwd': (C.c_int, [_P, _I64, _I32, _IT, _P, _I32, _P, _P, _I32, _P, _I64, _I32,
                                      _I32, _I64, _I64, _I32, _I32, _P]),
    'pcfd_jet_linear_bwd_dx': (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I64, _I32, _IT, _P, _I64, _I32, _P, _I32,
                                         _I32, _I64, _I64, _I32, _I32, _P]),
    'pcfd_jet_linear_bwd_dw_workspace_bytes': (_SZ, [_I32, _I64, _I64, _I32, _I32]),
    'pcfd_jet_linear_bwd_dw': (C.c_int, [_P, _I64, _I32, _P, _I64, _I32, _IT, _P, _I32, _P, _P, _I32,
                                         _I32, _I64, _I64, _I32, _I32, _P, _SZ, _P]),
    'pcfd_jet_linear_engine': (C.c_int, [_I32, _P, _I64, _I32, _P, _I32, _P, _I64, _I32, _IT, _I32, _I32, _I64, _I32, _I32]),
    'pcfd_ffma_jet_linear_fwd': (C.c_int, [_P, _I64, _I32, _IT, _P, _I32, _P, _P, _I32, _P, _I64, _I32,
                                           _I32, _I64, _I64, _I32, _I32, _P]),
    'pcfd_ffma_jet_linear_bwd_dx': (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I64, _I32, _IT, _P, _I64, _I32, _P, _I32,
                                              _I32, _I64, _I64, _I32, _I32, _P]),
    'pcfd_ffma_jet_linear_bwd_dw': (C.c_int, [_P, _I64, _I32, _P, _I64, _I32, _IT, _P, _I32, _P, _P, _I32,
                                              _I32, _I64, _I64, _I32, _I32, _P, _SZ, _P]),
    'pcfd_segmax_fwd': (C.c_int, [_P, _I32, _I32, _P, _I64, _I32, _I32, _P, _I32, _P, _P]),
    'pcfd_segmax_fwd_z': (C.c_int, [_P, _I32, _I32, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _I32, _P]),
    'pcfd_segmax_bwd': (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _I64, _I32, _I32, _P, _I32, _P]),
    'pcfd_pool_layer_bwd_supported': (C.c_int, [_I64, _I32, _I32, _I32, _IT, _I32]),
    'pcfd_pool_layer_bwd_workspace_bytes': (_SZ, [_I64, _I32, _I32, _I32]),
    'pcfd_pool_layer_bwd': (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _I64, _I32, _I32, _P, _I32, _IT, _I32, _P, _I32, _P, _I32,
                                      _P, _P, _I32, _P, _SZ, _P]),
    'pcfd_pool_compact': (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _I64, _I32, _P, _P, _I32, _P]),
    'pcfd_fps': (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P]),
    'pcfd_fps_workspace_bytes': (C.c_size_t, [_I32, _I32, _I32]),
    'pcfd_fps_ws': (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P, C.c_size_t, _P]),
    'pcfd_ball_query': (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _F, _I32, _P, _P, _P]),
    'pcfd_sa_edges': (C.c_int, [_P, _I64, _I32, _I64, _P, _P]),
    'pcfd_sa_cached_geometry': (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    'pcfd_sa_gather': (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _I64, _I32, _F, _P, _I32, _P]),
    'pcfd_sa_scatter_bwd': (C.c_int, [_P, _I32, _P, _I64, _I32, _I32, _P, _I32, _P]),
    'pcfd_gather_cols': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, _I64, C.POINTER(C.c_int32), _I32, _P, _I32,
                                   _I64, _I64, _I32, _P]),
    'pcfd_seed_jet': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, C.POINTER(C.c_int32), _I32, _I32, _P, _I64, _I32, _P]),
    'pcfd_residual_workspace_bytes': (_SZ, [_I32, _I64, _I64, _I64]),
    'pcfd_residual_loss': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I32,
                                     C.POINTER(ResidualParams), _P, _P, _P, _P, _SZ, _P]),
    'pcfd_residual_loss_w': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I32,
                                       C.POINTER(ResidualParams), _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'pcfd_residual_step': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I32,
                                     C.POINTER(ResidualParams), _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'pcfd_residual_fields': (C.c_int, [_P, _I32, _I64, _I32, _P, _I64, _P, _I64, _I32, C.POINTER(ResidualParams), _P, _P]),
    'pcfd_residual_eval': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(ResidualParams), _P, _P, _P]),
    'pcfd_mean_squares': (C.c_int, [_P, _I64, _I32, _P, _P]),
    'pcfd_dp_flags_len': (_I32, []),
    'pcfd_dp_adam_step': (C.c_int, [C.POINTER(DpPeers), _P, _P, _P, _P, _F, _F, _F, _F, _I64, _P, _P]),
    'pcfd_adam_step': (C.c_int, [_P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _I64, _P]),
    'pcfd_relobralo_update': (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _F, _F, _F, _F, C.c_uint64, _P, _P]),
    'pcfd_sdf_feature': (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    'pcfd_sdf_scratch_bytes': (_SZ, [_I32, _I32]),
    'pcfd_boundary_one_hot': (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _I32, _I32, _P]),
    'pcfd_gather_blocks': (C.c_int, [_P, _I64, _P, _I64, _P, _P]),
    'pcfd_gather_blocks_multi': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _I32, _P, _I64, _P]),
    'pcfd_zero': (C.c_int, [_P, _I64, _P]),
    'pcfd_advance_seed': (C.c_int, [_P, _P]),
}

_lib = None
launches = 0  # number of C-ABI calls that enqueue kernels (bench.py reports it as gpu_launches)


def lib_path() -> str:
    return LIB_PATH


def load() -> C.CDLL:
    """Load the native library (once).  Raises PcfdError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcfdError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                        '(nvcc, sm_100a). There is no fallback implementation.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.pcfd_abi_version() != 1:
        raise PcfdError('libpcfd_sm100.so ABI version mismatch; rebuild')
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = ERRORS.get(status, f'CUDA error {status - 100}' if status >= 100 else 'unknown')
        raise PcfdError(f'{what} failed: status {status} ({msg})')
