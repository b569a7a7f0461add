"""ctypes binding of libst_b200.so (C ABI declared in include/st_b200.h).

The library is the product's only compute backend: there is no PyTorch/CPU fallback.  If the
shared object is missing, importing this module raises (run `python -m soft_truncation_b200.build`
or `__graft_entry__.build()`).
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libst_b200.so')

c_int, c_i64, c_u64, c_f, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p


class GemmArgs(ctypes.Structure):
  """Mirror of `st_gemm_args` (include/st_b200.h)."""
  _fields_ = [
      ('a_mode', ctypes.c_int32), ('b_mode', ctypes.c_int32), ('in_dtype', ctypes.c_int32),
      ('out_dtype', ctypes.c_int32), ('backend', ctypes.c_int32), ('accumulate', ctypes.c_int32),
      ('split_k', ctypes.c_int32),
      ('M', ctypes.c_int32), ('N', ctypes.c_int32), ('K', ctypes.c_int32), ('batch', ctypes.c_int32),
      ('sAm', c_i64), ('sAk', c_i64), ('sAb', c_i64),
      ('sBn', c_i64), ('sBk', c_i64), ('sBb', c_i64),
      ('sCm', c_i64), ('sCb', c_i64),
      ('A', c_p), ('A2', c_p), ('B', c_p), ('B2', c_p), ('C', c_p),
      ('n_img', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('C1', ctypes.c_int32),
      ('C2', ctypes.c_int32), ('kh', ctypes.c_int32), ('kw', ctypes.c_int32),
      ('bias', c_p), ('rowbias', c_p), ('rows_per_rb', ctypes.c_int32), ('ld_rb', c_i64),
      ('residual', c_p), ('sRm', c_i64), ('sRb', c_i64),
      ('alpha', c_f),
      ('gn_part', c_p), ('gn_hw', ctypes.c_int32), ('gn_rows_out', ctypes.POINTER(ctypes.c_int32)),
      ('dz_x', c_p), ('dz_ldx', c_i64), ('dz_cst', c_p), ('dz_keep', c_p), ('dz_inv_keep', c_f),
      ('dz_act', ctypes.c_int32),
  ]


# name -> argument ctypes (every function returns int status unless noted)
SIGNATURES = {
    'st_gemm': [ctypes.POINTER(GemmArgs), c_p],
    'st_gemm_simt_fallbacks': [c_int],
    'st_attn_fwd_supported': [c_int, c_int, c_int],
    'st_attn_fwd': [c_p, c_p, c_p, c_int, c_int, c_int, c_f, c_p],
    'st_gn_stats': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p],
    'st_gn_finalize': [c_p, c_int, c_int, c_int, c_i64, c_f, c_p, c_p, c_p],
    'st_gn_apply': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f, c_u64, c_p,
                    c_p, c_p, c_p, c_int, c_i64, c_f, c_p, c_p, c_int, c_p],
    'st_gn_bwd_reduce': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f,
                         c_u64, c_p, c_p, c_int, c_p, c_p],
    'st_gn_bwd_params': [c_p, c_int, c_int, c_p, c_p, c_p],
    'st_gn_bwd_apply': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f,
                        c_u64, c_p, c_p, c_int, c_p, c_p, c_f, c_p, c_int, c_p, c_int, c_int, c_p, c_p, c_p, c_p],
    'st_gn_chunks': [c_int, c_int, c_int],
    'st_gn_fwd_fused_chunks': [c_int, c_int, c_int],
    'st_gn_fwd_fused': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_f, c_int, c_f, c_u64, c_p, c_p, c_p,
                        c_p, c_p, c_int, c_p],
    'st_gn_bwd_resident_chunks': [c_int, c_int, c_int, c_int],
    'st_gn_bwd_resident': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f,
                           c_u64, c_p, c_p, c_int, c_p, c_p, c_f, c_p, c_p, c_p, c_p],
    'st_gn_bwd_fused_chunks': [c_int, c_int, c_int, c_int, c_int],
    'st_gn_bwd_fused': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f,
                        c_u64, c_p, c_p, c_int, c_p, c_p, c_f, c_p, c_int, c_p, c_int, c_p, c_p],
    'st_gn_bwd_wave_plan': [c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32)],
    'st_gn_bwd_wave': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_int, c_f, c_u64, c_p,
                       c_p, c_int, c_int, c_p, c_p, c_f, c_p, c_int, c_p, c_int, c_p, c_p, c_int, c_p],
    'st_gn_bwd_consts': [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p, c_p],
    'st_gn_bwd_dz_apply': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_f, c_p,
                           c_int, c_p, c_int, c_int, c_p, c_p, c_p],
    'st_prep_batch': [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_f, c_f, c_p],
    'st_cast': [c_p, c_int, c_p, c_int, c_i64, c_p],
    'st_axpby': [c_p, c_p, c_p, c_int, c_f, c_f, c_i64, c_p],
    'st_silu': [c_p, c_p, c_int, c_i64, c_p],
    'st_silu_bwd': [c_p, c_p, c_p, c_int, c_i64, c_p],
    'st_resample2x': [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f, c_p],
    'st_colsum': [c_p, c_int, c_i64, c_i64, c_int, c_i64, c_f, c_p, c_int, c_p],
    'st_colsum_batched': [c_p, c_int, c_int, c_p],
    'st_transpose_conv_weights': [c_p, c_p, c_int, c_p, c_p, c_int, c_i64, c_p],
    'st_softmax_fwd': [c_p, c_p, c_int, c_i64, c_int, c_f, c_p],
    'st_softmax_bwd': [c_p, c_p, c_p, c_int, c_i64, c_int, c_f, c_p],
    'st_timestep_embedding': [c_p, c_p, c_p, c_int, c_int, c_p],
    'st_fourier_embedding': [c_p, c_p, c_p, c_int, c_int, c_p],
    'st_nchw_to_nhwc': [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_f, c_f, c_p],
    'st_nhwc_to_nchw': [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_p, c_p],
    'st_im2col_small': [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    'st_im2col': [c_p, c_p] + [c_int] * 11 + [c_p],
    'st_col2im': [c_p, c_p] + [c_int] * 11 + [c_p],
    'st_upfirdn2d': [c_p, c_p, c_int, c_p] + [c_int] * 14 + [c_p],
    'st_fused_bias_act': [c_p, c_p, c_p, c_p, c_int, c_i64, c_int, c_int, c_int, c_int, c_f, c_f, c_p],
    'st_dsm_perturb': [c_p, c_p, c_p, c_p, c_p, c_int, c_i64, c_p],
    'st_dsm_loss': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_i64, c_int, c_p],
    'st_sumsq': [c_p, c_i64, c_p, c_p],
    'st_adam_ema': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p] + [c_f] * 9 + [c_p, c_p],
    'st_set_dropout_seed_offset': [c_p],
    'st_pc_update': [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_i64, c_p],
    'st_batch_norms': [c_p, c_p, c_p, c_int, c_i64, c_p],
    'st_langevin_coeffs': [c_p, c_p, c_f, c_p, c_p, c_p, c_int, c_p],
    'st_version': [],
    'st_tc_available': [],
}


class StError(RuntimeError):
  pass


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(f'{LIB_PATH} is missing: build it with `python -m soft_truncation_b200.build` '
                      '(the CUDA library is the only backend; there is no CPU fallback)')
  lib = ctypes.CDLL(LIB_PATH)
  for name, args in SIGNATURES.items():
    fn = getattr(lib, name)          # AttributeError here = header/library mismatch
    fn.argtypes = args
    fn.restype = c_int
  for name in ('st_last_error', 'st_gemm_simt_fallback_reason'):
    getattr(lib, name).argtypes = []
    getattr(lib, name).restype = ctypes.c_char_p
  return lib


lib = _load()


launches = 0   # C-ABI calls issued so far (each enqueues at least one of our kernels); bench.py reads it


def check(status):
  global launches
  launches += 1
  if status != 0:
    raise StError(lib.st_last_error().decode() or f'libst_b200 error {status}')
