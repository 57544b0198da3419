"""ctypes binding of libmtp_b200.so (the C ABI declared in include/mtp_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTP_B200_LIB") or os.path.join(_HERE, "libmtp_b200.so")      # env override: experimental builds (tools/)

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


class MtpError(RuntimeError):
    pass


class Epilogue(ctypes.Structure):
    """struct mtp_epilogue (include/mtp_b200.h)."""
    _fields_ = [("mode", c_int), ("ldo", c_int), ("bias", c_void_p), ("out", c_void_p), ("out2", c_void_p),
                ("aux", c_void_p), ("row_scale", c_void_p), ("rows_per_group", c_int), ("pos_rows", c_int),
                ("accumulate", c_int), ("ps_h", c_int), ("ps_w", c_int), ("ps_cout", c_int), ("colsum", c_void_p), ("sumsq", c_void_p), ("hilo", c_int), ("out_lo_offset", c_int), ("b_static", c_int)]


class GemmDesc(ctypes.Structure):
    """struct mtp_gemm_desc (include/mtp_b200.h)."""
    _fields_ = [("A", c_void_p), ("lda", c_int), ("a_mn_major", c_int), ("B", c_void_p), ("ldb", c_int), ("b_mn_major", c_int),
                ("M", c_int), ("N", c_int), ("K", c_int), ("ep", ctypes.POINTER(Epilogue))]


EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID, EPI_F32_POS, EPI_F32, EPI_BF16_DGELU, EPI_BF16_PIXSHUF = range(7)

# name -> argtypes (restype is always int unless listed in _SPECIAL)
_SIGNATURES = {
    "mtp_gemm_bf16": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(Epilogue), c_int, c_void_p],
    "mtp_gemm_bf16_dual": [ctypes.POINTER(GemmDesc), ctypes.POINTER(GemmDesc), c_int, c_void_p],
    "mtp_set_pdl": [c_int],
    "mtp_set_sm_limit": [c_int],
    "mtp_gemm_last_config": [],
    "mtp_gemm_plan": [c_int] * 9 + [ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double)],
    "mtp_gemm_set_debug": [c_void_p],
    "mtp_gemm_set_debug_mode": [c_int],
    "mtp_gemm_set_max_stages": [c_int],
    "mtp_gemm_set_variant": [c_int],
    "mtp_layernorm_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p],
    "mtp_layernorm_bwd": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_scale_cast_bf16": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "mtp_colsum_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "mtp_cast_f32_bf16": [c_void_p, c_void_p, c_size_t, c_void_p],
    "mtp_add_bf16_into_f32": [c_void_p, c_void_p, c_size_t, c_void_p],
    "mtp_add_f32": [c_void_p, c_void_p, c_size_t, c_void_p],
    "mtp_rvsa_sampling_fwd": [c_void_p] * 9 + [c_int] * 5 + [c_void_p],
    "mtp_rvsa_attn_fwd": [c_void_p] * 7 + [c_int] * 5 + [c_void_p],
    "mtp_full_attn_fwd": [c_void_p] * 5 + [c_int] * 5 + [c_void_p],
    "mtp_patchify": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_patchify_u8": [c_void_p, c_int, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_tok_to_nchw": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_nchw_to_tok": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_maxpool2_tok_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_maxpool2_tok_bwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_sqloss_fwd_bwd": [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "mtp_sqloss_fwd_bwd_w": [c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_void_p],
    "mtp_rvsa_attn_bwd": [c_void_p] * 14 + [c_int] * 6 + [c_void_p],
    "mtp_rvsa_sampling_bwd": [c_void_p] * 13 + [c_int] * 5 + [c_void_p],
    "mtp_rvsa_attn_bwd_fused": [c_void_p] * 25 + [c_int] * 5 + [c_void_p],
    "mtp_full_attn_bwd": [c_void_p] * 10 + [c_int] * 5 + [c_void_p],
    "mtp_split_hilo": [c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p],
    "mtp_patchify_hilo": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_layernorm_fwd_hilo": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p],
    "mtp_rvsa_sampling_fwd_hilo": [c_void_p] * 9 + [c_int] * 5 + [c_void_p],
    "mtp_rvsa_attn_fwd_hilo": [c_void_p] * 6 + [c_int] * 5 + [c_void_p],
    "mtp_full_attn_fwd_hilo": [c_void_p] * 4 + [c_int] * 5 + [c_void_p],
    "mtp_tok_to_nchw_hilo": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_empty_launch": [c_void_p],
    "mtp_probe_launch": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_probe_launch2": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mtp_optim_step_begin": [c_void_p, c_void_p],
    "mtp_sumsq_f32": [c_void_p, c_size_t, c_void_p, c_void_p],
    "mtp_sumsq_bf16": [c_void_p, c_size_t, c_void_p, c_void_p],
    "mtp_adamw_step_mixed": [c_void_p, c_void_p, c_void_p, c_size_t] + [c_void_p] * 7 + [c_size_t, c_float, c_float, c_int, c_float, c_float, c_float, c_float, c_float, c_void_p],
    "mtp_adamw_step": [c_void_p] * 9 + [c_size_t, c_float, c_float, c_int, c_float, c_float, c_float, c_float, c_float, c_void_p],
}
_SIZE_FNS = {
    "mtp_rvsa_bwd_workspace_bytes": [c_int] * 5,
    "mtp_rvsa_sampling_bwd_workspace_bytes": [c_int] * 5,
    "mtp_full_attn_bwd_workspace_bytes": [c_int] * 4,
}

_lib = None


def load():
    """Load the shared library (once).  Raises MtpError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtpError(f"{LIB_PATH} not found: build it with `python -m mtp_b200.build` "
                       "(there is no CPU or PyTorch fallback for the backbone kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mtp_last_error.restype = ctypes.c_char_p
    lib.mtp_last_error.argtypes = []
    lib.mtp_version.restype = c_int
    lib.mtp_num_sms.restype = c_int
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    for name, args in _SIZE_FNS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_size_t
    _lib = lib
    return lib


def exported_symbols():
    return ["mtp_last_error", "mtp_version", "mtp_num_sms"] + list(_SIGNATURES) + list(_SIZE_FNS)


def check(rc, what=""):
    if rc != 0:
        raise MtpError(f"{what} failed ({rc}): {load().mtp_last_error().decode()}")


# ---- measurement aid (tools/step_breakdown.py): kernel families whose launches are replaced by an empty kernel, so that the time a
#      family costs INSIDE the graph-replayed step can be read off as a difference.  Results are garbage while a family is skipped.
FAMILIES = {
    "gemm": ("mtp_gemm_bf16", "mtp_gemm_bf16_dual"),
    "rvsa_attn_fwd": ("mtp_rvsa_attn_fwd",), "rvsa_sampling_fwd": ("mtp_rvsa_sampling_fwd",),
    "rvsa_attn_bwd": ("mtp_rvsa_attn_bwd", "mtp_rvsa_attn_bwd_fused"), "rvsa_sampling_bwd": ("mtp_rvsa_sampling_bwd",),
    "dense_attn_fwd": ("mtp_full_attn_fwd",), "dense_attn_bwd": ("mtp_full_attn_bwd",),
    "layernorm_fwd": ("mtp_layernorm_fwd",), "layernorm_bwd": ("mtp_layernorm_bwd",),
    "optimizer": ("mtp_adamw_step", "mtp_adamw_step_mixed", "mtp_sumsq_f32", "mtp_sumsq_bf16", "mtp_optim_step_begin"),
    "layout": ("mtp_patchify", "mtp_patchify_u8", "mtp_tok_to_nchw", "mtp_nchw_to_tok", "mtp_maxpool2_tok_fwd", "mtp_maxpool2_tok_bwd"),
    "casts_colsums": ("mtp_scale_cast_bf16", "mtp_colsum_bf16", "mtp_cast_f32_bf16", "mtp_add_bf16_into_f32", "mtp_add_f32"),
    "heads": ("mtp_sqloss_fwd_bwd", "mtp_sqloss_fwd_bwd_w"),
}
_skip = set()


def set_skipped_families(names):
    _skip.clear()
    for n in names:
        _skip.update(FAMILIES[n])


def call(name, *args):
    if _skip and name in _skip:
        name, args = "mtp_empty_launch", (args[-1],)       # the stream is the last argument of every launching entry point
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise MtpError(f"{name} failed ({rc}): {load().mtp_last_error().decode()}")
