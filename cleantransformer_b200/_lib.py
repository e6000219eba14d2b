"""ctypes loader for libct_b200.so — the thin C-ABI boundary (include/ct_b200.h).

The product path never falls back to CPU or eager PyTorch math: if the shared library is missing
or the device is not an sm_100 part, importing/using the ops raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CT_B200_LIB") or os.path.join(_HERE, "libct_b200.so")

_lock = threading.Lock()
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_float = ctypes.c_float
c_double = ctypes.c_double


class GemmArgs(ctypes.Structure):
    """Mirror of `ct_gemm_args` in include/ct_b200.h (field order and types must match)."""

    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("a_mn_major", ctypes.c_int32), ("b_mn_major", ctypes.c_int32),
        ("ab_dtype", ctypes.c_int32),
        ("A", c_void_p), ("lda", c_i64),
        ("B", c_void_p), ("ldb", c_i64),
        ("C", c_void_p), ("c_dtype", ctypes.c_int32), ("ldc", c_i64),
        ("alpha", c_float), ("beta", c_float),
        ("bias", c_void_p),
        ("act", ctypes.c_int32),
        ("preact", c_void_p), ("preact_dtype", ctypes.c_int32), ("ldp", c_i64),
        ("actgrad_src", c_void_p), ("actgrad_dtype", ctypes.c_int32),
        ("actgrad_act", ctypes.c_int32), ("ldg", c_i64),
        ("residual", c_void_p), ("res_dtype", ctypes.c_int32), ("ldr", c_i64),
        ("impl", ctypes.c_int32),
        ("row_stats", c_void_p),
    ]


class AttnArgs(ctypes.Structure):
    """Mirror of `ct_attn_args`."""

    _fields_ = [
        ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("Sq", ctypes.c_int32), ("Sk", ctypes.c_int32),
        ("D", ctypes.c_int32), ("dtype", ctypes.c_int32),
        ("q", c_void_p), ("q_sb", c_i64), ("q_sh", c_i64), ("q_ss", c_i64),
        ("k", c_void_p), ("k_sb", c_i64), ("k_sh", c_i64), ("k_ss", c_i64),
        ("v", c_void_p), ("v_sb", c_i64), ("v_sh", c_i64), ("v_ss", c_i64),
        ("o", c_void_p), ("o_sb", c_i64), ("o_sh", c_i64), ("o_ss", c_i64),
        ("lse2", c_void_p),
        ("scale", c_float), ("causal", ctypes.c_int32), ("causal_fill", c_float),
        ("kbias2", c_void_p), ("kb_sb", c_i64), ("kb_sh", c_i64),
        ("first_valid", c_void_p),
        ("impl", ctypes.c_int32),
        ("seq_len_dev", c_void_p),
        ("dropout_p", c_float), ("rng_stream", ctypes.c_uint32), ("rng_seed", ctypes.c_uint64),
        ("k_new", c_void_p), ("v_new", c_void_p), ("kn_sb", c_i64), ("kn_sh", c_i64), ("vn_sb", c_i64), ("vn_sh", c_i64),
    ]


class AttnBwdArgs(ctypes.Structure):
    """Mirror of `ct_attn_bwd_args`."""

    _fields_ = [
        ("f", AttnArgs),
        ("dout", c_void_p),
        ("dq", c_void_p), ("dq_sb", c_i64), ("dq_sh", c_i64), ("dq_ss", c_i64),
        ("dk", c_void_p), ("dk_sb", c_i64), ("dk_sh", c_i64), ("dk_ss", c_i64),
        ("dv", c_void_p), ("dv_sb", c_i64), ("dv_sh", c_i64), ("dv_ss", c_i64),
        ("delta", c_void_p), ("dq_accum", c_void_p),
    ]


class LnBwdArgs(ctypes.Structure):
    """Mirror of `ct_ln_bwd_args`."""

    _fields_ = [
        ("rows", c_i64), ("cols", c_i64),
        ("dy", c_void_p), ("dy_dtype", ctypes.c_int32),
        ("dy2", c_void_p), ("dy2_dtype", ctypes.c_int32),
        ("x", c_void_p), ("x_dtype", ctypes.c_int32),
        ("gamma", c_void_p), ("mean", c_void_p), ("rstd", c_void_p),
        ("dx_add", c_void_p), ("dx_add_dtype", ctypes.c_int32),
        ("dx", c_void_p), ("dx_dtype", ctypes.c_int32),
        ("dx2", c_void_p), ("dx2_dtype", ctypes.c_int32),
        ("dgamma", c_void_p), ("dbeta", c_void_p), ("dgb_accumulate", ctypes.c_int32),
        ("dxsum", c_void_p), ("dxsum_accumulate", ctypes.c_int32),
        ("workspace", c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


# name -> (restype, argtypes); every symbol declared in include/ct_b200.h
SIGNATURES = {
    "ct_version": (c_int, []),
    "ct_last_error": (c_int, [ctypes.c_char_p, ctypes.c_size_t]),
    "ct_device_check": (c_int, [c_int]),
    "ct_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "ct_get_option": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_int)]),
    "ct_layernorm_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                 c_int, c_void_p, c_void_p, c_i64, c_i64, c_float, c_void_p]),
    "ct_layernorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_int, c_void_p, ctypes.c_size_t, c_i64, c_i64, c_void_p]),
    "ct_layernorm_bwd_ex": (c_int, [ctypes.POINTER(LnBwdArgs), c_void_p]),
    "ct_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_double,
                              c_double, c_double, c_double, c_double, c_i64, c_int, c_float, c_void_p]),
    "ct_adamw_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_double, c_double, c_double, c_double, c_double, c_i64, c_int, c_float,
                               c_void_p]),
    "ct_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_float, c_float, c_float,
                            c_float, c_int, c_void_p]),
    "ct_sgd_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float,
                             c_int, c_void_p]),
    "ct_cast": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_void_p]),
    "ct_colsum": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_int, c_i64, c_i64, c_void_p]),
    "ct_act_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_i64, c_void_p]),
    "ct_act_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_i64,
                           c_void_p]),
    "ct_gemm": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "ct_gemm_bias_act": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                 c_int, c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ct_gemm_dgrad": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_i64,
                              c_i64, c_i64, c_int, c_void_p]),
    "ct_attn_fwd": (c_int, [ctypes.POINTER(AttnArgs), c_void_p]),
    "ct_attn_bwd": (c_int, [ctypes.POINTER(AttnBwdArgs), c_void_p]),
    "ct_attn_occupancy": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "ct_attn_mask_prep": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "ct_embedding_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ct_embedding_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ct_embedding_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_void_p, c_void_p,
                                           c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                           c_void_p, c_void_p, c_i64, c_i64, c_float, c_void_p]),
    "ct_cross_entropy_fwd": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_void_p, c_i64, c_void_p,
                                     c_void_p, c_i64, c_i64, c_i64, c_int, c_i64, c_void_p]),
    "ct_cross_entropy_fwd_stats": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                           c_i64, c_i64, c_i64, c_i64, c_int, c_i64, c_void_p]),
    "ct_scale_by_scalar": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_void_p]),
    "ct_comm_init": (c_int, [c_int, c_int, c_int, ctypes.c_size_t, ctypes.POINTER(c_void_p), c_void_p, c_void_p]),
    "ct_comm_connect": (c_int, [c_void_p, c_void_p]),
    "ct_allreduce_bucket": (c_int, [c_i64, c_i64, c_float, c_int, c_int, c_void_p]),
    "ct_broadcast": (c_int, [c_i64, c_i64, c_int, c_void_p]),
    "ct_comm_barrier": (c_int, [c_void_p]),
    "ct_comm_vmm_supported": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "ct_comm_vmm_init": (c_int, [c_int, c_int, c_int, ctypes.c_size_t, ctypes.POINTER(c_void_p),
                                 ctypes.POINTER(c_int)]),
    "ct_comm_vmm_connect": (c_int, [ctypes.POINTER(c_int)]),
    "ct_comm_mc_create": (c_int, [ctypes.POINTER(c_int)]),
    "ct_comm_mc_import": (c_int, [c_int]),
    "ct_comm_mc_add_device": (c_int, []),
    "ct_comm_mc_bind": (c_int, []),
    "ct_comm_info": (c_int, [ctypes.POINTER(c_int)]),
    "ct_embedding_bwd_allranks": (c_int, [c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int,
                                          c_void_p]),
    "ct_comm_finalize": (c_int, []),
    "ct_kv_append": (c_int, [c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int,
                             c_int, c_int, c_int, c_void_p]),
    "ct_attn_decode": (c_int, [ctypes.POINTER(AttnArgs), c_void_p]),
    "ct_kv_append_dev": (c_int, [c_void_p, c_i64, c_i64, c_i64, c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int,
                                 c_int, c_void_p, c_int, c_void_p]),
    "ct_greedy_step": (c_int, [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_int, c_i64, c_void_p, c_i64,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ct_dropout": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_i64, c_float, ctypes.c_uint64,
                           ctypes.c_uint32, c_void_p]),
    "ct_gemm_wgrad_bias": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_i64,
                                   c_i64, c_i64, c_int, c_void_p]),
}


def load():
    """Load the shared library (once) and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libct_b200.so not found at %s — run `python -m cleantransformer_b200.build` "
                "(there is no CPU or eager fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    load().ct_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, last_error()))
