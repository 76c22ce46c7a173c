"""ctypes binding of libasn_b200.so (the C ABI declared in include/asn_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared
object is missing, or a call fails, this module raises -- loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libasn_b200.so")

c_void_p, c_int, c_int64, c_size_t, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
PP = C.POINTER(c_void_p)

# name -> (restype, argtypes); mirrors include/asn_b200.h one to one
SIGNATURES = {
    "asn_abi_version": (c_int, []),
    "asn_build_id": (C.c_char_p, []),
    "asn_last_error": (C.c_char_p, []),
    "asn_sm_count": (c_int, [C.POINTER(c_int)]),
    "asn_launch_count": (c_int64, []),
    "asn_prof_enable": (c_int, [c_int]),
    "asn_prof_report": (c_int64, [C.c_char_p, c_int64]),
    "asn_prof_sequence": (c_int64, [C.c_char_p, c_int64]),
    "asn_fast_hist": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "asn_fast_hist_lut": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "asn_upsample_bilinear_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "asn_upsample_bwd_workspace_bytes": (c_size_t, [c_int] * 6),
    "asn_upsample_bilinear_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                          c_void_p, c_size_t, c_void_p]),
    "asn_upsample_argmax_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "asn_upsample2_argmax_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                        c_void_p]),
    "asn_upsample2_argmax_hist": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                          c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "asn_per_class_iu": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "asn_image_u8_to_bgr_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    "asn_label_u8_to_trainid_i64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "asn_softmax_ce_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p]),
    "asn_softmax_ce_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "asn_softmax_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "asn_softmax_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "asn_gan_loss_fwd_bwd": (c_int, [c_void_p, c_int64, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "asn_conv2d_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_float, c_int, c_void_p]),
    "asn_conv2d_dgrad_f32": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_int, c_void_p]),
    "asn_conv2d_wgrad_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "asn_lrelu_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "asn_aspp_np": (c_int, [c_int, c_int]),
    "asn_aspp_pack_weights": (c_int, [PP, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "asn_aspp_workspace_bytes": (c_size_t, [c_int] * 6),
    "asn_aspp_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             C.POINTER(c_int), c_int, c_void_p, c_size_t, c_void_p]),
    "asn_aspp_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, PP, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             C.POINTER(c_int), c_int, c_void_p, c_size_t, c_void_p]),
    "asn_fcd_wpack_bytes": (c_size_t, [c_int, c_int]),
    "asn_fcd_pack_weights": (c_int, [PP, c_int, c_int, c_void_p, c_void_p]),
    "asn_fcd_acts_bytes": (c_size_t, [c_int] * 5),
    "asn_fcd_workspace_bytes": (c_size_t, [c_int] * 5),
    "asn_fcd_act_layout": (c_int, [c_int] * 5 + [C.POINTER(c_int64)]),
    "asn_fcd_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_size_t, c_void_p]),
    "asn_fcd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, PP, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_size_t, c_void_p]),
    "asn_upsample_ce_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "asn_upsample_ce_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "asn_upsample_ce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "asn_fcd_workspace_bytes_lowres": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "asn_fcd_fwd_lowres": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_size_t, c_void_p]),
    "asn_fcd_bwd_lowres": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, PP, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "asn_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                             C.POINTER(c_float), c_int, c_float, c_float, c_int, c_float, c_void_p]),
    "asn_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                              c_int64, c_float, c_void_p]),
    "asn_gemm_bf16_tn": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "asn_gemm_bf16_nt_mn": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
}

ABI_VERSION = 2   # include/asn_b200.h: ASN_ABI_VERSION
_lib = None


class AsnError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library (once) and attach prototypes; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AsnError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(adaptsegnet_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.asn_abi_version() != ABI_VERSION:
        raise AsnError(f"ABI version mismatch: library reports {lib.asn_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().asn_last_error().decode("utf-8", "replace")
        raise AsnError(f"{what} failed (code {rc}): {msg}")


def ptr_array(ptrs) -> "C.Array":
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def int_array(vals) -> "C.Array":
    return (c_int * len(vals))(*vals)
