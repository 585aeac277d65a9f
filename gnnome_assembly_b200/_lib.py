"""ctypes binding of libgnnome_b200.so (the C ABI in include/gnnome_b200.h).

The product path has NO fallback: if the library is missing, `lib()` raises.  `lib()` only loads the
shared object (works on a CPU-only box, used by the symbol-export test); compute entry points need a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(_HERE, "libgnnome_b200.so")     # GG_LIB: A/B builds of the same ABI

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64

# name -> (restype, argtypes); mirrors include/gnnome_b200.h one to one
SIGNATURES = {
    "gg_version": (_i, []),
    "gg_last_error": (C.c_char_p, []),
    "gg_set_tc_mode": (_i, [_i]),
    "gg_debug_flags": (_i, [_i]),
    "gg_debug_trace": (_i, [_p, _i, _i]),
    "gg_launch_count": (_i64, []),
    "gg_profile_enable": (_i, [_i]),
    "gg_profile_report": (_i, [C.c_char_p, C.c_size_t]),
    "gg_plan_create": (_i, [_p, _p, _i64, _i64, _p, C.POINTER(_p)]),
    "gg_plan_create_ex": (_i, [_p, _p, _i64, _i64, _i, _p, C.POINTER(_p)]),
    "gg_plan_destroy": (_i, [_p]),
    "gg_plan_num_nodes": (_i64, [_p]),
    "gg_plan_num_edges": (_i64, [_p]),
    "gg_plan_perm": (_p, [_p]),
    "gg_plan_inv_perm": (_p, [_p]),
    "gg_plan_src": (_p, [_p]),
    "gg_plan_dst": (_p, [_p]),
    "gg_plan_in_ptr": (_p, [_p]),
    "gg_plan_out_ptr": (_p, [_p]),
    "gg_plan_out_eid": (_p, [_p]),
    "gg_plan_copy_array": (_i, [_p, _i, _p, _p]),
    "gg_subplan_slab_words": (C.c_size_t, [_i64, _i64]),
    "gg_subplan_count": (_i, [_p, _p, _i64, _p, C.POINTER(_i64)]),
    "gg_subplan_fill": (_i, [_p, _p, _p, C.POINTER(_p)]),
    "gg_linear_fwd": (_i, [_i64, _i, _i, _p, _p, _p, _i, _p, _p]),
    "gg_linear_bwd_data": (_i, [_i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "gg_linear_bwd_weight": (_i, [_i64, _i, _i, _p, _p, _p, _p, _p]),
    "gg_layer_fwd": (_i, [_p, _i, _i, _i] + [_p] * 18),
    "gg_layer_bwd": (_i, [_p, _i, _i, _i] + [_p] * 32),
    "gg_score_fwd": (_i, [_p, _i, _i] + [_p] * 11),
    "gg_score_bwd": (_i, [_p, _i, _i] + [_p] * 18),
    "gg_model_workspace_floats": (_i64, [_p, _p, _i]),
    "gg_model_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "gg_model_bwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _p, _p]),
    "gg_prep_edge_features": (_i, [_i64, _p, _p, _p, _p, _p]),
    "gg_prep_pe": (_i, [_p, _i, C.c_double, _p, _p, _p]),
    "gg_bce_metrics_fwd": (_i, [_i64, _p, _p, C.c_float, _p, _p]),
    "gg_bce_bwd": (_i, [_i64, _p, _p, C.c_float, _p, _p, _p]),
    "gg_gather_rows": (_i, [_i64, _i, _p, _p, _p, _p]),
    "gg_edge_mlp_bwd": (_i, [_i64, _i, _i, _i] + [_p] * 9),
    "gg_edge_mlp_fwd": (_i, [_i64, _i, _i, _i] + [_p] * 8),
    "gg_decode_walks": (_i, [_i64] + [_p] * 10 + [_i] + [_p] * 10),
    "gg_decode_commit": (_i, [_i64, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    "gg_decode_edge_weights": (_i, [_i64, _p, _p, _p, _p, _p, _p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"gnnome_assembly_b200: {LIB_PATH} is missing — build it with "
                "`python -m gnnome_assembly_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)        # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        if os.environ.get("GG_DEBUG_FLAGS") is not None:  # experiment switches, see include/gnnome_b200.h
            handle.gg_debug_flags(int(os.environ["GG_DEBUG_FLAGS"]))
        if os.environ.get("GG_TC_MODE") is not None:      # 0 = FFMA everywhere, 1 = tcgen05 3xTF32 where eligible
            handle.gg_set_tc_mode(int(os.environ["GG_TC_MODE"]))
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().gg_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def launch_count():
    return int(lib().gg_launch_count())


def profile(on):
    check(lib().gg_profile_enable(1 if on else 0), "gg_profile_enable")


def profile_report():
    """{kernel name: (launches, total_ms)} for the launches made while profiling was on."""
    import json
    buf = C.create_string_buffer(1 << 16)
    check(lib().gg_profile_report(buf, len(buf)), "gg_profile_report")
    return {k: (int(v[0]), float(v[1])) for k, v in json.loads(buf.value.decode()).items()}


def set_tc_mode(mode):
    """1 (default): tcgen05 3xTF32 GEMM where eligible; 0: FFMA GEMM everywhere.  Returns the previous mode."""
    return int(lib().gg_set_tc_mode(int(mode)))
