"""ctypes binding of libgsb.so (include/gsb.h).  Fails loudly when the library is missing:
there is no CPU fallback and no other backend."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libgsb.so"

GSB_OK, GSB_E_INVALID, GSB_E_CUDA, GSB_E_CAPACITY, GSB_E_UNSUPPORTED = 0, -1, -2, -3, -4
BIN_TWO_LEVEL, BIN_FLAT64 = 0, 1
MAX_VIEWS = 8
ABI_VERSION = 3
RAW_OPACITY, RAW_SCALE, RAW_ROTATION = 1, 2, 4


class GsbSettings(C.Structure):
    _fields_ = [("image_height", C.c_int32), ("image_width", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("sh_degree", C.c_int32), ("prefiltered", C.c_int32), ("debug", C.c_int32),
                ("raw_inputs", C.c_int32), ("forward_only", C.c_int32),
                ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
                ("campos", C.c_void_p), ("tanfov_dev", C.c_void_p)]


_LAYOUT_FIELDS = ["saved_bytes", "off_geom", "off_clamped", "off_counts", "off_point_list", "off_ranges",
                  "off_n_contrib", "off_final_T", "off_tile_order", "off_hit_count", "off_hits",
                  "saved_bytes_forward_only", "scratch_bytes", "off_rect", "off_tiles", "off_dkeys0",
                  "off_dkeys1", "off_dkeys2", "off_didx0", "off_didx1", "off_offsets", "off_blocksums",
                  "off_hist", "off_tkeys0", "off_tkeys1", "off_tvals_alt", "off_keys64_0", "off_keys64_1",
                  "off_ggrad"]


class GsbLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in _LAYOUT_FIELDS]


class GsbError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"libgsb {where} failed: {code} ({detail})")


_P = C.c_void_p
_PROTOS = {
    "gsb_abi_version": (C.c_int, []),
    "gsb_strerror": (C.c_char_p, [C.c_int]),
    "gsb_last_cuda_error": (C.c_char_p, []),
    "gsb_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_longlong, C.POINTER(GsbLayout)]),
    "gsb_preprocess_fwd": (C.c_int, [C.POINTER(GsbSettings), C.c_int, C.c_int] + [_P] * 7 + [_P, _P, _P, C.c_longlong, _P]),
    "gsb_bin_sort": (C.c_int, [C.POINTER(GsbSettings), C.c_int, _P, _P, C.c_longlong, C.c_int, _P, _P, _P]),
    "gsb_render_fwd": (C.c_int, [C.POINTER(GsbSettings), C.c_int, _P, C.c_longlong, _P, _P, _P, _P]),
    "gsb_forward": (C.c_int, [C.POINTER(GsbSettings), C.c_int, C.c_int] + [_P] * 7 + [_P] * 4 + [_P, _P, C.c_longlong, C.c_int, _P, _P, _P]),
    "gsb_read_counts": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_longlong, _P, _P]),
    "gsb_render_bwd": (C.c_int, [C.POINTER(GsbSettings), C.c_int, _P, _P, C.c_longlong, _P, _P, _P, _P]),
    "gsb_preprocess_bwd": (C.c_int, [C.POINTER(GsbSettings), C.c_int, C.c_int] + [_P] * 8 + [_P, _P, C.c_longlong] + [_P] * 8 + [C.c_int, _P]),
    "gsb_preprocess_bwd_views": (C.c_int, [C.c_int, _P, C.c_int, C.c_int] + [_P] * 7 + [_P, _P, _P, _P] + [_P] * 8 + [C.c_int, _P]),
    "gsb_backward": (C.c_int, [C.POINTER(GsbSettings), C.c_int, C.c_int] + [_P] * 8 + [_P, _P, C.c_longlong] + [_P] * 3 + [_P] * 8 + [C.c_int, _P]),
    "gsb_profile_enable": (C.c_int, [C.c_int]),
    "gsb_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "gsb_launch_count": (C.c_longlong, []),
    "gsb_adam_step": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_longlong,
                                C.c_float, C.c_longlong, _P, _P, _P, _P, _P, _P]),
    "gsb_adam_step_groups": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, _P,
                                       C.c_float, C.c_longlong, _P, _P, _P, _P, _P, _P]),
    "gsb_set_blend_variant": (C.c_int, [C.c_int]),
    "gsb_mark_visible": (C.c_int, [C.c_int, _P, _P, _P, _P, _P]),
    "gsb_exchange_config": (C.c_int, [C.c_int, C.c_int, C.c_longlong, _P]),
    "gsb_exchange_set_aux": (C.c_int, [_P, _P, _P]),
    "gsb_exchange_gather": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "gsb_mask_index_tmp_bytes": (C.c_size_t, [C.c_longlong]),
    "gsb_mask_to_index": (C.c_int, [C.c_longlong, _P, _P, _P, _P, _P]),
    "gsb_gather_rows": (C.c_int, [C.c_int, _P, _P, _P, C.c_longlong, _P, C.c_longlong, _P]),
    "gsb_knn_scratch_bytes": (C.c_size_t, [C.c_longlong]),
    "gsb_knn_dist2": (C.c_int, [C.c_longlong, _P, _P, _P, C.c_size_t, _P]),
    "gsb_debug_sorted_keys": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, C.c_longlong, _P, _P]),
    "gsb_radix_tmp_bytes": (C.c_size_t, [C.c_longlong, C.c_int]),
    "gsb_radix_sort_pairs_u32": (C.c_int, [C.c_longlong, _P, _P, _P, _P, C.c_int, _P, _P]),
    "gsb_radix_sort_pairs_u64": (C.c_int, [C.c_longlong, _P, _P, _P, _P, C.c_int, _P, _P]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def load() -> C.CDLL:
    """Load libgsb.so (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gaussianip_b200.build` "
            "(nvcc, sm_100a).  gaussianip_b200 has no CPU or non-CUDA fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)      # AttributeError if an include/gsb.h symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gsb_abi_version() != ABI_VERSION:
        raise RuntimeError("libgsb.so ABI version mismatch; rebuild")
    # A/B switch for measurements and for running the whole parity suite on the other forward blend kernel
    fv = os.environ.get("GSB_FWD_VARIANT")
    if fv:
        lib.gsb_set_blend_variant({"per_hit": 10, "transposed": 11, "gather4": 12, "precull": 13}[fv])
    _lib = lib
    return lib


def check(code: int, where: str) -> None:
    if code != GSB_OK:
        lib = load()
        detail = lib.gsb_strerror(code).decode()
        if code == GSB_E_CUDA:
            detail += ": " + lib.gsb_last_cuda_error().decode()
        raise GsbError(code, where, detail)


def layout(P: int, H: int, W: int, D_cap: int) -> GsbLayout:
    L = GsbLayout()
    check(load().gsb_layout(P, H, W, D_cap, C.byref(L)), "gsb_layout")
    return L


STAGES = ("preprocess_fwd", "depth_sort", "scan_emit", "tile_sort", "ranges", "render_fwd", "render_bwd",
          "preprocess_bwd")


def profile_enable(on: bool) -> None:
    check(load().gsb_profile_enable(int(on)), "gsb_profile_enable")


def profile_read():
    """{stage: (total device ms, calls)} accumulated since profile_enable(True)."""
    ms = (C.c_float * len(STAGES))()
    calls = (C.c_int * len(STAGES))()
    check(load().gsb_profile_read(ms, calls), "gsb_profile_read")
    return {s: (float(ms[i]), int(calls[i])) for i, s in enumerate(STAGES)}


def launch_count() -> int:
    return int(load().gsb_launch_count())
GATHER_MAX_TENSORS = 24
