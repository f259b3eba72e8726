"""ctypes binding of the C ABI in include/thevc_cuda.h (libthevc_cuda.so).

There is no CPU fallback: importing works anywhere (so that the symbol table can be checked on a
machine without a GPU) but creating a context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

TVC_OK = 0
DIST_SAD, DIST_SSE, DIST_HADS = 0, 1, 2
ME_FULL, ME_TZ = 0, 1
TU_DST, TU_SKIP, TU_BYPASS = 1, 2, 4
ME_RANGE = 64
ME_CAND = 129

ci = C.c_int
i32 = C.c_int32
u32 = C.c_uint32
vp = C.c_void_p


class Config(C.Structure):
    _fields_ = [("width", ci), ("height", ci), ("bit_depth", ci), ("max_cu", ci), ("num_slots", ci), ("device", ci)]


class DistJob(C.Structure):
    _fields_ = [("kind", i32), ("org_slot", i32), ("org_plane", i32), ("org_x", i32), ("org_y", i32),
                ("cur_slot", i32), ("cur_plane", i32), ("cur_x", i32), ("cur_y", i32),
                ("w", i32), ("h", i32), ("sub_shift", i32)]


class PU(C.Structure):
    _fields_ = [("x", i32), ("y", i32), ("w", i32), ("h", i32),
                ("ref_slot0", i32), ("mvx0", i32), ("mvy0", i32),
                ("ref_slot1", i32), ("mvx1", i32), ("mvy1", i32)]


class MeCenter(C.Structure):
    _fields_ = [("cx", i32), ("cy", i32)]


class MeJob(C.Structure):
    _fields_ = [("ref_index", i32), ("ref_slot", i32), ("x", i32), ("y", i32), ("w", i32), ("h", i32),
                ("mode", i32), ("fen", i32), ("search_range", i32),
                ("lx", i32), ("ty", i32), ("rx", i32), ("by", i32),
                ("predx", i32), ("predy", i32), ("startx", i32), ("starty", i32), ("lambda_cost", u32)]


class MeResult(C.Structure):
    _fields_ = [("mvx", i32), ("mvy", i32), ("sad", u32), ("n_sads", u32)]


class FracJob(C.Structure):
    _fields_ = [("ref_slot", i32), ("x", i32), ("y", i32), ("w", i32), ("h", i32), ("imvx", i32), ("imvy", i32),
                ("predx", i32), ("predy", i32), ("lambda_cost", u32), ("hadamard", i32)]


class FracResult(C.Structure):
    _fields_ = [("halfx", i32), ("halfy", i32), ("qtrx", i32), ("qtry", i32), ("cost_half", u32), ("cost", u32)]


class CensusPU(C.Structure):
    _fields_ = [("x", C.c_int16), ("y", C.c_int16), ("w", C.c_int16), ("h", C.c_int16), ("cu_x", C.c_int16), ("cu_y", C.c_int16)]


class MeFrameCfg(C.Structure):
    _fields_ = [("search_range", i32), ("fen", i32), ("hadamard", i32), ("use_tables", i32), ("do_frac", i32),
                ("lambda_cost", u32)]


ME_CENSUS = 593
# tvc_ubench selectors (include/thevc_cuda.h)
UB_VABSDIFF4, UB_IADD3, UB_IMAD, UB_LDS128, UB_DP2A, UB_HBM_WRITE = 0, 1, 2, 3, 4, 5

PHASES = ("me_tables", "me_search", "me_frac", "mc", "fwd_tq", "inv_tq", "other", "me_raster", "rdoq", "deblock", "intra")

# numpy views of the ABI structs (same layout) for bulk results
ME_RESULT_DTYPE = np.dtype([("mvx", "<i4"), ("mvy", "<i4"), ("sad", "<u4"), ("n_sads", "<u4")])
ME_PACKED_DTYPE = np.dtype([("mvx", "<i2"), ("mvy", "<i2"), ("halfx", "i1"), ("halfy", "i1"), ("qtrx", "i1"), ("qtry", "i1"), ("sad", "<u4"), ("cost", "<u4")])
INTRA_MODES = 35
GRID_WORDS = 289 + 289 + 81
GRID_JOB_DTYPE = np.dtype([("ref_slot", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("mvx", "<i4"), ("mvy", "<i4")])
INTRA_JOB_DTYPE = np.dtype([("log2_size", "<i4"), ("line_offset", "<i4"), ("org_offset", "<i4"), ("org_stride", "<i4"),
                            ("above", "<i4"), ("left", "<i4")])
FRAC_RESULT_DTYPE = np.dtype([("halfx", "<i4"), ("halfy", "<i4"), ("qtrx", "<i4"), ("qtry", "<i4"), ("cost_half", "<u4"), ("cost", "<u4")])
PU_DTYPE = np.dtype([(n, "<i4") for n in ("x", "y", "w", "h", "ref_slot0", "mvx0", "mvy0", "ref_slot1", "mvx1", "mvy1")])
TU_DTYPE = np.dtype([(n, "<i4") for n in ("plane", "x", "y", "log2_size", "flags", "scan_idx", "qp_per", "qp_rem", "base_per", "coef_offset")])


class TU(C.Structure):
    _fields_ = [("plane", i32), ("x", i32), ("y", i32), ("log2_size", i32), ("flags", i32), ("scan_idx", i32),
                ("qp_per", i32), ("qp_rem", i32), ("base_per", i32), ("coef_offset", i32)]


class QuantCfg(C.Structure):
    _fields_ = [("is_intra_slice", i32), ("sign_hide", i32), ("use_arl", i32)]


class EstBits(C.Structure):
    """tvc_est_bits = estBitsSbacStruct (TComTrQuant.h:59-72)"""
    _fields_ = [("sig_cg", i32 * 2 * 2), ("sig", i32 * 2 * 42), ("last_x", i32 * 32), ("last_y", i32 * 32),
                ("greater_one", i32 * 2 * 24), ("level_abs", i32 * 2 * 6), ("block_cbp", i32 * 2 * 15),
                ("block_root_cbp", i32 * 2 * 4), ("scan_zigzag", i32 * 2), ("scan_non_zigzag", i32 * 2)]


class RdoqTU(C.Structure):
    _fields_ = [("log2_size", i32), ("is_luma", i32), ("scan_idx", i32), ("qp_per", i32), ("qp_rem", i32),
                ("cbf_ctx", i32), ("est_index", i32), ("coef_offset", i32), ("lambda_", C.c_double)]


RDOQ_TU_DTYPE = np.dtype([(n, "<i4") for n in ("log2_size", "is_luma", "scan_idx", "qp_per", "qp_rem", "cbf_ctx", "est_index",
                                               "coef_offset")] + [("lambda_", "<f8")])

# every symbol include/thevc_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tvc_abi_version": (ci, []),
    "tvc_ctx_create": (ci, [C.POINTER(Config), C.POINTER(vp)]),
    "tvc_ctx_destroy": (None, [vp]),
    "tvc_ctx_set_stream": (ci, [vp, vp]),
    "tvc_sync": (ci, [vp]),
    "tvc_last_error": (C.c_char_p, [vp]),
    "tvc_launch_count": (C.c_uint64, [vp]),
    "tvc_pic_upload": (ci, [vp, ci, vp, ci, vp, vp, ci, ci]),
    "tvc_pic_download": (ci, [vp, ci, vp, ci, vp, vp, ci, ci]),
    "tvc_pic_extend_border": (ci, [vp, ci]),
    "tvc_pic_device_ptr": (ci, [vp, ci, ci, C.POINTER(vp), C.POINTER(ci)]),
    "tvc_pic_device_ptr_u8": (ci, [vp, ci, C.POINTER(vp), C.POINTER(ci)]),
    "tvc_pic_subtract": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci]),
    "tvc_pic_add_clip": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci]),
    "tvc_pic_remove_high_freq": (ci, [vp, ci, ci, ci, ci, ci, ci, ci]),
    "tvc_dist_block": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, C.POINTER(u32)]),
    "tvc_dist_batch": (ci, [vp, ci, vp, vp]),
    "tvc_dist_batch_dev": (ci, [vp, ci, vp, vp]),
    "tvc_filter_hor_luma": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, ci]),
    "tvc_filter_ver_luma": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, ci, ci]),
    "tvc_filter_hor_chroma": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, ci]),
    "tvc_filter_ver_chroma": (ci, [vp, vp, ci, vp, ci, ci, ci, ci, ci, ci]),
    "tvc_mc_batch": (ci, [vp, ci, ci, vp]),
    "tvc_mc_batch_dev": (ci, [vp, ci, ci, vp]),
    "tvc_mc_block": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp, vp, ci]),
    "tvc_me_prepass": (ci, [vp, ci, ci, vp, vp]),
    "tvc_me_reserve": (ci, [vp, ci]),
    "tvc_me_uses_tables": (ci, [vp]),
    "tvc_me_frame_packed": (ci, [vp, ci, ci, vp, vp, C.POINTER(MeFrameCfg), vp]),
    "tvc_me_bipred": (ci, [vp, ci, vp, ci, vp, ci, vp, vp]),
    "tvc_me_set_fused": (ci, [vp, ci]),
    "tvc_me_table_bytes": (C.c_size_t, [vp, ci]),
    "tvc_me_tables_dev": (ci, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "tvc_me_table_lookup": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp]),
    "tvc_me_search_batch": (ci, [vp, ci, ci, ci, vp, vp]),
    "tvc_me_search_batch_dev": (ci, [vp, ci, ci, ci, vp, vp]),
    "tvc_me_frac_batch": (ci, [vp, ci, ci, vp, vp]),
    "tvc_me_frac_batch_dev": (ci, [vp, ci, ci, vp, vp]),
    "tvc_me_census": (ci, [vp]),
    "tvc_me_frame": (ci, [vp, ci, ci, vp, vp, C.POINTER(MeFrameCfg), vp, vp]),
    "tvc_me_frame_dev": (ci, [vp, ci, ci, vp, vp, C.POINTER(MeFrameCfg), C.POINTER(vp), C.POINTER(vp)]),
    "tvc_me_ctu": (ci, [vp, ci, ci, ci, ci, MeCenter, C.POINTER(MeFrameCfg), vp, vp]),
    "tvc_me_ctu_async": (ci, [vp, ci, ci, ci, ci, ci, MeCenter, C.POINTER(MeFrameCfg)]),
    "tvc_me_ctu_fetch": (ci, [vp, ci, vp, vp]),
    "tvc_me_frame_stats": (ci, [vp, vp]),
    "tvc_fwd_transform_batch": (ci, [vp, ci, ci, vp, vp, C.c_size_t]),
    "tvc_fwd_tq_batch": (ci, [vp, ci, ci, vp, C.POINTER(QuantCfg), vp, vp, C.c_size_t, vp]),
    "tvc_inv_tq_batch": (ci, [vp, ci, ci, ci, ci, vp, vp, C.c_size_t]),
    "tvc_fwd_tq_batch_dev": (ci, [vp, ci, ci, vp, vp, C.POINTER(QuantCfg), vp, vp, vp]),
    "tvc_inv_tq_batch_dev": (ci, [vp, ci, ci, ci, ci, vp, vp, vp]),
    "tvc_rdoq_batch": (ci, [vp, ci, vp, ci, vp, C.POINTER(QuantCfg), vp, vp, vp, C.c_size_t, vp]),
    "tvc_rdoq_batch_dev": (ci, [vp, ci, vp, ci, vp, C.POINTER(QuantCfg), vp, vp, vp, C.c_size_t, vp]),
    "tvc_fwd_transform_batch_dev": (ci, [vp, ci, ci, vp, vp, vp]),
    "tvc_fwd_rdoq_batch": (ci, [vp, ci, ci, vp, vp, ci, vp, C.POINTER(QuantCfg), vp, vp, C.c_size_t, vp]),
    "tvc_fwd_rdoq_recon_batch": (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, C.POINTER(QuantCfg), vp, C.c_size_t, vp]),
    "tvc_fwd_rdoq_recon_batch16": (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, C.POINTER(QuantCfg), vp, C.c_size_t, vp]),
    "tvc_xRateDistOptQuant": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, C.c_double, vp, C.POINTER(u32)]),
    "tvc_xT": (ci, [vp, ci, vp, ci, vp, ci, ci]),
    "tvc_xIT": (ci, [vp, ci, vp, vp, ci, ci, ci]),
    "tvc_xDeQuant": (ci, [vp, vp, vp, ci, ci, ci, ci]),
    "tvc_deblock_pic": (ci, [vp, ci, vp, vp, ci, ci]),
    "tvc_sao_plane": (ci, [vp, ci, ci, ci, vp]),
    "tvc_pred_cost_batch": (ci, [vp, ci, ci, ci, vp, vp]),
    "tvc_pred_cost_batch_dev": (ci, [vp, ci, ci, ci, vp, vp]),
    "tvc_ctu_cost_grids": (ci, [vp, ci, ci, vp, vp]),
    "tvc_ctu_cost_grids_dev": (ci, [vp, ci, ci, vp, vp]),
    "tvc_pic_hash": (ci, [vp, ci, ci, vp]),
    "tvc_pic_ssd": (ci, [vp, ci, ci, vp]),
    "tvc_intra_rough_batch": (ci, [vp, ci, vp, vp, C.c_size_t, vp, C.c_size_t, vp]),
    "tvc_intra_rough_batch_dev": (ci, [vp, ci, vp, vp, vp, vp, vp, vp]),
    "tvc_intra_rough": (ci, [vp, ci, vp, vp, ci, ci, ci, vp, vp]),
    "tvc_prof_enable": (ci, [vp, ci]),
    "tvc_prof_read": (ci, [vp, vp, vp, ci]),
    "tvc_ubench": (ci, [vp, ci, C.POINTER(C.c_double)]),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """dlopen libthevc_cuda.so and bind every ABI symbol.  Fails loudly if the library is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError("libthevc_cuda.so is not built: run `python -m thevc_b200.build` (no CPU fallback exists)")
        _build.build()
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def ptr(a: np.ndarray, off: int = 0) -> C.c_void_p:
    """pointer to element `off` of a C-contiguous numpy array (off may address a view origin)."""
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data + off * a.itemsize)
