"""ctypes loaders for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package (it is the checker, never the product:
``thevc_b200`` does not import it and fails loudly without its CUDA library).

``oracle.lib()``  -> libhm_oracle.so, the C restatement of the reference arithmetic
                     (hm_oracle*.c; every function cites the reference file:line).
``oracle.ref()``  -> _ref/libhmref.so, the reference's OWN TLibCommon sources compiled by
                     oracle/Makefile plus the forwarding shim ref_shim.cpp; ``None`` when it
                     has not been built (it cannot be rebuilt where /root/reference is absent).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
vp = C.c_void_p
ci = C.c_int
cu = C.c_uint32


class CuGeom(C.Structure):
    _fields_ = [("pic_w", ci), ("pic_h", ci), ("cu_x", ci), ("cu_y", ci), ("max_cu", ci)]


class MeResult(C.Structure):
    _fields_ = [("mvx", ci), ("mvy", ci), ("sad", cu), ("n_sads", cu)]


class FracResult(C.Structure):
    _fields_ = [("halfx", ci), ("halfy", ci), ("qtrx", ci), ("qtry", ci), ("cost_half", cu), ("cost", cu)]


class EstBits(C.Structure):
    """estBitsSbacStruct (TComTrQuant.h:59-72) as 254 plain ints, same member order"""
    _fields_ = [("sig_cg", ci * 2 * 2), ("sig", ci * 2 * 42), ("last_x", ci * 32), ("last_y", ci * 32),
                ("greater_one", ci * 2 * 24), ("level_abs", ci * 2 * 6), ("block_cbp", ci * 2 * 15),
                ("block_root_cbp", ci * 2 * 4), ("scan_zigzag", ci * 2), ("scan_non_zigzag", ci * 2)]


class RdoqParam(C.Structure):
    _fields_ = [("log2_size", ci), ("is_luma", ci), ("scan_idx", ci), ("qp_per", ci), ("qp_rem", ci), ("bd", ci),
                ("cbf_ctx", ci), ("sign_hide", ci), ("use_arl", ci), ("lambda_", C.c_double)]


class QuantParam(C.Structure):
    _fields_ = [("qp_per", ci), ("qp_rem", ci), ("base_per", ci), ("is_intra_slice", ci),
                ("sign_hide", ci), ("use_arl", ci), ("bd", ci)]


def build(force: bool = False) -> None:
    """Compile libhm_oracle.so (always possible) and, when /root/reference is present,
    _ref/libhmref.so.  Building the checker is not using it."""
    so = os.path.join(HERE, "libhm_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("hm_oracle.c", "hm_oracle_me.c", "hm_oracle_tq.c", "hm_oracle_frame.c", "hm_oracle_rdoq.c", "hm_oracle_deblock.c", "hm_oracle_intra.c", "hm_oracle_hash.c", "hm_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "oracle"])
    ref_so = os.path.join(HERE, "_ref", "libhmref.so")
    shim = os.path.join(HERE, "ref_shim.cpp")
    if os.path.isdir("/root/reference/source") and (
            force or not os.path.exists(ref_so) or os.path.getmtime(shim) > os.path.getmtime(ref_so)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _ptr(a: np.ndarray, off: int = 0):
    """pointer to element `off` (may be negative relative to a view's start is NOT allowed;
    pass the base array and an absolute element offset)."""
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data + off * a.itemsize)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(HERE, "libhm_oracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    dist = [vp, ci, vp, ci, ci, ci]
    L.orc_sad.argtypes = dist + [ci, ci]; L.orc_sad.restype = cu
    L.orc_sad_generic.argtypes = dist + [ci]; L.orc_sad_generic.restype = cu
    L.orc_sse.argtypes = dist + [ci]; L.orc_sse.restype = cu
    L.orc_hads.argtypes = dist + [ci]; L.orc_hads.restype = cu
    L.orc_calc_had.argtypes = dist + [ci]; L.orc_calc_had.restype = cu
    L.orc_get_dist_part.argtypes = dist + [ci, ci]; L.orc_get_dist_part.restype = cu
    L.orc_mv_component_bits.argtypes = [ci]; L.orc_mv_component_bits.restype = cu
    L.orc_mv_bits.argtypes = [ci] * 5; L.orc_mv_bits.restype = cu
    L.orc_mv_cost.argtypes = [cu] + [ci] * 5; L.orc_mv_cost.restype = cu
    L.orc_lambda_motion_sad.argtypes = [C.c_double]; L.orc_lambda_motion_sad.restype = cu
    L.orc_filter_copy.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]; L.orc_filter_copy.restype = None
    L.orc_filter_hor_luma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]; L.orc_filter_hor_luma.restype = None
    L.orc_filter_ver_luma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci]; L.orc_filter_ver_luma.restype = None
    L.orc_filter_hor_chroma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]; L.orc_filter_hor_chroma.restype = None
    L.orc_filter_ver_chroma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci]; L.orc_filter_ver_chroma.restype = None
    mc = [vp, ci, ci, ci, ci, ci, vp, ci, ci, ci]
    L.orc_pred_inter_luma_blk.argtypes = mc; L.orc_pred_inter_luma_blk.restype = None
    L.orc_pred_inter_chroma_blk.argtypes = mc; L.orc_pred_inter_chroma_blk.restype = None
    L.orc_add_avg.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, ci]; L.orc_add_avg.restype = None
    L.orc_subtract.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci]; L.orc_subtract.restype = None
    L.orc_add_clip.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, ci]; L.orc_add_clip.restype = None
    L.orc_remove_high_freq.argtypes = [vp, ci, vp, ci, ci, ci]; L.orc_remove_high_freq.restype = None
    L.orc_extend_border.argtypes = [vp, ci, ci, ci, ci, ci]; L.orc_extend_border.restype = None
    gp = C.POINTER(CuGeom)
    L.orc_clip_mv.argtypes = [gp, C.POINTER(ci), C.POINTER(ci)]; L.orc_clip_mv.restype = None
    L.orc_set_search_range.argtypes = [gp, ci, ci, ci] + [C.POINTER(ci)] * 4; L.orc_set_search_range.restype = None
    L.orc_pattern_search.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, cu, ci, ci, C.POINTER(MeResult)]
    L.orc_pattern_search.restype = None
    L.orc_tz_search.argtypes = [gp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cu, ci, ci, ci, ci,
                                C.POINTER(MeResult)]
    L.orc_tz_search.restype = None
    L.orc_frac_search.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, cu, ci, ci, C.POINTER(FracResult)]
    L.orc_frac_search.restype = None
    L.orc_dct_matrix.argtypes = [ci, i16p]; L.orc_dct_matrix.restype = None
    L.orc_partial_butterfly.argtypes = [ci, i16p, i16p, ci, ci]; L.orc_partial_butterfly.restype = None
    L.orc_partial_butterfly_inverse.argtypes = [ci, i16p, i16p, ci, ci]; L.orc_partial_butterfly_inverse.restype = None
    L.orc_fast_forward_dst.argtypes = [i16p, i16p, ci]; L.orc_fast_forward_dst.restype = None
    L.orc_fast_inverse_dst.argtypes = [i16p, i16p, ci]; L.orc_fast_inverse_dst.restype = None
    L.orc_xTrMxN.argtypes = [i16p, i16p, ci, ci, ci, ci]; L.orc_xTrMxN.restype = None
    L.orc_xITrMxN.argtypes = [i16p, i16p, ci, ci, ci, ci]; L.orc_xITrMxN.restype = None
    L.orc_xT.argtypes = [ci, vp, ci, i32p, ci, ci, ci]; L.orc_xT.restype = None
    L.orc_xIT.argtypes = [ci, i32p, vp, ci, ci, ci, ci]; L.orc_xIT.restype = None
    L.orc_transform_skip.argtypes = [vp, ci, i32p, ci, ci, ci]; L.orc_transform_skip.restype = None
    L.orc_itransform_skip.argtypes = [i32p, vp, ci, ci, ci, ci]; L.orc_itransform_skip.restype = None
    L.orc_set_qp.argtypes = [ci, ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]; L.orc_set_qp.restype = None
    L.orc_scan.argtypes = [ci, ci, u32p]; L.orc_scan.restype = None
    L.orc_quant.argtypes = [i32p, i32p, vp, ci, ci, C.POINTER(QuantParam), u32p, C.POINTER(cu)]
    L.orc_quant.restype = None
    L.orc_dequant.argtypes = [i32p, i32p, ci, ci, ci, ci, ci]; L.orc_dequant.restype = None
    L.orc_rdoq.argtypes = [i32p, i32p, vp, C.POINTER(RdoqParam), C.POINTER(EstBits), u32p, C.POINTER(cu)]
    L.orc_rdoq.restype = None
    L.orc_rdoq_err_scale.argtypes = [ci, ci, ci]; L.orc_rdoq_err_scale.restype = C.c_double
    L.orc_deblock_pic.argtypes = [vp, ci, vp, vp, ci, ci, ci, vp, vp, ci, ci, ci]; L.orc_deblock_pic.restype = None
    L.orc_sao_plane.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, ci]; L.orc_sao_plane.restype = None
    L.orc_census.argtypes = [vp]; L.orc_census.restype = None
    L.orc_me_frame_ctu.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, cu, ci, ci, ci, ci, ci, vp, vp]
    L.orc_me_frame_ctu.restype = None
    L.orc_mc_batch.argtypes = [vp, ci, ci, vp, ci, vp, ci]; L.orc_mc_batch.restype = None
    L.orc_fwd_tq_batch.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, vp, vp]; L.orc_fwd_tq_batch.restype = None
    L.orc_fwd_rdoq_batch.argtypes = [vp, ci, ci, ci, vp, ci, ci, C.POINTER(EstBits), C.c_double, C.c_double, vp, vp]
    L.orc_fwd_rdoq_batch.restype = None
    L.orc_inv_tq_batch.argtypes = [vp, vp, vp, ci, ci, ci, vp, ci, vp]; L.orc_inv_tq_batch.restype = None
    L.orc_intra_filter_line.argtypes = [vp, ci, vp]; L.orc_intra_filter_line.restype = None
    L.orc_intra_mode_filtered.argtypes = [ci, ci]; L.orc_intra_mode_filtered.restype = ci
    L.orc_intra_pred_luma.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci]; L.orc_intra_pred_luma.restype = None
    L.orc_intra_rough.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp]; L.orc_intra_rough.restype = None
    for f in ("orc_md5_plane", "orc_crc_plane", "orc_checksum_plane"):
        getattr(L, f).argtypes = [vp, ci, ci, ci, ci, vp]; getattr(L, f).restype = None
    L.orc_ssd_plane.argtypes = [vp, ci, vp, ci, ci, ci]; L.orc_ssd_plane.restype = C.c_uint64
    _LIB = L
    return L


def ref():
    """The reference's own compiled code; None if oracle/_ref/libhmref.so is absent."""
    global _REF
    if _REF is not None:
        return _REF
    so = os.path.join(HERE, "_ref", "libhmref.so")
    if not os.path.exists(so):
        return None
    R = C.CDLL(so)
    R.ref_init.argtypes = [ci]; R.ref_init.restype = None
    R.ref_sad_me.argtypes = [vp, ci, vp, ci, ci, ci, ci]; R.ref_sad_me.restype = cu
    R.ref_dist_frac.argtypes = [vp, ci, vp, ci, ci, ci, ci]; R.ref_dist_frac.restype = cu
    R.ref_get_dist_part.argtypes = [vp, ci, vp, ci, ci, ci, ci]; R.ref_get_dist_part.restype = cu
    R.ref_calc_had.argtypes = [vp, ci, vp, ci, ci, ci]; R.ref_calc_had.restype = cu
    R.ref_component_bits.argtypes = [ci]; R.ref_component_bits.restype = cu
    R.ref_mv_cost.argtypes = [C.c_double, ci, ci, ci, ci, ci, C.POINTER(cu)]; R.ref_mv_cost.restype = cu
    R.ref_filter_hor_luma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci]; R.ref_filter_hor_luma.restype = None
    R.ref_filter_ver_luma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]; R.ref_filter_ver_luma.restype = None
    R.ref_filter_hor_chroma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci]; R.ref_filter_hor_chroma.restype = None
    R.ref_filter_ver_chroma.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci]; R.ref_filter_ver_chroma.restype = None
    R.ref_partial_butterfly.argtypes = [ci, i16p, i16p, ci, ci]; R.ref_partial_butterfly.restype = None
    R.ref_partial_butterfly_inverse.argtypes = [ci, i16p, i16p, ci, ci]; R.ref_partial_butterfly_inverse.restype = None
    R.ref_fast_forward_dst.argtypes = [i16p, i16p, ci]; R.ref_fast_forward_dst.restype = None
    R.ref_fast_inverse_dst.argtypes = [i16p, i16p, ci]; R.ref_fast_inverse_dst.restype = None
    R.ref_xTrMxN.argtypes = [i16p, i16p, ci, ci, ci]; R.ref_xTrMxN.restype = None
    R.ref_xITrMxN.argtypes = [i16p, i16p, ci, ci, ci]; R.ref_xITrMxN.restype = None
    R.ref_dct_matrix.argtypes = [ci, i16p]; R.ref_dct_matrix.restype = None
    R.ref_scan.argtypes = [ci, ci, u32p]; R.ref_scan.restype = None
    R.ref_xT.argtypes = [ci, vp, ci, i32p, ci, ci]; R.ref_xT.restype = None
    R.ref_xIT.argtypes = [ci, i32p, vp, ci, ci, ci]; R.ref_xIT.restype = None
    R.ref_transform_skip.argtypes = [vp, ci, i32p, ci, ci]; R.ref_transform_skip.restype = None
    R.ref_itransform_skip.argtypes = [i32p, vp, ci, ci, ci]; R.ref_itransform_skip.restype = None
    R.ref_set_qp.argtypes = [ci, ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]; R.ref_set_qp.restype = None
    R.ref_quant.argtypes = [i32p, i32p, i32p, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, C.POINTER(cu)]
    R.ref_quant.restype = None
    if hasattr(R, "ref_rdoq"):
        R.ref_rdoq.argtypes = [i32p, i32p, i32p, ci, ci, ci, ci, ci, ci, ci, ci, ci, C.c_double, C.POINTER(EstBits), C.POINTER(cu)]
        R.ref_rdoq.restype = None
        R.ref_est_bits_size.argtypes = []; R.ref_est_bits_size.restype = ci
    R.ref_dequant.argtypes = [i32p, i32p, ci, ci, ci, ci, ci]; R.ref_dequant.restype = None
    R.ref_extend_border.argtypes = [vp, ci, ci, ci, ci, ci]; R.ref_extend_border.restype = None
    if hasattr(R, "ref_pic_hash"):
        R.ref_pic_hash.argtypes = [ci, vp, vp, vp, ci, ci, vp]; R.ref_pic_hash.restype = None
    if hasattr(R, "ref_intra_rough"):
        R.ref_intra_rough.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp]; R.ref_intra_rough.restype = None
    ip = C.POINTER(ci)
    R.ref_me_setup.argtypes = [ci, ci, ci, ci, ci]; R.ref_me_setup.restype = None
    R.ref_set_search_range.argtypes = [ci, ci, ci, ci, ci, ip]; R.ref_set_search_range.restype = None
    R.ref_int_search.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, C.c_double, ci, ci, ci, ci, ip]
    R.ref_int_search.restype = None
    R.ref_frac_search.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, C.c_double, ci, ci, ip]; R.ref_frac_search.restype = None
    if hasattr(R, "ref_me_frame_ctu"):
        R.ref_me_frame_ctu.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, C.c_double, ci, vp, vp, vp]
        R.ref_me_frame_ctu.restype = None
    R.ref_mc_set_ref.argtypes = [vp, vp, vp, ci, ci]; R.ref_mc_set_ref.restype = None
    R.ref_mc_pu.argtypes = [ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]; R.ref_mc_pu.restype = None
    R.ref_add_avg.argtypes = [vp] * 6 + [ci, ci] + [vp] * 3; R.ref_add_avg.restype = None
    _REF = R
    return R


def ptr(a: np.ndarray, off: int = 0):
    return _ptr(a, off)
