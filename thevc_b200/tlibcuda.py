"""Object wrapper over the C ABI: the calls a Python user of TLibCuda makes.

All compute goes through libthevc_cuda.so; a missing library or a missing GPU raises ``TvcError``
(no fallback).  Host pictures mirror the reference's TComPicYuv layout (planar int16, margin
max_cu+16 around luma; TComPicYuv.cpp:71-127).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import (CensusPU, Config, DistJob, EstBits, FracJob, FracResult, MeCenter, MeFrameCfg, MeJob, MeResult, PU, QuantCfg,
                   RdoqTU, TU, ptr)


class TvcError(RuntimeError):
    pass


class HostPic:
    """Host-side TComPicYuv: three padded int16 planes.  ``y[r, c]`` views exclude the margin."""

    def __init__(self, width: int, height: int, max_cu: int = 64, alloc=None):
        """alloc(shape) -> zeroed C-contiguous int16 array; pass a pinned-memory allocator for
        asynchronous uploads (default: numpy)"""
        alloc = alloc or (lambda shape: np.zeros(shape, np.int16))
        self.w, self.h = width, height
        self.mx = self.my = max_cu + 16
        self.cmx = self.cmy = self.mx >> 1
        self.stride = width + 2 * self.mx
        self.cstride = (width >> 1) + 2 * self.cmx
        self.buf_y = alloc((height + 2 * self.my, self.stride))
        self.buf_u = alloc(((height >> 1) + 2 * self.cmy, self.cstride))
        self.buf_v = alloc(((height >> 1) + 2 * self.cmy, self.cstride))

    @property
    def y(self):
        return self.buf_y[self.my:self.my + self.h, self.mx:self.mx + self.w]

    @property
    def u(self):
        return self.buf_u[self.cmy:self.cmy + (self.h >> 1), self.cmx:self.cmx + (self.w >> 1)]

    @property
    def v(self):
        return self.buf_v[self.cmy:self.cmy + (self.h >> 1), self.cmx:self.cmx + (self.w >> 1)]

    def origin(self, plane: int) -> int:
        """element offset of pel (0,0) inside the plane buffer"""
        return (self.my * self.stride + self.mx) if plane == 0 else (self.cmy * self.cstride + self.cmx)

    def plane(self, p: int) -> np.ndarray:
        return (self.buf_y, self.buf_u, self.buf_v)[p]

    def extend_border(self) -> None:
        """TComPicYuv::extendPicBorder (edge replication into the margin), host-side."""
        for buf, w, h, mx, my in ((self.buf_y, self.w, self.h, self.mx, self.my),
                                  (self.buf_u, self.w >> 1, self.h >> 1, self.cmx, self.cmy),
                                  (self.buf_v, self.w >> 1, self.h >> 1, self.cmx, self.cmy)):
            core = buf[my:my + h, mx:mx + w]
            buf[my:my + h, :mx] = core[:, :1]
            buf[my:my + h, mx + w:] = core[:, -1:]
            buf[:my, :] = buf[my:my + 1, :]
            buf[my + h:, :] = buf[my + h - 1:my + h, :]


def _arr(struct_t, items: Sequence):
    a = (struct_t * len(items))()
    for i, it in enumerate(items):
        a[i] = it
    return a


class TLibCuda:
    """One context = one encoder/decoder instance on one GPU (the reference is one instance per
    process; multi-GPU = one process per GPU)."""

    def __init__(self, width: int, height: int, bit_depth: int = 8, num_slots: int = 8, device: int = 0,
                 max_cu: int = 64, stream: Optional[int] = None):
        self.L = capi.load()
        self.cfg = Config(width, height, bit_depth, max_cu, num_slots, device)
        h = C.c_void_p()
        rc = self.L.tvc_ctx_create(C.byref(self.cfg), C.byref(h))
        if rc != 0 or not h.value:
            raise TvcError("tvc_ctx_create failed (rc=%d): no CUDA device / out of memory -- there is no CPU fallback" % rc)
        self.h = h
        self.width, self.height, self.bit_depth = width, height, bit_depth
        self.ctus_x = (width + max_cu - 1) // max_cu
        self.ctus_y = (height + max_cu - 1) // max_cu
        if stream is not None:
            self._ck(self.L.tvc_ctx_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.tvc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise TvcError("libthevc_cuda error %d: %s" % (rc, self.L.tvc_last_error(self.h).decode()))

    # ------------------------------------------------------------------ pictures
    def upload(self, slot: int, pic: HostPic, with_margin: bool = True):
        self._ck(self.L.tvc_pic_upload(self.h, slot, ptr(pic.buf_y, pic.origin(0)), pic.stride,
                                       ptr(pic.buf_u, pic.origin(1)), ptr(pic.buf_v, pic.origin(1)), pic.cstride,
                                       1 if with_margin else 0))

    def download(self, slot: int, with_margin: bool = True, into: Optional[HostPic] = None) -> HostPic:
        pic = into if into is not None else HostPic(self.width, self.height, self.cfg.max_cu)
        self._ck(self.L.tvc_pic_download(self.h, slot, ptr(pic.buf_y, pic.origin(0)), pic.stride,
                                         ptr(pic.buf_u, pic.origin(1)), ptr(pic.buf_v, pic.origin(1)), pic.cstride,
                                         1 if with_margin else 0))
        return pic

    def extend_border(self, slot: int):
        self._ck(self.L.tvc_pic_extend_border(self.h, slot))

    def sync(self):
        self._ck(self.L.tvc_sync(self.h))

    def launch_count(self) -> int:
        return int(self.L.tvc_launch_count(self.h))

    def subtract(self, dst, a, b, plane, x, y, w, h):
        self._ck(self.L.tvc_pic_subtract(self.h, dst, a, b, plane, x, y, w, h))

    def add_clip(self, dst, a, b, plane, x, y, w, h):
        self._ck(self.L.tvc_pic_add_clip(self.h, dst, a, b, plane, x, y, w, h))

    def remove_high_freq(self, dst, a, plane, x, y, w, h):
        self._ck(self.L.tvc_pic_remove_high_freq(self.h, dst, a, plane, x, y, w, h))

    # ------------------------------------------------------------------ distortion
    def dist_block(self, kind: int, org: np.ndarray, org_off: int, so: int, cur: np.ndarray, cur_off: int, sc: int,
                   w: int, h: int, sub_shift: int = 0) -> int:
        out = C.c_uint32()
        self._ck(self.L.tvc_dist_block(self.h, kind, ptr(org, org_off), so, ptr(cur, cur_off), sc, w, h, sub_shift, C.byref(out)))
        return out.value

    def dist_batch(self, jobs: Sequence[DistJob]) -> np.ndarray:
        n = len(jobs)
        out = np.zeros(n, np.uint32)
        if n:
            arr = _arr(DistJob, jobs)
            self._ck(self.L.tvc_dist_batch(self.h, n, C.cast(arr, C.c_void_p), ptr(out)))
        return out

    # ------------------------------------------------------------------ interpolation drop-ins
    def filter_hor_luma(self, src, src_off, ss, dst, dst_off, ds, w, h, frac, is_last):
        self._ck(self.L.tvc_filter_hor_luma(self.h, ptr(src, src_off), ss, ptr(dst, dst_off), ds, w, h, frac, int(is_last)))

    def filter_ver_luma(self, src, src_off, ss, dst, dst_off, ds, w, h, frac, is_first, is_last):
        self._ck(self.L.tvc_filter_ver_luma(self.h, ptr(src, src_off), ss, ptr(dst, dst_off), ds, w, h, frac, int(is_first), int(is_last)))

    def filter_hor_chroma(self, src, src_off, ss, dst, dst_off, ds, w, h, frac, is_last):
        self._ck(self.L.tvc_filter_hor_chroma(self.h, ptr(src, src_off), ss, ptr(dst, dst_off), ds, w, h, frac, int(is_last)))

    def filter_ver_chroma(self, src, src_off, ss, dst, dst_off, ds, w, h, frac, is_first, is_last):
        self._ck(self.L.tvc_filter_ver_chroma(self.h, ptr(src, src_off), ss, ptr(dst, dst_off), ds, w, h, frac, int(is_first), int(is_last)))

    # ------------------------------------------------------------------ MC
    def mc_batch(self, dst_slot: int, pus: Sequence[PU]):
        if len(pus):
            arr = _arr(PU, pus)
            self._ck(self.L.tvc_mc_batch(self.h, dst_slot, len(pus), C.cast(arr, C.c_void_p)))

    def pred_cost_batch(self, cur_slot: int, kind: int, pus: Sequence[PU]) -> np.ndarray:
        """luma prediction + distortion (capi.DIST_SAD / DIST_HADS) of candidate motions against the original of cur_slot:
        xGetInterPredictionError per merge candidate / xGetTemplateCost's prediction + SAD per AMVP candidate"""
        out = np.zeros(len(pus), np.uint32)
        if len(pus):
            arr = _arr(PU, pus)
            self._ck(self.L.tvc_pred_cost_batch(self.h, cur_slot, kind, len(pus), C.cast(arr, C.c_void_p), ptr(out)))
        return out

    def ctu_cost_grids(self, cur_slot: int, jobs: np.ndarray):
        """prefix-sum cost grids of (CTU, reference, MV) jobs (capi.GRID_JOB_DTYPE): (I_sad4 [n,17,17], I_had4 [n,17,17], I_had8 [n,9,9])"""
        jobs = np.ascontiguousarray(jobs, capi.GRID_JOB_DTYPE)
        out = np.zeros((len(jobs), capi.GRID_WORDS), np.uint32)
        self._ck(self.L.tvc_ctu_cost_grids(self.h, cur_slot, len(jobs), ptr(jobs), ptr(out)))
        return out[:, :289].reshape(-1, 17, 17), out[:, 289:578].reshape(-1, 17, 17), out[:, 578:].reshape(-1, 9, 9)

    def mc_block(self, ref_slot: int, x: int, y: int, w: int, h: int, mvx: int, mvy: int, bi: bool):
        """one PU from one reference into dense arrays (Y w*h, U/V (w/2)*(h/2)); bi keeps 14-bit intermediates"""
        oy = np.zeros((h, w), np.int16); ou = np.zeros((h // 2, w // 2), np.int16); ov = np.zeros_like(ou)
        self._ck(self.L.tvc_mc_block(self.h, ref_slot, x, y, w, h, mvx, mvy, int(bi), ptr(oy), w, ptr(ou), ptr(ov), w // 2))
        return oy, ou, ov

    # ------------------------------------------------------------------ ME
    def me_prepass(self, cur_slot: int, ref_slots: Sequence[int], centers: Optional[np.ndarray] = None):
        """centers: int32 array [num_refs, num_ctus, 2] (cx, cy) or None"""
        refs = (C.c_int * len(ref_slots))(*ref_slots)
        cp = None
        if centers is not None:
            centers = np.ascontiguousarray(centers, np.int32)
            assert centers.shape == (len(ref_slots), self.ctus_x * self.ctus_y, 2)
            cp = ptr(centers)
        self._ck(self.L.tvc_me_prepass(self.h, cur_slot, len(ref_slots), C.cast(refs, C.c_void_p), cp))

    def me_table_bytes(self, num_refs: int) -> int:
        return int(self.L.tvc_me_table_bytes(self.h, num_refs))

    def me_table_lookup(self, ref_index, pu_x, pu_y, pu_w, pu_h, fen, cand_xy: np.ndarray) -> np.ndarray:
        cand = np.ascontiguousarray(cand_xy, np.int16).reshape(-1, 2)
        out = np.zeros(len(cand), np.uint32)
        self._ck(self.L.tvc_me_table_lookup(self.h, ref_index, pu_x, pu_y, pu_w, pu_h, int(fen), len(cand), ptr(cand), ptr(out)))
        return out

    def me_search_batch(self, cur_slot: int, jobs: Sequence[MeJob], use_tables: bool):
        n = len(jobs)
        res = (MeResult * n)()
        if n:
            arr = _arr(MeJob, jobs)
            self._ck(self.L.tvc_me_search_batch(self.h, cur_slot, int(use_tables), n, C.cast(arr, C.c_void_p), C.cast(res, C.c_void_p)))
        return list(res)

    def me_bipred(self, target_slot: int, target: np.ndarray, job: MeJob, hadamard: bool = True):
        """target: int16 [h, w] block (2 * org - pred of the other list); returns (MeResult, FracResult)"""
        blk = np.ascontiguousarray(target, np.int16)
        ri, rf = MeResult(), FracResult()
        self._ck(self.L.tvc_me_bipred(self.h, target_slot, ptr(blk), blk.shape[1], C.byref(job), int(hadamard), C.byref(ri), C.byref(rf)))
        return ri, rf

    def me_frac_batch(self, cur_slot: int, jobs: Sequence[FracJob]):
        n = len(jobs)
        res = (FracResult * n)()
        if n:
            arr = _arr(FracJob, jobs)
            self._ck(self.L.tvc_me_frac_batch(self.h, cur_slot, n, C.cast(arr, C.c_void_p), C.cast(res, C.c_void_p)))
        return list(res)

    def me_census(self) -> np.ndarray:
        """the 593 PU rectangles of a CTU in result order: int16 [593, 6] = x, y, w, h, cu_x, cu_y"""
        out = np.zeros((capi.ME_CENSUS, 6), np.int16)
        self._ck(self.L.tvc_me_census(ptr(out)))
        return out

    def me_frame(self, cur_slot: int, ref_slots: Sequence[int], pred_qpel: Optional[np.ndarray], lambda_cost: int,
                 search_range: int = 64, fen: bool = True, hadamard: bool = True, use_tables: bool = True,
                 do_frac: bool = True):
        """frame pre-pass; returns (int_results, frac_results) as structured arrays [num_refs, num_ctus, 593]"""
        nr, nctu = len(ref_slots), self.ctus_x * self.ctus_y
        refs = (C.c_int * nr)(*ref_slots)
        pp = None
        if pred_qpel is not None:
            pred_qpel = np.ascontiguousarray(pred_qpel, np.int32)
            assert pred_qpel.shape == (nr, nctu, 2)
            pp = ptr(pred_qpel)
        cfg = MeFrameCfg(search_range, int(fen), int(hadamard), int(use_tables), int(do_frac), lambda_cost)
        ires = np.zeros((nr, nctu, capi.ME_CENSUS), capi.ME_RESULT_DTYPE)
        fres = np.zeros((nr, nctu, capi.ME_CENSUS), capi.FRAC_RESULT_DTYPE)
        self._ck(self.L.tvc_me_frame(self.h, cur_slot, nr, C.cast(refs, C.c_void_p), pp, C.byref(cfg), ptr(ires),
                                     ptr(fres) if do_frac else None))
        return ires, fres

    def me_ctu(self, cur_slot: int, ref_index: int, ref_slot: int, ctu: int, pred_qpel, lambda_cost: int, search_range: int = 64,
               fen: bool = True, hadamard: bool = True, use_tables: bool = True):
        """census-wide integer + fractional search of one (CTU, reference) group with an explicit predictor"""
        ires = np.zeros(capi.ME_CENSUS, capi.ME_RESULT_DTYPE)
        fres = np.zeros(capi.ME_CENSUS, capi.FRAC_RESULT_DTYPE)
        cfg = MeFrameCfg(search_range, int(fen), int(hadamard), int(use_tables), 1, lambda_cost)
        self._ck(self.L.tvc_me_ctu(self.h, cur_slot, ref_index, ref_slot, ctu, MeCenter(int(pred_qpel[0]), int(pred_qpel[1])),
                                   C.byref(cfg), ptr(ires), ptr(fres)))
        return ires, fres

    def me_ctu_async(self, ticket: int, cur_slot: int, ref_index: int, ref_slot: int, ctu: int, pred_qpel, lambda_cost: int,
                     search_range: int = 64, fen: bool = True, hadamard: bool = True, use_tables: bool = True):
        cfg = MeFrameCfg(search_range, int(fen), int(hadamard), int(use_tables), 1, lambda_cost)
        self._ck(self.L.tvc_me_ctu_async(self.h, ticket, cur_slot, ref_index, ref_slot, ctu, MeCenter(int(pred_qpel[0]), int(pred_qpel[1])),
                                         C.byref(cfg)))

    def me_ctu_fetch(self, ticket: int):
        ires = np.zeros(capi.ME_CENSUS, capi.ME_RESULT_DTYPE)
        fres = np.zeros(capi.ME_CENSUS, capi.FRAC_RESULT_DTYPE)
        self._ck(self.L.tvc_me_ctu_fetch(self.h, ticket, ptr(ires), ptr(fres)))
        return ires, fres

    def me_frame_stats(self):
        st = np.zeros(3, np.uint64)
        self._ck(self.L.tvc_me_frame_stats(self.h, ptr(st)))
        if self.L.tvc_me_uses_tables(self.h):      # TVC_ME_FUSED=0: SAD tables in HBM
            return {"form": "sad-tables", "search_granules": int(st[0]), "raster_served_candidates": int(st[1]), "raster_candidates": int(st[2])}
        return {"form": "group-search", "candidate_sads": int(st[0]), "sample_differences": int(st[1]), "frac_jobs_served_at_cu_level": int(st[2])}

    # ------------------------------------------------------------------ TQ
    def fwd_transform_batch(self, resi_slot: int, tus: Sequence[TU], coef_elems: int) -> np.ndarray:
        coef = np.zeros(coef_elems, np.int32)
        arr = _arr(TU, tus)
        self._ck(self.L.tvc_fwd_transform_batch(self.h, resi_slot, len(tus), C.cast(arr, C.c_void_p), ptr(coef), coef_elems))
        return coef

    def fwd_tq_batch(self, resi_slot: int, tus: Sequence[TU], qc: QuantCfg, coef_elems: int, want_arl: bool = False):
        lev = np.zeros(coef_elems, np.int32)
        arl = np.zeros(coef_elems, np.int32) if want_arl else None
        abs_sum = np.zeros(len(tus), np.uint32)
        arr = _arr(TU, tus)
        self._ck(self.L.tvc_fwd_tq_batch(self.h, resi_slot, len(tus), C.cast(arr, C.c_void_p), C.byref(qc), ptr(lev),
                                         ptr(arl) if want_arl else None, coef_elems, ptr(abs_sum)))
        return lev, arl, abs_sum

    def inv_tq_batch(self, resi_slot: int, pred_slot: int, recon_slot: int, tus: Sequence[TU], levels: np.ndarray):
        levels = np.ascontiguousarray(levels, np.int32)
        arr = _arr(TU, tus)
        self._ck(self.L.tvc_inv_tq_batch(self.h, resi_slot, pred_slot, recon_slot, len(tus), C.cast(arr, C.c_void_p),
                                         ptr(levels), levels.size))

    def rdoq_batch(self, tus: Sequence[RdoqTU], est: Sequence[EstBits], qc: QuantCfg, coef: np.ndarray):
        """xRateDistOptQuant over a TU list: (levels, arl or None, abs_sum)"""
        coef = np.ascontiguousarray(coef, np.int32)
        lev = np.zeros(coef.size, np.int32)
        arl = np.zeros(coef.size, np.int32) if qc.use_arl else None
        abs_sum = np.zeros(len(tus), np.uint32)
        ta, ea = _arr(RdoqTU, tus), _arr(EstBits, est)
        self._ck(self.L.tvc_rdoq_batch(self.h, len(tus), C.cast(ta, C.c_void_p), len(est), C.cast(ea, C.c_void_p), C.byref(qc),
                                       ptr(coef), ptr(lev), ptr(arl) if arl is not None else None, coef.size, ptr(abs_sum)))
        return lev, arl, abs_sum

    def fwd_rdoq_batch(self, resi_slot: int, tus: Sequence[TU], rtus: Sequence[RdoqTU], est: Sequence[EstBits], qc: QuantCfg,
                       coef_elems: int):
        """transformNxN with RDOQ on (xT + xRateDistOptQuant), coefficients stay on the device"""
        lev = np.zeros(coef_elems, np.int32)
        arl = np.zeros(coef_elems, np.int32) if qc.use_arl else None
        abs_sum = np.zeros(len(tus), np.uint32)
        ta, ra, ea = _arr(TU, tus), _arr(RdoqTU, rtus), _arr(EstBits, est)
        self._ck(self.L.tvc_fwd_rdoq_batch(self.h, resi_slot, len(tus), C.cast(ta, C.c_void_p), C.cast(ra, C.c_void_p), len(est),
                                           C.cast(ea, C.c_void_p), C.byref(qc), ptr(lev), ptr(arl) if arl is not None else None,
                                           coef_elems, ptr(abs_sum)))
        return lev, arl, abs_sum

    def xRateDistOptQuant(self, coef: np.ndarray, n: int, is_luma: int, scan_idx: int, per: int, rem: int, cbf_ctx: int,
                          sign_hide: int, use_arl: int, lam: float, est: EstBits):
        coef = np.ascontiguousarray(coef, np.int32)
        q = np.zeros(n * n, np.int32)
        arl = np.zeros(n * n, np.int32)
        s = C.c_uint32(0)
        self._ck(self.L.tvc_xRateDistOptQuant(self.h, ptr(coef), ptr(q), ptr(arl), n, is_luma, scan_idx, per, rem, cbf_ctx,
                                              sign_hide, use_arl, lam, C.byref(est), C.byref(s)))
        return q, arl, s.value

    def xT(self, use_dst: int, resi: np.ndarray, off: int, stride: int, n: int) -> np.ndarray:
        coef = np.zeros(n * n, np.int32)
        self._ck(self.L.tvc_xT(self.h, use_dst, ptr(resi, off), stride, ptr(coef), n, n))
        return coef

    def xIT(self, use_dst: int, coef: np.ndarray, resi: np.ndarray, off: int, stride: int, n: int):
        coef = np.ascontiguousarray(coef, np.int32)
        self._ck(self.L.tvc_xIT(self.h, use_dst, ptr(coef), ptr(resi, off), stride, n, n))

    def xDeQuant(self, q: np.ndarray, n: int, per: int, rem: int) -> np.ndarray:
        q = np.ascontiguousarray(q, np.int32)
        out = np.zeros(n * n, np.int32)
        self._ck(self.L.tvc_xDeQuant(self.h, ptr(q), ptr(out), n, n, per, rem))
        return out

    def deblock_pic(self, slot: int, ver: Optional[np.ndarray], hor: Optional[np.ndarray], beta_offset_div2: int = 0,
                    tc_offset_div2: int = 0):
        """TComLoopFilter::loopFilterPic on a picture slot, in place; ver / hor: uint8 arrays [units, 4] (bs, qp, flags, 0)"""
        pv = ptr(np.ascontiguousarray(ver, np.uint8)) if ver is not None else None
        ph = ptr(np.ascontiguousarray(hor, np.uint8)) if hor is not None else None
        self._ck(self.L.tvc_deblock_pic(self.h, slot, pv, ph, beta_offset_div2, tc_offset_div2))

    def sao_plane(self, src_slot: int, dst_slot: int, plane: int, units: np.ndarray):
        """SAO apply of one component; units: int16 [num_ctus, 38] = (type, eo[5], bo[32]) per CTU"""
        units = np.ascontiguousarray(units, np.int16)
        assert units.shape == (self.ctus_x * self.ctus_y, 38)
        self._ck(self.L.tvc_sao_plane(self.h, src_slot, dst_slot, plane, ptr(units)))

    def intra_rough(self, log2_size: int, line: np.ndarray, org: np.ndarray, above: bool = True, left: bool = True,
                    want_preds: bool = False):
        """35-mode rough search of one N x N luma PU (estIntraPredQT, TEncSearch.cpp:2530-2537): SATD per mode
        [35] (+ the 35 predictions [35, N, N]); line: 4N+1 unfiltered reference samples, org: [N, N] (any row stride)"""
        n = 1 << log2_size
        line = np.ascontiguousarray(line, np.int16)
        assert line.shape == (4 * n + 1,) and org.dtype == np.int16 and org.shape == (n, n) and org.strides[1] == 2
        sad = np.zeros(capi.INTRA_MODES, np.uint32)
        preds = np.zeros((capi.INTRA_MODES, n, n), np.int16) if want_preds else None
        self._ck(self.L.tvc_intra_rough(self.h, log2_size, ptr(line), C.c_void_p(org.ctypes.data), org.strides[0] // 2, int(above), int(left),
                                        ptr(sad), ptr(preds) if want_preds else None))
        return (sad, preds) if want_preds else sad

    def intra_rough_batch(self, jobs: np.ndarray, lines: np.ndarray, org: np.ndarray) -> np.ndarray:
        """jobs: capi.INTRA_JOB_DTYPE [n]; lines / org: flat int16 arrays the jobs' offsets address; returns uint32 [n, 35]"""
        jobs = np.ascontiguousarray(jobs, capi.INTRA_JOB_DTYPE)
        lines, org = np.ascontiguousarray(lines, np.int16).reshape(-1), np.ascontiguousarray(org, np.int16).reshape(-1)
        sad = np.zeros((len(jobs), capi.INTRA_MODES), np.uint32)
        self._ck(self.L.tvc_intra_rough_batch(self.h, len(jobs), ptr(jobs), ptr(lines), lines.size, ptr(org), org.size, ptr(sad)))
        return sad

    def pic_hash(self, slot: int, method: int) -> np.ndarray:
        """calcMD5 (1) / calcCRC (2) / calcChecksum (3) of the picture in a slot: uint8 [3, 16]"""
        d = np.zeros((3, 16), np.uint8)
        self._ck(self.L.tvc_pic_hash(self.h, slot, method, ptr(d)))
        return d

    def pic_ssd(self, slot_a: int, slot_b: int) -> np.ndarray:
        """the three sums of squared differences of xCalculateAddPSNR: uint64 [3]"""
        s = np.zeros(3, np.uint64)
        self._ck(self.L.tvc_pic_ssd(self.h, slot_a, slot_b, ptr(s)))
        return s

    def prof_enable(self, on: bool = True):
        self._ck(self.L.tvc_prof_enable(self.h, int(on)))

    def prof_read(self, reset: bool = True):
        """{phase: (ms_sum, kernel_groups)} measured with CUDA events on the context stream"""
        ms = np.zeros(len(capi.PHASES), np.float64)
        n = np.zeros(len(capi.PHASES), np.uint64)
        self._ck(self.L.tvc_prof_read(self.h, ptr(ms), ptr(n), int(reset)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(capi.PHASES)}

    def ubench(self, which: int) -> float:
        v = C.c_double()
        self._ck(self.L.tvc_ubench(self.h, which, C.byref(v)))
        return v.value
