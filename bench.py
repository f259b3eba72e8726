#!/usr/bin/env python
"""bench.py -- hot-path throughput of one 1080p low-delay-P picture (BASELINE.json configs[1]).

A "step" is one pass of the ported hot path over one synthetic 1920x1080 P picture with 4 reference
pictures (cfg/encoder_lowdelay_P_main.cfg: SR 64, FEN 1, HadamardME 1, AMP 1, 8-bit):

  1. ME frame pre-pass  : for every PU of the partition census (593 per CTU) x 4 references the integer TZ
                          search (xTZSearch; SADs computed on demand from TMA-staged search windows, no SAD
                          tables in HBM) and the fractional search (xPatternSearchFracDIF);
  2. motion compensation: Y/U/V prediction of the whole picture from a PU partition with quarter-pel MVs;
  3. residual           : org - pred (three planes);
  4. transform + RDOQ   : every TU size (luma 4..32, chroma 4..16) tiled over the picture (the residual
                          quadtree visits each size); forward transform, then xRateDistOptQuant (RDOQ 1 in the
                          cfg) with sign-data hiding against a synthetic CABAC bit-estimate table;
  5. dequant + inverse transform + reconstruction for the same TUs.

`value`  = frames/s with every input resident in HBM (device-pointer ABI entry points).
`e2e`    = the same through the host-pointer C ABI: the new pictures (current original + newest reference
           reconstruction) and the PU / TU lists are copied host->device from pinned memory and the ME
           results, levels and the reconstruction are copied back, all inside the timed region.
`--impl reference` times the reference's own compiled search code (oracle/_ref/libhmref.so; the C restatement
for the 3 % that is MC / transform / RDOQ) on all host cores over a bounded sample of CTUs of the same workload.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "1080p LDP hot-path fps (ME pre-pass + interp/MC + T/RDOQ/IQ/IT per P picture)"
UNIT = "frames/s"
W, H, BD = 1920, 1080, 8
NUM_REFS = 4
SEARCH_RANGE = 64
LAMBDA = 57.908390             # sqrt-lambda domain value HM derives for QP 32 P pictures (any value works)
QP = 32
SLOT_CUR, SLOT_PRED, SLOT_RESI, SLOT_RESI2, SLOT_RECON = 0, 5, 6, 7, 8
NUM_SLOTS = 9


# --------------------------------------------------------------------------------------- workload
def chroma_qp(qp):
    """chroma QP mapping of TComTrQuant::setQPforQuant (TComTrQuant.cpp:192-222, table TComRom.cpp:380-386)"""
    tab = [29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37]
    if qp < 30:
        return qp
    if qp >= 44:
        return qp - 6
    return tab[qp - 30]


def make_tu_list(w, h, qp, bd):
    from thevc_b200.capi import TU_DTYPE
    per_l, rem_l = (qp + 6 * (bd - 8)) // 6, (qp + 6 * (bd - 8)) % 6
    qc = chroma_qp(qp) + 6 * (bd - 8)
    per_c, rem_c = qc // 6, qc % 6
    rows = []
    off = 0
    counts = [0, 0, 0, 0]
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        for plane in (0, 1, 2):
            if plane and log2 == 5:
                continue
            pw, ph = (w, h) if plane == 0 else (w // 2, h // 2)
            xs = np.arange(0, pw - n + 1, n, dtype=np.int32)
            ys = np.arange(0, ph - n + 1, n, dtype=np.int32)
            gx, gy = np.meshgrid(xs, ys)
            k = gx.size
            a = np.zeros(k, TU_DTYPE)
            a["plane"] = plane
            a["x"] = gx.ravel()
            a["y"] = gy.ravel()
            a["log2_size"] = log2
            a["flags"] = 0
            a["scan_idx"] = 0
            a["qp_per"] = per_l if plane == 0 else per_c
            a["qp_rem"] = rem_l if plane == 0 else rem_c
            a["base_per"] = a["qp_per"]
            a["coef_offset"] = off + np.arange(k, dtype=np.int64) * (n * n)
            off += k * n * n
            counts[log2 - 2] += k
            rows.append(a)
    return np.concatenate(rows), np.array(counts, np.int32), off


def make_rdoq_list(tus, lam_luma, lam_chroma):
    """tvc_rdoq_tu of every TU: inter TUs, diagonal scan, luma at transform depth 0 (root cbf), chroma cbf context 5,
    one bit-estimate table (index 0)"""
    from thevc_b200.capi import RDOQ_TU_DTYPE
    a = np.zeros(len(tus), RDOQ_TU_DTYPE)
    luma = tus["plane"] == 0
    a["log2_size"] = tus["log2_size"]
    a["is_luma"] = luma
    a["scan_idx"] = 0
    a["qp_per"], a["qp_rem"] = tus["qp_per"], tus["qp_rem"]
    a["cbf_ctx"] = np.where(luma, -1, 5)
    a["est_index"] = 0
    a["coef_offset"] = tus["coef_offset"]
    a["lambda_"] = np.where(luma, lam_luma, lam_chroma)
    return a


def make_pu_list(w, h, rng, ctu_filter=None):
    """uni-predicted PU partition of the picture: CTU k uses CU size 64 >> (k % 4) with 2Nx2N parts;
    quarter-pel MVs within +-16 pels (most are 2-D fractional: the separable two-pass case)."""
    from thevc_b200.capi import PU_DTYPE
    rows = []
    ctus_x, ctus_y = (w + 63) // 64, (h + 63) // 64
    for ctu in range(ctus_x * ctus_y):
        if ctu_filter is not None and ctu not in ctu_filter:
            rng.integers(0, 1, 1)
            continue
        x0, y0 = (ctu % ctus_x) * 64, (ctu // ctus_x) * 64
        s = 64 >> (ctu % 4)
        for yy in range(0, 64, s):
            for xx in range(0, 64, s):
                if x0 + xx + s > w or y0 + yy + s > h:
                    continue
                rows.append((x0 + xx, y0 + yy, s, s, 1 + (ctu + xx // s) % NUM_REFS, 0, 0, -1, 0, 0))
    a = np.array(rows, PU_DTYPE)
    mv = rng.integers(-64, 65, (len(a), 2)).astype(np.int32)
    a["mvx0"], a["mvy0"] = mv[:, 0], mv[:, 1]
    return a


class Workload:
    def __init__(self, seed, pinned):
        import synth
        from thevc_b200.tlibcuda import HostPic
        if pinned:
            import torch

            def alloc(shape):
                t = torch.zeros(shape, dtype=torch.int16).pin_memory()
                self._keep.append(t)
                return t.numpy()
        else:
            alloc = None
        self._keep = []
        self.alloc = alloc
        seq = synth.make_sequence(W, H, NUM_REFS + 1, seed=seed)
        self.pics = []
        for i in range(NUM_REFS + 1):          # slot 0 = current picture (newest), slots 1..4 = refs, nearest first
            f = seq[NUM_REFS - i]
            p = HostPic(W, H, alloc=alloc)
            p.y[:], p.u[:], p.v[:] = f
            p.extend_border()
            self.pics.append(p)
        self.ctus_x, self.ctus_y = (W + 63) // 64, (H + 63) // 64
        self.nctu = self.ctus_x * self.ctus_y
        rng = np.random.default_rng(seed + 1)
        # predictor guesses (quarter pels): the synthetic global motion (3,5) px/frame x temporal distance + noise
        self.pred = np.zeros((NUM_REFS, self.nctu, 2), np.int32)
        for r in range(NUM_REFS):
            # tests/synth.py crops bg[5t:, 3t:]: the block at (x, y) of the current picture sits at (x + 3 dt, y + 5 dt) in the
            # picture dt frames back, so the motion vector towards reference r (dt = r + 1) is +(3, 5) * dt pels
            self.pred[r, :, 0] = 3 * 4 * (r + 1)
            self.pred[r, :, 1] = 5 * 4 * (r + 1)
        # ... to within a quarter pel: what a pre-pass can know.  The hooked encoder's own searches (real AMVP predictors) visit 27
        # candidates on average on this content; guesses off by up to 1.5 pels (round 1 / early round 2: +-6) made it 72, +-1 gives ~45,
        # exact guesses 37 (oracle, 40-CTU sample).  detail.tz_workload_check compares the step with the encoder leg of the same run.
        self.pred += rng.integers(-1, 2, self.pred.shape).astype(np.int32)
        self.pus = make_pu_list(W, H, rng)
        self.tus, self.tu_counts, self.coef_elems = make_tu_list(W, H, QP, BD)
        import rdoq_cases
        self.lam_luma, self.lam_chroma = rdoq_cases.lambda_for(QP), rdoq_cases.lambda_for(QP) * 0.85
        self.rtus = make_rdoq_list(self.tus, self.lam_luma, self.lam_chroma)
        self.est = rdoq_cases.make_est(np.random.default_rng(seed + 2))      # oracle.EstBits == tvc_est_bits layout
        self.est_bytes = np.frombuffer(bytes(self.est), np.uint8).copy()
        if pinned:      # the lists a host builds per picture live in page-locked memory too (copied to the device where they lie)
            import torch

            def pin(a):
                t = torch.zeros(a.nbytes, dtype=torch.uint8).pin_memory()
                self._keep.append(t)
                v = t.numpy().view(a.dtype).reshape(a.shape)
                v[...] = a
                return v
            self.pus, self.tus, self.rtus, self.est_bytes, self.pred = (pin(a) for a in (self.pus, self.tus, self.rtus, self.est_bytes, self.pred))

    def pic_bytes(self):
        p = self.pics[0]
        return p.buf_y.nbytes + p.buf_u.nbytes + p.buf_v.nbytes


def census_valid_mask(census, ctus_x, ctus_y):
    """[nctu, 593] bool: PU lies inside the 1080p picture"""
    m = np.zeros((ctus_x * ctus_y, len(census)), bool)
    for ctu in range(ctus_x * ctus_y):
        x0, y0 = (ctu % ctus_x) * 64, (ctu // ctus_x) * 64
        m[ctu] = (x0 + census[:, 0] + census[:, 2] <= W) & (y0 + census[:, 1] + census[:, 3] <= H)
    return m


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self):
        """start / end of the timed region: only samples between the first and last mark are reported"""
        self.marks.append(time.perf_counter())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = self.marks[0] if self.marks else 0.0
        hi = self.marks[-1] + 0.05 if len(self.marks) > 1 else float("inf")
        inside = [ln for (ts, ln) in self.lines if lo <= ts <= hi]
        if not inside:                      # region shorter than one sampling period: nearest samples
            inside = [ln for (_, ln) in self.lines[-2:]]
        for ln in inside:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_ctu_work(args):
    """the reference functions (oracle port) over ONE CTU of the workload: ME for the census x 4 refs,
    MC of the CTU's PUs, T+Q and IQ+IT+recon of the CTU's TUs.  Returns seconds per phase."""
    seed, ctu = args
    import oracle
    orc = oracle.lib()
    wl = _CPU_WL.get(seed)
    if wl is None:
        wl = Workload(seed, pinned=False)
        _CPU_WL[seed] = wl
    from oracle import ptr as optr
    cur = wl.pics[0]
    x0, y0 = (ctu % wl.ctus_x) * 64, (ctu // wl.ctus_x) * 64
    lc = orc.orc_lambda_motion_sad(LAMBDA)
    refs = (C.c_void_p * NUM_REFS)(*[optr(wl.pics[1 + r].buf_y, wl.pics[1 + r].origin(0)).value for r in range(NUM_REFS)])
    ires = (oracle.MeResult * (NUM_REFS * 593))()
    fres = (oracle.FracResult * (NUM_REFS * 593))()
    pred = np.ascontiguousarray(wl.pred[:, ctu, :].reshape(-1), np.int32)
    R = _cpu_reference()
    t0 = time.perf_counter()
    if R is not None:
        # the reference's OWN compiled TEncSearch (oracle/_ref/libhmref.so): xSetSearchRange + xTZSearch + xPatternSearchFracDIF
        oi = _CPU_WL.setdefault(("oi", seed), np.zeros((NUM_REFS * 593, 4), np.int32))
        of = _CPU_WL.setdefault(("of", seed), np.zeros((NUM_REFS * 593, 5), np.int32))
        R.ref_me_frame_ctu(optr(cur.buf_y, cur.origin(0)), refs, NUM_REFS, cur.stride, W, H, x0, y0, optr(pred), LAMBDA, SEARCH_RANGE,
                           optr(_CPU_WL["census"]), optr(oi), optr(of))
    else:
        orc.orc_me_frame_ctu(optr(cur.buf_y, cur.origin(0)), refs, NUM_REFS, cur.stride, W, H, x0, y0, optr(pred), lc, SEARCH_RANGE, 1, 1, 1, BD,
                             ires, fres)
    t1 = time.perf_counter()
    # MC
    pus = wl.pus[(wl.pus["x"] >= x0) & (wl.pus["x"] < x0 + 64) & (wl.pus["y"] >= y0) & (wl.pus["y"] < y0 + 64)]
    pus = np.ascontiguousarray(pus)
    from thevc_b200.tlibcuda import HostPic
    st = _CPU_WL.setdefault(("scratch", seed), [HostPic(W, H) for _ in range(4)])
    predp, resi, resi2, recon = st
    planes = (C.c_void_p * ((NUM_REFS + 1) * 3))()
    for s in range(NUM_REFS + 1):
        for pl in range(3):
            planes[s * 3 + pl] = optr(wl.pics[s].plane(pl), wl.pics[s].origin(pl)).value

    def tri(pic):
        return (C.c_void_p * 3)(*[optr(pic.plane(pl), pic.origin(pl)).value for pl in range(3)])
    orc.orc_mc_batch(planes, cur.stride, cur.cstride, tri(predp), len(pus), optr(pus.view(np.int32)), BD)
    t2 = time.perf_counter()
    for pl in range(3):
        sh = 1 if pl else 0
        stp = cur.stride if pl == 0 else cur.cstride
        o = cur.origin(pl) + (y0 >> sh) * stp + (x0 >> sh)
        bw, bh = min(64, W - x0) >> sh, min(64, H - y0) >> sh
        orc.orc_subtract(optr(cur.plane(pl), o), stp, optr(predp.plane(pl), o), stp, optr(resi.plane(pl), o), stp, bw, bh)
    t3 = time.perf_counter()
    tus = wl.tus
    sh = (tus["plane"] > 0).astype(np.int32)
    m = ((tus["x"] << sh) >= x0) & ((tus["x"] << sh) < x0 + 64) & ((tus["y"] << sh) >= y0) & ((tus["y"] << sh) < y0 + 64)
    sel = np.ascontiguousarray(tus[m])
    sizes = 1 << (2 * sel["log2_size"].astype(np.int64))
    sel["coef_offset"] = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    levels = np.zeros(int(sizes.sum()), np.int32)
    abs_sum = np.zeros(len(sel), np.uint32)
    orc.orc_fwd_rdoq_batch(tri(resi), cur.stride, cur.cstride, len(sel), optr(sel.view(np.int32)), 1, BD, C.byref(wl.est), wl.lam_luma,
                           wl.lam_chroma, optr(levels), optr(abs_sum))
    t4 = time.perf_counter()
    orc.orc_inv_tq_batch(tri(resi2), tri(predp), tri(recon), cur.stride, cur.cstride, len(sel), optr(sel.view(np.int32)), BD, optr(levels))
    t5 = time.perf_counter()
    n_sads = int(sum(r.n_sads for r in ires)) if R is None else 0
    return {"me": t1 - t0, "mc": t2 - t1, "sub": t3 - t2, "fwd_tq": t4 - t3, "inv_tq": t5 - t4, "total": t5 - t0, "n_sads": n_sads}


_CPU_WL = {}


def _cpu_reference():
    """libhmref.so (the reference's own sources compiled by oracle/Makefile) set up for the ME loops, or None when it was never
    built: the CPU arms then time the C restatement (measured here: the reference's code is 1.3-1.6x faster than the port)"""
    if "ref" not in _CPU_WL:
        import oracle
        R = oracle.ref()
        if R is not None and hasattr(R, "ref_me_frame_ctu"):
            R.ref_init(BD)
            R.ref_me_setup(W, H, SEARCH_RANGE, 1, 1)
            census = np.zeros((593, 6), np.int16)
            oracle.lib().orc_census(census.ctypes.data_as(C.c_void_p))
            _CPU_WL["census"] = census
        else:
            R = None
        _CPU_WL["ref"] = R
    return _CPU_WL["ref"]


def cpu_kind():
    return "reference" if _cpu_reference() is not None else "port"


CPU_WHAT = ("census ME x4 refs by the reference's own compiled TEncSearch (oracle/_ref/libhmref.so; 97 % of the time) + MC + T + RDOQ + IQ/IT "
            "by the C restatement")


def cpu_sample_ctus(n, nctu, ctus_x):
    """n interior CTUs spread over the picture (every census PU valid)"""
    rows = (H // 64)            # full CTU rows
    cand = [r * ctus_x + c for r in range(rows) for c in range(ctus_x)]
    step = max(1, len(cand) // n)
    return [cand[(i * step + 7) % len(cand)] for i in range(n)]


def frame_ctu_equiv(valid):
    """number of full-CTU equivalents in one picture (partial bottom CTUs weighted by valid census PUs)"""
    return float(valid.sum()) / valid.shape[1]


def run_cpu_baseline(seed, n_ctus, cores):
    """oracle port on `cores` processes over n_ctus CTUs; returns fps and detail"""
    import multiprocessing as mp
    import oracle
    oracle.build()
    wl_dims = ((W + 63) // 64, (H + 63) // 64)
    ctus = cpu_sample_ctus(n_ctus, wl_dims[0] * wl_dims[1], wl_dims[0])
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_ctu_work((seed, c)) for c in ctus]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_warm, [seed] * cores)
            t0 = time.perf_counter()
            res = pool.map(_cpu_ctu_work, [(seed, c) for c in ctus], chunksize=1)
    wall = time.perf_counter() - t0
    return wall, res, ctus


def _cpu_warm(seed):
    if seed not in _CPU_WL:
        _CPU_WL[seed] = Workload(seed, pinned=False)
    return 0


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores: the motion search (97 % of the
    time) is the reference's own compiled TEncSearch code (oracle/_ref/libhmref.so), MC / transform / RDOQ the C restatement;
    without libhmref.so (never built) everything is the restatement and `kind` says "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    import thevc_b200.capi as capi  # noqa: F401  (struct dtypes only; no GPU use on this arm)
    census = np.zeros((593, 6), np.int16)
    oracle.lib().orc_census(census.ctypes.data_as(C.c_void_p))
    valid = census_valid_mask(census, (W + 63) // 64, (H + 63) // 64)
    equiv = frame_ctu_equiv(valid)
    per_step = args.ref_ctus_per_core * cores
    _cpu_warm(args.seed)
    times = []
    import multiprocessing as mp
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_warm, [args.seed] * cores)
        for s in range(args.warmup + args.steps):
            ctus = cpu_sample_ctus(per_step * (args.warmup + args.steps), 0, (W + 63) // 64)[s * per_step:(s + 1) * per_step]
            t0 = time.perf_counter()
            pool.map(_cpu_ctu_work, [(args.seed, c) for c in ctus], chunksize=1)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    fps = (per_step * len(times) / equiv) / total
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16/int32", "data": "synthetic", "impl": "reference",
            "config": workload_config(),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                             "sample": "%d CTUs per step of the %.1f CTU-equivalents of a picture: %s" % (per_step, equiv, CPU_WHAT)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config():
    # distinct device buffers one step streams through (no SAD tables in the round-2 form): per census job a tvc_me_job (72 B), a
    # tvc_me_result (16), a tvc_frac_job (44), a tvc_frac_result (24) and 5 B of flags / list entries; coefficients and levels as
    # int32 over every TU size; the RDOQ scratch (ncu: 449 MB written per step); 13 int16 / u8 picture planes
    n_jobs = NUM_REFS * ((W + 63) // 64) * ((H + 63) // 64) * 593
    rec_mb = n_jobs * (72 + 16 + 44 + 24 + 5) / 1e6
    return {"workload": "cfg/encoder_lowdelay_P_main.cfg 1920x1080 8-bit synthetic, 1 P picture x 4 refs, SR 64, FEN 1, HAD 1: "
                        "ME pre-pass (integer TZ + fractional search for the 593-PU census of 510 CTUs x 4 refs) + MC + residual + T + RDOQ + IQ/IT",
            "width": W, "height": H, "num_refs": NUM_REFS, "search_range": SEARCH_RANGE, "qp": QP,
            "l2": "inputs larger than L2, no explicit flush: one step streams %.0f MB of job / result records, 90 MB of coefficient + level "
                  "buffers, the RDOQ scratch and 13 picture planes (8.1 MB int16 / 2.7 MB u8 each) -- more than 0.4 GB of distinct buffers "
                  "against the 126 MB L2; ncu measures about 1 GB of DRAM traffic per step (profiles/traffic.json)" % rec_mb,
            "parallelism": "independent sequences per GPU, no collective"}


# --------------------------------------------------------------------------------------- whole-encoder leg
ENC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppEncoderStatic")
ENC_CUDA = os.path.join(ROOT, "build", "hm", "TAppEncoderCuda")
LDP_CFG = os.path.join(ROOT, "build", "hm", "cfg", "encoder_lowdelay_P_main.cfg")
HM_HOOKS = os.environ.get("TVC_BENCH_HM_HOOKS", "me,frac,tables,frame,candgrid")      # the fast configuration of the HM shim


def _write_yuv(path, frames, seed):
    import synth
    seq = synth.make_sequence(W, H, frames, seed=seed)
    with open(path, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())


def _enc_args(yuv, frames, out):
    return ["-c", LDP_CFG, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(frames), "--SEIpictureDigest=1", "-o", os.devnull, "-b", out]


def run_hm_encode(frames, device_index=0):
    """BASELINE.json metric (i): end-to-end 1080p low-delay-P encode.  The unmodified reference encoder
    (oracle/_ref/bin/TAppEncoderStatic) and the reference encoder with the TLibCuda hooks (build/hm/TAppEncoderCuda:
    integer + fractional ME of the CU loop served by look-up from census-wide device batches, merge / AMVP candidate costs from
    device cost grids; CABAC, RDO, transforms stay the host's own code) encode the same synthetic 1920x1080 sequence, concurrently,
    one process each.  >= 17 pictures: four references are active from POC 4 on and the encoder recycles its picture buffers.
    fps = frames / wall time of the whole process (start-up, I picture and file I/O included); bitstream md5 compared with each
    other and, for the 17-picture default, with the committed md5 of the reference's single run (tests/golden/hm_md5.json)."""
    import hashlib
    import tempfile
    for pth in (ENC_REF, ENC_CUDA, LDP_CFG):
        if not os.path.exists(pth):
            return {"unavailable": "%s not built (needs /root/reference at build time)" % os.path.relpath(pth, ROOT)}
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        _write_yuv(yuv, frames, 20261018)
        env = dict(os.environ, TVC_HM=HM_HOOKS, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device_index)))
        t0 = time.perf_counter()
        pr = subprocess.Popen([ENC_REF] + _enc_args(yuv, frames, os.path.join(d, "ref.bin")), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        pc = subprocess.Popen([ENC_CUDA] + _enc_args(yuv, frames, os.path.join(d, "cuda.bin")), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        done = {}

        def wait(name, p):
            out, err = p.communicate()
            done[name] = (time.perf_counter() - t0, p.returncode, out, err or "")
        th = [threading.Thread(target=wait, args=("ref", pr)), threading.Thread(target=wait, args=("cuda", pc))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if done["ref"][1] != 0 or done["cuda"][1] != 0:
            return {"error": "encoder exit codes ref=%d cuda=%d: %s" % (done["ref"][1], done["cuda"][1], done["cuda"][3][-300:])}

        def md5(pth):
            return hashlib.md5(open(pth, "rb").read()).hexdigest()

        def ets(out):
            return [int(ln.split("[ET")[1].split("]")[0]) for ln in out.splitlines() if ln.startswith("POC") and "[ET" in ln]
        m_ref, m_cuda = md5(os.path.join(d, "ref.bin")), md5(os.path.join(d, "cuda.bin"))
        golden = None
        try:
            g = json.load(open(os.path.join(ROOT, "tests", "golden", "hm_md5.json"))).get("ldp_1080_%d" % frames)
            golden = g["md5"] if g else None
        except Exception:
            pass
        served = [ln for ln in done["cuda"][3].splitlines() if ln.startswith("TLibCuda")]
        out = {"frames": frames, "config": "cfg/encoder_lowdelay_P_main.cfg 1920x1080 synthetic, 1 I + %d P pictures, one process per encoder, "
                                           "TVC_HM=%s" % (frames - 1, HM_HOOKS),
               "reference_fps": frames / done["ref"][0], "ours_fps": frames / done["cuda"][0], "speedup": done["ref"][0] / done["cuda"][0],
               "reference_wall_s": done["ref"][0], "ours_wall_s": done["cuda"][0],
               "reference_picture_seconds": ets(done["ref"][2]), "ours_picture_seconds": ets(done["cuda"][2]),
               "bitstream_md5_equal": m_ref == m_cuda, "bitstream_md5": m_cuda, "golden_md5": golden,
               "golden_md5_equal": (m_cuda == golden) if golden else None, "hooks": served}
        if m_ref != m_cuda or (golden and m_cuda != golden):
            out["error"] = "bitstream md5 differs: ours %s reference %s golden %s" % (m_cuda, m_ref, golden)
        for ln in served:           # the encoder's own TZ work per search, for the workload check of the bench line
            if ln.startswith("TLibCuda TZ work:"):
                out["tz_candidates_per_search_mean"] = float(ln.split()[3])
        return out


def _pinned_encoder(args):
    """worker of run_cpu_encoder_baseline: own sequence, then one unmodified reference encoder pinned to one core"""
    core, frames, seed, workdir, start_at = args
    yuv, out = os.path.join(workdir, "in_%d.yuv" % core), os.path.join(workdir, "ref_%d.bin" % core)
    _write_yuv(yuv, frames, seed)
    while time.time() < start_at:          # all encoders start together, after every sequence is on disk
        time.sleep(0.01)
    t0 = time.perf_counter()
    cmd = [ENC_REF] + _enc_args(yuv, frames, out)
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    os.remove(yuv)
    return core, r.returncode, dt


def run_cpu_encoder_baseline(frames):
    """BASELINE.md 3.4 / north star: the reference encoder on the GPU box's own host cores, one process pinned per core, each on an
    independent synthetic sequence of the same configuration; aggregate and per-core fps with the core count.  A reported
    baseline, bounded to `frames` pictures per process so that the default bench stays within minutes."""
    import multiprocessing as mp
    import tempfile
    if not os.path.exists(ENC_REF) or not os.path.exists(LDP_CFG):
        return {"unavailable": "oracle/_ref/bin/TAppEncoderStatic not built (needs /root/reference at build time)"}
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = list(range(os.cpu_count() or 1))
    with tempfile.TemporaryDirectory() as d:
        start_at = time.time() + 6.0 + 0.4 * frames
        with mp.get_context("fork").Pool(len(cores)) as pool:
            res = pool.map(_pinned_encoder, [(c, frames, 20261018 + 104729 * (i + 1), d, start_at) for i, c in enumerate(cores)], chunksize=1)
    bad = [r for r in res if r[1] != 0]
    if bad:
        return {"error": "reference encoder failed on cores %s" % [r[0] for r in bad]}
    walls = [r[2] for r in res]
    return {"cores": len(cores), "frames_per_process": frames, "kind": "reference",
            "config": "cfg/encoder_lowdelay_P_main.cfg 1920x1080 synthetic, 1 I + %d P pictures per process, one unmodified reference encoder pinned per "
                      "host core, independent sequences" % (frames - 1),
            "aggregate_fps": len(cores) * frames / max(walls), "per_core_fps": frames / (sum(walls) / len(walls)),
            "wall_s_max": max(walls), "wall_s_mean": sum(walls) / len(walls)}


# --------------------------------------------------------------------------------------- N > 1 host logic
def rank_seed(seed, rank):
    """independent synthetic sequence per rank (an LDP sequence is one dependency chain)"""
    return seed + 7919 * rank


def max_over_ranks(v, world, device="cuda"):
    if world > 1:
        import torch
        import torch.distributed as dist
        tt = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())
    return v


def aggregate_fps(ms_per_step, world):
    """whole-job throughput: every rank finishes one picture per step"""
    return world * 1e3 / ms_per_step


# --------------------------------------------------------------------------------------- GPU arm
def intra_rough_leg(t, wl, local, cpu_sample=64):
    """SURVEY 8f-2 beside the P-picture step: the 35-mode rough search of EVERY intra PU of one 1080p picture (all five PU sizes
    of every CTU: what an I picture's estIntraPredQT calls add up to) as one device batch; reference samples = the picture's own
    neighbouring rows / columns (an I picture predicts from its reconstruction: same data volume, same arithmetic)."""
    import torch
    from thevc_b200 import capi
    from thevc_b200.capi import ptr
    pic = wl.pics[0]
    plane = np.ascontiguousarray(pic.y)                      # H x W int16
    Hh, Ww = plane.shape
    jobs, lines, off = [], [], 0
    for log2n in (6, 5, 4, 3, 2):
        n = 1 << log2n
        for y in range(0, Hh - n + 1, n):
            for x in range(0, Ww - n + 1, n):
                yy, xx = max(y, 1), max(x, 1)
                left = plane[np.clip(np.arange(yy + 2 * n - 1, yy - 1, -1), 0, Hh - 1), xx - 1]
                above = plane[yy - 1, np.clip(np.arange(xx, xx + 2 * n), 0, Ww - 1)]
                lines.append(np.concatenate([left, plane[yy - 1:yy, xx - 1], above]))
                jobs.append((log2n, off, y * Ww + x, Ww, 1, 1))
                off += 4 * n + 1
    jobs = np.array(jobs, capi.INTRA_JOB_DTYPE)
    lines = np.concatenate(lines).astype(np.int16)
    d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).cuda(local)
    d_lines, d_org = torch.from_numpy(lines).cuda(local), torch.from_numpy(plane.reshape(-1).copy()).cuda(local)
    d_sad = torch.zeros(len(jobs) * 35, dtype=torch.int32, device="cuda")

    def run():
        rc = t.L.tvc_intra_rough_batch_dev(t.h, len(jobs), C.c_void_p(d_jobs.data_ptr()), C.c_void_p(d_lines.data_ptr()),
                                           C.c_void_p(d_org.data_ptr()), C.c_void_p(d_sad.data_ptr()), None, None)
        if rc:
            raise RuntimeError("tvc_intra_rough_batch_dev: %s" % t.L.tvc_last_error(t.h).decode())
    for _ in range(3):
        run()
    t.prof_enable(True)
    t.prof_read(reset=True)
    reps = 10
    for _ in range(reps):
        run()
    ms = t.prof_read(reset=True)["intra"][0] / reps
    t.prof_enable(False)
    pels = float(sum((1 << (2 * int(j["log2_size"]))) for j in jobs)) * 35
    out = {"pus": int(len(jobs)), "ms": ms, "satd_gpel_per_s": pels / (ms * 1e-3) / 1e9,
           "algorithmic_GBps": (pels / 35 * 2 + lines.nbytes + len(jobs) * (24 + 140)) / (ms * 1e-3) / 1e9}
    # the oracle (C restatement of predIntraLumaAng + calcHAD) on a strided sample of the same PUs, one core
    if cpu_sample <= 0:
        return out
    try:
        import time
        import oracle
        O = oracle.lib()
        got = d_sad.cpu().numpy().view(np.uint32).reshape(-1, 35)
        idx = np.linspace(0, len(jobs) - 1, cpu_sample).astype(int)
        sad = np.zeros(35, np.uint32)
        t0 = time.perf_counter()
        spels = 0.0
        for i in idx:
            j = jobs[i]
            n = 1 << int(j["log2_size"])
            ln = np.ascontiguousarray(lines[j["line_offset"]:j["line_offset"] + 4 * n + 1])
            O.orc_intra_rough(oracle.ptr(ln), C.c_void_p(plane.ctypes.data + 2 * int(j["org_offset"])), Ww, int(j["log2_size"]), 1, 1, BD,
                              oracle.ptr(sad), None)
            if not np.array_equal(sad, got[i]):
                raise RuntimeError("intra rough search: device result of PU %d differs from the oracle" % i)
            spels += n * n * 35
        dt = time.perf_counter() - t0
        out["cpu_oracle_gpel_per_s_1core"] = spels / dt / 1e9
        out["cpu_sample"] = "%d PUs of all sizes, device result checked against each" % len(idx)
    except ImportError:
        pass
    return out



def gpu_arm(args):
    import torch
    from thevc_b200 import TLibCuda, capi
    from thevc_b200.capi import MeFrameCfg, QuantCfg, ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        # stdout carries ONE JSON line: NCCL prints its version banner (and, with NCCL_DEBUG=INFO, its log) to stdout while the
        # communicator comes up, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.cuda.set_device(local)
    if world > 1:
        # one rank per GPU on one host: every rank keeps to its own share of the cores (the host side of the end-to-end step --
        # list validation threads, staging copies -- otherwise competes with seven other ranks for the same cores)
        try:
            cores = sorted(os.sched_getaffinity(0))
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            per = max(1, len(cores) // max(1, lw))
            mine = cores[local * per:(local + 1) * per] or cores
            os.sched_setaffinity(0, set(mine))
        except (AttributeError, OSError, ValueError):
            pass
    stream = torch.cuda.Stream()
    wl = Workload(rank_seed(args.seed, rank), pinned=True)
    t = TLibCuda(W, H, BD, num_slots=NUM_SLOTS, device=local, stream=stream.cuda_stream)
    L, h = t.L, t.h
    lc = int(np.floor(65536.0 * np.sqrt(LAMBDA)))       # TComRdCost::setLambda (TComRdCost.cpp:167-173)
    census = t.me_census()
    valid = census_valid_mask(census, wl.ctus_x, wl.ctus_y)

    for s, p in enumerate(wl.pics):
        t.upload(s, p)
    refs = (C.c_int * NUM_REFS)(*range(1, NUM_REFS + 1))
    mcfg = MeFrameCfg(SEARCH_RANGE, 1, 1, 1, 1, lc)
    qc = QuantCfg(0, 1, 0)
    n_pu, n_tu = len(wl.pus), len(wl.tus)

    def dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda(local)
    pus_d, tus_d, rtus_d, est_d = dev(wl.pus), dev(wl.tus), dev(wl.rtus), dev(wl.est_bytes)
    levels_d = torch.zeros(wl.coef_elems, dtype=torch.int32, device="cuda")
    coef_d = torch.zeros(wl.coef_elems, dtype=torch.int32, device="cuda")
    abs_d = torch.zeros(n_tu, dtype=torch.int32, device="cuda")
    counts = wl.tu_counts
    pi, pf = C.c_void_p(), C.c_void_p()
    pred = wl.pred
    planes_wh = [(0, W, H), (1, W // 2, H // 2), (2, W // 2, H // 2)]

    def ck(rc):
        if rc:
            raise RuntimeError("libthevc_cuda error %d: %s" % (rc, L.tvc_last_error(h).decode()))

    def step_device():
        ck(L.tvc_me_frame_dev(h, SLOT_CUR, NUM_REFS, refs, ptr(pred), C.byref(mcfg), C.byref(pi), C.byref(pf)))
        ck(L.tvc_mc_batch_dev(h, SLOT_PRED, n_pu, C.c_void_p(pus_d.data_ptr())))
        for pl, pw, ph in planes_wh:
            ck(L.tvc_pic_subtract(h, SLOT_RESI, SLOT_CUR, SLOT_PRED, pl, 0, 0, pw, ph))
        ck(L.tvc_fwd_transform_batch_dev(h, SLOT_RESI, n_tu, C.c_void_p(tus_d.data_ptr()), ptr(counts), C.c_void_p(coef_d.data_ptr())))
        ck(L.tvc_rdoq_batch_dev(h, n_tu, C.c_void_p(rtus_d.data_ptr()), 1, C.c_void_p(est_d.data_ptr()), C.byref(qc),
                                C.c_void_p(coef_d.data_ptr()), C.c_void_p(levels_d.data_ptr()), None, wl.coef_elems,
                                C.c_void_p(abs_d.data_ptr())))
        ck(L.tvc_inv_tq_batch_dev(h, SLOT_RESI2, SLOT_PRED, SLOT_RECON, n_tu, C.c_void_p(tus_d.data_ptr()), ptr(counts),
                                  C.c_void_p(levels_d.data_ptr())))

    # ---- host-pointer (e2e) path buffers, pinned
    def pinned(shape, dtype):
        tt = torch.zeros(shape, dtype=dtype).pin_memory()
        return tt, tt.numpy()
    # results as the CU loop consumes them: 16 bytes per ME job (tvc_me_packed), levels as int16
    _k1, me_h = pinned((NUM_REFS * wl.nctu * 593, 4), torch.int32)
    _k3, levels_h = pinned((wl.coef_elems,), torch.int16)
    _k4, abs_h = pinned((n_tu,), torch.int32)
    recon_h = type(wl.pics[0])(W, H, alloc=wl.alloc)

    def step_e2e():
        # per-picture input traffic of a low-delay encoder: the new original and the newest reconstruction (the three older
        # references were uploaded when they were the newest)
        t.upload(0, wl.pics[0])
        t.upload(1, wl.pics[1])
        ck(L.tvc_me_frame_packed(h, SLOT_CUR, NUM_REFS, refs, ptr(pred), C.byref(mcfg), ptr(me_h)))
        ck(L.tvc_mc_batch(h, SLOT_PRED, n_pu, ptr(wl.pus)))
        for pl, pw, ph in planes_wh:
            ck(L.tvc_pic_subtract(h, SLOT_RESI, SLOT_CUR, SLOT_PRED, pl, 0, 0, pw, ph))
        # transform + RDOQ, levels / uiAbsSum to the host, dequant + inverse + reconstruction from the device copy of the levels
        ck(L.tvc_fwd_rdoq_recon_batch16(h, SLOT_RESI, SLOT_RESI2, SLOT_PRED, SLOT_RECON, n_tu, ptr(wl.tus), ptr(wl.rtus), 1, ptr(wl.est_bytes),
                                        C.byref(qc), ptr(levels_h), wl.coef_elems, ptr(abs_h)))
        t.download(SLOT_RECON, into=recon_h)

    h2d = 2 * wl.pic_bytes() + wl.pred.nbytes + wl.pus.nbytes + wl.tus.nbytes + wl.rtus.nbytes + wl.est_bytes.nbytes
    d2h = me_h.nbytes + levels_h.nbytes + abs_h.nbytes + wl.pic_bytes()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None

    # ---- device-resident timing (value)
    if clocks:
        clocks.start()
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        t.prof_enable(True)
        t.prof_read(reset=True)
        launches0 = t.launch_count()
        barrier()
        if clocks:
            clocks.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        if clocks:
            clocks.mark()
        ms_total = e0.elapsed_time(e1)
        clk = clocks.stop() if clocks else None
        launches = t.launch_count() - launches0
        phases = t.prof_read(reset=True)
        t.prof_enable(False)
    ms_total = max_over_ranks(ms_total, world)
    ms_step = ms_total / args.steps
    value = aggregate_fps(ms_step, world)

    # integer-ME candidate counts of this workload (for the algorithmic byte count of the search kernel);
    # run once more outside the timed region, results to the host
    ires_flat = np.zeros((NUM_REFS * wl.nctu * 593, 4), np.int32)
    fres_flat = np.zeros((NUM_REFS * wl.nctu * 593, 6), np.int32)
    ck(L.tvc_me_frame(h, SLOT_CUR, NUM_REFS, refs, ptr(pred), C.byref(mcfg), ptr(ires_flat), ptr(fres_flat)))
    n_sads = ires_flat[:, 3].astype(np.int64).reshape(NUM_REFS, wl.nctu, 593)
    me_stats = t.me_frame_stats()       # exact work counters of that call (table granules, raster candidates)

    # ---- e2e timing (host buffers through the host-pointer ABI)
    with torch.cuda.stream(stream):
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        k_e2e = max(1, min(args.steps, 5))
        for _ in range(k_e2e):
            step_e2e()
        e3.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e2.elapsed_time(e3), world) / k_e2e

    if rank != 0:
        t.close()
        return

    # ---- integer-pipe micro-benchmarks of THIS run (the measured vabsdiff4 rate is the ME denominator)
    ub = {}
    try:
        ub = {"vabsdiff4_ginstr_per_s": t.ubench(capi.UB_VABSDIFF4), "iadd3_ginstr_per_s": t.ubench(capi.UB_IADD3),
              "imad_ginstr_per_s": t.ubench(capi.UB_IMAD), "dp2a_ginstr_per_s": t.ubench(capi.UB_DP2A),
              "lds128_ginstr_per_s": t.ubench(capi.UB_LDS128), "hbm_write_GBps": t.ubench(capi.UB_HBM_WRITE)}
        ub["vabsdiff4_per_clk_per_sm"] = ub["vabsdiff4_ginstr_per_s"] * 1e9 / (148 * (clk["sm_mhz"] or 1965.0) * 1e6) if clk else None
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(ub, open(os.path.join(ROOT, "gpurun_out", "ubench.json"), "w"), indent=1)
    except Exception as ex:
        ub = {"error": repr(ex)[:200]}

    # ---- roofline (per launch = per step for these frame-level kernels)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_valid = int(valid.sum()) * NUM_REFS
    n_groups = NUM_REFS * wl.nctu
    fused = me_stats.get("form") == "group-search"
    pu_area = (census[:, 2].astype(np.int64) * census[:, 3])
    cw, chh = census[:, 2].astype(np.int64), census[:, 3].astype(np.int64)
    frac_bytes = float((valid * (((cw + 8) * (chh + 8) + pu_area) * 2 + 24)[None, :]).sum()) * NUM_REFS
    samples_tq = float(sum((1 << (2 * int(l))) * int(c) for l, c in zip((2, 3, 4, 5), counts)))
    alg_bytes = {
        "me_frac": frac_bytes,
        "mc": float(sum(int(p["w"]) * int(p["h"]) for p in wl.pus)) * 1.5 * 2 * 2,
        "fwd_tq": samples_tq * (2 + 4),
        "rdoq": samples_tq * (4 + 4) + float(n_tu) * 40,
        "inv_tq": samples_tq * (4 + 2 + 2 + 2),
    }
    bound = {"me_frac": "integer pipe", "mc": "hbm", "fwd_tq": "hbm", "rdoq": "latency (dependent FP64 chain)", "inv_tq": "hbm"}
    int_ops = {}          # algorithmic vabsdiff4 thread-instructions per step (4 pels each)
    if fused:
        # group search: the staged window + CTU of every (CTU, reference) in, one job record in and one result out per PU;
        # arithmetic = the sample differences of the candidates the reference's TZ search evaluates, four per vabsdiff4
        alg_bytes["me_search"] = n_groups * (208 * 192 + 80 * 64) + n_valid * (72 + 16)
        bound["me_search"] = "issue slots (per-lane SAD: shared-memory loads + funnel shifts + VABSDIFF4, and the TZ control flow); 17 of 32 lanes active on average"
        int_ops["me_search"] = float(me_stats["sample_differences"]) / 4.0
    else:
        table_bytes = t.me_table_bytes(NUM_REFS)
        alg_bytes["me_tables"] = table_bytes + NUM_REFS * wl.nctu * (208 * 192 + 64 * 64)
        alg_bytes["me_search"] = float(me_stats["search_granules"]) * 16 + n_valid * (72 + 16 + 8)
        alg_bytes["me_raster"] = float(me_stats["raster_candidates"]) * 1024 + n_valid * 8
        bound.update({"me_tables": "hbm", "me_search": "hbm (latency)", "me_raster": "shared memory"})
        int_ops["me_tables"] = float(NUM_REFS * wl.nctu) * 129 * 129 * 1024.0
    ph_ms = {k: (v[0] / args.steps) for k, v in phases.items()}
    traffic = {}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            traffic = json.load(open(traffic_file))
        except Exception:
            traffic = {}
    tkey = {k: (k + "_group" if (fused and k == "me_search") else k) for k in alg_bytes}
    vpeak = ub.get("vabsdiff4_ginstr_per_s")
    kernels = {}
    for k in alg_bytes:
        if ph_ms.get(k, 0) <= 0:
            continue
        a = alg_bytes[k] / (ph_ms[k] * 1e-3) / 1e9
        tr = traffic.get(tkey[k]) if isinstance(traffic.get(tkey[k]), dict) else {}
        kernels[k] = {"ms": ph_ms[k], "share_of_step": ph_ms[k] / ms_step, "bound": bound[k], "algorithmic_GB": alg_bytes[k] / 1e9,
                      "achieved_GBps": a, "frac_of_hbm_peak": a / peak, "traffic": tr.get("dram_bytes_per_launch"),
                      "issue_slots_busy_pct": tr.get("issue_slots_busy_pct"), "ncu_capture": tr.get("capture")}
        if k in int_ops and vpeak:
            r = int_ops[k] / (ph_ms[k] * 1e-3) / 1e9
            kernels[k]["integer_pipe"] = {"algorithmic_vabsdiff4_ginstr_per_s": r, "measured_peak_ginstr_per_s": vpeak, "frac": r / vpeak,
                                          "note": "4 sample differences per vabsdiff4; only the candidates the reference's search counts (speculative rounds excluded)"}
    # the roofline object names the kernel that takes the most TIME of the step (round-1 VERDICT weak #3); every kernel is in
    # detail.kernels with its own bound.  `traffic` is the DRAM byte count of one ncu --set full capture of that kernel
    # (profiles/traffic.json, written from the capture named in `ncu_capture`), or null when this build has no capture yet.
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    groups_per_step = max(1, phases[dom][1] // args.steps)
    kd = kernels[dom]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kd["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kd["frac_of_hbm_peak"],
                "traffic": kd["traffic"], "peak_source": peak_src, "launch_ms": kd["ms"] / groups_per_step, "launches_per_step": groups_per_step,
                "algorithmic_bytes_per_launch": alg_bytes[dom] / groups_per_step, "limiter": kd["bound"],
                "issue_slots_busy_pct": kd["issue_slots_busy_pct"], "ncu_capture": kd["ncu_capture"],
                "selection": "kernel with the largest share of the step's device time (%.0f %%)" % (100 * kd["share_of_step"])}
    if "integer_pipe" in kd:
        roofline["integer_pipe"] = kd["integer_pipe"]
    # whole step against HBM: all algorithmic bytes over the step time
    roofline["step"] = {"algorithmic_GB": sum(alg_bytes[k] for k in kernels) / 1e9, "ms": ms_step,
                        "frac_of_hbm_peak": sum(alg_bytes[k] for k in kernels) / (ms_step * 1e-3) / 1e9 / peak}
    # the integer-ME stage as a whole (SURVEY 8d metric (ii)): sample differences the reference's TZ search evaluates per second
    sad_pels = float(NUM_REFS * wl.nctu) * 129 * 129 * 4096
    satd_pels = float((valid * pu_area[None, :]).sum()) * NUM_REFS * 18
    sub = {
        "phase_ms_per_step": ph_ms,
        "phase_GBps": {k: kernels[k]["achieved_GBps"] for k in kernels},
        "me_sad_table_gpel_per_s": sad_pels / (ph_ms["me_tables"] * 1e-3) / 1e9 if ph_ms.get("me_tables", 0) > 0 else None,
        "me_frac_satd_gpel_per_s": satd_pels / (ph_ms["me_frac"] * 1e-3) / 1e9 if ph_ms["me_frac"] > 0 else None,
        "kernels": kernels,
        "ubench": ub,
        "me_work": me_stats,
        "tz_candidates_per_step": int(n_sads.sum()),
        "tz_candidates_per_search_mean": float(n_sads.sum()) / max(1, n_valid),
        "me_jobs_per_step": n_valid,
    }

    # k_me_group launches its CTAs longest group first, by the durations the previous picture-level call of the same shape measured
    # (a CTU that was expensive against a reference in the last picture usually is again).  The bench repeats ONE picture, the best
    # case for that history; this side leg times the same call right after a call of another shape (two references), i.e. with no
    # usable history (references interleaved, launch order otherwise), so that both numbers are on record.
    if fused:
        try:
            refs2 = (C.c_int * 2)(refs[0], refs[1])
            ms_hist, ms_nohist = [], []
            t.prof_enable(True)
            for _ in range(4):
                ck(L.tvc_me_frame_dev(h, SLOT_CUR, 2, refs2, None, C.byref(mcfg), C.byref(pi), C.byref(pf)))
                t.prof_read(reset=True)
                ck(L.tvc_me_frame_dev(h, SLOT_CUR, NUM_REFS, refs, ptr(pred), C.byref(mcfg), C.byref(pi), C.byref(pf)))
                ms_nohist.append(t.prof_read(reset=True)["me_search"][0])
                ck(L.tvc_me_frame_dev(h, SLOT_CUR, NUM_REFS, refs, ptr(pred), C.byref(mcfg), C.byref(pi), C.byref(pf)))
                ms_hist.append(t.prof_read(reset=True)["me_search"][0])
            t.prof_enable(False)
            sub["me_group_launch_order"] = {"policy": "longest group first by the previous call's CTA durations (TVC_GROUP_ORDER=4); first call / new shape: references interleaved",
                                            "me_search_ms_with_history": float(np.median(ms_hist)), "me_search_ms_without_history": float(np.median(ms_nohist)),
                                            "note": "the timed steps repeat one picture, so they run with history"}
        except Exception as ex:
            sub["me_group_launch_order"] = {"error": repr(ex)[:300]}

    try:
        sub["intra_rough"] = intra_rough_leg(t, wl, local, cpu_sample=0 if args.no_cpu else 64)
    except Exception as ex:          # a side leg must never take the hot-path line down with it
        sub["intra_rough"] = {"error": repr(ex)[:300]}

    # ---- CPU baseline (oracle port, one core, bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu:
        wall, res, ctus = run_cpu_baseline(args.seed, args.cpu_ctus, 1)
        equiv = frame_ctu_equiv(valid)
        per_ctu = wall / len(ctus)
        cpu = {"value": 1.0 / (per_ctu * equiv), "unit": UNIT, "cores": 1, "kind": cpu_kind(),
               "sample": "%d interior CTUs of %.1f CTU-equivalents per picture (%.1f s of CPU work): %s" % (len(ctus), equiv, wall, CPU_WHAT),
               "phase_s_per_ctu": {k: float(np.mean([r[k] for r in res])) for k in ("me", "mc", "fwd_tq", "inv_tq")}}
        # metric (ii): ME distortion throughput as the reference counts it -- sample differences of the SADs its TZ search evaluates
        # (w x (h >> iSubShift) per candidate) over its ME time (integer + fractional search; the reference driver times them together),
        # against the same count over the device's raster + search + fractional-search time (the device additionally fills the full
        # +-64 tables: me_sad_table_gpel_per_s)
        sad_rows = np.where(chh > 8, chh >> 1, chh)                     # FEN: even rows only when the PU is taller than 8
        per_pu = (cw * sad_rows)[None, None, :]
        tz_pels = (n_sads * per_pu)                                     # [ref, ctu, pu]
        me_s = float(sum(r["me"] for r in res))
        cpu["me_tz_sad_gpel_per_s"] = float(tz_pels[:, ctus, :].sum()) / me_s / 1e9
        dev_ms = ph_ms.get("me_search", 0.0) + ph_ms.get("me_raster", 0.0) + ph_ms.get("me_frac", 0.0)
        if dev_ms > 0:
            sub["me_tz_sad_gpel_per_s"] = float(tz_pels.sum()) / (dev_ms * 1e-3) / 1e9

    # ---- whole-encoder leg (metric (i) of BASELINE.json): after the context is gone (the encoder process opens its own)
    t.close()
    t = None
    failed = None
    if world == 1 and args.hm_frames > 0:
        try:
            sub["hm_encode"] = run_hm_encode(args.hm_frames, local)
        except Exception as ex:
            sub["hm_encode"] = {"error": repr(ex)[:300]}
        if "error" in sub["hm_encode"]:
            failed = "hm_encode: " + sub["hm_encode"]["error"]
        enc_mean = sub["hm_encode"].get("tz_candidates_per_search_mean")
        if enc_mean:
            # round-1 VERDICT weak #1: the step's synthetic predictor guesses must give the TZ search the work the real encoder's
            # searches have on the same kind of content -- within a factor of two, or the line is not worth quoting
            ratio = sub["tz_candidates_per_search_mean"] / enc_mean
            sub["tz_workload_check"] = {"bench_step_mean": sub["tz_candidates_per_search_mean"], "hooked_encoder_mean": enc_mean,
                                        "ratio": ratio, "ok": 0.5 <= ratio <= 2.0}
            if not 0.5 <= ratio <= 2.0:         # reported, not fatal: tests/test_config_matrix.py holds the assertion
                print("bench.py: workload check: %.1f TZ candidates per search in the step against %.1f in the encoder" % (
                    sub["tz_candidates_per_search_mean"], enc_mean), file=sys.stderr)
        if not args.no_cpu and args.cpu_enc_frames > 0:
            try:
                sub["cpu_baseline_encoder"] = run_cpu_encoder_baseline(args.cpu_enc_frames)
            except Exception as ex:
                sub["cpu_baseline_encoder"] = {"error": repr(ex)[:300]}
            h, c = sub["hm_encode"], sub["cpu_baseline_encoder"]
            if "ours_fps" in h and "aggregate_fps" in c:
                # metric (i) side by side: one hooked encoder process on 1 GPU + 1 host core, the reference on 1 core, the reference on every core
                sub["encode_fps"] = {"ours_1gpu_1process": h["ours_fps"], "reference_1core": h["reference_fps"],
                                     "reference_all_cores_aggregate": c["aggregate_fps"], "reference_per_core_under_full_load": c["per_core_fps"],
                                     "host_cores": c["cores"], "note": "ours / reference_1core over %d pictures; the all-core baseline runs %d pictures per "
                                     "process" % (h["frames"], c["frames_per_process"])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16/int32", "data": "synthetic", "config": workload_config(),
            "e2e": {"value": aggregate_fps(ms_e2e, world), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e, "steps": k_e2e},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "detail": sub}
    print(json.dumps(line))
    if failed:           # the line above is complete, but a broken encoder leg (exit code, md5 mismatch) fails the bench
        sys.stdout.flush()
        print("bench.py: " + failed, file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--cpu-ctus", type=int, default=48, help="CTUs of the CPU-baseline sample (1 core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--hm-frames", type=int, default=17, help="pictures of the whole-encoder 1080p leg (0 = skip)")
    ap.add_argument("--cpu-enc-frames", type=int, default=5, help="pictures per process of the all-cores reference-encoder baseline (0 = skip)")
    ap.add_argument("--ref-ctus-per-core", type=int, default=4, help="--impl reference: CTUs per core and step")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
