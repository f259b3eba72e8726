"""Generate tests/golden/rdoq_golden.npz from the REFERENCE ITSELF: TComTrQuant::xRateDistOptQuant of the
unmodified /root/reference sources (oracle/_ref/libhmref.so, ref_rdoq in oracle/ref_shim.cpp).  Run in the build
container; the .npz is committed so that the RDOQ pin also holds on the GPU box.

    python tests/golden/make_rdoq_golden.py

Each case stores its inputs (bit-estimate table, parameters, coefficients) and the reference's outputs."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
import rdoq_cases as rc  # noqa: E402

# columns of the parameter table
COLS = ("bd", "log2", "is_luma", "qp", "per", "rem", "is_intra", "intra_dir", "tr_idx", "scan_idx", "cbf_ctx", "sign_hide",
        "use_arl", "est", "offset", "abs_sum")


def main():
    oracle.build()
    R, L = oracle.ref(), oracle.lib()
    assert R is not None, "oracle/_ref/libhmref.so is not built"
    rng = np.random.default_rng(7202)
    ests = [rc.make_est(rng) for _ in range(4)]
    est_arr = np.stack([np.frombuffer(bytes(e), np.int32) for e in ests])
    params, lambdas, coefs, levels, arls = [], [], [], [], []
    off = 0
    for bd in (8, 10):
        R.ref_init(bd)
        for log2, count in ((2, 24), (3, 20), (4, 10), (5, 6)):
            n = 1 << log2
            for k in range(count):
                is_luma = 1 if log2 == 5 else int(rng.integers(0, 2))
                qp = int(rng.choice([10, 22, 27, 32, 37, 45]))
                per, rem = C.c_int(), C.c_int()
                L.orc_set_qp(qp, is_luma, 6 * (bd - 8), 0, C.byref(per), C.byref(rem))
                icu = int(rng.integers(0, 2)); ldir = int(rng.choice([1, 10, 26, 34])); tr_idx = int(rng.integers(0, 3))
                sh, arl = int(rng.integers(0, 2)), int(rng.integers(0, 2))
                lam = rc.lambda_for(qp) * float(rng.uniform(0.5, 2.0))
                ei = int(rng.integers(0, len(ests)))
                coef = rc.make_coef(rng, log2, per.value, bd, k % 6)
                q = np.zeros(n * n, np.int32); a = np.zeros(n * n, np.int32); s = C.c_uint32(0)
                R.ref_rdoq(coef.copy(), q, a, n, qp, 6 * (bd - 8), is_luma, icu, ldir, tr_idx, sh, arl, lam, C.byref(ests[ei]), C.byref(s))
                params.append((bd, log2, is_luma, qp, per.value, rem.value, icu, ldir, tr_idx, rc.scan_idx_for(icu, is_luma, n, ldir),
                               rc.cbf_ctx_for(icu, is_luma, tr_idx), sh, arl, ei, off, s.value))
                lambdas.append(lam); coefs.append(coef); levels.append(q); arls.append(a)
                off += n * n
    path = os.path.join(HERE, "rdoq_golden.npz")
    np.savez_compressed(path, est=est_arr, params=np.array(params, np.int64), lambdas=np.array(lambdas, np.float64),
                        coef=np.concatenate(coefs), levels=np.concatenate(levels), arl=np.concatenate(arls))
    print("wrote", path, os.path.getsize(path), "bytes,", len(params), "cases")


if __name__ == "__main__":
    main()
