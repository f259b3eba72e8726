"""Picture hashes and PSNR sums.  CPU: the C restatement against the reference's own compiled calcMD5 / calcCRC /
calcChecksum (libhmref.so) and against independent statements (hashlib's MD5 of the packed samples, a bit-serial CRC and the byte
sum written straight from the text of TComPicYuvMD5.cpp in numpy / Python).  GPU (-m gpu): tvc_pic_hash / tvc_pic_ssd through the
C ABI against the oracle on whole pictures, 8- and 10-bit."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle
from oracle import ptr
import synth

METHODS = {1: ("orc_md5_plane", 16), 2: ("orc_crc_plane", 2), 3: ("orc_checksum_plane", 4)}


def _oracle_digest(orc, method, planes, bd):
    name, n = METHODS[method]
    out = np.zeros((3, 16), np.uint8)
    for k, pl in enumerate(planes):
        pl = np.ascontiguousarray(pl)
        getattr(orc, name)(ptr(pl), pl.shape[1], pl.shape[0], pl.shape[1], bd, ptr(out[k]))
    return out


def _planes(rng, w, h, bd):
    return [rng.integers(0, 1 << bd, s).astype(np.int16) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_against_independent_statements(orc, bd):
    rng = np.random.default_rng(5 + bd)
    for (w, h) in ((8, 4), (72, 40), (208, 120)):
        for pl in _planes(rng, w, h, bd):
            d = np.zeros(16, np.uint8)
            orc.orc_md5_plane(ptr(pl), pl.shape[1], pl.shape[0], pl.shape[1], bd, ptr(d))
            packed = pl.astype("<u2").tobytes() if bd > 8 else pl.astype(np.uint8).tobytes()
            assert bytes(d) == hashlib.md5(packed).digest()
            orc.orc_checksum_plane(ptr(pl), pl.shape[1], pl.shape[0], pl.shape[1], bd, ptr(d))
            yy, xx = np.mgrid[0:pl.shape[0], 0:pl.shape[1]]
            mask = ((xx & 0xff) ^ (yy & 0xff) ^ (xx >> 8) ^ (yy >> 8)).astype(np.int64) & 0xff
            s = int((((pl.astype(np.int64) & 0xff) ^ mask).sum() + ((((pl.astype(np.int64) >> 8) ^ mask).sum()) if bd > 8 else 0)) & 0xffffffff)
            assert [int(v) for v in d[:4]] == [(s >> 24) & 255, (s >> 16) & 255, (s >> 8) & 255, s & 255]
            if pl.size <= 72 * 40:          # bit-serial CRC in Python: small planes only
                orc.orc_crc_plane(ptr(pl), pl.shape[1], pl.shape[0], pl.shape[1], bd, ptr(d))
                crc, msb = 0xffff, bd - 1
                for v in pl.reshape(-1):
                    for b in range(bd):
                        top = (crc >> 15) & 1
                        crc = (((crc << 1) + ((int(v) >> (msb - (b & msb))) & 1)) & 0xffff) ^ (top * 0x1021)
                for _ in range(16):
                    top = (crc >> 15) & 1
                    crc = ((crc << 1) & 0xffff) ^ (top * 0x1021)
                assert (int(d[0]), int(d[1])) == (crc >> 8, crc & 255)


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_against_reference_hashes(orc, hmref, bd):
    if not hasattr(hmref, "ref_pic_hash"):
        pytest.skip("libhmref.so predates ref_pic_hash")
    hmref.ref_init(bd)
    rng = np.random.default_rng(15 + bd)
    for (w, h) in ((64, 64), (208, 120), (416, 240), (72, 40)):
        y, u, v = _planes(rng, w, h, bd)
        for method, (_, n) in METHODS.items():
            d = np.zeros((3, 16), np.uint8)
            hmref.ref_pic_hash(method, ptr(y), ptr(u), ptr(v), w, h, ptr(d))
            assert np.array_equal(_oracle_digest(orc, method, (y, u, v), bd)[:, :n], d[:, :n]), (w, h, method)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_gpu_pic_hash_and_ssd(orc, bd):
    from thevc_b200 import TLibCuda
    W, H = 416, 240
    t = TLibCuda(W, H, bd, num_slots=3)
    try:
        rng = np.random.default_rng(25 + bd)
        a, b = synth.random_pic(rng, W, H, bd), synth.random_pic(rng, W, H, bd)
        t.upload(0, a); t.upload(1, b)
        for method in METHODS:
            assert np.array_equal(t.pic_hash(0, method), _oracle_digest(orc, method, (a.y, a.u, a.v), bd)), method
            assert np.array_equal(t.pic_hash(1, method), _oracle_digest(orc, method, (b.y, b.u, b.v), bd)), method
        got = t.pic_ssd(0, 1)
        for k, (p, q) in enumerate(((a.y, b.y), (a.u, b.u), (a.v, b.v))):
            p, q = np.ascontiguousarray(p), np.ascontiguousarray(q)
            assert int(got[k]) == int(orc.orc_ssd_plane(ptr(p), p.shape[1], ptr(q), q.shape[1], p.shape[1], p.shape[0]))
        assert [int(v) for v in t.pic_ssd(0, 0)] == [0, 0, 0]
        with pytest.raises(Exception):
            t.pic_hash(0, 4)
    finally:
        t.close()
