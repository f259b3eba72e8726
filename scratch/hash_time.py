"""time of tvc_pic_hash / tvc_pic_ssd on a 1080p picture (wall clock around the synchronous calls)"""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, synth
from thevc_b200 import TLibCuda
t = TLibCuda(1920, 1080, 8, num_slots=3)
rng = np.random.default_rng(1)
a, b = synth.random_pic(rng, 1920, 1080, 8), synth.random_pic(rng, 1920, 1080, 8)
t.upload(0, a); t.upload(1, b)
for name, f in (("md5", lambda: t.pic_hash(0, 1)), ("crc", lambda: t.pic_hash(0, 2)), ("checksum", lambda: t.pic_hash(0, 3)), ("ssd", lambda: t.pic_ssd(0, 1))):
    f()
    t0 = time.perf_counter()
    for _ in range(5): f()
    print("%-9s %.3f ms per 1080p picture" % (name, (time.perf_counter() - t0) / 5 * 1e3))
