import sys, os, json
sys.path.insert(0, os.getcwd())
from thevc_b200 import TLibCuda
t = TLibCuda(416, 240, 8, num_slots=1)
names = ["vabsdiff4.add", "iadd", "imad", "lds128", "dp2a"]
out = {}
for i, n in enumerate(names):
    out[n] = round(t.ubench(i), 1)
hbm_write = round(t.ubench(5), 1)
print(json.dumps({"hbm_write_only_GBps": hbm_write}))
print(json.dumps({"ginstr_per_s_thread_level": out, "per_sm_per_clk_at_1965MHz": {k: round(v * 1e9 / (148 * 1.965e9), 1) for k, v in out.items()}}))
