"""Bring-up: cycle counters of conv2_pair_kernel (library built with TB_NVCC_EXTRA=-DTB_CONV2_STATS)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trex_b200
from trex_b200 import _capi
from trex_b200.weights import random_v118_3_state_dict

prec = sys.argv[1]
n = 23408
net = trex_b200.VINetwork(100, max_images=n, device=0, precision=prec)
net.load_weights(random_v118_3_state_dict(100, seed=0))
crops = torch.randint(0, 255, (n, 80, 80), dtype=torch.uint8, device="cuda")
probs = torch.empty((n, 100), dtype=torch.float32, device="cuda")
L = _capi.lib()
f = L.tbdbg_conv2_stats
f.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    net.predict_device(crops.data_ptr(), n, 0, probs.data_ptr(), 0, 0)
f(None, 1)
R = 5
for _ in range(R):
    net.predict_device(crops.data_ptr(), n, 0, probs.data_ptr(), 0, 0)
f(out, 1)
v = [x / R for x in out]
nmma = 148 if os.environ.get("TB_VI_CONV2_PAIR", "2") != "2" else 74
print(prec, "variant", os.environ.get("TB_VI_CONV2_PAIR", "2"))
print("  MMA thread (per CTA, us @1.92GHz): total %.0f  wait own band %.0f  wait peer band %.0f  wait acc free %.0f" % tuple(x / nmma / 1920 for x in v[0:4]))
print("  epilogue phases: tmem ld %.0f  math+shuffle %.0f  stores %.0f" % (v[8] / 148 / 1920, v[9] / 148 / 1920, v[10] / 148 / 1920))
print("  epilogue warp 2 (per CTA): total %.0f  wait acc full %.0f  hand-over barriers %.0f  tiles %.0f" % (v[4] / 148 / 1920, v[5] / 148 / 1920, v[6] / 148 / 1920, v[7] / 148))
