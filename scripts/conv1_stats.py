"""Bring-up: cycle counters of conv1_tc_pipe_kernel (library built with TB_NVCC_EXTRA=-DTB_CONV2_STATS; set TB_VI_CONV2_PAIR=0 so conv2 does not add to them)."""
import ctypes, os, sys
import torch
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
f = _capi.lib().tbdbg_conv2_stats
f.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    net.predict_device(crops.data_ptr(), n, 0, probs.data_ptr(), 0, 0)
f(None, 1)
R = 5
for _ in range(R):
    net.predict_device(crops.data_ptr(), n, 0, probs.data_ptr(), 0, 0)
f(out, 1)
v = [x / R / 148 / 1920 for x in out]
print(prec, "conv1 (per CTA, us @1.92 GHz)")
print("  MMA thread: total %.0f  wait plane %.0f  wait acc free %.0f" % (v[0], v[1], v[3]))
print("  epilogue warp 2: total %.0f  wait acc full %.0f" % (v[4], v[5]))
print("  producer thread 0: total %.0f  wait plane free %.0f" % (v[11], v[12]))
