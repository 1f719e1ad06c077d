"""Per-layer phase timeline of the latency-mode trunk (csrc/conv_lat.cu), CTA 0, from globaltimer stamps:
    MZB_LAT_TRACE=1 python profiles/prof_lat_trace.py 24
phases per layer: poll+load (the previous layer's flag-in-data words polled into the row buffer + this layer's weights landed) |
w-issue (next item's TMA boxes) | math (18 k16 steps per warp) | to-smem (partials) | reduce+epilogue (stores) | end (barrier)"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MZB_LAT_TRACE", "1")
from muzero_breakout_b200 import _lib
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
h = torch.rand(n, 256, 4, 5)
for _ in range(3):
    nets.prediction(h)
torch.cuda.synchronize()
buf = np.zeros(8 * 64, dtype=np.uint64)
L = _lib.lib(); L.mz_lat_trace.argtypes = [ctypes.c_void_p]
L.mz_lat_trace(buf.ctypes.data)
t = buf.reshape(8, 64).astype(np.int64)
t = t[[0, 1, 2, 6, 7, 3, 4, 5]]      # program order: slot 6 = next layer's weight prefetch issued, 7 = MMAs done
nl = 28
names = ["-", "poll+load", "w-issue", "math", "to-smem", "reduce+epi", "end"]
d = np.stack([t[i + 1, :nl] - t[i, :nl] for i in range(7)])
for l in range(nl):
    print(f"layer {l:2d}: " + "  ".join(f"{names[i]} {d[i, l]:6d}" for i in range(7)) + f"   total {t[7, l] - t[0, l]:6d} ns")
print("mean (layers 2..): " + "  ".join(f"{names[i]} {d[i, 2:].mean():7.0f}" for i in range(7)) + f"   layer period {(t[0, nl - 1] - t[0, 2]) / (nl - 3):.0f} ns")
