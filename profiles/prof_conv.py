"""One 3x3 256->256 convolution with residual, timed alone (CUDA events): python profiles/prof_conv.py [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.networks import ACT, BF16, OP_CONV, Program
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.randn(n, 20, 256, device="cuda").bfloat16(); y = torch.empty_like(x); r = torch.randn_like(x)
w = (torch.randn(9, 4, 256, 64, device="cuda") / 48).bfloat16()
sc, sh = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
p = Program(n)
p.add(op=OP_CONV, dtype=BF16, H=4, W=5, cin=256, cout=256, ksize=3, act=ACT["relu"], use_tc=1, w_layout=1, src=x, dst=y, res=r, w=w, scale=sc, shift=sh)
for _ in range(5): p.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): p.run()
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 50 * 1e3
print(f"n={n} debug={os.environ.get('MZB_TC_DEBUG','0')}: {us:.1f} us per conv, {n*130*256*256*2/us/1e6:.0f} TFLOP/s valid-tap")

if int(os.environ.get("MZB_TC_DEBUG", "0")) & 8:
    import ctypes, numpy as np
    from muzero_breakout_b200 import _lib
    buf = np.zeros(8 * 64, np.uint64)
    L = _lib.lib(); L.mz_conv_trace.argtypes = [ctypes.c_void_p]
    L.mz_conv_trace(buf.ctypes.data)
    t = buf.reshape(8, 64).astype(np.int64)
    t0 = t[0, 0]
    names = ["mma:wait_tempty", "mma:start_issue", "mma:issued", "epi:start", "epi:res_loaded", "epi:tfull", "epi:drained", "epi:stored"]
    for it in range(6):
        if t[0, it] == 0: break
        print("tile", it, " ".join(f"{n}={(t[k, it] - t0) / 1e3:7.2f}" for k, n in enumerate(names)))
