"""The training path's per-layer tcgen05 convolution (conv_tc_kernel through Program / mz_run) at a minibatch: CUDA-event time per launch,
forward form (bias + fp32 side output) and data-gradient form.   python profiles/prof_train_conv.py [samples]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200 import train
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
w = torch.randn(256, 256, 3, 3) * 0.02
blk = train.ResidualBlockTrain(w, torch.zeros(256), torch.ones(256), torch.zeros(256), w, torch.zeros(256), torch.ones(256), torch.zeros(256))
x = torch.randn(n, 4, 5, 256, device="cuda").bfloat16()
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print(f"n={n} forward conv (Program build + launch, eager): {timed(lambda: blk._conv(x, 0)):.1f} us per call")
print(f"n={n} dgrad conv: {timed(lambda: blk.dgrad[0](x)):.1f} us per call")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): blk._conv(x, 0)
print(f"n={n} forward conv inside a CUDA graph (20 back to back): {timed(g.replay, 20) / 20:.1f} us per launch")
if int(os.environ.get("MZB_TC_DEBUG", "0")) & 8:
    import ctypes, numpy as np
    from muzero_breakout_b200 import _lib
    blk._conv(x, 0); torch.cuda.synchronize()
    buf = np.zeros(8 * 64, np.uint64)
    L = _lib.lib(); L.mz_conv_trace.argtypes = [ctypes.c_void_p]
    L.mz_conv_trace(buf.ctypes.data)
    t = buf.reshape(8, 64).astype(np.int64)
    t0 = t[0, 0]
    names = ["mma:wait_tempty", "mma:start_issue", "mma:issued", "epi:start", "epi:res_loaded", "epi:tfull", "epi:drained", "epi:stored"]
    fine = buf[32:48].astype(np.int64).reshape(4, 4)
    if fine[0, 0]:
        for c in range(4):
            print(f"  epilogue warp 2, tile 0, chunk {c}: start {(fine[c, 0] - t0) / 1e3:6.2f}  tmem_wait done {(fine[c, 1] - t0) / 1e3:6.2f}  staged {(fine[c, 2] - t0) / 1e3:6.2f}  stored {(fine[c, 3] - t0) / 1e3:6.2f}")
    for it in range(3):
        if t[0, it] == 0: break
        print("tile", it, " ".join(f"{n}={(t[k, it] - t0) / 1e3:7.2f}" for k, n in enumerate(names)))
