"""k-step issue timestamps of the first tile (cluster 0, layer 0): MZB_STACK_TRACE=2 (TMA producer) / 3 (MMA issuer)"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200 import _lib
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
h = torch.rand(n, 256, 4, 5)
for _ in range(3): nets.prediction(h)
torch.cuda.synchronize()
buf = np.zeros(6 * 64, np.uint64)
L = _lib.lib(); L.mz_stack_trace.argtypes = [ctypes.c_void_p]
L.mz_stack_trace(buf.ctypes.data)
t = buf[:16].astype(np.int64)
print("mode", os.environ.get("MZB_STACK_TRACE"), "n", n, "k-step timestamps (ns since first):", (t - t[0]).tolist())
