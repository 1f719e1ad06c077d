"""Per-layer timestamps of cluster 0 in the fused trunk kernel: MZB_STACK_TRACE=1 python profiles/prof_stack.py [n]"""
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
t = buf.reshape(6, 64).astype(np.int64); t0 = t[0, 0]
names = ["prod:tile0", "prod:flags_ok", "mma:first_data", "mma:tile0_issued", "epi:tfull", "epi:flag_released"]
for layer in range(8):
    print("layer", layer, " ".join(f"{nm}={(t[k, layer] - t0) / 1e3:7.2f}" for k, nm in enumerate(names)))
print("last layer 27:", " ".join(f"{nm}={(t[k, 27] - t0) / 1e3:7.2f}" for k, nm in enumerate(names)))
