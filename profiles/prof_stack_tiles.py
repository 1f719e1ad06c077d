"""Per-tile timeline of CTA pair 0's MMA issuer in the fused trunk kernel under sustained load:
    MZB_STACK_TRACE=4 python profiles/prof_stack_tiles.py [n]
columns: wait for a free accumulator | wait for the first operands | MMA issue time, per k-step"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200 import _lib
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
h = torch.rand(n, 256, 4, 5)
for _ in range(40): nets.prediction(h)
torch.cuda.synchronize()
buf = np.zeros(6 * 64, np.uint64)
L = _lib.lib(); L.mz_stack_trace.argtypes = [ctypes.c_void_p]
L.mz_stack_trace(buf.ctypes.data)
t = buf.reshape(96, 4)
ks = (t[:, 3] >> np.uint64(40)).astype(np.int64)
t = t.astype(np.int64)
t[:, 3] = t[:, 2] + (t[:, 3] & ((1 << 40) - 1))
t0 = t[0, 0]
tot_issue = tot_wait_acc = tot_wait_data = 0
for i in range(96):
    wa, wd, iss = t[i, 1] - t[i, 0], t[i, 2] - t[i, 1], t[i, 3] - t[i, 2]
    gap = t[i, 0] - t[i - 1, 3] if i else 0
    if i < 24 or i % 8 == 0:
        print(f"tile {i:2d} ksteps {ks[i]:2d} start {(t[i,0]-t0)/1e3:8.2f} us  gap {gap/1e3:5.2f}  wait_acc {wa/1e3:5.2f}  wait_data {wd/1e3:5.2f}  issue {iss/1e3:6.2f} = {iss/max(ks[i],1):5.0f} ns/kstep")
    tot_issue += iss; tot_wait_acc += wa; tot_wait_data += wd
span = t[95, 3] - t[0, 0]
print(f"96 tiles: span {span/1e3:.1f} us, issue {tot_issue/1e3:.1f} ({100*tot_issue/span:.0f} %), wait_acc {tot_wait_acc/1e3:.1f} ({100*tot_wait_acc/span:.0f} %), wait_data {tot_wait_data/1e3:.1f} ({100*tot_wait_data/span:.0f} %), ksteps {ks.sum()}, {tot_issue/ks.sum():.0f} ns per k-step while issuing, {span/ks.sum():.0f} ns per k-step overall")
