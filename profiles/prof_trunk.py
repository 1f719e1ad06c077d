"""Sustained time of one MCTS simulation step's network program (dynamics + prediction, 59 trunk layers + heads) at n samples, with
the SM clock and board power sampled while it runs:
    MZB_STACK_SLICE=4096 MZB_STACK_ROT=13 python profiles/prof_trunk.py 4096
(the env switches are read once per process: run one process per setting).  "cyc/kstep" = SM cycles per k-step (four 256x256x16
tcgen05.mma) of a CTA pair averaged over the whole step: 512 would be the tensor pipe never idle."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
m.search(h, None, 0)
plan = next(iter(m._plans.values()))
prog = plan.sim_prog
for _ in range(5):
    prog.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); prog.run(); b.record(); torch.cuda.synchronize()
reps = max(10, int(secs * 1e3 / a.elapsed_time(b)))

lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
t = threading.Thread(target=lambda: lines.extend(proc.stdout), daemon=True)
t.start()
time.sleep(0.3)
n0 = len(lines)
a.record()
for _ in range(reps):
    prog.run()
b.record(); torch.cuda.synchronize()
n1 = len(lines)
ms = a.elapsed_time(b) / reps
a.record()
for _ in range(3):
    m.search(h, None, 0)
b.record(); torch.cuda.synchronize()
sms = a.elapsed_time(b) / 3
proc.terminate(); t.join(timeout=2)
clk, pw, cap = [], [], 0
for ln in lines[n0 + 3:max(n1, n0 + 4)]:           # skip the ramp
    f = [x.strip() for x in ln.split(",")]
    try:
        clk.append(float(f[0])); pw.append(float(f[1])); cap += f[2].lower().startswith("active")
    except (ValueError, IndexError):
        pass
mhz = float(np.median(clk)) if clk else float("nan")
# k-steps of one CTA pair per simulation step: 57 residual-trunk layers + the fused policy/value head layer = 58 3x3 layers of 130 (pixel, tap)
# pairs x 4 channel chunks per 256-sample group pair, + the 1x1 reward layer (20 x 4), spread over 74 pairs
ksteps = (58 * 130 * 4 + 20 * 4) * ((B + 255) // 256) / 74
flop = 983895040 * B
env = " ".join(f"{k[4:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("MZB_"))
print(f"n={B} [{env}] kernels={prog.n_kernels} sim-step {ms:.3f} ms = {flop / ms / 1e9:.0f} TFLOP/s valid-tap | SM {mhz:.0f} MHz, {np.median(pw) if pw else 0:.0f} W, "
      f"power-capped {cap}/{len(clk)} samples | {ms * 1e3 * mhz / ksteps:.0f} cyc/kstep | search {sms:.1f} ms = {B * 50 / sms:.0f} k sims/s")
