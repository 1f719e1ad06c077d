"""Sustained time of one MCTS simulation step's network program (dynamics + prediction, 60 conv layers) at n samples:
    MZB_FUSE_MAX_SAMPLES=100000 MZB_STACK_ROT=13 python profiles/prof_trunk.py 4096
(the env switches are read once per process: run one process per setting)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
a.record()
for _ in range(reps):
    prog.run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
flop = 983895040 * B
a.record()
for _ in range(3):
    m.search(h, None, 0)
b.record(); torch.cuda.synchronize()
sms = a.elapsed_time(b) / 3
print(f"n={B} fuse_max={os.environ.get('MZB_FUSE_MAX_SAMPLES', 'default')} rot={os.environ.get('MZB_STACK_ROT', 'default')} "
      f"kernels={prog.n_kernels} sim-step {ms:.3f} ms = {flop / ms / 1e9:.0f} TFLOP/s valid-tap; search {sms:.1f} ms = {B * 50 / sms:.0f} k sims/s")
