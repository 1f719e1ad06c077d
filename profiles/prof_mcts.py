"""Profiling driver: a few MCTS simulation steps without the CUDA graph, for `ncu` (see profiles/README.md).
usage: python profiles/prof_mcts.py [trees] [sims]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": S, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": False, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
for _ in range(2):
    v, n = m.search(h, None, 0)
torch.cuda.synchronize()
print("ok", int(n.sum()), "launches per search", next(iter(m._plans.values())).kernels_per_search)
