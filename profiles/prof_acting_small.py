"""Whole acting moves at config.yaml's default size (24 envs x 50 simulations) and the representation network alone:
    python profiles/prof_acting_small.py 24"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.acting import Actor
from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
env = BreakoutEnvironment(dict(n_parallel=B, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0,
                               output_device="cuda", reset_rng="device", seed=5))
moves = 8
actor = Actor(env, m, temperature=1.0, seed=0, max_moves=moves, check_done_every=1 << 30, record_frames=False)
actor.run_episode()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    a.record()
    actor.run_episode()
    b.record(); torch.cuda.synchronize()
    print(f"n={B}: acting move {a.elapsed_time(b) / moves:.2f} ms (rep input + representation net + search + sampling + env step + record)")
x = torch.rand(B, 64, 16, 20, device="cuda")
out = torch.empty(B, 256, 4, 5, device="cuda")
prog = nets.representation_program(B, x, out)
for _ in range(3):
    prog.run()
torch.cuda.synchronize()
a.record()
for _ in range(20):
    prog.run()
b.record(); torch.cuda.synchronize()
print(f"n={B}: representation network {a.elapsed_time(b) / 20 * 1e3:.0f} us, {prog.n_kernels} launches")
h = torch.rand(B, 256, 4, 5, device="cuda")
m.search(h, None, 0)
a.record()
for _ in range(5):
    m.search(h, None, 0)
b.record(); torch.cuda.synchronize()
print(f"n={B}: search {a.elapsed_time(b) / 5:.2f} ms")
