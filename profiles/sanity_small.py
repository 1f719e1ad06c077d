"""Smallest end-to-end exercise of every kernel (for compute-sanitizer memcheck): env reset/step/ingest/gray, one tiny
search (tree + fused trunk + heads + scale), one acting move."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.acting import Actor
from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict

ENV = dict(n_parallel=37, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)
env = BreakoutEnvironment(ENV)
torch.manual_seed(0)
s, _ = env.reset()
d = torch.zeros(37, dtype=torch.bool)
for t in range(20):
    s, r, d, v = env.step(s.clone() if t % 5 == 0 else s, torch.randint(0, 3, (37,)), d)
env.gray(s)
for prec in ("bf16", "f32"):
    nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=prec)
    cfg = {"num_simulations": 5, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
           "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": False, "output_device": "cuda"}}
    m = MCTSSearchVec(cfg, nets, None)
    v, n = m.search(torch.rand(7, 256, 4, 5), None, 0)
    assert int(n.sum()) == 35
    h = nets.representation(torch.rand(3, 64, 16, 20))
envc = BreakoutEnvironment(dict(ENV, n_parallel=7, output_device="cuda"))
out = Actor(envc, m, max_moves=2).run_episode()
torch.cuda.synchronize()
print("sanity ok", out["action"].shape)
