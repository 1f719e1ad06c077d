"""One EAGER training-loop iteration of the drop-in agent (minibatch 512 x K = 5) between cudaProfilerStart / Stop, for an ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/prof_train_eager_once.py
(the graphed iteration replays the same kernels; per-launch times under ncu are serialised and cold-cache)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.agent import MuZeroAgent
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG
from muzero_breakout_b200.train import k_step_rollout, loss_fn

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4, device="cuda")
g = torch.Generator(device=dev).manual_seed(11)
frames = torch.rand((mb, 64, 16, 20), device=dev, generator=g)
actions = torch.randint(0, 3, (mb, K), device=dev, generator=g)
obs_r = torch.randint(-1, 2, (mb, K), device=dev, generator=g).float()
val_t = (torch.rand((mb, K), device=dev, generator=g) - 0.5) * 8
visits = torch.randint(1, 30, (mb, K, 3), device=dev, generator=g).float()
supports = torch.linspace(-5, 5, 11, device=dev)
torch.manual_seed(0)
agent = MuZeroAgent(cfg)
agent.train_mode()


def step():
    agent.optimizer.zero_grad()
    pr, pv, pp = k_step_rollout(agent, frames[:, :32], frames[:, 32:], actions, K)
    loss_fn(obs_r, pr, val_t, pv, visits, pp, supports, K)[0].backward()
    agent.optimizer.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
