"""Where one training-loop iteration of the drop-in agent spends its time: wall (CUDA events), GPU-busy time and the top kernels
(torch.profiler, CUPTI) at minibatch 512 x K = 5.   python profiles/prof_train_breakdown.py [minibatch] [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.agent import MuZeroAgent
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG
from muzero_breakout_b200.train import loss_fn

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


def rollout():
    h = agent.create_hidden_state_root(frames)
    pol, val, rew = [], [], []
    for k in range(K):
        p_, v_ = agent.evaluate_state(h)
        planes = torch.nn.functional.one_hot(actions[:, k], 3).float().view(-1, 3, 1, 1).expand(-1, -1, 4, 5)
        h, r_ = agent.hidden_state_transition(h, planes)
        pol.append(p_); val.append(v_); rew.append(r_)
    return torch.stack(rew, 1), torch.stack(val, 1), torch.stack(pol, 1)


def step():
    agent.optimizer.zero_grad()
    pr, pv, pp = rollout()
    loss = loss_fn(obs_r, pr, val_t, pv, visits, pp, supports, K)[0]
    loss.backward()
    agent.optimizer.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    step()
b.record(); torch.cuda.synchronize()
print(f"wall per step {a.elapsed_time(b) / 3:.1f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"GPU busy {tot / 1e3:.1f} ms in {sum(e.count for e in ev)} kernels / copies")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:28]:
    print(f"{e.device_time_total / 1e3:8.2f} ms {e.count:5d} x {e.device_time_total / e.count:8.1f} us  {e.key[:110]}")
