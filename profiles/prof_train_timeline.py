"""Timeline of ONE graphed training iteration (train.GraphedTrainStep replay, minibatch 512 x K = 5) from CUPTI kernel records: span, busy time
per stream, the phases (representation forward, unroll forward, backward, batched weight-gradient flush, Adam) by the first / last launch of
marker kernels, and the top kernels.   python profiles/prof_train_timeline.py [minibatch] [K]"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.agent import MuZeroAgent
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG
from muzero_breakout_b200.train import GraphedTrainStep

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
step = GraphedTrainStep(agent, supports, K)
args = (frames[:, :32], frames[:, 32:], actions, obs_r, val_t, visits)
for _ in range(3):
    step(*args)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(*args)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
span = (max(e.time_range.end for e in ev) - t0) / 1e3
print(f"one replay: {len(ev)} kernels / copies, span {span:.2f} ms")
busy = defaultdict(float)
for e in ev:
    busy[getattr(e, "device_resource_id", getattr(e, "stream", 0))] += (e.time_range.end - e.time_range.start) / 1e3
print("busy ms per stream:", {k: round(v, 2) for k, v in sorted(busy.items(), key=lambda kv: -kv[1])}, "sum", round(sum(busy.values()), 2))
# union of busy intervals = time with at least one kernel running
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
u, cs, ce = 0.0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > ce:
        u += ce - cs
        cs, ce = a, b
    else:
        ce = max(ce, b)
u += ce - cs
print(f"at least one kernel running: {u / 1e3:.2f} ms of {span:.2f}")


def phase(sub):
    hit = [e for e in ev if sub in e.name]
    if not hit:
        return None
    return (hit[0].time_range.start - t0) / 1e3, (hit[-1].time_range.end - t0) / 1e3, len(hit)


for name in ("nchw_in_kernel", "conv_tc_kernel<128", "pool2_fwd", "scale_fwd", "linear_fwd", "loss_kernel", "linear_bwd_data", "scale_bwd", "bn_bwd_apply",
             "pool2_bwd", "wgrad_transpose", "wgrad_kernel", "wgrad_reduce", "planes_dw_kernel", "adam"):
    ph = phase(name)
    if ph:
        print(f"  {name:22s} first {ph[0]:7.2f} ms  last {ph[1]:7.2f} ms  x{ph[2]}")
tot = defaultdict(lambda: [0.0, 0])
for e in ev:
    k = e.name.split("(")[0][-70:]
    tot[k][0] += (e.time_range.end - e.time_range.start) / 1e3
    tot[k][1] += 1
for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {ms:7.2f} ms {n:5d} x {1e3 * ms / n:7.1f} us  {k}")
