"""Where a simulation step goes at a SMALL batch (config.yaml default: 24 roots): every launch of the per-simulation
program timed on its own (CUDA events around `reps` back-to-back runs of that launch), the tree step, the whole program,
and the graph-replayed search.
    python profiles/prof_small.py 24
    MZB_LAT_MAX_SAMPLES=0 python profiles/prof_small.py 24        # tcgen05 trunk at every batch size"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200 import _lib
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, MzOp, PackedNetworks, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
m.search(h, None, 0)
plan = next(iter(m._plans.values()))
prog = plan.sim_prog
st = torch.cuda.current_stream().cuda_stream
OPS = ["conv", "pool", "scale", "head", "nchw_in", "nhwc_out"]


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


total = 0.0
for kind, item, cnt in prog._segs:
    if kind == "stack":
        us = timed(lambda: item.run(st))
        print(f"  trunk  {cnt:2d} layers ({'latency mode' if item.lat else 'tcgen05'}): {us:8.1f} us = {us / cnt:.2f} us/layer")
        total += us
    else:
        for i in range(cnt):
            one = (MzOp * 1)(item[i])
            us = timed(lambda: _lib.check(_lib.lib().mz_run(one, 1, B, st)))
            o = item[i]
            print(f"  {OPS[o.op]:8s} k{o.ksize} {o.cin}->{o.cout or o.nout}: {us:8.1f} us")
            total += us
us = timed(lambda: plan.tree.step(3, plan.reward, plan.value, plan.pi, 1))
print(f"  tree step: {us:8.1f} us")
total += us
print(f"sum of launches timed alone: {total:.1f} us;  whole program back to back: {timed(prog.run):.1f} us")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    m.search(h, None, 0)
b.record(); torch.cuda.synchronize()
sms = a.elapsed_time(b) / 5
print(f"n={B}: search (graph replay) {sms:.2f} ms = {sms * 1e3 / 50:.1f} us per simulation, {B * 50 / sms:.1f} k simulations/s")
