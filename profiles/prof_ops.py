"""Per-op CUDA-event timing of one MCTS simulation step (no profiler): python profiles/prof_ops.py [trees]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, MzOp, PackedNetworks, Program, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
m.search(h, None, 0)
plan = next(iter(m._plans.values()))
names = {0: "conv", 1: "pool", 2: "scale", 3: "head", 4: "nchw_in", 5: "nhwc_out"}
tot = {}
ops = plan.sim_prog.ops
# whole program first (steady state), then op by op back to back in program order
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
print("whole sim program: %.3f ms" % timed(plan.sim_prog.run))
evs = []
for rep in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
    evs[0].record()
    for i, o in enumerate(ops):
        p = Program(B); p.ops = [o]; p.keep = plan.sim_prog.keep
        p.run(); evs[i + 1].record()
    torch.cuda.synchronize()
for i, o in enumerate(ops):
    t = evs[i].elapsed_time(evs[i + 1])
    key = names[o.op] + ("" if o.op else f"_k{o.ksize}_n{o.cout}")
    tot.setdefault(key, []).append(t)
for k, v in tot.items():
    print(f"{k:16s} n={len(v):3d} total {sum(v):7.3f} ms  avg {sum(v)/len(v)*1e3:7.1f} us  min {min(v)*1e3:7.1f} max {max(v)*1e3:7.1f}")
t0 = timed(lambda: plan.tree.step(1, plan.reward, plan.value, plan.pi, 1))
print("tree step: %.1f us" % (t0 * 1e3))
