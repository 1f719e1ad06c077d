"""Where a search() call spends its time (CUDA events, no profiler): python profiles/prof_search.py [trees]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision=os.environ.get("MZB_PREC", "f16"))
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
for _ in range(3): m.search(h, None, 0)
plan = next(iter(m._plans.values()))
def timed(fn, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("search() call           %.2f ms" % timed(lambda: m.search(h, None, 0)))
print("graph.replay() alone    %.2f ms" % timed(plan.graph.replay))
print("sim program x50 (eager) %.2f ms" % timed(lambda: [plan.sim_prog.run() for _ in range(50)], 2))
print("one sim program         %.3f ms" % timed(plan.sim_prog.run, 10))
def tree_pass():
    plan.tree.root(plan.value, plan.pi, plan.noise, 0.175, 1)
    for s in range(50):
        plan.tree.step(s, plan.reward, plan.value, plan.pi, 1)
print("tree root + 50 steps    %.2f ms" % timed(tree_pass, 2))
print("root program            %.3f ms" % timed(plan.root_prog.run, 5))
t0 = time.perf_counter(); 
for _ in range(5): m.search(h, None, 0)
torch.cuda.synchronize(); print("search() wall           %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
