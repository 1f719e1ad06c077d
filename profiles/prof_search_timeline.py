"""Kernel timeline of ONE search() CUDA-graph replay (CUPTI through torch.profiler): per kernel type the busy time and the idle gap in front of
it -- where the time between the kernels of a simulation step goes.   python profiles/prof_search_timeline.py [roots]"""
import os, sys, json, tempfile, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.src.mcts import MCTSSearchVec
from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, PackedNetworks, random_state_dict
from torch.profiler import ProfilerActivity, profile

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nets = PackedNetworks(random_state_dict(seed=0, bn_jitter=0.2), DEFAULT_MODEL_CFG, precision="f16")
cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": DEFAULT_MODEL_CFG,
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "use_graph": True, "output_device": "cuda"}}
m = MCTSSearchVec(cfg, nets, None)
h = torch.rand(B, 256, 4, 5, device="cuda")
for _ in range(6):
    m.search(h, None, 0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); m.search(h, None, 0); b.record(); torch.cuda.synchronize()
print(f"search (events, no profiler): {a.elapsed_time(b):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        m.search(h, None, 0)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
ev = ev[len(ev) // 2:]                      # the second search
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = collections.defaultdict(float); gap = collections.defaultdict(float); cnt = collections.Counter()
prev_end = ev[0]["ts"]
for e in ev:
    name = e["name"].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").split("(")[0][:44]
    busy[name] += e["dur"]; cnt[name] += 1
    gap[name] += max(0.0, e["ts"] - prev_end)
    prev_end = max(prev_end, e["ts"] + e["dur"])
print(f"{len(ev)} kernels, span {span / 1e3:.2f} ms, busy {sum(busy.values()) / 1e3:.2f} ms, idle {(span - sum(busy.values())) / 1e3:.2f} ms")
for k in sorted(busy, key=lambda k: -busy[k]):
    print(f"{k:42s} x{cnt[k]:4d}  busy {busy[k] / 1e3:8.2f} ms ({busy[k] / cnt[k]:8.1f} us each)  gap before {gap[k] / 1e3:6.2f} ms ({gap[k] / cnt[k]:5.1f} us each)")
