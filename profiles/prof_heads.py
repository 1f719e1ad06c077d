import os, sys, torch
sys.path.insert(0, '/root/repo')
from muzero_breakout_b200.src.networks import F16, OP_HEAD, Program
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.rand(n, 20, 256).half().cuda(); x2 = torch.rand(n, 20, 256).half().cuda()
w = {k: torch.rand(o, f).cuda() for k, (o, f) in {"r": (11, 5120), "p": (3, 2560), "v": (11, 2560)}.items()}
b = {k: torch.rand(v.shape[0]).cuda() for k, v in w.items()}
o1, o2, o3 = torch.empty(n, device="cuda"), torch.empty(n, 3, device="cuda"), torch.empty(n, device="cuda")
pr = Program(n); pr.add(op=OP_HEAD, dtype=F16, H=4, W=5, cin=256, nout=11, head_mode=1, src=x, w=w["r"], shift=b["r"], out=o1)
pp = Program(n)
pp.add(op=OP_HEAD, dtype=F16, H=4, W=5, cin=128, cout=256, nout=3, head_mode=2, src=x2, w=w["p"], shift=b["p"], out=o2)
pp.add(op=OP_HEAD, dtype=F16, H=4, W=5, cin=128, cout=256, nout=11, head_mode=1, src=x2.view(-1)[128:], w=w["v"], shift=b["v"], out=o3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, p in (("reward head", pr), ("policy+value heads", pp)):
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    print(f"{name}: n={n} simt={os.environ.get('MZB_HEAD_SIMT','0')} median {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us")
ts = []
for it in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o1[:8].zero_(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"(event pair around a one-CTA fill kernel: {sorted(ts)[6]:.1f} us)")
