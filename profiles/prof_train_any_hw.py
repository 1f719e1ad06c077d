import os, sys, copy, torch
os.environ["MZB_TRAIN_ANY_HW"] = "1"
sys.path.insert(0, "/root/repo")
from muzero_breakout_b200 import train
from muzero_breakout_b200.src.agent import ResidualBlock
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for (H, W, n) in ((4, 5, 96), (8, 10, 64), (16, 20, 32), (16, 20, 512)):
    torch.manual_seed(3)
    blocks = torch.nn.ModuleList([ResidualBlock(256, "relu") for _ in range(1)]).cuda().train()
    with torch.no_grad():
        for m in blocks.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(256) * 0.4 + 0.3); m.bias.copy_(torch.where(torch.arange(256) % 2 == 0, 3.0, -3.0))
    x = torch.rand(n, 256, H, W, device="cuda")
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    dy = torch.randn(n, 256, H, W, device="cuda")
    ref = copy.deepcopy(blocks)
    assert train.trunk_supported(blocks, xa)
    ya = train.trunk_forward(blocks, xa)
    r16 = lambda t: t + (t.bfloat16().float() - t).detach()
    yb = xb
    for m in ref:
        xin = r16(yb)
        h = m.act(m.bn1(torch.nn.functional.conv2d(xin, r16(m.conv1.weight), m.conv1.bias, padding=1)))
        yb = m.act(m.bn2(torch.nn.functional.conv2d(r16(h), r16(m.conv2.weight), m.conv2.bias, padding=1)) + xin)
    (ya * dy).sum().backward(); (yb * dy).sum().backward()
    out = {"y": rel(ya, yb), "dx": rel(xa.grad, xb.grad)}
    for (na, pa), (nb, pb) in zip(blocks.named_parameters(), ref.named_parameters()):
        if "conv" in na and "bias" in na: continue
        out[na] = rel(pa.grad, pb.grad)
    print(H, W, n, {k: f"{v:.2e}" for k, v in out.items()})
