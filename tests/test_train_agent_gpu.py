"""GPU tests of the learner-side drop-in (muzero-breakout_b200/src/agent.py, train.trunk_forward / accelerate_agent): the ResidualBlock
trunks of the dynamics / prediction networks trained by this library's kernels inside loss.backward(), against torch autograd on the same
modules (fp32 / TF32 off) and against the fp32 oracle restatement of the reference networks (oracle/networks.py).
Tolerances: the kernels compute in bf16 operands with fp32 accumulation (mixed-precision training, like autocast): against torch with the
same rounding points emulated the gradients agree to 3e-2 in relative L2 norm; against pure fp32 the forward agrees to 2-5e-2 of range and
the gradients in direction (cosine >= 0.9 ... 0.98: reduced-precision forwards flip ReLU masks) -- written next to each assert."""
import copy

import pytest
import torch

from oracle.networks import DEFAULT_MODEL_CFG, OracleAgent

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0))


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_trunk_forward_matches_torch_autograd():
    from muzero_breakout_b200 import train
    from muzero_breakout_b200.src.agent import ResidualBlock
    torch.manual_seed(3)
    blocks = torch.nn.ModuleList([ResidualBlock(256, "relu") for _ in range(3)]).cuda().train()
    with torch.no_grad():
        for m in blocks.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                # mask-stable BatchNorm parameters (as tests/test_train_gpu.py): pre-activations sit at +-3 sigma, so the bf16 forward
                # flips no ReLU mask and the comparison isolates the kernels' arithmetic
                m.weight.copy_(torch.rand(256) * 0.4 + 0.3); m.bias.copy_(torch.where(torch.arange(256) % 2 == 0, 3.0, -3.0))
    x = torch.rand(96, 256, 4, 5, device="cuda")
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    dy = torch.randn(96, 256, 4, 5, device="cuda")
    assert train.trunk_supported(blocks, xa)
    ref = copy.deepcopy(blocks)
    n0 = train._lib.launch_count()
    ya = train.trunk_forward(blocks, xa)
    # reference 1: the same modules on torch ops with the kernels' rounding points emulated (bf16 convolution operands, straight-through
    # gradient) -- isolates the kernels' arithmetic from the ReLU-mask flips any reduced-precision forward causes
    r16 = lambda t: t + (t.to(train.FWD_DTYPE).float() - t).detach()       # the forward operands' element type (fp16 by default)
    yb = xb
    for m in ref:
        xin = r16(yb)
        h = m.act(m.bn1(torch.nn.functional.conv2d(xin, r16(m.conv1.weight), m.conv1.bias, padding=1)))
        yb = m.act(m.bn2(torch.nn.functional.conv2d(r16(h), r16(m.conv2.weight), m.conv2.bias, padding=1)) + xin)
    (ya * dy).sum().backward(); (yb * dy).sum().backward()
    assert train._lib.launch_count() - n0 >= 3 * 10, "the library kernels did not run"
    assert float((ya.detach() - yb.detach()).abs().max() / yb.detach().abs().max()) <= 5e-3
    assert _rel(xa.grad, xb.grad) <= 3e-2 and _cos(xa.grad, xb.grad) >= 0.999, (_rel(xa.grad, xb.grad), _cos(xa.grad, xb.grad))
    for (name, pa), pb in zip(blocks.named_parameters(), ref.parameters()):
        if name.endswith("conv1.bias") or name.endswith("conv2.bias"):
            assert float(pa.grad.abs().max()) == 0.0          # exactly zero: a train-mode BatchNorm follows (torch returns rounding noise)
            continue
        assert _rel(pa.grad, pb.grad) <= 3e-2 and _cos(pa.grad, pb.grad) >= 0.999, (name, _rel(pa.grad, pb.grad))
    # reference 2: pure fp32 autograd (what the reference trainer computes): bf16 forward errors flip ReLU masks, so element-wise
    # agreement is looser; the directions agree
    ref32 = copy.deepcopy(blocks)
    for p_ in ref32.parameters():
        p_.grad = None
    xc = x.clone().requires_grad_()
    yc = xc
    for m in ref32:
        h = m.act(m.bn1(m.conv1(yc)))
        yc = m.act(m.bn2(m.conv2(h)) + yc)
    (yc * dy).sum().backward()
    assert float((ya.detach() - yc.detach()).abs().max() / yc.detach().abs().max()) <= 2e-2
    assert _cos(xa.grad, xc.grad) >= 0.98
    for (name, pa), pc in zip(blocks.named_parameters(), ref32.parameters()):
        if "conv" in name and name.endswith("weight"):
            assert _cos(pa.grad, pc.grad) >= 0.98, (name, _cos(pa.grad, pc.grad))
    for ma, mb in zip(blocks.modules(), ref.modules()):
        if isinstance(ma, torch.nn.BatchNorm2d):              # running statistics updated in place like nn.BatchNorm2d
            assert int(ma.num_batches_tracked) == int(mb.num_batches_tracked) == 1
            assert torch.allclose(ma.running_mean, mb.running_mean, atol=2e-3) and torch.allclose(ma.running_var, mb.running_var, rtol=2e-2, atol=1e-3)
    # eval mode / no_grad: plain torch ops, no library launches
    blocks.eval()
    n1 = train._lib.launch_count()
    with torch.no_grad():
        blocks[0](x)
    assert train._lib.launch_count() == n1


def _rollout(agent, frames, actions, K):
    """train_torch.py:487-528 _k_step_rollout on an agent (same call order: evaluate_state, then hidden_state_transition, K times)."""
    h = agent.create_hidden_state_root(frames)
    pol, val, rew = [], [], []
    for k in range(K):
        p, v = agent.evaluate_state(h)
        planes = torch.nn.functional.one_hot(actions[:, k], 3).float().view(-1, 3, 1, 1).expand(-1, -1, 4, 5)
        h, r = agent.hidden_state_transition(h, planes)
        pol.append(p); val.append(v); rew.append(r)
    return torch.stack(rew, 1), torch.stack(val, 1), torch.stack(pol, 1)


def test_dropin_agent_training_step_vs_oracle_agent():
    from muzero_breakout_b200 import train
    from muzero_breakout_b200.src.agent import MuZeroAgent
    cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4, device="cuda")
    torch.manual_seed(0)
    oracle = OracleAgent(cfg, device="cuda")
    torch.manual_seed(0)
    agent = MuZeroAgent(cfg)
    # same module tree, same keys, same creation order: identical initial weights from the same seed
    sa, so = agent.state_dict(), oracle.state_dict()
    assert list(sa.keys()) == list(so.keys())
    assert all(torch.equal(sa[k], so[k]) for k in sa)
    assert isinstance(agent.optimizer, train.Adam)
    agent.train_mode(); oracle.train()
    B, K = 24, 2
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = torch.rand(B, 64, 16, 20, device="cuda", generator=g)
    actions = torch.randint(0, 3, (B, K), device="cuda", generator=g)
    w = [torch.randn(B, K, n, device="cuda", generator=g) for n in (11, 11, 3)]
    agent.optimizer.zero_grad()
    outs_a = _rollout(agent, frames, actions, K)
    outs_o = _rollout(oracle, frames, actions, K)
    sum((o * wi).sum() for o, wi in zip(outs_a, w)).backward()
    sum((o * wi).sum() for o, wi in zip(outs_o, w)).backward()
    err = {name: float((a.detach() - o.detach()).abs().max() / o.detach().abs().max()) for a, o, name in zip(outs_a, outs_o, ("reward", "value", "policy"))}
    # yardstick: torch's own mixed precision (autocast bf16) of the same fp32 modules on the same batch.  Reduced-precision forwards flip
    # ReLU masks, and through 2 x 14 blocks x K steps at random-init weights that decorrelates the earliest layers' gradients for ANY 16-bit
    # path; the library's gradients must be as close to fp32 autograd as autocast's are (a wrong kernel gives a cosine near 0)
    ac = copy.deepcopy(oracle)
    for p_ in ac.parameters():
        p_.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs_c = _rollout(ac, frames, actions, K)
    sum((o.float() * wi).sum() for o, wi in zip(outs_c, w)).backward()
    # forward outputs: 6 + 2 x 14 bf16 blocks per unroll step at random-init weights with batch statistics over 24 samples and the min / max
    # normalisation of `_scale_state` in between -- the same yardstick: as close to the fp32 modules as torch's autocast is (x 1.5), or 1e-1
    err_ac = {name: float((c.detach().float() - o.detach()).abs().max() / o.detach().abs().max()) for c, o, name in zip(outs_c, outs_o, ("reward", "value", "policy"))}
    print(f"forward outputs vs the fp32 modules (max error / range): library {err}, torch autocast bf16 {err_ac}")
    for name in err:
        assert err[name] <= max(1e-1, 1.5 * err_ac[name]), f"{name}: {err[name]:.3f} (autocast {err_ac[name]:.3f})"
    # second yardstick: TF32 convolutions, what the reference's own training runs on a GPU (torch's cuDNN default)
    tf = copy.deepcopy(oracle)
    for p_ in tf.parameters():
        p_.grad = None
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        outs_t = _rollout(tf, frames, actions, K)
        sum((o * wi).sum() for o, wi in zip(outs_t, w)).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    err_tf = {name: float((c.detach() - o.detach()).abs().max() / o.detach().abs().max()) for c, o, name in zip(outs_t, outs_o, ("reward", "value", "policy"))}
    print(f"forward outputs, torch TF32: {err_tf}")
    worst, worst_ac, worst_tf = 1.0, 1.0, 1.0
    for (name, pa), po, pc, pt in zip(agent.named_parameters(), oracle.parameters(), ac.parameters(), tf.parameters()):
        if "res_blocks" in name and name.endswith("weight") and "conv" in name:   # the tensors the library kernels produce
            c, c_ac, c_tf = _cos(pa.grad, po.grad), _cos(pc.grad, po.grad), _cos(pt.grad, po.grad)
            worst, worst_ac, worst_tf = min(worst, c), min(worst_ac, c_ac), min(worst_tf, c_tf)
            assert c >= min(0.95, c_ac - 0.1), f"{name}: cosine vs fp32 autograd {c:.4f}, autocast bf16 reaches {c_ac:.4f}"
    print(f"worst cosine similarity of a trunk convolution's weight gradient vs fp32 autograd: library {worst:.4f}, torch autocast bf16 {worst_ac:.4f}, "
          f"torch TF32 {worst_tf:.4f}")
    # every other weight (stems, ConvBlocks, Linear heads, 128-channel blocks, BatchNorm affine parameters): library kernels as well
    # (train_layers.py), same yardstick
    worst_o, worst_o_ac, worst_name = 1.0, 1.0, ""
    for (name, pa), po, pc in zip(agent.named_parameters(), oracle.parameters(), ac.parameters()):
        trunk_conv = "res_blocks" in name and name.endswith("weight") and "conv" in name
        if trunk_conv or pa.grad is None or float(po.grad.abs().max()) < 1e-12 or (name.endswith("bias") and "conv" in name and "rep_net.blocks.0." not in name
                                                                                     and "rep_net.blocks.3." not in name):
            continue                                        # conv biases under a BatchNorm: exactly zero here, rounding noise in torch
        c, c_ac = _cos(pa.grad, po.grad), _cos(pc.grad, po.grad)
        if c < worst_o:
            worst_o, worst_o_ac, worst_name = c, c_ac, name
        assert c >= min(0.9, c_ac - 0.1), f"{name}: cosine vs fp32 autograd {c:.4f}, autocast bf16 reaches {c_ac:.4f}"
    print(f"worst cosine of any other parameter's gradient vs fp32 autograd: library {worst_o:.4f} ({worst_name}; torch autocast bf16 there {worst_o_ac:.4f})")
    v0 = [p._version for p in agent.parameters()]
    agent.optimizer.step()
    assert all(p._version > v for p, v in zip(agent.parameters(), v0))
    # the acting side packs the learner's state_dict as before
    from muzero_breakout_b200.src.networks import PackedNetworks
    agent.eval_mode()
    nets = PackedNetworks(agent, cfg, precision="f16")
    pol, val = nets.prediction(torch.rand(5, 256, 4, 5))
    assert torch.isfinite(pol).all() and torch.isfinite(val).all()


def test_accelerate_agent_patches_a_reference_shaped_module():
    from muzero_breakout_b200 import train
    cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4)
    torch.manual_seed(0)
    agent = OracleAgent(cfg, device="cuda")
    agent.optimizer = torch.optim.Adam(agent.parameters(), lr=2e-4, weight_decay=1e-4)
    plain = copy.deepcopy(agent)
    train.accelerate_agent(agent)
    assert isinstance(agent.optimizer, train.Adam) and agent.optimizer.lr == 2e-4 and agent.optimizer.weight_decay == 1e-4
    agent.train(); plain.train()
    h = torch.rand(16, 256, 4, 5, device="cuda")
    n0 = train._lib.launch_count()
    pa, va = agent.evaluate_state(h)
    pb, vb = plain.evaluate_state(h)
    assert train._lib.launch_count() - n0 >= 14 * 4
    assert float((pa.detach() - pb.detach()).abs().max() / pb.detach().abs().max()) <= 5e-2
    assert float((va.detach() - vb.detach()).abs().max() / vb.detach().abs().max()) <= 5e-2
    (pa.sum() + va.sum()).backward()
    assert agent.pred_net.res_blocks[0].conv1.weight.grad is not None


def _minibatch(n, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    states = torch.rand((n, 32, 16, 20), device="cuda", generator=g)
    planes = torch.randint(0, 3, (n, 32, 1, 1), device="cuda", generator=g).float().expand(-1, -1, 16, 20) / 3
    acts = torch.randint(0, 3, (n, K), device="cuda", generator=g)
    rew = torch.randint(-1, 2, (n, K), device="cuda", generator=g).float()
    val = (torch.rand((n, K), device="cuda", generator=g) - 0.5) * 8
    vis = torch.randint(1, 30, (n, K, 3), device="cuda", generator=g).float()
    return states, planes.contiguous(), acts, rew, val, vis


def test_graphed_train_step_equals_eager_steps():
    """train.GraphedTrainStep (the loop body of train_torch.py:385-417 as one CUDA-graph replay, optimizer step count on the device)
    against the same iterations run eagerly on a second agent with the same initial weights.  The captured kernels are the ones the eager
    path launches, so the FIRST step -- identical weights and inputs -- must give the same losses (1e-5) and the same gradients (flat
    gradient buffer, relative L2 <= 5e-2, see below), and it must count as exactly
    ONE update although capture needs a warm-up pass.  Later steps are only loosely comparable: Adam's first update is -lr * sign(g) for
    EVERY parameter, noise-level gradients included, so two runs that differ in the last bit of a gradient drift apart (bounded at 10 %)."""
    from muzero_breakout_b200 import train
    from muzero_breakout_b200.src.agent import MuZeroAgent
    cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4, device="cuda")
    K, n = 3, 48
    sup = torch.linspace(-5, 5, 11, device="cuda")
    torch.manual_seed(5); a = MuZeroAgent(cfg); a.train_mode()
    torch.manual_seed(5); b = MuZeroAgent(cfg); b.train_mode()
    w0 = b.dyn_net.res_blocks[3].conv1.weight.detach().clone()
    step = train.GraphedTrainStep(b, sup, K)
    la, lb = [], []
    for it in range(3):
        mb = _minibatch(n, K, 100 + it)
        a.optimizer.zero_grad()
        pr, pv, pp = train.k_step_rollout(a, mb[0], mb[1], mb[2], K)
        out = train.loss_fn(mb[3], pr, mb[4], pv, mb[5], pp, sup, K)
        out[0].backward()
        a.optimizer.step()
        la.append([float(o.detach()) for o in out])
        lb.append([float(o) for o in step(*mb)])
        if it == 0:
            ga, gb = a.optimizer.flat_grad, b.optimizer.flat_grad
            # not bit-equal: a last-bit difference in a torch / cuDNN layer (other algorithm under capture) that lands on a bf16 rounding
            # boundary of a trunk operand becomes a 4e-3 difference there and may flip ReLU masks downstream (measured 1e-2 overall)
            assert _rel(gb, ga) <= 5e-2, f"gradients of the first step: graph vs eager {_rel(gb, ga):.2e}"
            assert b.optimizer.step_count == 1 and int(b.optimizer._dev_state[0]) == 1, "the capture warm-up must not count as an update"
            wa, wb = a.dyn_net.res_blocks[3].conv1.weight.detach(), b.dyn_net.res_blocks[3].conv1.weight.detach()
            assert float((wb - w0).abs().max()) > 1e-5, "the graph did not update the parameters"
            same = ((wa - wb).abs() <= 1e-7).float().mean()          # the first Adam update is -lr * sign(g): equal wherever the sign agrees
            assert float(same) >= 0.95, f"trunk weights after the first update: {float(same):.4f} of the elements equal"
    print(f"losses eager {la} graph {lb}")
    assert step.replays == 3 and b.optimizer.step_count == 3 == a.optimizer.step_count
    assert int(b.optimizer._dev_state[0]) == 3, "device step counter"
    for it, (x, y) in enumerate(zip(la, lb)):
        for u, v in zip(x, y):
            assert abs(u - v) <= (1e-5 if it == 0 else 1e-1) * max(1.0, abs(u)), (la, lb)
    bn_a, bn_b = a.pred_net.res_blocks[5].bn2, b.pred_net.res_blocks[5].bn2
    assert int(bn_b.num_batches_tracked) == int(bn_a.num_batches_tracked) == 3 * K
    assert _rel(bn_b.running_mean, bn_a.running_mean) <= 1e-1 and _rel(bn_b.running_var, bn_a.running_var) <= 1e-1
    assert b.dyn_net.res_blocks[3].conv1.weight._version > 0
