"""GPU parity tests of the training step's loss and optimizer ends (muzero-breakout_b200/train.py -> mz_loss / mz_adam,
csrc/train.cu) against the reference's golden vectors (tests/golden/train.npz: loss_fn of train_torch.py:33-66 under autograd,
torch.optim.Adam of networks.py:268) and the numpy oracle (oracle/train_oracle.py).  Floating point: the north star's fp32
tolerance 1e-5 relative for the losses and gradients; Adam moments bit-exact, parameters within 1 ulp."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as T

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "train.npz"))
K = int(G["K"])
RTOL = 1e-5


def _cuda(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_() if grad else t


def _run_loss(pr, pv, pp, obs, val, vis, supports, K):
    from muzero_breakout_b200.train import loss_fn
    tpr, tpv, tpp = _cuda(pr, True), _cuda(pv, True), _cuda(pp, True)
    out = loss_fn(observed_reward=_cuda(obs), predicted_reward=tpr, bootstrapped_reward=_cuda(val), predicted_value=tpv,
                  visit_counts=_cuda(vis), predicted_policy=tpp, target_transformation=_cuda(supports), K=K)
    out[0].backward()
    return np.array([float(x.detach()) for x in out], np.float32), tpr.grad.cpu().numpy(), tpv.grad.cpu().numpy(), tpp.grad.cpu().numpy()


def _close(got, want, name):
    assert got.shape == want.shape, name
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max(), (name, float(np.abs(got - want).max()), float(np.abs(want).max()))


@pytest.mark.parametrize("case", ["a", "b"])
def test_loss_fn_matches_reference_goldens(case):
    g = lambda k: G[f"{case}_{k}"]
    losses, d_r, d_v, d_p = _run_loss(g("pred_reward"), g("pred_value"), g("pred_policy"), g("obs_reward"), g("value_target"), g("visits"),
                                      G["supports"], K)
    np.testing.assert_allclose(losses, g("losses"), rtol=RTOL, atol=0)
    _close(d_r, g("d_reward"), "d_reward"); _close(d_v, g("d_value"), "d_value"); _close(d_p, g("d_policy"), "d_policy")


def test_loss_fn_minibatch_size_vs_oracle_and_deterministic():
    """config.yaml's minibatch (512 x K=5) plus a ragged row count; a (B, K, n) grad_output scaling; two calls bit-identical."""
    rng = np.random.default_rng(3)
    for B in (512, 77):
        pr, pv = (rng.standard_normal((B, K, 11)).astype(np.float32) * 2 for _ in range(2))
        pp = rng.standard_normal((B, K, 3)).astype(np.float32) * 2
        obs = rng.choice(np.array([0, 1, -1, 5, 6], np.float32), size=(B, K))
        val = ((rng.random((B, K)) - 0.5) * 40).astype(np.float32)
        vis = rng.multinomial(50, [0.2, 0.5, 0.3], size=(B, K)).astype(np.float32)
        got = _run_loss(pr, pv, pp, obs, val, vis, G["supports"], K)
        want_l = np.array(T.loss_fn(obs, pr, val, pv, vis, pp, G["supports"], K), np.float32)
        np.testing.assert_allclose(got[0], want_l, rtol=RTOL, atol=0)
        for a, b, name in zip(got[1:], T.loss_grads(obs, pr, val, pv, vis, pp, G["supports"], K), ("d_reward", "d_value", "d_policy")):
            _close(a, b, name)
        again = _run_loss(pr, pv, pp, obs, val, vis, G["supports"], K)
        assert all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(got, again))


def test_loss_fn_accepts_the_reference_call_forms_and_rejects_cpu():
    from muzero_breakout_b200.train import loss_fn

    class Transforms:                                    # stands in for utils.ScalarTransforms (has .supports and the bound method)
        supports = _cuda(G["supports"])

        def supports_representation(self, x):
            raise AssertionError("the fused kernel computes the targets itself")

    g = lambda k: _cuda(G[f"a_{k}"])
    tr = Transforms()
    a = loss_fn(g("obs_reward"), g("pred_reward"), g("value_target"), g("pred_value"), g("visits"), g("pred_policy"), tr.supports_representation, K)
    b = loss_fn(g("obs_reward"), g("pred_reward"), g("value_target"), g("pred_value"), g("visits"), g("pred_policy"), tr, K)
    assert all(float(x) == float(y) for x, y in zip(a, b))
    with pytest.raises(RuntimeError):
        loss_fn(*(torch.from_numpy(G[f"a_{k}"]) for k in ("obs_reward", "pred_reward", "value_target", "pred_value", "visits", "pred_policy")),
                torch.from_numpy(G["supports"]), K)


def _ulp(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_adam_matches_torch_goldens():
    from muzero_breakout_b200 import _lib
    L = _lib.lib()
    p, m, v = _cuda(G["adam_p0"]), torch.zeros(4099, device="cuda"), torch.zeros(4099, device="cuda")
    for s in range(5):
        grad = _cuda(G[f"adam_g{s}"])
        _lib.check(L.mz_adam(4099, p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), float(G["adam_lr"]), 0.9, 0.999, 1e-8, 1e-4, s + 1,
                             torch.cuda.current_stream().cuda_stream))
        assert np.array_equal(m.cpu().numpy(), G[f"adam_m{s + 1}"]) and np.array_equal(v.cpu().numpy(), G[f"adam_v{s + 1}"]), f"moments, step {s + 1}"
        u = _ulp(p.cpu().numpy(), G[f"adam_p{s + 1}"])
        assert u.max() <= 1 and (u != 0).mean() < 1e-2, f"parameters, step {s + 1}: max {u.max()} ulp"
        p.copy_(_cuda(G[f"adam_p{s + 1}"]))


def test_adam_class_on_a_module_vs_torch_optimizer():
    """The flat-buffer optimizer on a small module with ragged parameter sizes, 6 steps of real backward passes, against
    torch.optim.Adam on a CPU copy (weight decay 1e-4 as networks.py:268)."""
    from muzero_breakout_b200.train import Adam
    torch.manual_seed(1)
    make = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3, padding=1), torch.nn.BatchNorm2d(7), torch.nn.ReLU(), torch.nn.Flatten(),
                                       torch.nn.Linear(7 * 4 * 5, 11))
    ref = make()
    net = make().cuda()
    net.load_state_dict(ref.state_dict())
    opt_ref = torch.optim.Adam(ref.parameters(), lr=2e-4, weight_decay=1e-4)
    opt = Adam(net, lr=2e-4, weight_decay=1e-4)
    assert all(p.data_ptr() >= opt.flat_param.data_ptr() for p in net.parameters())
    x = torch.randn(6, 16, 3, 4, 5)
    for s in range(6):
        opt_ref.zero_grad(); opt.zero_grad()
        ref(x[s]).square().mean().backward()
        net(x[s].cuda()).square().mean().backward()
        v0 = [p._version for p in net.parameters()]
        opt_ref.step(); opt.step()
        assert all(p._version > v for p, v in zip(net.parameters(), v0)), "step() must bump the version counters (MCTSSearchVec re-pack check)"
    for (n, a), b in zip(net.named_parameters(), ref.parameters()):
        assert float((a.detach().cpu() - b.detach()).abs().max()) <= 2e-5 * float(b.detach().abs().max()), n
    # checkpoint interchange (train_torch.py:624,652): torch.optim.Adam's state_dict layout both ways
    sd, sd_ref = opt.state_dict(), opt_ref.state_dict()
    assert sd.keys() == sd_ref.keys() and sd["param_groups"][0]["params"] == sd_ref["param_groups"][0]["params"]
    for i, st in sd_ref["state"].items():
        assert float(sd["state"][i]["step"]) == float(st["step"]) == 6 and sd["state"][i]["exp_avg"].shape == st["exp_avg"].shape
        assert torch.allclose(sd["state"][i]["exp_avg"].cpu(), st["exp_avg"], rtol=1e-2, atol=1e-8)      # GPU (TF32 conv) vs CPU gradients
    fresh_ref = torch.optim.Adam(make().parameters(), lr=1.0)
    fresh_ref.load_state_dict({"state": {i: {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()} for i, st in sd["state"].items()},
                               "param_groups": sd["param_groups"]})        # torch accepts what we emit
    assert fresh_ref.param_groups[0]["lr"] == 2e-4
    net2 = make().cuda()
    opt2 = Adam(net2, lr=1.0, weight_decay=0.0)
    opt2.load_state_dict(sd_ref)                                            # and we accept what torch emits
    assert opt2.step_count == 6 and opt2.lr == 2e-4 and opt2.weight_decay == 1e-4
    for i, (p_, o) in enumerate(zip(opt2.params, opt2._offsets)):
        assert torch.equal(opt2.exp_avg_sq[o:o + p_.numel()].view(p_.shape).cpu(), sd_ref["state"][i]["exp_avg_sq"])


def test_adam_full_parameter_count_vs_oracle():
    """42 205 081 parameters (rep + dyn + pred), two updates: moments bit-exact, parameters within 1 ulp."""
    from muzero_breakout_b200 import _lib
    L = _lib.lib()
    n = 42_205_081
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(n, dtype=np.float32) * 0.05
    p, m, v = _cuda(p0), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    op, om, ov = p0, np.zeros(n, np.float32), np.zeros(n, np.float32)
    for s in range(2):
        g = rng.standard_normal(n, dtype=np.float32) * 1e-3
        grad = _cuda(g)
        _lib.check(L.mz_adam(n, p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), 2e-4, 0.9, 0.999, 1e-8, 1e-4, s + 1,
                             torch.cuda.current_stream().cuda_stream))
        op, om, ov = T.adam_step(op, g, om, ov, s + 1)
        assert np.array_equal(m.cpu().numpy(), om) and np.array_equal(v.cpu().numpy(), ov)
        assert _ulp(p.cpu().numpy(), op).max() == 0, "same operation order as the oracle: bit-exact"


def test_train_abi_argument_errors():
    from muzero_breakout_b200 import _lib
    L = _lib.lib()
    assert L.mz_loss(0, 5, 11, 3, *([None] * 13)) != 0 and b"rows" in L.mzb_last_error()
    assert L.mz_adam(16, None, None, None, None, 2e-4, 0.9, 0.999, 1e-8, 1e-4, 1, None) != 0
    t = torch.zeros(32, device="cuda")
    assert L.mz_adam(16, t.data_ptr() + 4, t.data_ptr(), t.data_ptr(), t.data_ptr(), 2e-4, 0.9, 0.999, 1e-8, 1e-4, 1, None) != 0
    assert L.mz_adam(16, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), 2e-4, 0.9, 0.999, 1e-8, 1e-4, 0, None) != 0


@pytest.mark.parametrize("case", [(512, 4, 5, 256, 256, 3), (77, 4, 5, 256, 128, 3), (33, 4, 5, 256, 256, 1), (6, 8, 10, 256, 256, 3), (4, 16, 20, 128, 256, 3), (2048, 4, 5, 256, 256, 3)],
                         ids=lambda c: "n%d_%dx%d_c%d-%d_k%d" % c)
def test_conv_dgrad_on_tensor_cores_vs_autograd(case):
    """The input gradient of the networks' convolutions (autograd of nn.Conv2d, networks.py:11,24-25) through the tcgen05 kernel with re-packed
    weights, against torch autograd in fp32 on the same bf16-rounded operands (the tolerance of the forward convolution test)."""
    from muzero_breakout_b200.train import ConvDgrad
    n, H, W, cin, cout, k = case
    g = torch.Generator().manual_seed(n + cin + cout + k)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16()
    dy = torch.randn(n, cout, H, W, generator=g).bfloat16()
    x = torch.zeros(n, cin, H, W, requires_grad=True)
    torch.nn.functional.conv2d(x, w.float(), padding=k // 2).backward(dy.float())
    got = ConvDgrad(w)(dy.permute(0, 2, 3, 1).contiguous().cuda()).permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    err = float((got - x.grad).abs().max() / x.grad.abs().max())
    assert err <= 2e-4, f"dgrad rel err {err:.2e}"


@pytest.mark.parametrize("dts", [(torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float16), (torch.float16, torch.float16)], ids=["bf16", "dy_bf16_x_f16", "f16"])
@pytest.mark.parametrize("case", [(512, 4, 5, 3), (77, 4, 5, 3), (1, 4, 5, 3), (300, 4, 5, 1), (40, 8, 10, 3), (2560, 4, 5, 3)], ids=lambda c: "n%d_%dx%d_k%d" % c)
def test_conv_wgrad_on_tensor_cores_vs_autograd(case, dts):
    """The weight gradient of the trunks' 256 -> 256 convolutions (tcgen05, K = samples x pixels, csrc/wgrad.cu) against torch autograd in
    fp32 on the same 16-bit-rounded operands; deterministic (fixed-order split reduction).  dy bf16 with x fp16 is the form of a training step:
    fp16 activations from the forward pass (rounded to bf16 inside their transpose: the reference sees that rounding too), bf16 gradients."""
    from muzero_breakout_b200.train import conv_wgrad
    n, H, W, k = case
    g = torch.Generator().manual_seed(n + H + k)
    x = torch.randn(n, 256, H, W, generator=g).to(dts[1])
    dy = torch.randn(n, 256, H, W, generator=g).to(dts[0])
    w = torch.zeros(256, 256, k, k, requires_grad=True)
    x_seen = x.to(dts[0]) if dts[0] != dts[1] else x
    torch.nn.functional.conv2d(x_seen.float(), w, padding=k // 2).backward(dy.float())
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    got = conv_wgrad(cl(dy), cl(x), k)
    assert torch.isfinite(got).all()
    err = float((got.cpu() - w.grad).abs().max() / w.grad.abs().max())
    assert err <= 2e-4, f"wgrad rel err {err:.2e}"
    assert torch.equal(got, conv_wgrad(cl(dy), cl(x), k))


@pytest.mark.parametrize("case", [(512, 4, 5, 256, True, "relu"), (77, 4, 5, 256, False, "relu"), (3, 16, 20, 128, True, "none"), (64, 4, 5, 64, False, "leaky_relu"),
                                  (2560, 4, 5, 256, True, "relu")], ids=lambda c: "n%d_%dx%d_c%d_res%d_%s" % c)
def test_bn_train_forward_backward_vs_torch(case):
    """Training-mode BatchNorm2d (+ residual) + activation, forward (output, saved statistics, running statistics) and backward (d input, d gamma,
    d beta, d residual) against torch's nn.BatchNorm2d in train mode on the CPU in fp32: the north star's fp32 tolerance, 1e-5 of each tensor's range."""
    from muzero_breakout_b200.train import bn_train_backward, bn_train_forward
    n, H, W, C, use_res, act = case
    g = torch.Generator().manual_seed(n + C)
    z = (torch.randn(n, C, H, W, generator=g) * 1.5 + torch.randn(1, C, 1, 1, generator=g)).requires_grad_()
    res16 = torch.randn(n, C, H, W, generator=g).bfloat16() if use_res else None
    res = res16.float().requires_grad_() if use_res else None
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5); bn.bias.copy_(torch.randn(C, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1); bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bn.train()
    pre = bn(z) + (res if use_res else 0)
    want = {"relu": torch.relu, "none": lambda t: t, "leaky_relu": torch.nn.functional.leaky_relu}[act](pre)
    dy = torch.randn(n, C, H, W, generator=g)
    want.backward(dy)

    cl = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().cuda()
    back = lambda t: t.float().cpu().permute(0, 3, 1, 2)
    rm, rv = rm0.cuda(), rv0.cuda()
    gam, bet = bn.weight.detach().cuda(), bn.bias.detach().cuda()
    y16, y32, mean, invstd = bn_train_forward(cl(z), gam, bet, res=cl(res16) if use_res else None, act=act, running_mean=rm, running_var=rv)
    dz, dz16, dgamma, dbeta, dres = bn_train_backward(cl(z), cl(dy), gam, bet, mean, invstd, res=cl(res16) if use_res else None, act=act)
    def tol(a, b, t, name, keep=None):
        d = (a - b).abs()
        if keep is not None:
            d = d * keep
        assert float(d.max()) <= t * float(b.abs().max()), f"{name}: {float(d.max() / b.abs().max()):.2e}"

    # an activation mask may legitimately flip where the pre-activation is within rounding of zero: those elements are left out
    keep = (pre.detach().abs() > 1e-4).float()
    tol(back(y32), want.detach(), 1e-5, "y")
    tol(back(y16), want.detach(), 5e-3, "y (bf16)")
    tol(rm.cpu(), bn.running_mean, 1e-5, "running_mean"); tol(rv.cpu(), bn.running_var, 1e-5, "running_var")
    tol(back(dz), z.grad, 2e-5, "dz", keep)
    tol(back(dz16), z.grad, 5e-3, "dz (bf16)", keep)
    tol(dgamma.cpu(), bn.weight.grad, 2e-5, "dgamma"); tol(dbeta.cpu(), bn.bias.grad, 2e-5, "dbeta")
    if use_res:
        tol(back(dres), res.grad, 1e-5, "dres", keep)
    assert keep.mean() > 0.999
    if use_res:
        # the form of a training step with fp16 forward operands: an fp16 residual, fp16 y, bf16 dz (mz_bn_train_bwd_mixed)
        r16 = res16.float().half()
        if torch.equal(r16.float(), res16.float()):        # bf16 values that fp16 holds exactly: same reference
            y16h, y32h, mean_h, invstd_h = bn_train_forward(cl(z), gam, bet, res=cl(r16), act=act, running_mean=rm0.cuda(), running_var=rv0.cuda(),
                                                            out_dtype=torch.float16)
            dz_h, dz16_h, dg_h, db_h, dres_h = bn_train_backward(cl(z), cl(dy), gam, bet, mean_h, invstd_h, res=cl(r16), act=act)
            assert y16h.dtype == torch.float16 and dz16_h.dtype == torch.bfloat16
            assert torch.equal(y32h, y32) and torch.equal(dz_h, dz) and torch.equal(dz16_h, dz16) and torch.equal(dg_h, dgamma) and torch.equal(dres_h, dres)
            tol(back(y16h), want.detach(), 1e-3, "y (fp16)")


@pytest.mark.parametrize("n", [512, 77])
def test_residual_block_training_step_vs_torch(n):
    """Forward + backward of one ResidualBlock in train mode (networks.py:19-35) through this library's kernels only (tcgen05 convolutions forward /
    dgrad / wgrad + the BatchNorm kernels) against the torch restatement in fp32 on the same bf16-rounded weights and input.  bf16 storage of the two
    intermediate activations and of the two d z operands bounds the agreement (same 3e-2-of-range bound as the bf16 acting networks)."""
    from muzero_breakout_b200.train import ResidualBlockTrain
    from oracle.networks import _Res
    torch.manual_seed(n)
    blk = _Res(256, "relu")
    with torch.no_grad():
        for m in (blk.conv1, blk.conv2):
            m.weight.copy_(m.weight.bfloat16().float())
        for m in (blk.bn1, blk.bn2):
            # pre-activations well away from zero (half of the channels always active, half always masked): with pre-activations AT zero the
            # bf16 rounding of the stored activations flips ReLU masks, and a flipped element moves a gradient by O(1) -- a property of the
            # precision, measured separately below, not of the kernels' arithmetic
            m.weight.copy_(torch.rand(256) * 0.4 + 0.3); m.bias.copy_(torch.where(torch.arange(256) % 2 == 0, 3.0, -3.0))
    blk.train()
    x = torch.rand(n, 256, 4, 5).bfloat16().float().requires_grad_()
    y = blk(x)
    dy = torch.randn(n, 256, 4, 5)
    y.backward(dy)
    sd = {k: v.detach() for k, v in blk.state_dict().items()}
    ours = ResidualBlockTrain(sd["conv1.weight"], sd["conv1.bias"], sd["bn1.weight"], sd["bn1.bias"], sd["conv2.weight"], sd["conv2.bias"], sd["bn2.weight"],
                              sd["bn2.bias"])
    cl = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().cuda()
    y16, y32 = ours.forward(cl(x).bfloat16())
    dx, grads = ours.backward(cl(dy))
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    back = lambda t: t.float().cpu().permute(0, 3, 1, 2)
    errs = {"y": rel(back(y32), y.detach()), "dx": rel(back(dx), x.grad)}
    for k, gk in grads.items():
        mod, par = k.split(".")
        errs[k] = rel(gk.cpu(), getattr(getattr(blk, mod), par).grad)
    print(errs)
    assert max(errs.values()) <= 1e-2, errs      # measured 1e-3 ... 7e-3
    for i, bn in enumerate((blk.bn1, blk.bn2)):
        assert rel(ours.running_mean[i].cpu(), bn.running_mean) <= 2e-2 and rel(ours.running_var[i].cpu(), bn.running_var) <= 2e-2
    # the convolution biases feed a BatchNorm: their gradient is zero up to rounding, which is why the block does not compute it
    assert float(blk.conv1.bias.grad.abs().max()) <= 1e-3 * float(blk.bn1.bias.grad.abs().max())


def test_trunk_training_step_chain_and_graph_replay():
    """Three chained ResidualBlocks, forward + backward, eager vs CUDA-graph replay bit-identical, and against the torch restatement (mask-stable
    BatchNorm parameters as above; the bf16 rounding points accumulate over the blocks)."""
    from muzero_breakout_b200.train import TrunkTrain
    from oracle.networks import _Res
    torch.manual_seed(5)
    blocks = torch.nn.Sequential(*[_Res(256, "relu") for _ in range(3)])
    with torch.no_grad():
        for blk in blocks:
            for m in (blk.conv1, blk.conv2):
                m.weight.copy_(m.weight.bfloat16().float())
            for m in (blk.bn1, blk.bn2):
                m.weight.copy_(torch.rand(256) * 0.4 + 0.3); m.bias.copy_(torch.where(torch.arange(256) % 2 == 0, 3.0, -3.0))
    blocks.train()
    n = 130
    x = torch.rand(n, 256, 4, 5).bfloat16().float().requires_grad_()
    y = blocks(x)
    dy = torch.randn(n, 256, 4, 5)
    y.backward(dy)
    sd = {k: v.detach() for k, v in blocks.state_dict().items()}
    trunk = TrunkTrain.from_state_dict(sd, [f"{i}." for i in range(3)])
    cl = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().cuda()
    x16, dyc = cl(x).bfloat16(), cl(dy)
    y16, y32, dx, grads = trunk.step(x16, dyc)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    back = lambda t: t.float().cpu().permute(0, 3, 1, 2)
    errs = {"y": rel(back(y32), y.detach()), "dx": rel(back(dx), x.grad)}
    for i, g in enumerate(grads):
        for k, gk in g.items():
            mod, par = k.split(".")
            errs[f"{i}.{k}"] = rel(gk.cpu(), getattr(getattr(blocks[i], mod), par).grad)
    print(errs)
    assert max(errs.values()) <= 3e-2, errs
    trunk2 = TrunkTrain.from_state_dict(sd, [f"{i}." for i in range(3)])
    for _ in range(2):                                             # capture, then a second replay
        gy16, gy32, gdx, ggrads = trunk2.step(x16, dyc, graph=True)
    assert torch.equal(gy16, y16) and torch.equal(gdx, dx)
    assert all(torch.equal(a[k], b[k]) for a, b in zip(ggrads, grads) for k in a)
