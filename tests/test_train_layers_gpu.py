"""GPU parity of the training-step layers around the residual trunks (muzero-breakout_b200/train_layers.py, csrc/train_layers.cu, the
generalised weight gradient of csrc/wgrad.cu) against torch's own float32 ops / autograd on the same inputs.  Reference layers:
src/networks.py:7-17 (ConvBlock), :43-92 (representation stems and pools), :117-122,295 (dynamics ConvBlock + action planes),
:138-149,200-223 (heads), :314-328 (_scale_state)."""
import os
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _nchw_cl(n, c, h, w, gen, scale=1.0):
    """an NCHW-shaped float32 tensor with channels_last strides (what the bridges exchange)"""
    return (torch.randn(n, h, w, c, device="cuda", generator=gen) * scale).permute(0, 3, 1, 2)


@pytest.mark.parametrize("cout,cin,k,n,H,W", [(256, 256, 3, 70, 4, 5), (128, 256, 3, 70, 4, 5), (128, 256, 1, 130, 4, 5), (256, 256, 1, 64, 4, 5),
                                              (256, 128, 3, 9, 16, 20), (128, 128, 3, 9, 16, 20), (128, 64, 3, 5, 16, 20), (128, 128, 3, 33, 8, 10)])
def test_wgrad_any_shape_vs_torch(cout, cin, k, n, H, W):
    from muzero_breakout_b200 import train
    g = torch.Generator(device="cuda").manual_seed(cout + cin + k + n)
    dy = torch.randn(n, H, W, cout, device="cuda", generator=g).bfloat16()
    x = torch.randn(n, H, W, cin, device="cuda", generator=g).bfloat16()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2), padding=k // 2)
    dw = train.conv_wgrad(dy, x, k)
    assert dw.shape == ref.shape
    assert _rel(dw, ref) <= 2e-4, _rel(dw, ref)
    # in-place accumulation, into a gradient tensor with extra input channels (the dynamics ConvBlock's 256 + 3)
    into = torch.full((cout, cin + 3, k, k), 0.5, device="cuda")
    train.conv_wgrad(dy, x, k, into)
    assert _rel(into[:, :cin] - 0.5, ref) <= 2e-4
    assert float((into[:, cin:] - 0.5).abs().max()) == 0.0
    # an fp16 forward activation next to bf16 gradients (converted inside the transpose)
    dw2 = train.conv_wgrad(dy, x.float().half(), k)
    ref2 = torch.nn.grad.conv2d_weight(x.float().half().bfloat16().float().permute(0, 3, 1, 2), (cout, cin, k, k), dy.float().permute(0, 3, 1, 2), padding=k // 2)
    assert _rel(dw2, ref2) <= 2e-4


def test_pool_linear_scale_vs_torch_autograd():
    from muzero_breakout_b200 import _lib, train_layers as TL
    g = torch.Generator(device="cuda").manual_seed(3)
    n0 = _lib.launch_count()
    # AvgPool2d(2, 2)
    xa = _nchw_cl(7, 256, 16, 20, g).requires_grad_()
    xb = xa.detach().clone().requires_grad_()
    pool = nn.AvgPool2d(kernel_size=(2, 2), stride=2)
    assert TL.pool_supported(pool, xa)
    ya, yb = TL.pool_forward(xa), pool(xb)
    w = torch.randn_like(yb)
    (ya * w).sum().backward(); (yb * w).sum().backward()
    assert ya.shape == yb.shape and _rel(ya, yb) <= 1e-6 and _rel(xa.grad, xb.grad) <= 1e-6
    # Flatten + Linear heads: 11 outputs of 256 x 20 features, 3 outputs of 128 x 20; ragged sample counts
    for n, C_, O in ((37, 256, 11), (130, 128, 3), (3, 128, 11)):
        lin_a = nn.Linear(C_ * 20, O).cuda()
        lin_b = nn.Linear(C_ * 20, O).cuda()
        lin_b.load_state_dict(lin_a.state_dict())
        xa = _nchw_cl(n, C_, 4, 5, g).requires_grad_()
        xb = xa.detach().clone().requires_grad_()
        assert TL.flatten_linear_supported(lin_a, xa)
        oa, ob = TL.flatten_linear(lin_a, xa), lin_b(torch.flatten(xb, 1))
        w = torch.randn_like(ob)
        (oa * w).sum().backward(); (ob * w).sum().backward()
        assert _rel(oa, ob) <= 1e-5, _rel(oa, ob)
        assert _rel(xa.grad, xb.grad) <= 1e-5 and _rel(lin_a.weight.grad, lin_b.weight.grad) <= 1e-5 and _rel(lin_a.bias.grad, lin_b.bias.grad) <= 1e-5
        # second call accumulates into the existing .grad like autograd does
        (TL.flatten_linear(lin_a, xa.detach()) * w).sum().backward(); (lin_b(torch.flatten(xb.detach(), 1)) * w).sum().backward()
        assert _rel(lin_a.weight.grad, lin_b.weight.grad) <= 1e-5 and _rel(lin_a.bias.grad, lin_b.bias.grad) <= 1e-5
    # _scale_state (networks.py:314-328) incl. the gradient through min and max
    def ref_scale(h):
        flat = h.view(h.shape[0], -1)
        s_min = flat.min(dim=1, keepdim=True)[0].view(-1, 1, 1, 1)
        s_max = flat.max(dim=1, keepdim=True)[0].view(-1, 1, 1, 1)
        return (h - s_min) / (s_max - s_min + 1e-8)
    xa = _nchw_cl(19, 256, 4, 5, g).requires_grad_()
    xb = xa.detach().contiguous().requires_grad_()
    assert TL.scale_supported(xa)
    ya, yb = TL.scale_state(xa), ref_scale(xb)
    w = torch.randn_like(yb)
    (ya * w).sum().backward(); (yb * w).sum().backward()
    assert _rel(ya, yb) <= 1e-6 and float(ya.min()) == 0.0 and abs(float(ya.max()) - 1.0) <= 1e-6
    assert _rel(xa.grad, xb.grad) <= 1e-5, _rel(xa.grad, xb.grad)
    assert _lib.launch_count() - n0 >= 2 + 3 * 4 + 2, "the library kernels did not run"


def test_convblock_and_conv_bridges_vs_torch_autograd():
    """ConvBlock (with and without action planes) and a plain convolution: forward + backward through the bridges against the float32 modules.
    fp16 forward operands, bf16 gradient operands: errors ~1e-3 / ~5e-3 of the norm; a wrong kernel gives O(1)."""
    from muzero_breakout_b200 import _lib, train_layers as TL
    from muzero_breakout_b200.src.agent import ConvBlock
    g = torch.Generator(device="cuda").manual_seed(5)
    torch.manual_seed(5)
    cases = [("head 3x3 256->128", 256, 128, 3, 0, 50, 4, 5), ("head 1x1 256->128", 256, 128, 1, 0, 50, 4, 5), ("head 1x1 256->256", 256, 256, 1, 0, 33, 4, 5),
             ("dynamics 259->256", 256, 256, 3, 3, 50, 4, 5)]
    for name, cin, cout, k, extra, n, H, W in cases:
        a = ConvBlock("relu", cin + extra, cout, 1, kernel_size=k, padding=k // 2).cuda().train()
        b = ConvBlock("relu", cin + extra, cout, 1, kernel_size=k, padding=k // 2).cuda().train()
        b.load_state_dict(a.state_dict())
        xa = _nchw_cl(n, cin, H, W, g).requires_grad_()
        xb = xa.detach().clone().requires_grad_()
        planes = None
        if extra:
            acts = torch.randint(0, extra, (n,), device="cuda", generator=g)
            planes = F.one_hot(acts, extra).float().view(n, extra, 1, 1).expand(-1, -1, H, W)       # the reference's expanded view (strides 0)
        assert TL.convblock_supported(a, xa, planes), name
        n0 = _lib.launch_count()
        ya = TL.convblock_forward(a, xa, planes)
        yb = b(xb if planes is None else torch.cat([xb, planes], dim=1))
        w = torch.randn_like(yb)
        (ya * w).sum().backward(); (yb * w).sum().backward()
        assert _lib.launch_count() - n0 >= 8, "the library kernels did not run"
        assert _rel(ya, yb) <= 5e-3, (name, _rel(ya, yb))
        assert _rel(xa.grad, xb.grad) <= 2e-2, (name, "dx", _rel(xa.grad, xb.grad))
        assert _rel(a.conv.weight.grad, b.conv.weight.grad) <= 2e-2, (name, "dw", _rel(a.conv.weight.grad, b.conv.weight.grad))
        if extra:
            assert _rel(a.conv.weight.grad[:, cin:], b.conv.weight.grad[:, cin:]) <= 2e-2, (name, "dw planes")
        assert _rel(a.bn.weight.grad, b.bn.weight.grad) <= 2e-2 and _rel(a.bn.bias.grad, b.bn.bias.grad) <= 2e-2, name
        assert float(a.conv.bias.grad.abs().max()) == 0.0                    # a train-mode BatchNorm follows
        assert int(a.bn.num_batches_tracked) == 1
        assert torch.allclose(a.bn.running_mean, b.bn.running_mean, atol=2e-3) and torch.allclose(a.bn.running_var, b.bn.running_var, rtol=2e-2, atol=1e-3)
    # plain convolutions: the representation network's stems at 16x20 (the first one's input needs no gradient)
    for name, cin, cout, needs in (("stem 64->128", 64, 128, False), ("stem 128->256", 128, 256, True)):
        a = nn.Conv2d(cin, cout, 3, 1, 1).cuda()
        b = nn.Conv2d(cin, cout, 3, 1, 1).cuda()
        b.load_state_dict(a.state_dict())
        xa = torch.randn(6, cin, 16, 20, device="cuda", generator=g) if not needs else _nchw_cl(6, cin, 16, 20, g)
        xa.requires_grad_(needs)
        xb = xa.detach().clone().requires_grad_(needs)
        assert TL.conv_supported(a, xa), name
        ya, yb = TL.conv_forward(a, xa), b(xb)
        w = torch.randn_like(yb)
        (ya * w).sum().backward(); (yb * w).sum().backward()
        assert _rel(ya, yb) <= 2e-3, (name, _rel(ya, yb))
        assert _rel(a.weight.grad, b.weight.grad) <= 1e-2, (name, "dw", _rel(a.weight.grad, b.weight.grad))
        assert _rel(a.bias.grad, b.bias.grad) <= 1e-5, (name, "db", _rel(a.bias.grad, b.bias.grad))
        if needs:
            assert _rel(xa.grad, xb.grad) <= 1e-2, (name, "dx", _rel(xa.grad, xb.grad))


def test_planes_conv_kernels_vs_torch():
    """the action-plane channels alone, with arbitrary (not one-hot) plane values and a contiguous plane tensor"""
    from muzero_breakout_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(9)
    n, H, W, A, cout, cin_total = 37, 4, 5, 3, 256, 259
    planes = torch.randn(n, A, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin_total, 3, 3, device="cuda", generator=g)
    z0 = torch.randn(n, H, W, cout, device="cuda", generator=g)
    z = z0.clone()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.mz_planes_conv_fwd(n, H, W, A, cout, cin_total, 256, planes.data_ptr(), *planes.stride(), w.data_ptr(), z.data_ptr(), st))
    ref = F.conv2d(planes, w[:, 256:], padding=1).permute(0, 2, 3, 1)
    assert _rel(z - z0, ref) <= 1e-5
    dz = torch.randn(n, H, W, cout, device="cuda", generator=g)
    dw = torch.zeros_like(w)
    scratch = torch.empty(L.mz_planes_wgrad_scratch_bytes(n, cout) // 4, device="cuda")
    _lib.check(L.mz_planes_conv_wgrad(n, H, W, A, cout, cin_total, 256, planes.data_ptr(), *planes.stride(), dz.data_ptr(), dw.data_ptr(), 1,
                                      scratch.data_ptr(), st))
    refw = torch.nn.grad.conv2d_weight(planes, (cout, A, 3, 3), dz.permute(0, 3, 1, 2), padding=1)
    assert _rel(dw[:, 256:], refw) <= 1e-5 and float(dw[:, :256].abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,k,n,H,W", [(256, 256, 3, 70, 4, 5), (256, 128, 1, 300, 4, 5), (128, 128, 3, 9, 16, 20), (64, 128, 3, 5, 16, 20), (256, 256, 3, 33, 8, 10)])
def test_conv_epilogue_batchnorm_statistics(cin, cout, k, n, H, W):
    """mz_op.bn_partial: the training-form convolution writes the per-channel sums / sums of squares of its float32 output per 32-row group
    (ragged sample counts: rows outside the tensor must not count); BatchNorm from those partial sums = BatchNorm with its own reduction pass."""
    from muzero_breakout_b200 import train, train_layers as TL
    g = torch.Generator(device="cuda").manual_seed(cin + cout + n)
    torch.manual_seed(n)
    conv = nn.Conv2d(cin, cout, k, 1, k // 2).cuda()
    kern = TL.ConvKernels(conv, None, False).refresh(conv)
    x16 = torch.randn(n, H, W, cin, device="cuda", generator=g).to(train.FWD_DTYPE)
    z, stats = kern.fwd(x16, True)
    assert stats is not None and stats.shape[1:] == (2, cout)
    zz = z.double().reshape(-1, cout)
    assert _rel(stats[:, 0].sum(0), zz.sum(0)) <= 1e-6 and _rel(stats[:, 1].sum(0), (zz * zz).sum(0)) <= 1e-6
    assert torch.equal(z, kern.fwd(x16))                           # the output itself does not depend on the side computation
    gam, bet = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    rm = [torch.zeros(cout, device="cuda") for _ in range(2)]
    rv = [torch.ones(cout, device="cuda") for _ in range(2)]
    a = train.bn_train_forward(z, gam, bet, None, "relu", running_mean=rm[0], running_var=rv[0], stats=stats)
    b = train.bn_train_forward(z, gam, bet, None, "relu", running_mean=rm[1], running_var=rv[1])
    for i, (u, v) in enumerate(zip(a, b)):                         # (y 16-bit, y float32, mean, invstd)
        tol = 2.0 ** -7 if i == 0 else 2e-6                        # the 16-bit copy: one rounding step where a value sits on a boundary
        assert float((u.float() - v.float()).abs().max()) <= tol * max(1.0, float(v.float().abs().max())), i
    assert torch.allclose(rm[0], rm[1], atol=1e-6) and torch.allclose(rv[0], rv[1], rtol=1e-5)


@pytest.mark.parametrize("cout,cin_total,cin,k", [(256, 256, 256, 3), (128, 256, 256, 1), (256, 259, 256, 3), (128, 64, 64, 3)])
def test_pack_conv_equals_the_torch_formulas(cout, cin_total, cin, k):
    """mz_pack_conv (one launch) against the permute / flip / convert chains it replaces: the forward operand [tap][cin/64][cout][64] and the
    data-gradient operand [tap][cout/64][cin][64] of the transposed, tap-flipped filter -- bit for bit."""
    from muzero_breakout_b200 import train
    torch.manual_seed(cout + cin_total + k)
    w = torch.randn(cout, cin_total, k, k, device="cuda")
    fwd = torch.empty((k * k, cin // 64, cout, 64), dtype=train.FWD_DTYPE, device="cuda")
    dg = torch.empty((k * k, cout // 64, cin, 64), dtype=torch.bfloat16, device="cuda")
    train.pack_conv(w, cin, None, fwd, dg)
    ws = w[:, :cin]
    want_fwd = train.ResidualBlockTrain._pack(ws, "cuda", train.FWD_DTYPE)
    wt = train.ConvDgrad.dgrad_filter(ws)                                  # (cin, cout, k, k)
    want_dg = wt.permute(0, 2, 3, 1).reshape(cin, k * k, cout // 64, 64).permute(1, 2, 0, 3).contiguous().to(torch.bfloat16)
    assert torch.equal(fwd, want_fwd) and torch.equal(dg, want_dg)
