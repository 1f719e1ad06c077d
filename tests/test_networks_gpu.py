"""GPU tests of the network kernels (csrc/nets.cu, csrc/conv_tc.cu through mz_run / PackedNetworks).

Tolerances (max abs error / max abs reference value, i.e. relative to the tensor's range):
  fp32 path  vs outputs of the reference networks (golden)          1e-5   (BASELINE.json north_star)
  fp16 tensor-core pipeline (the search's default: fp16 operands, fp32 accumulation, fp16 + e4m3 residual stream)
      vs the reference networks (golden + the fp32 oracle at 300 samples)        1e-3   (BASELINE.json north_star)
      -- for the fused trunk (conv_stack.cu), one launch per layer (conv_tc.cu) and the latency trunk (conv_lat.cu)
  tcgen05 conv vs fp32 torch conv on the same 16-bit operands        2e-4 on the fp32 side output
  bf16 pipeline vs the same pipeline on CUDA cores (same bf16 rounding points)   2e-2 (a few bf16 ulps)
  bf16 pipeline vs the fp32 reference networks (golden)              reported, bounded at 1e-2 (bf16 has an 8-bit
      mantissa: the convolution operands alone cost 3-6e-3; see DESIGN.md "precision")
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import perturb_bn
from oracle.networks import OracleAgent

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu() if hasattr(a, "detach") else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if hasattr(b, "detach") else torch.as_tensor(b).double()
    assert a.shape == b.shape, f"{tuple(a.shape)} vs {tuple(b.shape)}"
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def agent():
    torch.manual_seed(0)
    a = OracleAgent()
    perturb_bn(a, 1)
    a.eval_mode()
    return a


@pytest.fixture(scope="module")
def rec(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts_real.npz"))
    return {k: g[k] for k in g.files}


def _planes(actions):
    p = torch.zeros(len(actions), 3, 4, 5)
    p[torch.arange(len(actions)), torch.as_tensor(actions)] = 1
    return p


def test_fp32_path_matches_reference_outputs(agent, rec):
    from muzero_breakout_b200.src.networks import PackedNetworks
    nets = PackedNetworks(agent, agent.cfg, precision="f32")
    tol = 1e-5
    hidden = nets.representation(torch.from_numpy(rec["rep_in"]))
    assert rel(hidden, rec["hidden"]) <= tol, f"root latent {rel(hidden, rec['hidden']):.2e}"
    pol, val = nets.prediction(torch.from_numpy(rec["hidden"]))
    assert rel(pol, rec["root_policy_logits"]) <= tol and rel(val, rec["root_value_logits"]) <= tol
    h2, rew = nets.dynamics(torch.from_numpy(rec["hidden"]), _planes(rec["dyn_actions"]))
    assert rel(h2, rec["dyn_h"]) <= tol, f"dynamics latent {rel(h2, rec['dyn_h']):.2e}"
    assert rel(rew, rec["dyn_reward_logits"]) <= tol
    pol2, val2 = nets.prediction(torch.from_numpy(rec["dyn_h"]))
    assert rel(pol2, rec["pred_policy_logits"]) <= tol and rel(val2, rec["pred_value_logits"]) <= tol
    v = nets.inverted_softmax_expectation(torch.from_numpy(rec["root_value_logits"]).cuda())
    assert rel(v, rec["v_root"]) <= tol


def test_fp32_path_ragged_batches(agent):
    """batch sizes that are not multiples of any tile (1, 7, 61) against the torch fp32 oracle."""
    from muzero_breakout_b200.src.networks import PackedNetworks
    nets = PackedNetworks(agent, agent.cfg, precision="f32")
    g = torch.Generator().manual_seed(3)
    for n in (1, 7, 61):
        h = torch.rand(n, 256, 4, 5, generator=g)
        acts = torch.randint(0, 3, (n,), generator=g)
        with torch.no_grad():
            oh, orew = agent.hidden_state_transition(h, _planes(acts))
            opol, oval = agent.evaluate_state(h)
        h2, rew = nets.dynamics(h, _planes(acts))
        pol, val = nets.prediction(h)
        for got, want, what in ((h2, oh, "latent"), (rew, orew, "reward"), (pol, opol, "policy"), (val, oval, "value")):
            assert rel(got, want) <= 1e-5, f"n={n} {what}: {rel(got, want):.2e}"


CONV_CASES = [  # (n, H, W, cin, cout, ksize, residual, act_bias)
    (1, 4, 5, 256, 256, 3, False, False), (25, 4, 5, 256, 256, 3, True, False), (26, 4, 5, 256, 256, 3, True, True),
    (300, 4, 5, 256, 256, 3, True, False), (7, 4, 5, 256, 128, 3, False, False), (9, 4, 5, 256, 256, 1, False, False),
    (9, 4, 5, 256, 128, 1, False, False), (5, 16, 20, 64, 128, 3, False, False), (4, 16, 20, 128, 256, 3, True, False),
    (5, 8, 10, 256, 256, 3, True, False), (3000, 4, 5, 256, 256, 3, True, False),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "n%d_%dx%d_c%d-%d_k%d_r%d_a%d" % tuple(int(x) for x in c))
def test_tcgen05_conv_vs_torch(case):
    """One convolution op on the tensor cores against torch's fp32 conv on the same bf16-rounded operands."""
    from muzero_breakout_b200.src.networks import ACT, BF16, OP_CONV, Program
    n, H, W, cin, cout, k, use_res, use_ab = case
    g = torch.Generator().manual_seed(n * 131 + cin + cout + k)
    x = (torch.randn(n, cin, H, W, generator=g)).bfloat16()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16()
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(n, cout, H, W, generator=g).bfloat16() if use_res else None
    ab = torch.randn(3, H * W, cout, generator=g) * 0.2 if use_ab else None
    idx = torch.randint(0, 3, (n,), generator=g, dtype=torch.int32)
    want = F.conv2d(x.float(), w.float(), padding=k // 2)
    if use_ab:
        want = want + ab[idx.long()].view(n, H, W, cout).permute(0, 3, 1, 2)
    want = want * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if use_res:
        want = want + res.float()
    want = torch.relu(want)

    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    dst = torch.full((n, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    dst32 = torch.full((n, H, W, cout), float("nan"), dtype=torch.float32, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().cuda()
    wt = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin // 64, 64).permute(1, 2, 0, 3).contiguous().cuda()   # tile-contiguous
    for use_tc, layout in ((1, 0), (1, 1), (0, 0)):
        prog = Program(n)
        prog.add(op=OP_CONV, dtype=BF16, H=H, W=W, cin=cin, cout=cout, ksize=k, act=ACT["relu"], use_tc=use_tc, w_layout=layout,
                 src=nhwc(x), dst=dst, res=nhwc(res) if use_res else None, dst_f32=dst32, w=wt if layout else wp, scale=scale.cuda(), shift=shift.cuda(),
                 act_bias=ab.cuda() if use_ab else None, act_idx=idx.cuda() if use_ab else None)
        prog.run()
        torch.cuda.synchronize()
        got32 = dst32.permute(0, 3, 1, 2).cpu()
        got16 = dst.permute(0, 3, 1, 2).float().cpu()
        assert torch.isfinite(got32).all(), f"use_tc={use_tc}: unwritten / non-finite outputs"
        e32, e16 = rel(got32, want), rel(got16, want)
        assert e32 <= 2e-4, f"use_tc={use_tc}: fp32 side output rel err {e32:.2e}"
        assert e16 <= 5e-3, f"use_tc={use_tc}: bf16 output rel err {e16:.2e}"
        dst.fill_(float("nan")); dst32.fill_(float("nan"))


def test_bf16_pipeline_tensor_cores_vs_cuda_cores_and_reference(agent, rec):
    from muzero_breakout_b200.src.networks import PackedNetworks
    tc = PackedNetworks(agent, agent.cfg, precision="bf16", use_tc=True)
    cc = PackedNetworks(agent, agent.cfg, precision="bf16", use_tc=False)
    h = torch.from_numpy(rec["hidden"])
    planes = _planes(rec["dyn_actions"])
    out = {}
    for name, nets in (("tc", tc), ("cc", cc)):
        h2, rew = nets.dynamics(h, planes)
        pol, val = nets.prediction(h)
        hid = nets.representation(torch.from_numpy(rec["rep_in"]))
        out[name] = dict(h2=h2, rew=rew, pol=pol, val=val, hid=hid)
    same = {k: rel(out["tc"][k], out["cc"][k]) for k in out["tc"]}
    print("bf16 tensor-core vs CUDA-core pipelines (rel to range):", {k: f"{v:.2e}" for k, v in same.items()})
    ref = dict(h2=rec["dyn_h"], rew=rec["dyn_reward_logits"], pol=rec["root_policy_logits"], val=rec["root_value_logits"], hid=rec["hidden"])
    errs = {k: rel(out["tc"][k], ref[k]) for k in ref}
    print("bf16 tensor-core pipeline vs fp32 reference (rel to range):", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in same.items():      # same rounding points; a flipped bf16 rounding (ulp 3.9e-3 of the value) propagates
        assert e <= 2e-2, f"{k}: tensor-core vs CUDA-core bf16 pipelines differ by {e:.2e}"
    for k, e in errs.items():
        assert e <= 1e-2, f"{k}: bf16 pipeline deviates {e:.2e} from the fp32 reference"


@pytest.mark.parametrize("n", [5, 130, 1000, 1537, 3000, 4096, 9473])      # <= 1536 samples: output-channel-split items (two CTA pairs per pixel tile)
def test_fused_trunk_launch_equals_layer_by_layer(agent, n):
    """csrc/conv_stack.cu (whole residual trunk in one persistent launch, layers ordered by per-group device
    counters) must give bit-identical results to one launch per convolution: same tiles, same arithmetic."""
    from muzero_breakout_b200.src.networks import PackedNetworks
    fused = PackedNetworks(agent, agent.cfg, precision="bf16")
    plain = PackedNetworks(agent, agent.cfg, precision="bf16")
    plain.fuse_stacks = False
    fused.lat_max = 0                          # n = 5 would otherwise take the latency-mode trunk (tests/test_conv_lat_gpu.py)
    assert fused.fuse_stacks
    g = torch.Generator().manual_seed(n)
    h = torch.rand(n, 256, 4, 5, generator=g)
    acts = torch.randint(0, 3, (n,), generator=g)
    for rep in range(2):                       # twice: the done-counters are re-zeroed on the stream each run
        a_h, a_r = fused.dynamics(h, _planes(acts))
        b_h, b_r = plain.dynamics(h, _planes(acts))
        a_p, a_v = fused.prediction(h)
        b_p, b_v = plain.prediction(h)
        for x, y, what in ((a_h, b_h, "latent"), (a_r, b_r, "reward"), (a_p, b_p, "policy"), (a_v, b_v, "value")):
            assert torch.equal(x, y), f"n={n} rep={rep} {what}: fused and per-layer launches differ (max {float((x - y).abs().max()):.3e})"


def _f16_errs(nets, rec):
    h = torch.from_numpy(rec["hidden"])
    h2, rew = nets.dynamics(h, _planes(rec["dyn_actions"]))
    pol, val = nets.prediction(h)
    pol2, val2 = nets.prediction(torch.from_numpy(rec["dyn_h"]))
    hid = nets.representation(torch.from_numpy(rec["rep_in"]))
    return dict(h2=rel(h2, rec["dyn_h"]), rew=rel(rew, rec["dyn_reward_logits"]), pol=rel(pol, rec["root_policy_logits"]),
                val=rel(val, rec["root_value_logits"]), pol2=rel(pol2, rec["pred_policy_logits"]), val2=rel(val2, rec["pred_value_logits"]),
                hid=rel(hid, rec["hidden"]))


@pytest.mark.parametrize("path", ["fused_trunk", "per_layer", "latency_trunk"])
def test_f16_pipeline_vs_reference(agent, rec, path):
    """precision="f16", the mode the search and bench.py default to: fp16 operands on the tensor cores, fp32 accumulation, residual
    stream as fp16 + e4m3 correction.  Against the reference's own fp32 outputs (golden): within the north star's 1e-3 on every
    output, for each of the three kernels that can run the trunks."""
    from muzero_breakout_b200.src.networks import PackedNetworks
    nets = PackedNetworks(agent, agent.cfg, precision="f16")
    if path == "fused_trunk":
        nets.lat_max = 0                       # 5 samples would otherwise take the latency trunk
    elif path == "per_layer":
        nets.fuse_stacks = False
    errs = _f16_errs(nets, rec)
    print(f"f16 {path} vs fp32 reference (rel to range):", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e <= 1e-3, f"{path} {k}: f16 pipeline deviates {e:.2e} from the fp32 reference"


def test_f16_fused_trunk_vs_oracle_300(agent):
    """the same bound on a batch that spans several 128-sample groups and an odd group count (300 = 3 groups), against the fp32
    torch restatement of the reference networks evaluated here."""
    from muzero_breakout_b200.src.networks import PackedNetworks
    nets = PackedNetworks(agent, agent.cfg, precision="f16")
    n = 300
    g = torch.Generator().manual_seed(5)
    h = torch.rand(n, 256, 4, 5, generator=g)
    acts = torch.randint(0, 3, (n,), generator=g)
    with torch.no_grad():
        oh, orew = agent.hidden_state_transition(h, _planes(acts))
        opol, oval = agent.evaluate_state(h)
    for rep in range(2):                       # twice: the second launch runs at the next epoch of the dependency counters
        h2, rew = nets.dynamics(h, _planes(acts))
        pol, val = nets.prediction(h)
        errs = dict(h2=rel(h2, oh), rew=rel(rew, orew), pol=rel(pol, opol), val=rel(val, oval))
        print("f16 fused trunk, 300 samples, vs fp32 oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
        for k, e in errs.items():
            assert e <= 1e-3, f"rep {rep} {k}: {e:.2e}"


def test_residual_stream_correction_plane():
    """mz_op.res_lo / dst_lo: a convolution that keeps the e4m3 correction of its 16-bit output, followed by one that adds
    (output + correction) as its residual, against the same two convolutions with an exact fp32 residual.  bf16 makes the
    difference visible: without the plane the residual carries 2^-9 relative error, with it 2^-13."""
    from muzero_breakout_b200 import _lib
    from muzero_breakout_b200.src.networks import ACT, BF16, OP_CONV, Program
    for (n, H, W, c, use_tc) in ((200, 4, 5, 256, 1), (3, 16, 20, 128, 1), (9, 4, 5, 256, 0)):
        g = torch.Generator().manual_seed(n)
        x1 = torch.randn(n, c, H, W, generator=g).bfloat16()
        x2 = torch.randn(n, c, H, W, generator=g).bfloat16()
        w1 = (4 * torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).bfloat16()
        w2 = (0.05 * torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).bfloat16()
        shift = torch.randn(c, generator=g) * 0.1
        y1 = F.conv2d(x1.float(), w1.float(), padding=1) + shift.view(1, -1, 1, 1)                   # stream value, exact
        want = F.conv2d(x2.float(), w2.float(), padding=1) + shift.view(1, -1, 1, 1) + y1
        nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
        tile = lambda w: (w.permute(0, 2, 3, 1).reshape(c, 9, c // 64, 64).permute(1, 2, 0, 3) if use_tc else w.permute(0, 2, 3, 1).reshape(c, -1)).contiguous().cuda()
        s16 = torch.full((n, H, W, c), float("nan"), dtype=torch.bfloat16, device="cuda")
        out16 = torch.full_like(s16, float("nan"))
        out32 = torch.full((n, H, W, c), float("nan"), dtype=torch.float32, device="cuda")
        lo = torch.zeros(int(_lib.lib().mz_conv_lo_bytes(n, H, W, c, 3)), dtype=torch.uint8, device="cuda")
        errs = {}
        for with_lo in (True, False):
            prog = Program(n)
            common = dict(op=OP_CONV, dtype=BF16, H=H, W=W, cin=c, cout=c, ksize=3, act=ACT["none"], use_tc=use_tc, w_layout=use_tc, scale=None, shift=shift.cuda())
            prog.add(src=nhwc(x1), dst=s16, w=tile(w1), dst_lo=lo if with_lo else None, **common)
            prog.add(src=nhwc(x2), dst=out16, res=s16, res_lo=lo if with_lo else None, dst_f32=out32, w=tile(w2), **common)
            prog.run()
            torch.cuda.synchronize()
            errs[with_lo] = rel(out32.permute(0, 3, 1, 2).cpu(), want)
        print(f"n={n} {H}x{W} c={c} use_tc={use_tc}: residual with correction plane {errs[True]:.2e}, without {errs[False]:.2e}")
        assert errs[True] <= 1.5e-4 and errs[False] > 3 * errs[True], errs


@pytest.mark.parametrize("n", [1, 15, 33, 300, 2500, 9500])
@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_mma_heads_vs_fp64(n, dt):
    """The 16-bit heads on the warp MMA (csrc/nets.cu head_mma_kernel: Flatten + Linear + softmax / support expectation / inverse transform,
    networks.py:147-149,207-209,221-223 + utils.py:74-81): a dense 256-channel head (reward), and the policy (3 outputs) + value (11) pair
    reading the two 128-channel halves of one buffer in ONE launch, against an fp64 evaluation of the same 16-bit activations and fp32
    weights -- the kernel's only rounding is its fp32 accumulation (weights enter as hi + lo halves), so the bound is 2e-6 of the logits' range (bf16: 3e-5);
    batch sizes cover single and ragged CTAs and both samples-per-CTA variants."""
    from muzero_breakout_b200.src.networks import BF16, F16, OP_HEAD, Program
    g = torch.Generator().manual_seed(n)
    tdt, code = (torch.float16, F16) if dt == "f16" else (torch.bfloat16, BF16)
    x = (torch.rand(n, 20, 256, generator=g) * 2).to(tdt).cuda()
    ws = {k: ((torch.rand(o, f, generator=g) * 2 - 1) / f ** 0.5 * 8).cuda() for k, (o, f) in {"r": (11, 5120), "p": (3, 2560), "v": (11, 2560)}.items()}
    bs = {k: (torch.rand(w.shape[0], generator=g) - 0.5).cuda() for k, w in ws.items()}
    out = {k: torch.full((n,) if k != "p" else (n, 3), float("nan"), device="cuda") for k in ws}
    lg = {k: torch.full((n, ws[k].shape[0]), float("nan"), device="cuda") for k in ws}
    prog = Program(n)
    prog.add(op=OP_HEAD, dtype=code, H=4, W=5, cin=256, nout=11, head_mode=1, src=x, w=ws["r"], shift=bs["r"], out=out["r"], out_logits=lg["r"])
    prog.add(op=OP_HEAD, dtype=code, H=4, W=5, cin=128, cout=256, nout=3, head_mode=2, src=x, w=ws["p"], shift=bs["p"], out=out["p"], out_logits=lg["p"])
    prog.add(op=OP_HEAD, dtype=code, H=4, W=5, cin=128, cout=256, nout=11, head_mode=1, src=x.view(-1)[128:], w=ws["v"], shift=bs["v"], out=out["v"],
             out_logits=lg["v"])
    prog.run()
    torch.cuda.synchronize()
    xd = x.double()
    feats = {"r": xd.reshape(n, 5120), "p": xd[:, :, :128].reshape(n, 2560), "v": xd[:, :, 128:].reshape(n, 2560)}
    sup = torch.arange(-5, 6, dtype=torch.float64, device="cuda")
    for k in ws:
        ref = feats[k] @ ws[k].double().T + bs[k].double()
        tol = 2e-6 if dt == "f16" else 3e-5                    # bf16 weights enter as 8 + 8 significant bits
        assert rel(lg[k], ref) <= tol, f"{k} logits {rel(lg[k], ref):.2e}"
        p = torch.softmax(ref, 1)
        if k == "p":
            assert rel(out[k], p) <= 5e-6 + tol
        else:
            e = (p * sup).sum(1)
            y = torch.sign(e) * ((e.abs() + 0.999) ** 2 - 1)
            assert float((out[k].double() - y).abs().max()) <= (2e-5 + 10 * tol) * max(1.0, float(y.abs().max()))
