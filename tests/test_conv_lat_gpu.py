"""GPU tests of the latency-mode residual trunk (csrc/conv_lat.cu, mz_lat_build / mz_lat_run): the small-batch form of the
trunk that config.yaml's default acting stage (24 roots) runs on.

Tolerances (max abs error / max abs reference value):
  one layer vs torch's fp32 conv on the same 16-bit operands     2e-4 on the fp32 side output, 5e-3 on the bf16 output
  whole networks, latency trunk vs tcgen05 trunk                  2e-2 (same bf16 rounding points; a flipped rounding propagates)
  whole networks vs the fp32 reference outputs (golden)           3e-2 for bf16 storage, 2.5e-3 for fp16 storage (as in test_networks_gpu.py)
"""
import ctypes as C
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
    assert a.shape == b.shape
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _planes(actions):
    p = torch.zeros(len(actions), 3, 4, 5)
    p[torch.arange(len(actions)), torch.as_tensor(actions)] = 1
    return p


def _run_layers(ops, n, dtype):
    """mz_lat_build + mz_lat_run on a list of MzOp records, straight through the C ABI."""
    from muzero_breakout_b200 import _lib
    from muzero_breakout_b200.src.networks import MzOp
    L = _lib.lib()
    lb = L.mz_lat_layer_bytes()
    raw = (C.c_uint8 * (len(ops) * lb + 64))()
    host = (C.addressof(raw) + 63) & ~63
    split = L.mz_lat_build((MzOp * len(ops))(*ops), len(ops), host, len(ops) * lb)
    assert split in (0, 1), _lib.lib().mzb_last_error().decode()
    blob = torch.frombuffer((C.c_uint8 * (len(ops) * lb)).from_address(host), dtype=torch.uint8).clone().cuda()
    done = torch.zeros((L.mz_lat_scratch_bytes(n) + 3) // 4, dtype=torch.int32, device="cuda")   # zeroed once, then owned by the launches
    act_idx = next((o.act_idx for o in ops if o.act_idx), None)
    for _ in range(2):                                        # twice: the second launch runs at epoch 1 on the same scratch
        _lib.check(L.mz_lat_run(blob.data_ptr(), len(ops), split, n, act_idx, done.data_ptr(), dtype, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return done.cpu()


LAYER_CASES = [  # (n, residual, act_bias, act, dtype[, ksize])
    (24, False, False, "relu", "bf16", 1), (7, True, False, "relu", "f16", 1),
    (1, False, False, "relu", "bf16"), (3, True, False, "relu", "bf16"), (24, True, True, "relu", "bf16"), (25, True, False, "none", "bf16"),
    (27, False, True, "silu", "bf16"), (61, True, True, "relu", "bf16"), (200, True, False, "relu", "bf16"), (24, True, True, "relu", "f16"),
]


@pytest.mark.parametrize("case", LAYER_CASES, ids=lambda c: "n%d_r%d_a%d_%s_%s_k%d" % (c[0], c[1], c[2], c[3], c[4], c[5] if len(c) > 5 else 3))
def test_one_layer_vs_torch(case):
    """One 3x3 (or 1x1) 256->256 convolution (+ action bias, folded BN, residual, activation) against torch's fp32 conv on the same
    16-bit operands; every output element compared, buffers NaN-prefilled.  n = 61 / 200 need several items per CTA."""
    from muzero_breakout_b200.src.networks import ACT, BF16, F16, OP_CONV, MzOp
    n, use_res, use_ab, act, prec = case[:5]
    k = case[5] if len(case) > 5 else 3
    td, dt = (torch.bfloat16, BF16) if prec == "bf16" else (torch.float16, F16)
    g = torch.Generator().manual_seed(n * 7 + use_res + 2 * use_ab)
    x = torch.randn(n, 256, 4, 5, generator=g).to(td)
    scale = torch.rand(256, generator=g) + 0.5              # the BatchNorm scale is folded into the 16-bit weights (mz_op.scale == NULL)
    w = (torch.randn(256, 256, k, k, generator=g) / (16.0 * k) * scale.view(-1, 1, 1, 1)).to(td)
    shift = torch.randn(256, generator=g) * 0.1
    res = torch.randn(n, 256, 4, 5, generator=g).to(td) if use_res else None
    ab = torch.randn(3, 20, 256, generator=g) * 0.2 if use_ab else None
    idx = torch.randint(0, 3, (n,), generator=g, dtype=torch.int32)
    want = F.conv2d(x.float(), w.float(), padding=k // 2)
    if use_ab:
        want = want + ab[idx.long()].view(n, 4, 5, 256).permute(0, 3, 1, 2)
    want = want + shift.view(1, -1, 1, 1)
    if use_res:
        want = want + res.float()
    want = {"relu": torch.relu, "none": lambda t: t, "silu": F.silu}[act](want)

    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    keep = dict(src=nhwc(x), res=nhwc(res) if use_res else None, shift=shift.cuda(), ab=ab.cuda() if use_ab else None,
                idx=idx.cuda(), w=w.permute(0, 2, 3, 1).reshape(256, k * k, 4, 64).permute(1, 2, 0, 3).contiguous().cuda(),
                dst=torch.full((n, 4, 5, 256), float("nan"), dtype=td, device="cuda"),
                dst32=torch.full((n, 4, 5, 256), float("nan"), dtype=torch.float32, device="cuda"))
    op = MzOp()
    for k, v in dict(op=OP_CONV, dtype=dt, H=4, W=5, cin=256, cout=256, ksize=k, act=ACT[act], use_tc=1, w_layout=1, src=keep["src"], dst=keep["dst"],
                     res=keep["res"], dst_f32=keep["dst32"], w=keep["w"], scale=None, shift=keep["shift"], act_bias=keep["ab"],
                     act_idx=keep["idx"] if use_ab else None).items():
        setattr(op, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    done = _run_layers([op], n, dt)
    assert int(done[-(2 + 2 * ((n + 2) // 3))]) == 2, "the launch epoch advances once per launch"
    got32 = keep["dst32"].permute(0, 3, 1, 2).cpu()
    got16 = keep["dst"].permute(0, 3, 1, 2).float().cpu()
    assert torch.isfinite(got32).all() and torch.isfinite(got16).all(), "unwritten / non-finite outputs"
    e32, e16 = rel(got32, want), rel(got16, want)
    assert e32 <= 2e-4, f"fp32 side output rel err {e32:.2e}"
    assert e16 <= (5e-3 if prec == "bf16" else 1e-3), f"16-bit output rel err {e16:.2e}"


@pytest.fixture(scope="module")
def agent():
    torch.manual_seed(0)
    a = OracleAgent()
    perturb_bn(a, 1)
    a.eval_mode()
    return a


@pytest.mark.parametrize("n", [2, 24, 27, 61])
def test_networks_latency_trunk_vs_tcgen05_trunk(agent, n):
    """dynamics + prediction with the trunks in latency mode against the same networks on the tcgen05 trunk (lat_max = 0)
    and against the fp32 torch oracle; run twice (the second run is the next launch epoch on the same hand-off buffers)."""
    from muzero_breakout_b200.src.networks import PackedNetworks, lat_max_samples
    assert n <= lat_max_samples()
    lat = PackedNetworks(agent, agent.cfg, precision="bf16")
    tc = PackedNetworks(agent, agent.cfg, precision="bf16")
    tc.lat_max = 0
    g = torch.Generator().manual_seed(n)
    h = torch.rand(n, 256, 4, 5, generator=g)
    acts = torch.randint(0, 3, (n,), generator=g)
    with torch.no_grad():
        oh, orew = agent.hidden_state_transition(h, _planes(acts))
        opol, oval = agent.evaluate_state(h)
    from muzero_breakout_b200 import _lib
    for rep in range(2):
        n0 = _lib.launch_count()
        a = lat.dynamics(h, _planes(acts)) + lat.prediction(h)
        n1 = _lib.launch_count()
        b = tc.dynamics(h, _planes(acts)) + tc.prediction(h)
        # latency form: layout in, ONE launch for the dynamics network (trunk + reward ConvBlock + reward head + _scale_state), layout out;
        # layout in, ONE launch for the prediction network (trunk + policy / value ConvBlocks + both heads); tcgen05 form: layout in, trunk +
            # reward ConvBlock, reward head, _scale_state, layout out; layout in, trunk + both head ConvBlocks, the two heads in one launch = 8 launches
        assert n1 - n0 == 5 and _lib.launch_count() - n1 == 8, (n1 - n0, _lib.launch_count() - n1)
        for x, y, o, what in zip(a, b, (oh, orew, opol, oval), ("latent", "reward", "policy", "value")):
            assert torch.isfinite(x).all()
            assert rel(x, y) <= 2e-2, f"n={n} {what}: latency trunk vs tcgen05 trunk {rel(x, y):.2e}"
            assert rel(x, o) <= 3e-2, f"n={n} {what}: latency trunk vs fp32 oracle {rel(x, o):.2e}"


def test_f16_latency_trunk_vs_reference(golden_dir):
    """precision="f16" (the search's default) at the golden batch (5 roots): latency trunk against the reference's fp32 outputs,
    within the north star's 1e-3 (max |err| / max |ref| per tensor)."""
    from muzero_breakout_b200.src.networks import PackedNetworks
    torch.manual_seed(0)
    a = OracleAgent()
    perturb_bn(a, 1)
    a.eval_mode()
    rec = np.load(os.path.join(golden_dir, "mcts_real.npz"))
    nets = PackedNetworks(a, a.cfg, precision="f16")
    h = torch.from_numpy(rec["hidden"])
    h2, rew = nets.dynamics(h, _planes(rec["dyn_actions"]))
    pol, val = nets.prediction(h)
    errs = dict(h2=rel(h2, rec["dyn_h"]), rew=rel(rew, rec["dyn_reward_logits"]), pol=rel(pol, rec["root_policy_logits"]), val=rel(val, rec["root_value_logits"]))
    print("f16 latency trunk vs fp32 reference (rel to range):", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e <= 1e-3, f"{k}: {e:.2e}"
