"""CPU emulation of the rounding points of the 16-bit tensor-core network path (no GPU needed).

Which storage scheme of the residual trunks meets the north star's 1e-3 (max |err| / max |ref|, against the fp32
reference outputs in tests/golden/mcts_real.npz)?  Every conv is evaluated in fp64 on operands rounded the way
the kernels round them:
    weights        -> 16-bit after folding the BN scale (fold=1) or before (fold=0: scale applied in fp32)
    conv inputs    -> 16-bit (the TMA operand)
    residual       -> "16": read back from the 16-bit activation buffer
                      "32": carried in fp32 (the stream x_k is kept in fp32, the 16-bit copy is only the operand)
                      "16+8": 16-bit value + an 8-bit (e4m3, scaled by 2^11) correction of its rounding error
Usage: python profiles/emulate_precision.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from common import perturb_bn                      # noqa: E402
from oracle.networks import OracleAgent, scale_state   # noqa: E402


def rnd(x, dt):
    return x.to(dt).to(torch.float64) if dt is not None else x


def e4m3(x):
    return x.to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


def store(x, dt, res_mode):
    """-> (operand copy, residual value) of a stream tensor x (fp64 holding an fp32-exact epilogue result)"""
    x = x.to(torch.float32).to(torch.float64)
    op = rnd(x, dt)
    if res_mode == "16":
        return op, op
    if res_mode == "32":
        return op, x
    if res_mode == "16+8":
        lo = e4m3((x - op) * 2048.0) / 2048.0
        return op, op + lo
    raise ValueError(res_mode)


def conv(x_op, m, bn, dt, fold, k):
    w = m.weight.double()
    b = m.bias.double()
    if bn is not None:
        a = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
        shift = (b - bn.running_mean.double()) * a + bn.bias.double()
    else:
        a, shift = torch.ones_like(b), b
    if fold:
        wq = rnd(w * a.view(-1, 1, 1, 1), dt)
        y = F.conv2d(x_op, wq, padding=k // 2)
    else:
        y = F.conv2d(x_op, rnd(w, dt), padding=k // 2) * a.view(1, -1, 1, 1)
    return y + shift.view(1, -1, 1, 1)


def res_block(x_op, x_res, blk, dt, fold, res_mode):
    y = torch.relu(conv(x_op, blk.conv1, blk.bn1, dt, fold, 3))
    y = rnd(y.float().double(), dt)
    z = torch.relu(conv(y, blk.conv2, blk.bn2, dt, fold, 3) + x_res)
    return store(z, dt, res_mode)


def dynamics(agent, h, planes, dt, fold, res_mode):
    d = agent.dyn_net
    x = torch.cat([rnd(h.double(), dt), planes.double()], 1)
    # action planes are exact (0/1) and their weights stay fp32 in the bias table: emulate by not rounding those weight columns
    m, bn = d.conv_block.conv, d.conv_block.bn
    a = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    shift = (m.bias.double() - bn.running_mean.double()) * a + bn.bias.double()
    w = m.weight.double()
    if fold:
        wq = torch.cat([rnd(w[:, :256] * a.view(-1, 1, 1, 1), dt), w[:, 256:] * a.view(-1, 1, 1, 1)], 1)
        y = F.conv2d(x, wq, padding=1)
    else:
        wq = torch.cat([rnd(w[:, :256], dt), w[:, 256:]], 1)
        y = F.conv2d(x, wq, padding=1) * a.view(1, -1, 1, 1)
    y = torch.relu(y + shift.view(1, -1, 1, 1))
    x_op, x_res = store(y, dt, res_mode)
    for blk in d.res_blocks:
        x_op, x_res = res_block(x_op, x_res, blk, dt, fold, res_mode)
    # reward head reads the 16-bit operand copy; _scale_state reads the fp32 side output of the last layer
    hd = d.reward_head
    r = torch.relu(conv(x_op, hd[0].conv, hd[0].bn, dt, fold, 1))
    r = rnd(r.float().double(), dt)
    rew = F.linear(r.flatten(1), hd[2].weight.double(), hd[2].bias.double())
    last32 = x_res if res_mode != "16" else None
    return last32, x_op, rew


def prediction(agent, h, dt, fold, res_mode):
    p = agent.pred_net
    x_op, x_res = store(h.double(), dt, res_mode)
    for blk in p.res_blocks:
        x_op, x_res = res_block(x_op, x_res, blk, dt, fold, res_mode)
    out = []
    for hd, k in ((p.policy_head, 3), (p.value_head, 1)):
        r = torch.relu(conv(x_op, hd[0].conv, hd[0].bn, dt, fold, k))
        r = rnd(r.float().double(), dt)
        out.append(F.linear(r.flatten(1), hd[2].weight.double(), hd[2].bias.double()))
    return out


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max())


def rel2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def main():
    torch.manual_seed(0)
    agent = OracleAgent()
    perturb_bn(agent, 1)
    agent.eval()
    g = np.load(os.path.join(ROOT, "tests", "golden", "mcts_real.npz"))
    h = torch.from_numpy(g["hidden"])
    acts = torch.from_numpy(g["dyn_actions"]).long()
    planes = torch.zeros(len(acts), 3, 4, 5)
    planes[torch.arange(len(acts)), acts] = 1
    ref = dict(h2=g["dyn_h"], rew=g["dyn_reward_logits"], pol=g["root_policy_logits"], val=g["root_value_logits"])
    print(f"{'dtype':6} {'fold':4} {'res':5} | " + " ".join(f"{k:>18}" for k in ref) + "   (max/max | l2/l2)")
    with torch.no_grad():
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "f16")):
            for fold in (0, 1):
                for res_mode in ("16", "32", "16+8"):
                    last32, x_op, rew = dynamics(agent, h, planes, dt, fold, res_mode)
                    # the last layer's fp32 side output feeds _scale_state in every mode (conv_stack.cu dst_f32)
                    if last32 is None:
                        # res "16": side output is the fp32 epilogue value before rounding -> emulate by re-running the last block in fp32 store
                        last32 = x_op
                    h2 = rnd(scale_state(last32).float().double(), dt)
                    pol, val = prediction(agent, h, dt, fold, res_mode)
                    got = dict(h2=h2, rew=rew, pol=pol, val=val)
                    print(f"{name:6} {fold:4d} {res_mode:5} | " + " ".join(f"{rel(got[k], ref[k]):8.2e} {rel2(got[k], ref[k]):8.2e}" for k in ref))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def representation(agent, x, dt, fold, res_mode):
    """rep net (networks.py:38-99): plain convs (bias only), residual blocks, two average pools, _scale_state"""
    r = agent.rep_net
    x_op, x_res = store(x.double(), dt, res_mode)
    for m in r.blocks:
        if isinstance(m, torch.nn.Conv2d):
            x_op, x_res = store(conv(x_op, m, None, dt, fold, 3), dt, res_mode)
        elif isinstance(m, torch.nn.AvgPool2d):
            # pool reads the residual-precision stream when there is one
            x_op, x_res = store(F.avg_pool2d(x_res, 2, 2), dt, res_mode)
            last_pool32 = F.avg_pool2d(x_res, 2, 2) if False else None
        else:
            x_op, x_res = res_block(x_op, x_res, m, dt, fold, res_mode)
    return x_op, x_res


def main_rep():
    torch.manual_seed(0)
    agent = OracleAgent()
    perturb_bn(agent, 1)
    agent.eval()
    g = np.load(os.path.join(ROOT, "tests", "golden", "mcts_real.npz"))
    x = torch.from_numpy(g["rep_in"])
    with torch.no_grad():
        for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "f16")):
            for res_mode in ("16", "32", "16+8"):
                x_op, x_res = representation(agent, x, dt, 1, res_mode)
                # the pool kernel's fp32 side output feeds _scale_state: "16" -> pooled from 16-bit inputs but kept fp32
                hid = rnd(scale_state(x_res).float().double(), dt)
                print(f"rep {name:5} {res_mode:5} hid {rel(hid, g['hidden']):8.2e} {rel2(hid, g['hidden']):8.2e}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "rep":
    main_rep()
