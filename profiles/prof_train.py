"""The training step's two ends on their own (csrc/train.cu, rb_gather_input): a few launches each, for the launch list / ncu capture.
    python profiles/prof_train.py
    ncu --set full --clock-control none --import-source on -k regex:'adam_kernel|loss_kernel|rb_gather_input' -c 6 -o gpurun_out/r1_train python profiles/prof_train.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200 import _lib
from muzero_breakout_b200.replay_buffer import ReplayBuffer
from muzero_breakout_b200.train import loss_fn

L = _lib.lib()
dev = torch.device("cuda:0")
n = 42_205_081
p, g = torch.randn(n, device=dev) * 0.05, torch.randn(n, device=dev) * 1e-3
m, v = torch.zeros_like(p), torch.zeros_like(p)
st = torch.cuda.current_stream().cuda_stream
B, K = 512, 5
pr, pv, pp = torch.randn(B, K, 11, device=dev), torch.randn(B, K, 11, device=dev), torch.randn(B, K, 3, device=dev)
obs, val = torch.randint(-1, 2, (B, K), device=dev).float(), (torch.rand(B, K, device=dev) - 0.5) * 20
vis, sup = torch.randint(1, 30, (B, K, 3), device=dev).float(), torch.linspace(-5, 5, 11, device=dev)
T, E = 64, 1024
rec = dict(action=torch.randint(0, 3, (T, E), device=dev), reward=torch.zeros(T, E, device=dev), value=torch.rand(T, E, device=dev),
           visits=torch.randint(0, 51, (T, E, 3), device=dev), frames=torch.rand(T, E, 1, 16, 20, device=dev),
           recorded=torch.ones(T, E, dtype=torch.bool, device=dev), initial_gray=torch.rand(E, 1, 16, 20, device=dev))
rb = ReplayBuffer(32, K, 60000, 0.985, 512, device=dev, max_moves=261)
rb.save_episode(rec)
idx = torch.randperm(rb.length, device=dev)[:B]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


step = [0]


def adam():
    step[0] += 1
    _lib.check(L.mz_adam(n, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), 2e-4, 0.9, 0.999, 1e-8, 1e-4, step[0], st))


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
us = timed(adam, reps)
print(f"mz_adam, {n} parameters: {us:.1f} us = {n * 28 / us / 1e3:.0f} GB/s (28 B per parameter)")
us = timed(lambda: loss_fn(obs, pr, val, pv, vis, pp, sup, K), reps)
print(f"loss_fn, {B * K} rows: {us:.1f} us")
us = timed(lambda: rb.repnet_input(idx), reps)
print(f"repnet_input, {B} samples: {us:.1f} us = {B * 120 * 1024 / us / 1e3:.0f} GB/s (40 KB read + 80 KB written per sample)")
us = timed(lambda: torch.cat((rb.get_batched_states(idx).view(B, -1, 16, 20),
                              torch.ones((B, 32, 16, 20), device=dev) * (rb.get_batched_past_actions(idx) / 3)[:, :, None, None].expand(-1, -1, 16, 20)), dim=1), reps)
print(f"  the same from two gathers + _encode_actions + cat in torch: {us:.1f} us")

# weight / data gradients of one trunk convolution at the training minibatch (512 samples, 4x5 latent, 3x3 256 -> 256)
from muzero_breakout_b200.train import ConvDgrad, conv_wgrad
nb = 512
xa = torch.randn(nb, 4, 5, 256, device=dev).bfloat16()
dya = torch.randn(nb, 4, 5, 256, device=dev).bfloat16()
ns = L.mz_wgrad_padded_samples(nb)
dy_t, x_t = torch.empty(256, 20, ns, dtype=torch.bfloat16, device=dev), torch.empty(256, 20, ns, dtype=torch.bfloat16, device=dev)
partial = torch.empty(L.mz_wgrad_partial_bytes(3, nb) // 4, dtype=torch.float32, device=dev)
dw = torch.empty(256, 256, 3, 3, device=dev)
flop = 130 * nb * 256 * 256 * 2          # in-bounds (pixel, tap) pairs only, as for the forward convolution
us = timed(lambda: (_lib.check(L.mz_wgrad_transpose(nb, 20, 256, dya.data_ptr(), dy_t.data_ptr(), st)),
                    _lib.check(L.mz_wgrad_transpose(nb, 20, 256, xa.data_ptr(), x_t.data_ptr(), st))), reps)
print(f"wgrad transposes (2 x {nb * 20 * 256 * 2 / 1e6:.1f} MB): {us:.1f} us")
us = timed(lambda: _lib.check(L.mz_conv_wgrad(nb, 4, 5, 3, 1, dy_t.data_ptr(), x_t.data_ptr(), partial.data_ptr(), dw.data_ptr(), st)), reps)
print(f"mz_conv_wgrad, {nb} samples 3x3 256->256: {us:.1f} us = {flop / us / 1e6:.0f} TFLOP/s of in-bounds-tap work (kernel + split reduction)")
wdg = ConvDgrad(torch.randn(256, 256, 3, 3) / 48)
us = timed(lambda: wdg(dya), reps)
print(f"ConvDgrad, {nb} samples: {us:.1f} us = {flop / us / 1e6:.0f} TFLOP/s (forward kernel, incl. Program build + 2 output allocations)")
w32 = torch.randn(256, 256, 3, 3, device=dev)
xn, dyn = xa.permute(0, 3, 1, 2).float().contiguous(), dya.permute(0, 3, 1, 2).float().contiguous()
us = timed(lambda: torch.nn.grad.conv2d_weight(xn, w32.shape, dyn, padding=1), reps)
print(f"  torch (cuDNN fp32/TF32) conv2d_weight on the same shapes: {us:.1f} us")
xb, dyb = xa.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last), dya.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
us = timed(lambda: torch.nn.grad.conv2d_weight(xb, w32.shape, dyb, padding=1), reps)
print(f"  torch (cuDNN bf16, channels_last) conv2d_weight: {us:.1f} us")

# the K unroll steps share their weights: one wgrad launch per layer over all 5 x 512 (dY, X) pairs of a training step
nb5 = 2560
xa5, dya5 = torch.randn(nb5, 4, 5, 256, device=dev).bfloat16(), torch.randn(nb5, 4, 5, 256, device=dev).bfloat16()
dy_t5, x_t5 = torch.empty(256, 20, nb5, dtype=torch.bfloat16, device=dev), torch.empty(256, 20, nb5, dtype=torch.bfloat16, device=dev)
partial5 = torch.empty(L.mz_wgrad_partial_bytes(3, nb5) // 4, dtype=torch.float32, device=dev)
us_t = timed(lambda: (_lib.check(L.mz_wgrad_transpose(nb5, 20, 256, dya5.data_ptr(), dy_t5.data_ptr(), st)),
                      _lib.check(L.mz_wgrad_transpose(nb5, 20, 256, xa5.data_ptr(), x_t5.data_ptr(), st))), reps)
us = timed(lambda: _lib.check(L.mz_conv_wgrad(nb5, 4, 5, 3, 1, dy_t5.data_ptr(), x_t5.data_ptr(), partial5.data_ptr(), dw.data_ptr(), st)), reps)
print(f"mz_conv_wgrad, {nb5} samples: {us:.1f} us = {flop * 5 / us / 1e6:.0f} TFLOP/s (+ transposes {us_t:.1f} us)")
xb5, dyb5 = xa5.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last), dya5.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
us = timed(lambda: torch.nn.grad.conv2d_weight(xb5, w32.shape, dyb5, padding=1), reps)
print(f"  torch (cuDNN bf16, channels_last) conv2d_weight, {nb5} samples: {us:.1f} us")

# one ResidualBlock training step (forward + backward, 512 samples) through this library's kernels vs torch / cuDNN on the same GPU
from muzero_breakout_b200.train import ResidualBlockTrain
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.networks import _Res          # profiling script only: the torch restatement of networks.py:19-35 as the cuDNN comparison
blk = _Res(256, "relu").cuda().train()
sd = {k: v.detach() for k, v in blk.state_dict().items()}
ours = ResidualBlockTrain(sd["conv1.weight"], sd["conv1.bias"], sd["bn1.weight"], sd["bn1.bias"], sd["conv2.weight"], sd["conv2.bias"], sd["bn2.weight"], sd["bn2.bias"])
x16 = torch.rand(nb, 4, 5, 256, device=dev).bfloat16()
dyf = torch.randn(nb, 4, 5, 256, device=dev)
us = timed(lambda: (ours.forward(x16), ours.backward(dyf)), reps)
print(f"ResidualBlockTrain forward + backward, {nb} samples: {us:.1f} us (22 launches + allocations, no buffer reuse yet)")
xt = x16.permute(0, 3, 1, 2).float().contiguous().requires_grad_()
dyt = dyf.permute(0, 3, 1, 2).contiguous()
def torch_step():
    for p_ in blk.parameters(): p_.grad = None
    xt.grad = None
    blk(xt).backward(dyt)
us = timed(torch_step, reps)
print(f"  torch (cuDNN fp32/TF32) forward + backward of the same block: {us:.1f} us")
blk16 = _Res(256, "relu").cuda().train().to(memory_format=torch.channels_last)
xt16 = x16.permute(0, 3, 1, 2).float().contiguous(memory_format=torch.channels_last).requires_grad_()
dyt16 = dyt.contiguous(memory_format=torch.channels_last)
def torch_step16():
    for p_ in blk16.parameters(): p_.grad = None
    xt16.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = blk16(xt16)
    o.backward(dyt16.to(o.dtype))
us = timed(torch_step16, reps)
print(f"  torch autocast bf16, channels_last: {us:.1f} us")

# the same step captured once as a CUDA graph (static shapes; every launch is stream-ordered, no host sync inside)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        ours.forward(x16); ours.backward(dyf)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    ours.forward(x16)
    g_dx, g_grads = ours.backward(dyf)
us = timed(graph.replay, reps)
print(f"ResidualBlockTrain forward + backward as a CUDA graph replay: {us:.1f} us")

# a 14-block trunk (dynamics / prediction network body) as one training step, CUDA-graph replay vs torch
from muzero_breakout_b200.train import TrunkTrain
trunk_t = torch.nn.Sequential(*[_Res(256, "relu") for _ in range(14)]).cuda().train()
sd14 = {k: v.detach() for k, v in trunk_t.state_dict().items()}
trunk = TrunkTrain.from_state_dict(sd14, [f"{i}." for i in range(14)])
us = timed(lambda: trunk.step(x16, dyf, graph=True), max(reps // 2, 2))
print(f"TrunkTrain 14 ResidualBlocks forward + backward, {nb} samples, CUDA-graph replay: {us / 1e3:.2f} ms")
def torch_trunk():
    for p_ in trunk_t.parameters(): p_.grad = None
    xt.grad = None
    trunk_t(xt).backward(dyt)
us = timed(torch_trunk, max(reps // 2, 2))
print(f"  torch (cuDNN fp32/TF32): {us / 1e3:.2f} ms")
trunk_t16 = trunk_t.to(memory_format=torch.channels_last)
def torch_trunk16():
    for p_ in trunk_t16.parameters(): p_.grad = None
    xt16.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = trunk_t16(xt16)
    o.backward(dyt16.to(o.dtype))
us = timed(torch_trunk16, max(reps // 2, 2))
print(f"  torch autocast bf16, channels_last: {us / 1e3:.2f} ms")
