"""A few launches of the training-step tensor-core / BatchNorm kernels for the ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel|wgrad_transpose|wgrad_reduce|bn_' -c 12 -o gpurun_out/r1_wgrad python profiles/prof_wgrad.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.train import bn_train_backward, bn_train_forward, conv_wgrad

dev = torch.device("cuda:0")
for n in (512, 2560):
    x, dy = torch.randn(n, 4, 5, 256, device=dev).bfloat16(), torch.randn(n, 4, 5, 256, device=dev).bfloat16()
    for _ in range(2):
        conv_wgrad(dy, x, 3)
z = torch.randn(512, 4, 5, 256, device=dev)
gam, bet = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
y16, y32, mean, invstd = bn_train_forward(z, gam, bet, res=x[:512].contiguous(), act="relu")
bn_train_backward(z, torch.randn_like(z), gam, bet, mean, invstd, res=x[:512].contiguous(), act="relu")
torch.cuda.synchronize()
print("done")
