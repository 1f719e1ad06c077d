"""Three eager ResidualBlockTrain steps (forward + backward, 512 samples) for the launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_block_step_launches.csv python profiles/prof_block_step.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_breakout_b200.train import ResidualBlockTrain

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
mk = lambda *shape, s=1.0: torch.randn(*shape, generator=g) * s
blk = ResidualBlockTrain(mk(256, 256, 3, 3, s=0.02), mk(256, s=0.01), torch.rand(256, generator=g) + 0.5, mk(256, s=0.1),
                         mk(256, 256, 3, 3, s=0.02), mk(256, s=0.01), torch.rand(256, generator=g) + 0.5, mk(256, s=0.1))
x16 = torch.rand(512, 4, 5, 256, device=dev).bfloat16()
dy = torch.randn(512, 4, 5, 256, device=dev)
for _ in range(3):
    blk.forward(x16)
    blk.backward(dy)
torch.cuda.synchronize()
print("done")
