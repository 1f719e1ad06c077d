"""One training-loop iteration (train_torch.py:385-417) at config.yaml's minibatch 512 x K = 5: the learner-side drop-in against the
unmodified reference modules under torch + cuDNN on the same GPU -- bench.py's `train_step` leg alone.
    python profiles/prof_train_step.py [minibatch] [K]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
print(json.dumps(bench.bench_train_step(torch.device("cuda", 0), mb, K), indent=1))
