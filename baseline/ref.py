"""Import shim for the UNMODIFIED reference checkout (baseline/_ref, else /root/reference).

Used by bench.py's reference legs and by tests/test_reference_dropin_gpu.py; never by the product (nothing under
muzero-breakout_b200/ imports this).  SURVEY.md Appendix A:
  * matplotlib is imported-but-unused by the reference (environment/parallel_breakout.py:8, src/mcts.py:6) and is not
    installed -> stub modules;
  * "cuda" is hard-coded (src/networks.py:249, src/mcts.py:190, train_torch.py:15) -> redirected to "cpu" when no GPU is
    visible (the CPU arm hides the GPU with CUDA_VISIBLE_DEVICES="" before torch is imported).
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_dir():
    """Directory of the reference checkout, or None."""
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(d, "src", "mcts.py")):
            return d
    return None


def install(ahead_of_repo: bool = False) -> str:
    """Make `import src.mcts`, `environment.parallel_breakout`, `train_torch` ... resolve to the reference.  Returns its directory."""
    import torch
    import torch.nn as nn

    d = ref_dir()
    if d is None:
        raise RuntimeError("no reference checkout: run __graft_entry__.build() in the build container (copies /root/reference to baseline/_ref)")
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if d not in sys.path:
        sys.path.insert(0, d) if ahead_of_repo else sys.path.append(d)
    if not torch.cuda.is_available() and not getattr(torch.Tensor.to, "_redirected", False):
        fix = lambda x: "cpu" if isinstance(x, str) and x.startswith("cuda") else x
        _t, _m = torch.Tensor.to, nn.Module.to

        def t_to(s, *a, **k):
            return _t(s, *map(fix, a), **{q: fix(v) for q, v in k.items()})

        def m_to(s, *a, **k):
            return _m(s, *map(fix, a), **{q: fix(v) for q, v in k.items()})

        t_to._redirected = True
        torch.Tensor.to, nn.Module.to = t_to, m_to
    return d


def load_cfg() -> dict:
    import torch
    import yaml

    cfg = yaml.safe_load(open(os.path.join(ref_dir(), "config.yaml")))["parameters"]
    if not torch.cuda.is_available():
        cfg["model"]["device"] = "cpu"
    return cfg
