"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

Used only by tests/golden/gen_golden.py to produce the committed golden vectors; nothing that runs
on the GPU box imports this (the reference does not exist there).  SURVEY.md Appendix A / C.

* matplotlib is imported-but-unused by the reference and is not installed -> stub modules.
* "cuda" is hard-coded (src/networks.py:249, src/mcts.py:190) -> redirected to "cpu" when there is
  no GPU.
* RNG injection: src.mcts's module-global `torch` is replaced by a proxy whose
  `distributions.Dirichlet(..).sample()` returns row i of a supplied (B,3) tensor on its i-th call
  (patch point src/mcts.py:114) and whose `randint(n,(1,))` returns u32(seed, tree, ctr[tree]++) % n
  (patch point src/mcts.py:297); the tree index is captured by wrapping ucb_action (its `idx` arg).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn
import yaml

REF = "/root/reference"
M64 = (1 << 64) - 1


def rng_u32(seed: int, tree: int, ctr: int) -> int:
    """Same function as oracle/mcts_oracle.c:mto_rng_u32 and csrc/tree.cu:mz_rng_u32."""
    z = (seed + 0x9E3779B97F4A7C15 * (((tree & 0xFFFFFFFF) << 32) | (ctr & 0xFFFFFFFF))) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return (z >> 32) & 0xFFFFFFFF


def install():
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if not torch.cuda.is_available() and not getattr(torch.Tensor.to, "_redirected", False):
        fix = lambda x: "cpu" if isinstance(x, str) and x.startswith("cuda") else x
        _t, _m = torch.Tensor.to, nn.Module.to

        def t_to(s, *a, **k):
            return _t(s, *map(fix, a), **{q: fix(v) for q, v in k.items()})

        def m_to(s, *a, **k):
            return _m(s, *map(fix, a), **{q: fix(v) for q, v in k.items()})

        t_to._redirected = True
        torch.Tensor.to, nn.Module.to = t_to, m_to


def load_cfg():
    cfg = yaml.safe_load(open(REF + "/config.yaml"))["parameters"]
    if not torch.cuda.is_available():
        cfg["model"]["device"] = "cpu"
    return cfg


class _Dirichlet:
    def __init__(self, proxy):
        self.proxy = proxy

    def sample(self):
        row = self.proxy.noise[self.proxy.noise_calls]
        self.proxy.noise_calls += 1
        return row.clone()


class TorchProxy:
    """Stands in for the module-global `torch` of src.mcts."""

    def __init__(self, seed: int, noise: torch.Tensor):
        self.seed, self.noise = seed, noise
        self.noise_calls = 0
        self.ctr = {}
        self.cur_idx = None
        self.ucb_calls = 0
        proxy = self

        class _Dist:
            @staticmethod
            def Dirichlet(_alpha):
                return _Dirichlet(proxy)

        self.distributions = _Dist()

    def randint(self, n, size):
        assert tuple(size) == (1,)
        c = self.ctr.get(self.cur_idx, 0)
        self.ctr[self.cur_idx] = c + 1
        return torch.tensor([rng_u32(self.seed, self.cur_idx, c) % n])

    def __getattr__(self, name):
        return getattr(torch, name)


def injected_search(mcts, hidden, mask, seed, noise, trace=None):
    """Run the unmodified MCTSSearchVec.search with the two RNG draws injected.  `trace`, if a list,
    receives one dict per simulation with what _expand_nodes returned (rewards, values, policies)
    and which (prev_node, action, node) each tree expanded."""
    import src.mcts as ref_mcts

    proxy = TorchProxy(seed, noise)
    cls = type(mcts)
    orig_ucb, orig_expand = cls.ucb_action, cls._expand_nodes

    def ucb(self, subtree, action_mask, idx):
        proxy.cur_idx = idx
        proxy.ucb_calls += 1
        return orig_ucb(self, subtree, action_mask, idx)

    def expand(self, expand_buffer, last_nodes):
        out = orig_expand(self, expand_buffer, last_nodes)
        if trace is not None:
            states, rewards, policies, values = out
            trace.append(dict(last_nodes=list(last_nodes), rewards=rewards.detach().cpu().clone(),
                              values=values.detach().cpu().clone(), policies=policies.detach().cpu().clone()))
        return out

    saved = ref_mcts.torch
    ref_mcts.torch, cls.ucb_action, cls._expand_nodes = proxy, ucb, expand
    try:
        value, visits = mcts.search(hidden, mask, 0)
    finally:
        ref_mcts.torch, cls.ucb_action, cls._expand_nodes = saved, orig_ucb, orig_expand
    return value, visits, proxy
