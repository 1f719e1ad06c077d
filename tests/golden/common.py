"""Helpers shared by the golden-vector generator (reference side) and the tests (oracle / CUDA side)."""
from __future__ import annotations

import numpy as np
import torch


def pack_state(state) -> np.ndarray:
    """(B,3,16,20) 0/1 float state -> (B,120) uint8 bit-packed."""
    s = state.detach().cpu().numpy() if hasattr(state, "detach") else np.asarray(state)
    assert np.all((s == 0) | (s == 1)), "state is not 0/1"
    return np.packbits(s.reshape(s.shape[0], -1).astype(np.uint8), axis=1)


def unpack_state(packed: np.ndarray) -> np.ndarray:
    B = packed.shape[0]
    return np.unpackbits(packed, axis=1)[:, :960].reshape(B, 3, 16, 20).astype(np.float32)


class FakeNet:
    """Deterministic stand-in for MuZeroAgent with the call surface MCTSSearchVec uses
    (src/mcts.py:95,194-195).  Latents are (B,8) float32; every op is element-wise per sample.
    mode: "varied" (spread rewards/values/priors), "deep" (optimistic values -> long chains),
    "flat" (constant outputs -> exact pUCT ties everywhere), "mild" (small rewards/values, near-uniform
    priors, like a random-init network -> bushy trees where the prior term decides)."""

    def __init__(self, mode: str = "varied"):
        self.mode = mode
        self.support = torch.arange(-5, 6, dtype=torch.float32)

    def _logits(self, target, sharp):
        return -sharp * (self.support[None, :] - target[:, None]) ** 2

    def hidden_state_transition(self, h, planes):
        a = planes[:, :, 0, 0].argmax(dim=1).to(torch.float32)
        h2 = torch.frac(h * 1.6180339 + (a[:, None] + 1.0) * 0.7548777 + 0.1234)
        if self.mode == "flat":
            return h2, torch.zeros(h.shape[0], 11)
        if self.mode == "deep":
            return h2, self._logits(h2[:, 0] * 0.4, 2.0)
        if self.mode == "mild":
            return h2, self._logits((h2[:, 0] - 0.5) * 0.2, 0.7)
        return h2, self._logits(h2[:, 0] * 3.0 - 1.0, 1.5)

    def evaluate_state(self, h):
        B = h.shape[0]
        if self.mode == "flat":
            return torch.zeros(B, 3), torch.zeros(B, 11)
        if self.mode == "deep":
            return (h[:, 1:4] - 0.5) * 6.0, self._logits(1.0 + h[:, 4] * 2.0, 2.0)
        if self.mode == "mild":
            return (h[:, 1:4] - 0.5) * 1.0, self._logits((h[:, 4] - 0.5) * 0.4, 0.7)
        return (h[:, 1:4] - 0.5) * 3.0, self._logits(h[:, 4] * 4.0 - 2.0, 1.0)

    # oracle-side convenience (the reference side uses its own ScalarTransforms)
    def inverted_softmax_expectation(self, logits):
        from oracle.networks import inverted_softmax_expectation

        return inverted_softmax_expectation(logits)


def perturb_bn(agent: torch.nn.Module, seed: int = 1) -> None:
    """Give every BatchNorm non-trivial affine + running stats (seeded), so eval-mode BN is not
    the identity in the network goldens.  Applied identically to the reference agent (generator)
    and to the oracle / packed CUDA weights (tests)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in agent.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                n = m.num_features
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)


def dirichlet_noise(B: int, seed: int) -> torch.Tensor:
    """(B,3) Dirichlet(0.25) rows from a seeded CPU generator via normalised gammas."""
    g = torch.Generator().manual_seed(seed)
    x = torch._standard_gamma(torch.full((B, 3), 0.25), generator=g).clamp_min(1e-30)
    return (x / x.sum(dim=1, keepdim=True)).to(torch.float32)
