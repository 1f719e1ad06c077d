"""oracle/train_oracle.py (CPU restatement of loss_fn / supports_representation / Adam) against tests/golden/train.npz,
which tests/golden/gen_golden.py produced by running the reference's own loss_fn (train_torch.py:33-66) under autograd and
torch.optim.Adam as networks.py:268 builds it."""
import os

import numpy as np
import pytest

from oracle import train_oracle as T

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "train.npz"))
K = int(G["K"])


@pytest.mark.parametrize("case", ["a", "b"])
def test_supports_representation(case):
    """Bit-exact except where torch's vectorised CPU sqrt is 1 ulp off the correctly rounded result (0.7 % of inputs, probed;
    np.sqrt, CUDA's sqrt.rn — what the reference runs on its own device — and the kernel are correctly rounded): those rows
    move the two coefficients by one ulp of the compact value (< 5e-7)."""
    for name, src in (("target_reward", "obs_reward"), ("target_value", "value_target")):
        got = T.supports_representation(G[f"{case}_{src}"].reshape(-1), G["supports"])
        want = G[f"{case}_{name}"].reshape(-1, 11)
        exact = (got.view(np.uint32) == want.view(np.uint32)).all(axis=1)
        assert exact.mean() >= 0.98 and np.abs(got - want).max() <= 5e-7, (name, exact.mean(), np.abs(got - want).max())
        assert np.array_equal(got != 0, want != 0), name


@pytest.mark.parametrize("case", ["a", "b"])
def test_loss_and_gradients(case):
    args = (G[f"{case}_obs_reward"], G[f"{case}_pred_reward"], G[f"{case}_value_target"], G[f"{case}_pred_value"], G[f"{case}_visits"],
            G[f"{case}_pred_policy"], G["supports"], K)
    got = np.array(T.loss_fn(*args), np.float32)
    np.testing.assert_allclose(got, G[f"{case}_losses"], rtol=1e-5, atol=0)          # fp32 tolerance of the north star
    for g, name in zip(T.loss_grads(*args), ("d_reward", "d_value", "d_policy")):
        want = G[f"{case}_{name}"]
        assert np.abs(g - want).max() <= 1e-5 * np.abs(want).max(), name


def test_adam_moments_bit_exact_parameters_within_one_ulp():
    p, m, v = G["adam_p0"], np.zeros_like(G["adam_p0"]), np.zeros_like(G["adam_p0"])
    for s in range(5):
        p, m, v = T.adam_step(p, G[f"adam_g{s}"], m, v, s + 1, lr=float(G["adam_lr"]))
        assert np.array_equal(m, G[f"adam_m{s + 1}"]) and np.array_equal(v, G[f"adam_v{s + 1}"]), f"moments differ at step {s + 1}"
        ulp = np.abs(p.view(np.int32).astype(np.int64) - G[f"adam_p{s + 1}"].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1 and (ulp != 0).mean() < 1e-2, f"parameters differ at step {s + 1}: max {ulp.max()} ulp"
        p = G[f"adam_p{s + 1}"].copy()      # continue from torch's parameters so a 1-ulp difference does not compound
