"""TEST INFRASTRUCTURE ONLY — CPU restatement of the loss and optimizer ends of the reference's training step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(muzero-breakout_b200/) never does.

* supports_representation  — utils.py:30-64 (ScalarTransforms), with _invertible_transform_normal_to_compact :21-24
* loss_fn                  — train_torch.py:33-66 (three F.kl_div(log_softmax(pred), target, "batchmean") + (1/K) * sum)
* loss_grads               — what loss.backward() (train_torch.py:515) leaves in the three logit tensors
* adam_step                — torch.optim.Adam(lr, weight_decay=1e-4) of networks.py:268, single-tensor path of
                             torch/optim/adam.py (third-party arithmetic, torch 2.11: not pinned by any reference test)

Plain per-row Python/numpy loops in fp32 (rows are few); pinned by tests/golden/train.npz, which
tests/golden/gen_golden.py produced by running the reference's own loss_fn under autograd and torch.optim.Adam.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


def to_compact(x: np.float32) -> np.float32:
    """utils.py:24: sign(x) * (sqrt(|x| + 1) - 1 + eps * x), each op rounded to fp32."""
    x = F32(x)
    r = F32(F32(F32(np.sqrt(F32(abs(x) + F32(1)))) - F32(1)) + F32(F32(0.001) * x))
    return F32(np.sign(x) * r)


def supports_representation(target: np.ndarray, supports: np.ndarray) -> np.ndarray:
    """utils.py:30-64 for a flat array of scalars -> (rows, n) two-hot coefficient vectors."""
    n = len(supports)
    out = np.zeros((len(target), n), F32)
    for r, x in enumerate(target):
        t = to_compact(x)
        lo = int(np.searchsorted(supports, t, side="right")) - 1          # :46
        lo = min(max(lo, 0), n - 2)                                        # :47
        s_lo, s_hi = F32(supports[lo]), F32(supports[lo + 1])
        p_lo = F32(F32(s_hi - t) / F32(F32(s_hi - s_lo) + F32(1e-10)))     # :56
        p_hi = F32(F32(1) - p_lo)                                          # :57
        out[r, lo] = p_lo                                                  # :61
        out[r, lo + 1] = p_hi                                              # :62
    return out


def _log_softmax(z: np.ndarray) -> np.ndarray:
    z = z.astype(F32)
    s = z - z.max(axis=-1, keepdims=True)
    return (s - np.log(np.exp(s).sum(axis=-1, keepdims=True, dtype=F32))).astype(F32)


def _kl_batchmean(logp: np.ndarray, t: np.ndarray) -> float:
    """F.kl_div(logp, t, reduction="batchmean") = sum(xlogy(t, t) - t * logp) / rows; the sum is taken in double here
    (torch's own fp32 summation order is an implementation detail; tests compare at 1e-5)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        xlogy = np.where(t == 0, F32(0), t * np.log(t)).astype(F32)
    return float((xlogy.astype(np.float64) - t.astype(np.float64) * logp.astype(np.float64)).sum() / logp.shape[0])


def _targets(observed_reward, value_target, visit_counts, supports):
    t_r = supports_representation(np.asarray(observed_reward, F32).reshape(-1), supports)
    t_v = supports_representation(np.asarray(value_target, F32).reshape(-1), supports)
    v = np.asarray(visit_counts, F32).reshape(-1, np.shape(visit_counts)[-1])
    t_p = (v / v.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)       # train_torch.py:58
    return t_r, t_v, t_p


def loss_fn(observed_reward, predicted_reward, bootstrapped_reward, predicted_value, visit_counts, predicted_policy, supports, K):
    """train_torch.py:33-66 -> (total, reward_loss, value_loss, policy_loss) as fp32 scalars."""
    supports = np.asarray(supports, F32)
    t_r, t_v, t_p = _targets(observed_reward, bootstrapped_reward, visit_counts, supports)
    n, a = predicted_reward.shape[-1], predicted_policy.shape[-1]
    rl = F32(_kl_batchmean(_log_softmax(np.asarray(predicted_reward, F32).reshape(-1, n)), t_r))
    vl = F32(_kl_batchmean(_log_softmax(np.asarray(predicted_value, F32).reshape(-1, n)), t_v))
    pl = F32(_kl_batchmean(_log_softmax(np.asarray(predicted_policy, F32).reshape(-1, a)), t_p))
    return F32(F32(1.0 / K) * F32(F32(rl + vl) + pl)), rl, vl, pl


def loss_grads(observed_reward, predicted_reward, bootstrapped_reward, predicted_value, visit_counts, predicted_policy, supports, K):
    """d total / d logits for the three heads: (1/K)/rows * (softmax(z) * sum(t) - t)  (kl_div + log_softmax backward)."""
    supports = np.asarray(supports, F32)
    ts = _targets(observed_reward, bootstrapped_reward, visit_counts, supports)
    out = []
    for z, t in zip((predicted_reward, predicted_value, predicted_policy), ts):
        z2 = np.asarray(z, np.float64).reshape(-1, np.shape(z)[-1])
        sm = np.exp(z2 - z2.max(axis=-1, keepdims=True))
        sm /= sm.sum(axis=-1, keepdims=True)
        g = (sm * t.astype(np.float64).sum(axis=-1, keepdims=True) - t) * ((1.0 / K) / z2.shape[0])
        out.append(g.astype(F32).reshape(np.shape(z)))
    return tuple(out)


def _fma(a, b, c):
    """fp32 fused multiply-add (the 48-bit product is exact in double; one rounding to fp32)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4):
    """One torch.optim.Adam update (torch/optim/adam.py _single_tensor_adam; defaults = networks.py:268 + config.yaml:28).
    `step` counts this update.  Returns new (param, exp_avg, exp_avg_sq).  The FMA contractions are the ones torch's CPU
    kernels make (probed, tests/golden/gen_golden.py gen_train)."""
    p, g, m, v = (np.asarray(x, F32) for x in (param, grad, exp_avg, exp_avg_sq))
    if weight_decay != 0:
        g = _fma(F32(weight_decay), p, g)                                   # grad.add(param, alpha=weight_decay)
    m = _fma(F32(1 - beta1), (g - m).astype(F32), m)                        # exp_avg.lerp_(grad, 1 - beta1)
    v = _fma((F32(1 - beta2) * g).astype(F32), g, (v * F32(beta2)).astype(F32))   # mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = ((np.sqrt(v).astype(F32) / F32(math.sqrt(bc2))).astype(F32) + F32(eps)).astype(F32)
    p = (p + ((F32(-(lr / bc1)) * m).astype(F32) / denom).astype(F32)).astype(F32)   # addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v
