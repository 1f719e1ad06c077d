"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's replay buffer (replay_buffer.py) in numpy.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(muzero-breakout_b200/replay_buffer.py -> csrc/replay.cu) never does.

Pinned against tests/golden/replay.npz, which tests/golden/gen_golden.py produced by running the UNMODIFIED
reference classes (ObservationTrajectory + ReplayBuffer) fed the way train_torch.py feeds them.

What it restates (reference file:line):
  * trajectory padding, train_torch.py:313-332: 32 zero actions / rewards / values / visit-count rows and 31 copies of
    the initial gray frame in front of the real moves (the frame list is one shorter than the others);
  * ReplayBuffer.save_observation_trajectory, replay_buffer.py:96-165: one sample per start s in
    range(length - K + 1): past actions [s, s+32), the 32 frames [s, s+32) of the (shorter) frame list, then K future
    actions / rewards / visit counts / values from list index s+32;
  * the value targets, :136-151: td_steps = 10; target(kk) = values[b] * discount**K + sum_{k<10} discount**k * r[c+k] when
    the bootstrap index b = c + 10 is inside the trajectory, else the plain discounted tail sum; every product and add is
    a separately rounded fp32 op (0-d fp32 tensors; the Python-double power is rounded to fp32 when it meets the tensor);
    NB the bootstrap is discounted by discount**K (K = 5), not **10 -- reproduced as written;
  * reward_sum, replay_buffer.py:34: sequential fp32 sum of the trajectory's rewards, stored once per sample (:122);
  * FIFO eviction at max_length, :154-163; get_reward_sums, :212-216.

Design: a plain Python list of per-sample numpy records (the reference's own shape), nothing shared with the device
ring of csrc/replay.cu.
"""
from __future__ import annotations

import numpy as np

TD_STEPS = 10          # replay_buffer.py:137


def pad_trajectory(init_frame, frames, action, reward, visits, value, hist=32):
    """train_torch.py:313-332 + :204-208 -> the five lists of an ObservationTrajectory as arrays."""
    T = len(action)
    A = np.concatenate([np.zeros(hist, np.int64), np.asarray(action, np.int64)])
    R = np.concatenate([np.zeros(hist, np.float32), np.asarray(reward, np.float32)])
    V = np.concatenate([np.zeros(hist, np.float32), np.asarray(value, np.float32)])
    N = np.concatenate([np.zeros((hist, 3), np.float32), np.asarray(visits, np.float32).reshape(T, 3)])
    init = np.asarray(init_frame, np.float32).reshape(1, 1, 16, 20)
    S = np.concatenate([np.repeat(init, hist - 1, 0), np.asarray(frames, np.float32).reshape(T, 1, 16, 20)])
    return A, S, R, N, V, T


def reward_sum(reward):
    acc = np.float32(0.0)                      # int 0 + fp32 tensor (replay_buffer.py:34)
    for r in np.asarray(reward, np.float32):
        acc = np.float32(acc + r)
    return float(acc)


def value_targets(R, V, T, s, K, discount, hist=32):
    """replay_buffer.py:136-151 for the sample starting at s (state_start = s + hist)."""
    g = [np.float32(discount ** k) for k in range(max(K, TD_STEPS) + T + 1)]
    state_start = s + hist
    max_length = hist + T
    out = np.zeros(K, np.float32)
    bootstrap = state_start + TD_STEPS
    for j, cur in enumerate(range(state_start, state_start + K)):
        if bootstrap < max_length:
            vt = np.float32(V[bootstrap] * g[K])
            for k, r in enumerate(R[cur:bootstrap]):
                vt = np.float32(vt + np.float32(g[k] * r))
        else:
            vt = np.float32(0.0)
            for k, r in enumerate(R[cur:max_length]):
                vt = np.float32(vt + np.float32(g[k] * r))
        out[j] = vt
        bootstrap += 1
    return out


class ReplayOracle:
    def __init__(self, seq_len, K, max_length, discount, num_rewards_to_sum):
        self.hist, self.K, self.max_length, self.discount, self.n_sum = seq_len, K, max_length, discount, num_rewards_to_sum
        self.samples = []          # dicts of numpy arrays, oldest first

    def __len__(self):
        return len(self.samples)

    def save(self, init_frame, frames, action, reward, visits, value):
        A, S, R, N, V, T = pad_trajectory(init_frame, frames, action, reward, visits, value, self.hist)
        rs = reward_sum(reward)
        h, K = self.hist, self.K
        for s in range(T - K + 1):
            a = s + h
            self.samples.append(dict(
                past_actions=A[s:a].copy(), future_actions=A[a:a + K].copy(), states=S[s:a].copy(),
                rewards=R[a:a + K].copy(), visit_counts=N[a:a + K].copy(), value_buffer=V[a:a + K].copy(),
                values=value_targets(R, V, T, s, K, self.discount, h), reward_sum=rs))
            if len(self.samples) > self.max_length:
                self.samples.pop(0)

    def batch(self, field, idxs):
        return np.stack([self.samples[int(i)][field] for i in idxs])

    def reward_sums(self):
        return [x["reward_sum"] for x in self.samples][-self.n_sum:]
