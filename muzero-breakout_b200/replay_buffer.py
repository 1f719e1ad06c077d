"""Device-resident replay buffer: same classes, constructor and methods as the reference's replay_buffer.py
(`ObservationTrajectory` :4-73, `ReplayBuffer` :76-232), backed by csrc/replay.cu through the C ABI (rb_append /
rb_gather, include/mzb200.h).  SURVEY.md section 8f row 3.

    reference                                             here
    ObservationTrajectory (Python lists, 32 padded rows)  same dataclass (the caller's per-env record, train_torch.py:204-208)
    save_observation_trajectory(trajectory) :96-165       same call -> one rb_append of that trajectory
                                                          save_episode(record): every trajectory of an on-device acting episode
                                                          (acting.Actor.run_episode) in ONE rb_append, no host round trip
    get_batched_{past_actions,future_actions,states,      same calls -> rb_gather; minibatch(idx) returns all six from one launch
      rewards,visit_counts,values}(batch_idxs) :167-210
    get_reward_sums() :212-216, empty_buffer() :218-229,  same
      __len__ / .length / .max_length

A trajectory is stored once (T+1 entries) instead of one 32-frame window per sample; the value targets (:136-152) are
computed by the append kernel with the reference's fp32 rounding sequence (bit-exact, tests/test_replay_gpu.py).
There is no CPU fallback: every call needs a CUDA device and libmzb200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class ObservationTrajectory:
    """Per-environment record the acting loop appends to (reference replay_buffer.py:4-73): five Python lists that
    start with the padding rows of train_torch.py:313-332 (the frame list is one shorter), `length` = real moves."""
    actions: list
    states: list
    rewards: list
    visit_counts: list
    values: list
    length: int
    reward_sum: int

    def add_observation(self, action, state, reward, visit_counts, values):
        for lst, x in ((self.actions, action), (self.states, state), (self.rewards, reward),
                       (self.visit_counts, visit_counts), (self.values, values)):
            lst.append(x)
        self.reward_sum += reward
        self.length += 1

    def get_actions(self):
        return torch.tensor(self.actions)

    def get_states(self):
        return torch.stack(self.states)

    def get_rewards(self):
        return torch.tensor(self.rewards)

    def get_visit_counts(self):
        return torch.stack(self.visit_counts)

    def get_values(self):
        return torch.tensor(self.values)

    def get_reward_sum(self):
        return self.reward_sum


class RbRing(C.Structure):
    """mirror of struct rb_ring (include/mzb200.h)"""
    _fields_ = [("cap_samples", C.c_int32), ("cap_entries", C.c_int32), ("K", C.c_int32), ("hist", C.c_int32),
                ("frame", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p), ("value", C.c_void_p),
                ("visits", C.c_void_p), ("meta", C.c_void_p), ("reward_sum", C.c_void_p), ("target", C.c_void_p),
                ("state", C.c_void_p), ("gpow", C.c_float * 16)]


def _p(t):
    return None if t is None else t.data_ptr()


class ReplayBuffer:
    FRAME = 320

    def __init__(self, seq_len: int, K: int, max_length: int, discount: float, num_rewards_to_sum: int,
                 device=None, output_device: str = "cuda", max_moves: int = 512):
        self.hist_seq_len, self.K, self.max_length = int(seq_len), int(K), int(max_length)
        self.discount, self.num_rewards_to_sum = discount, int(num_rewards_to_sum)
        self.max_moves = int(max_moves)
        self.output_device = output_device
        self._device = device
        self._ring = None
        self._host_counts = [0, 0]       # mirror of the device ring state (entries written, samples appended)
        self._counts_valid = True

    # ---------------------------------------------------------------- device state
    def _alloc(self):
        if self._ring is not None:
            return
        _lib.require_cuda()
        L = _lib.lib()
        dev = torch.device(self._device if self._device is not None else f"cuda:{torch.cuda.current_device()}")
        self._dev = dev
        cap, K = self.max_length, self.K
        ce = int(L.rb_entries_for(cap, K, self.max_moves))
        if ce <= 0 or ce >= 2 ** 31:
            raise ValueError(f"bad replay geometry: max_length={cap} K={K} max_moves={self.max_moves}")
        f32, i32 = torch.float32, torch.int32
        self._t = dict(
            frame=torch.empty((ce, self.FRAME), dtype=f32, device=dev), action=torch.zeros(ce, dtype=i32, device=dev),
            reward=torch.zeros(ce, dtype=f32, device=dev), value=torch.zeros(ce, dtype=f32, device=dev),
            visits=torch.zeros((ce, 3), dtype=f32, device=dev), meta=torch.zeros(cap, dtype=torch.int64, device=dev),
            reward_sum=torch.zeros(cap, dtype=f32, device=dev), target=torch.zeros((cap, K), dtype=f32, device=dev),
            state=torch.zeros(2, dtype=torch.int64, device=dev))
        self._status = torch.zeros(1, dtype=i32, device=dev)
        r = RbRing(cap, ce, K, self.hist_seq_len, *[_p(self._t[k]) for k in
                   ("frame", "action", "reward", "value", "visits", "meta", "reward_sum", "target", "state")])
        for k in range(16):
            r.gpow[k] = float(self.discount) ** k           # Python double power, rounded to fp32 on assignment (:142-148)
        self._ring = r
        self._host_counts, self._counts_valid = [0, 0], True

    def _stream(self):
        return torch.cuda.current_stream(self._dev).cuda_stream

    def _sync_counts(self):
        if self._ring is not None and not self._counts_valid:
            e, s = self._t["state"].tolist()
            self._host_counts, self._counts_valid = [int(e), int(s)], True
            self._raise_on_status()

    def _raise_on_status(self):
        st = int(self._status.item())
        if st:
            self._status.zero_()
            raise IndexError(f"replay buffer kernel status {st} (1 = bad trajectory length, 2 = sample index out of range)")

    @property
    def length(self) -> int:
        self._sync_counts()
        return min(self._host_counts[1], self.max_length)

    def __len__(self):
        return self.length

    # ---------------------------------------------------------------- append
    def _append(self, action, reward, value, visits, frames, init_frame, pad_action, recorded, lengths, min_length):
        L = _lib.lib()
        T, B = int(action.shape[0]), int(init_frame.shape[0])
        if T > self.max_moves:
            raise ValueError(f"trajectory of {T} moves exceeds max_moves={self.max_moves}")
        plan = torch.empty(int(L.rb_plan_bytes(B)), dtype=torch.uint8, device=self._dev)
        _lib.check(L.rb_append(C.byref(self._ring), B, T, _p(action), _p(reward), _p(value), _p(visits), _p(frames), _p(init_frame),
                               int(pad_action), _p(recorded), _p(lengths), int(min_length), self.max_moves, _p(plan), _p(self._status),
                               self._stream()))

    def save_observation_trajectory(self, observation_trajectory: ObservationTrajectory):
        """replay_buffer.py:96-165.  The padding rows must be the constant rows _pad_initial_state creates
        (train_torch.py:313-332): one repeated action, one repeated initial frame."""
        self._alloc()
        ot, h, dev = observation_trajectory, self.hist_seq_len, self._dev
        T = len(ot.actions) - h
        if T != ot.length or len(ot.states) != h - 1 + T:
            raise ValueError("ObservationTrajectory lists do not have the 32-row padding of train_torch.py:313-332")
        pad = ot.actions[:h]
        if any(int(a) != int(pad[0]) for a in pad) or any(s is not ot.states[0] and not torch.equal(s, ot.states[0]) for s in ot.states[1:h - 1]):
            raise ValueError("padding rows are not constant; the trajectory-compressed ring cannot represent them")
        f32 = dict(dtype=torch.float32, device=dev)

        def col(items, **kw):
            if not items:
                return torch.zeros(0, **kw)
            return torch.stack([torch.as_tensor(x) for x in items]).to(**kw)

        action = col(ot.actions[h:], dtype=torch.int64, device=dev).view(T, 1)
        reward = col(ot.rewards[h:], **f32).view(T, 1)
        value = col(ot.values[h:], **f32).view(T, 1)
        visits = col(ot.visit_counts[h:], dtype=torch.int64, device=dev).view(T, 1, 3)
        frames = col(ot.states[h - 1:], **f32).reshape(T, 1, self.FRAME)
        init = torch.as_tensor(ot.states[0]).to(**f32).reshape(1, self.FRAME).contiguous()
        lengths = torch.tensor([T], dtype=torch.int32, device=dev)
        self._append(action.contiguous(), reward.contiguous(), value.contiguous(), visits.contiguous(), frames.contiguous(), init,
                     int(pad[0]), None, lengths, 0)
        if self._counts_valid and T >= self.K:
            self._host_counts[0] += T + 1
            self._host_counts[1] += T - self.K + 1

    def save_episode(self, record: dict, min_length: int | None = None):
        """All trajectories of one on-device acting episode (acting.Actor.run_episode's record: action/reward/value (T,B),
        visits (T,B,3), frames (T,B,1,16,20), recorded (T,B) bool, initial_gray (B,1,16,20)) in one rb_append.
        min_length defaults to K+2, the caller-side filter of train_torch.py:223-225 (`length > K + 1`)."""
        self._alloc()
        dev = self._dev
        action = record["action"].to(device=dev, dtype=torch.int64).contiguous()
        T, B = action.shape
        reward = record["reward"].to(device=dev, dtype=torch.float32).contiguous()
        value = record["value"].to(device=dev, dtype=torch.float32).contiguous()
        visits = record["visits"].to(device=dev, dtype=torch.int64).contiguous()
        frames = record["frames"].to(device=dev, dtype=torch.float32).reshape(T, B, self.FRAME).contiguous()
        init = record["initial_gray"].to(device=dev, dtype=torch.float32).reshape(B, self.FRAME).contiguous()
        recorded = record["recorded"].to(device=dev).contiguous().view(torch.uint8)
        self._append(action, reward, value, visits, frames, init, 0, recorded, None, self.K + 2 if min_length is None else min_length)
        self._counts_valid = False           # lengths live on the device; .length reads the ring state when asked

    # ---------------------------------------------------------------- gather
    def _gather(self, batch_idxs, want):
        self._alloc()
        L, dev, K, h = _lib.lib(), self._dev, self.K, self.hist_seq_len
        idx = torch.as_tensor(batch_idxs).to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        n = int(idx.numel())
        shapes = dict(past_actions=((n, h), torch.int64), future_actions=((n, K), torch.int64),
                      states=((n, h, 1, 16, 20), torch.float32), rewards=((n, K), torch.float32),
                      visit_counts=((n, K, 3), torch.float32), values=((n, K), torch.float32),
                      value_buffer=((n, K), torch.float32), reward_sums=((n,), torch.float32))
        out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=dev) for k in want}      # fresh per call (callers keep them)
        _lib.check(L.rb_gather(C.byref(self._ring), n, _p(idx), *[_p(out.get(k)) for k in
                               ("past_actions", "future_actions", "states", "rewards", "visit_counts", "values", "value_buffer", "reward_sums")],
                               _p(self._status), self._stream()))
        if self.output_device == "cpu":
            out = {k: v.cpu() for k, v in out.items()}
            self._raise_on_status()
        return out

    def get_batched_past_actions(self, batch_idxs):
        return self._gather(batch_idxs, ("past_actions",))["past_actions"]

    def get_batched_future_actions(self, batch_idxs):
        return self._gather(batch_idxs, ("future_actions",))["future_actions"]

    def get_batched_states(self, batch_idxs):
        return self._gather(batch_idxs, ("states",))["states"]

    def get_batched_rewards(self, batch_idxs):
        return self._gather(batch_idxs, ("rewards",))["rewards"]

    def get_batched_visit_counts(self, batch_idxs):
        return self._gather(batch_idxs, ("visit_counts",))["visit_counts"]

    def get_batched_values(self, batch_idxs):
        return self._gather(batch_idxs, ("values",))["values"]

    def minibatch(self, batch_idxs):
        """The six tensors train_torch.py:455-484 _prepare_minibatch assembles, from ONE launch:
        (past_actions, states, visit_counts, future_actions, rewards, values)."""
        o = self._gather(batch_idxs, ("past_actions", "states", "visit_counts", "future_actions", "rewards", "values"))
        return o["past_actions"], o["states"], o["visit_counts"], o["future_actions"], o["rewards"], o["values"]

    def repnet_input(self, batch_idxs, n_actions=3):
        """`torch.cat((input_states, input_actions_encoded), dim=1)` of train_torch.py:500 for the samples `batch_idxs`, (n, 2*seq_len, 16, 20),
        from ONE launch (rb_gather_input): the frame window + `_encode_actions` (:279-293) of the past actions, no intermediate tensors."""
        self._alloc()
        dev, h = self._dev, self.hist_seq_len
        idx = torch.as_tensor(batch_idxs).to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        n = int(idx.numel())
        out = torch.empty((n, 2 * h, 16, 20), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().rb_gather_input(C.byref(self._ring), n, _p(idx), int(n_actions), _p(out), _p(self._status), self._stream()))
        if self.output_device == "cpu":
            out = out.cpu()
            self._raise_on_status()
        return out

    def get_reward_sums(self):
        """replay_buffer.py:212-216: the reward sums of the newest num_rewards_to_sum samples, as Python floats."""
        n = min(self.length, self.num_rewards_to_sum)
        if n == 0:
            return []
        idx = torch.arange(-n, 0, dtype=torch.int64)
        out = self._gather(idx, ("reward_sums",))["reward_sums"].cpu().tolist()
        self._raise_on_status()
        return out

    def empty_buffer(self):
        if self._ring is not None:
            self._t["state"].zero_()
        self._host_counts, self._counts_valid = [0, 0], True

    # ---------------------------------------------------------------- the reference's list attributes (checkpoint code reads them, train_torch.py:627-636)
    def _materialise(self, field):
        n = self.length
        if n == 0:
            return []
        return list(self._gather(torch.arange(n), (field,))[field].cpu().unbind(0))

    past_actions_buffer = property(lambda self: self._materialise("past_actions"))
    future_actions_buffer = property(lambda self: self._materialise("future_actions"))
    state_buffer = property(lambda self: self._materialise("states"))
    reward_buffer = property(lambda self: self._materialise("rewards"))
    visit_counts_buffer = property(lambda self: self._materialise("visit_counts"))
    value_buffer = property(lambda self: self._materialise("value_buffer"))
    bootstrapped_values = property(lambda self: self._materialise("values"))

    @property
    def reward_sums(self):
        n = self.length
        return [] if n == 0 else self._gather(torch.arange(n), ("reward_sums",))["reward_sums"].cpu().tolist()

    def state_dict(self):
        self._alloc()
        self._sync_counts()
        return {k: v.cpu() for k, v in self._t.items()} | {"geometry": (self.hist_seq_len, self.K, self.max_length, self.max_moves)}

    def load_state_dict(self, sd):
        if tuple(sd["geometry"]) != (self.hist_seq_len, self.K, self.max_length, self.max_moves):
            raise ValueError("replay buffer geometry differs from the checkpoint's")
        self._alloc()
        for k, v in self._t.items():
            v.copy_(sd[k])
        self._counts_valid = False
