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
        self._pending = {}               # the reference's list attributes assigned by a checkpoint load (train_torch.py:659-668)
        self._min_entries = 0

    # ---------------------------------------------------------------- device state
    def _alloc(self):
        if self._ring is not None:
            return
        _lib.require_cuda()
        L = _lib.lib()
        dev = torch.device(self._device if self._device is not None else f"cuda:{torch.cuda.current_device()}")
        self._dev = dev
        cap, K = self.max_length, self.K
        ce = max(int(L.rb_entries_for(cap, K, self.max_moves)), int(self._min_entries))
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

    @length.setter
    def length(self, n):
        """train_torch.py:666 assigns it after the lists while loading a checkpoint: accepted when it agrees with them (the ring state is the
        authority); `rb.length = 0` alone empties the buffer."""
        n = int(n)
        lists = [v for k, v in self._pending.items() if k != "reward_sums"]
        if lists and any(len(v) != n for v in lists):
            raise ValueError(f"length = {n} disagrees with the assigned buffers ({[len(v) for v in lists]} samples)")
        if not self._pending and n != self.length:
            if n != 0:
                raise ValueError("length can only be assigned together with the sample lists (checkpoint load) or set to 0")
            self.empty_buffer()

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
        idx = torch.as_tensor(batch_idxs)
        self._check_indices(idx)
        idx = idx.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
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

    def _check_indices(self, idx):
        """The reference raises IndexError on an out-of-range sample index (list indexing, replay_buffer.py:174).  Host indices are checked
        here whenever the sample count is known without a device read (always, unless the last append was a device-side save_episode);
        device-resident indices are checked by the kernel, whose status word is read by the next host-facing call (.length,
        get_reward_sums, any output_device="cpu" getter) -- deferred, like every asynchronous CUDA error."""
        if not idx.is_cuda and idx.numel() and self._counts_valid:
            n = min(self._host_counts[1], self.max_length)
            lo, hi = int(idx.min()), int(idx.max())
            if lo < -n or hi >= n:
                raise IndexError(f"sample index out of range for a replay buffer of {n} samples")

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
        idx = torch.as_tensor(batch_idxs)
        self._check_indices(idx)
        idx = idx.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
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

    # Assigning them (the reference's _load_weights, train_torch.py:659-668, assigns all eight one after the other) restores the buffer:
    # the assignments are collected and, once the set is complete, _restore() rebuilds the device rings from the per-sample lists.
    _LIST_FIELDS = {"past_actions_buffer": "past_actions", "future_actions_buffer": "future_actions", "state_buffer": "states",
                    "reward_buffer": "rewards", "visit_counts_buffer": "visit_counts", "value_buffer": "value_buffer",
                    "bootstrapped_values": "values"}

    def _list_property(attr, field):              # noqa: N805 (class-body helper)
        def get(self):
            return self._materialise(field)

        def set_(self, value):
            self._pending[attr] = list(value)
            self._restore_if_complete()
        return property(get, set_)

    for _attr, _field in _LIST_FIELDS.items():
        locals()[_attr] = _list_property(_attr, _field)
    del _attr, _field, _list_property

    @property
    def reward_sums(self):
        n = self.length
        return [] if n == 0 else self._gather(torch.arange(n), ("reward_sums",))["reward_sums"].cpu().tolist()

    @reward_sums.setter
    def reward_sums(self, value):
        self._pending["reward_sums"] = [float(v) for v in value]
        self._restore_if_complete()

    def _restore_if_complete(self):
        need = set(self._LIST_FIELDS) | {"reward_sums"}
        if need <= set(self._pending):
            pend, self._pending = self._pending, {}
            self._restore(pend)

    def _restore(self, pend):
        """Rebuild the rings from the reference's per-sample lists (sample j: 32 past actions, 32-frame window, K future actions / rewards /
        visit counts / values / value targets, reward sum; oldest first).  Consecutive samples that are each other's shift by one move
        (samples s, s+1 of one trajectory, the way save_observation_trajectory appends them, replay_buffer.py:106-152) are merged into one
        stored trajectory whose first sample starts at padded-list index 32 (beyond the padding rows, which the lists no longer identify);
        a sample that continues nothing starts a trajectory of its own.  Every get_batched_* output is then exactly the assigned row."""
        n = len(pend["past_actions_buffer"])
        if any(len(v) != n for v in pend.values()):
            raise ValueError("the assigned replay-buffer lists have different lengths")
        if n > self.max_length:
            raise ValueError(f"{n} samples assigned to a buffer of max_length {self.max_length}")
        self.empty_buffer()
        if n == 0:
            return
        h, K, FR = self.hist_seq_len, self.K, self.FRAME
        st = lambda key, dt, shape: torch.stack([torch.as_tensor(x) for x in pend[key]]).to(dt).reshape((n,) + shape).cpu()
        past, fut = st("past_actions_buffer", torch.int64, (h,)), st("future_actions_buffer", torch.int64, (K,))
        frames = st("state_buffer", torch.float32, (h, FR))
        rew, vis = st("reward_buffer", torch.float32, (K,)), st("visit_counts_buffer", torch.float32, (K, 3))
        val, tgt = st("value_buffer", torch.float32, (K,)), st("bootstrapped_values", torch.float32, (K,))
        rsum = torch.tensor(pend["reward_sums"], dtype=torch.float32)
        # sample j continues sample j-1 iff every window is its shift by one move
        cont = torch.zeros(n, dtype=torch.bool)
        if n > 1:
            c = (past[1:, :-1] == past[:-1, 1:]).all(1) & (past[1:, -1] == fut[:-1, 0]) & (fut[1:, :-1] == fut[:-1, 1:]).all(1)
            c &= (frames[1:, :-1] == frames[:-1, 1:]).flatten(1).all(1) & (rew[1:, :-1] == rew[:-1, 1:]).all(1)
            c &= (val[1:, :-1] == val[:-1, 1:]).all(1) & (vis[1:, :-1] == vis[:-1, 1:]).flatten(1).all(1) & (rsum[1:] == rsum[:-1])
            cont[1:] = c
        first = torch.nonzero(~cont).flatten()                           # first sample of every run
        run_of = torch.cumsum((~cont).long(), 0) - 1                      # run index of every sample
        d = torch.arange(n) - first[run_of]                               # position in its run
        run_len = torch.diff(torch.cat([first, torch.tensor([n])]))
        if int(run_len.max()) + h + K > self.max_moves:
            raise ValueError("a restored trajectory exceeds max_moves")
        ent_cnt = h + K + run_len                                         # entries 0 .. 31 + K + m of a run of m samples
        ent_base = torch.cumsum(ent_cnt, 0) - ent_cnt
        total_e = int(ent_cnt.sum())
        self._min_entries = total_e + int(_lib.lib().rb_entries_for(self.max_length, K, self.max_moves))
        if self._ring is not None and (self._ring.cap_entries < self._min_entries or self._ring.cap_samples != self.max_length):
            self._ring = None                                             # needs a larger entry ring than the acting loop's / max_length was reassigned (:667)
        self._alloc()
        t, dev = self._t, self._dev
        e_frame = torch.zeros((total_e, FR)); e_act = torch.zeros(total_e, dtype=torch.int32)
        e_rew = torch.zeros(total_e); e_val = torch.zeros(total_e); e_vis = torch.zeros((total_e, 3))
        fb, ar = ent_base[:, None], torch.arange
        # first sample of a run (start 32): frames -> entries 2..33, past actions -> entries 1..32, futures -> entries 33..32+K
        e_frame[(fb + 2 + ar(h)[None, :]).flatten()] = frames[first].reshape(-1, FR)
        e_frame[ent_base] = frames[first, 0]; e_frame[ent_base + 1] = frames[first, 0]
        e_act[(fb + 1 + ar(h)[None, :]).flatten()] = past[first].flatten().int()
        fut_idx = (fb + h + 1 + ar(K)[None, :]).flatten()
        e_act[fut_idx] = fut[first].flatten().int(); e_rew[fut_idx] = rew[first].flatten(); e_val[fut_idx] = val[first].flatten()
        e_vis[fut_idx] = vis[first].reshape(-1, 3)
        # every later sample of a run adds one move: its newest frame -> entry 33 + d, its last future row -> entry 32 + K + d
        later = torch.nonzero(cont).flatten()
        if later.numel():
            eb = ent_base[run_of[later]]
            e_frame[eb + h + 1 + d[later]] = frames[later, -1]
            li = eb + h + K + d[later]
            e_act[li] = fut[later, -1].int(); e_rew[li] = rew[later, -1]; e_val[li] = val[later, -1]; e_vis[li] = vis[later, -1]
        meta = ent_base[run_of] | ((h + d) << 32) | ((ent_cnt[run_of] - 1) << 48)
        for key, src in (("frame", e_frame), ("action", e_act), ("reward", e_rew), ("value", e_val), ("visits", e_vis)):
            t[key][:total_e].copy_(src.to(dev))
        t["meta"][:n].copy_(meta.to(dev)); t["target"][:n].copy_(tgt.to(dev)); t["reward_sum"][:n].copy_(rsum.to(dev))
        t["state"].copy_(torch.tensor([total_e, n], dtype=torch.int64))
        self._host_counts, self._counts_valid = [total_e, n], True

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
