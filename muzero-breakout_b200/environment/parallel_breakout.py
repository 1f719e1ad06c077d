"""Drop-in for the reference's `environment/parallel_breakout.py` (class BreakoutEnvironment :59,
abstract base MuZeroEnvironment :11) backed by the sm_100a kernels of libmzb200.so (csrc/env.cu).

Same constructor, attributes and call signatures, so the reference's train_torch.py loads it through
config.yaml (`environment_path: muzero_breakout_b200.environment.parallel_breakout`,
train_torch.py:93-94):

    reset() -> (state f32 (B,3,16,20), 0)                                         :107-139
    step(state, action i64 (B,), done_mask bool (B,)) -> (next_state, reward f32 (B,),
         done_mask [the SAME object, mutated in place], valid_actions f32 (B,3))   :158-254
    get_valid_actions(state, paddle_pos_new)                                      :141-155
    .batch (assignable, train_torch.py:448,452), .ball_dx int64 (B,), .ball_dy float32 (B,)

Differences, all opt-in through extra cfg keys (the defaults reproduce the reference):
  * the authoritative state is a structure-of-arrays in HBM (include/mzb200.h); the dense frame a
    caller passes to step() is only re-parsed ("ingested") when it is not the tensor step()/reset()
    returned last;
  * cfg["output_device"]: "cpu" (default, like the reference: host tensors, device<->host copies per
    call) or "cuda" (device-resident tensors, no host round trip);
  * cfg["reset_rng"]: "torch" (default: the reference's four torch CPU RNG calls in its order, so
    torch.manual_seed(s) gives the reference's initial state) or "device" (counter-based, no host RNG);
  * step(..., want_gray=True) additionally returns convert_to_grayscale(next_state)
    (train_torch.py:334-358) from the same kernel.
There is no CPU implementation: a CUDA device is required.
"""
from __future__ import annotations

import ctypes
from abc import ABC, abstractmethod
from typing import Tuple

import torch

from .. import _lib


class MuZeroEnvironment(ABC):
    """Interface of the reference's abstract base (parallel_breakout.py:11-56)."""

    @abstractmethod
    def reset(self):
        ...

    @abstractmethod
    def step(self, state, action, done_mask):
        ...

    @abstractmethod
    def get_valid_actions(self, state, paddle_pos_new):
        ...

    @property
    @abstractmethod
    def action_space_size(self) -> int:
        ...

    @property
    @abstractmethod
    def state_shape(self) -> Tuple[int, ...]:
        ...


def _ptr(t):
    return None if t is None else t.data_ptr()


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> freshly allocated pinned host tensor (async copy; caller synchronises the stream)."""
    out = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    out.copy_(t, non_blocking=True)
    return out


def _status_message(bits: int) -> str:
    msgs = []
    if bits & 1:
        msgs.append("ball left the grid (the reference raises IndexError, parallel_breakout.py:243)")
    if bits & 2:
        msgs.append("ingested state is not a valid Breakout frame (0/1 planes, one ball pixel, 6-cell paddle on row 15)")
    if bits & 4:
        msgs.append("action outside {0,1,2}")
    return "; ".join(msgs)


class BreakoutEnvironment(MuZeroEnvironment):
    def __init__(self, cfg: dict, width: int = 10, height: int = 15, paddle_width: int = 6, brick_rows: int = 3,
                 device: str = "cuda"):
        # like the reference (:76-80) the geometry is fixed: cfg["resolution"] / cfg["brick_rows"] and the
        # width/height/brick_rows arguments are ignored; only paddle_width=6 is supported by the kernels
        if paddle_width != 6:
            raise ValueError("only paddle_width=6 (the reference default, parallel_breakout.py:72) is built")
        self.height, self.width = 16, 20
        self.paddle_width = 6
        self.brick_rows = 3
        self.batch = cfg["n_parallel"]
        self.paddle_hit_reward = cfg["paddle_hit_reward"]
        self.brick_hit_reward = cfg["brick_hit_reward"]
        self.game_lost_reward = cfg["game_lost_reward"]
        self.game_won_reward = cfg["game_won_reward"]
        self.CHANNEL_PADDLE, self.CHANNEL_BALL, self.CHANNEL_BRICKS = 0, 1, 2
        self._action_space_size = 3
        self.device = cfg.get("output_device", "cpu")          # where returned tensors live
        self.reset_rng = cfg.get("reset_rng", "torch")
        self.seed = int(cfg.get("seed", 0))
        self._cuda = torch.device(cfg.get("cuda_device", "cuda"))
        self._episode = 0
        self._B = 0                                             # allocated batch
        self._hdr = self._bricks = self._status = None
        self._last_state = None                                 # tensor handed out last (identity check)
        self._last_version = -1
        self._host_io = {}                                      # (B, want_state, want_gray) -> packed staging buffers of the host-tensor step
        self._rewards = (ctypes.c_float * 4)(self.paddle_hit_reward, self.brick_hit_reward,
                                             self.game_lost_reward, self.game_won_reward)

    # ------------------------------------------------------------------ reference surface
    @property
    def action_space_size(self) -> int:
        return self._action_space_size

    @property
    def state_shape(self) -> Tuple[int, ...]:
        return (self.batch, 3, self.height, self.width)

    def reset(self):
        _lib.require_cuda()
        B = self._alloc()
        L, st = _lib.lib(), self._stream()
        state = torch.empty((B, 3, 16, 20), dtype=torch.float32, device=self._cuda)
        if self.reset_rng == "torch":
            # the reference's RNG calls, same order (:116,:126,:127,:136); one batched randint(0,2,(B,))
            # consumes the global CPU generator exactly like its B scalar draws
            draws = torch.stack([torch.randint(-6, 8, (B,)), torch.randint(1, 19, (B,)), torch.randint(-3, -1, (B,)),
                                 torch.randint(0, 2, (B,))]).to(self._cuda, non_blocking=True)
            _lib.check(L.bk_env_reset(B, _ptr(self._hdr), _ptr(self._bricks), _ptr(draws[0]), _ptr(draws[1]),
                                      _ptr(draws[2]), _ptr(draws[3]), _ptr(state), st))
        else:
            _lib.check(L.bk_env_reset_device_rng(B, _ptr(self._hdr), _ptr(self._bricks), self.seed, self._episode,
                                                 _ptr(state), st))
        self._episode += 1
        out = self._hand_out(state)
        if self.device == "cpu":
            torch.cuda.current_stream(self._cuda).synchronize()
        return out, 0

    def get_valid_actions(self, state: torch.Tensor, paddle_pos_new: torch.Tensor) -> torch.Tensor:
        valid = torch.ones((paddle_pos_new.shape[0], 3), device=paddle_pos_new.device)
        valid[paddle_pos_new == 0, 0] = 0
        valid[paddle_pos_new + self.paddle_width >= self.width, -1] = 0
        return valid

    def step(self, state: torch.Tensor, action: torch.Tensor, done_mask: torch.Tensor, want_gray: bool = False,
             want_state: bool = True):
        _lib.require_cuda()
        B = self._alloc()
        L, st = _lib.lib(), self._stream()
        dev = self._cuda
        if state is not None and not self._is_ours(state):
            self._ingest(state)
        if self.device == "cpu" and not action.is_cuda and not done_mask.is_cuda:
            return self._step_host(B, action, done_mask, want_gray, want_state)
        act = action if (action.is_cuda and action.dtype == torch.int64 and action.is_contiguous()) else \
            action.to(dev, torch.int64, non_blocking=True).contiguous()
        if act.numel() != B or done_mask.numel() != B:
            raise ValueError(f"action/done_mask must have {B} elements")
        if done_mask.is_cuda and done_mask.dtype in (torch.bool, torch.uint8) and done_mask.is_contiguous():
            done_dev = done_mask                                 # updated in place on the device
        else:
            done_dev = done_mask.to(dev, torch.bool, non_blocking=True)
        nxt = torch.empty((B, 3, 16, 20), dtype=torch.float32, device=dev) if want_state else None
        gray = torch.empty((B, 1, 16, 20), dtype=torch.float32, device=dev) if want_gray else None
        reward = torch.empty((B,), dtype=torch.float32, device=dev)
        valid = torch.empty((B, 3), dtype=torch.float32, device=dev)
        _lib.check(L.bk_env_step(B, _ptr(self._hdr), _ptr(self._bricks), _ptr(act), _ptr(done_dev), _ptr(nxt),
                                 _ptr(reward), _ptr(valid), _ptr(gray), self._rewards, _ptr(self._status), st))
        if self.device == "cpu":
            reward, valid = _to_host(reward), _to_host(valid)
            gray = _to_host(gray) if want_gray else None
        out_state = self._hand_out(nxt) if want_state else None
        if done_dev is not done_mask:
            done_mask.copy_(done_dev)                            # in place: callers alias it (train_torch.py:179)
        if self.device == "cpu":
            self.check()                                         # syncs; raises like the reference would
        if want_gray:
            return out_state, reward, done_mask, valid, gray
        return out_state, reward, done_mask, valid

    def _step_host(self, B, action, done_mask, want_gray, want_state):
        """The reference's call (host tensors in and out) as ONE library call: one pinned staging buffer up (done | action | status = 0),
        the step kernel, one packed buffer down (frames | gray | reward | valid | done | action | status), one synchronisation -- instead
        of two uploads, five downloads and a status read.  The outputs are views of a freshly allocated pinned block (torch's caching
        host allocator), so tensors a caller keeps from earlier steps are never overwritten."""
        if action.numel() != B or done_mask.numel() != B:
            raise ValueError(f"action/done_mask must have {B} elements")
        key = (B, bool(want_state), bool(want_gray))
        lay = self._host_io.get(key)
        if lay is None:
            off = (ctypes.c_size_t * 8)()
            total = int(_lib.lib().bk_env_io_layout(B, int(want_state), int(want_gray), off))
            off = [int(o) for o in off]
            assert all(o % 4 == 0 for o in off)
            lay = dict(off=[o // 4 for o in off], words=total // 4, io=torch.empty(total, dtype=torch.uint8, device=self._cuda),
                       hin=torch.zeros(total - off[4], dtype=torch.uint8).pin_memory())
            self._host_io[key] = lay
        act = action if (action.dtype == torch.int64 and action.is_contiguous()) else action.to(torch.int64).contiguous()
        in_place = done_mask.dtype in (torch.bool, torch.uint8) and done_mask.is_contiguous()
        dm = done_mask if in_place else done_mask.to(torch.bool).contiguous()
        off = lay["off"]                                                     # in 4-byte words
        hout = torch.empty(lay["words"], dtype=torch.float32, pin_memory=True)
        rc = _lib.lib().bk_env_step_host(B, self._hdr.data_ptr(), self._bricks.data_ptr(), lay["io"].data_ptr(), lay["hin"].data_ptr(), hout.data_ptr(),
                                         act.data_ptr(), dm.data_ptr(), int(want_state), int(want_gray), self._rewards, self._stream())
        if rc:
            if rc < 0:
                _lib.check(rc)
            raise IndexError(_status_message(rc))
        if not in_place:
            done_mask.copy_(dm)                                              # in place: callers alias it (train_torch.py:179)
        out_state = None
        if want_state:
            out_state = hout[off[0]:off[0] + B * 960].view(B, 3, 16, 20)
            self._last_state, self._last_version = out_state, out_state._version
        else:
            self._last_state = None
        reward = hout[off[2]:off[2] + B]
        valid = hout[off[3]:off[3] + 3 * B].view(B, 3)
        if want_gray:
            return out_state, reward, done_mask, valid, hout[off[1]:off[1] + B * 320].view(B, 1, 16, 20)
        return out_state, reward, done_mask, valid

    # ------------------------------------------------------------------ extras
    @property
    def ball_dx(self) -> torch.Tensor:
        return self._velocity()[0]

    @ball_dx.setter
    def ball_dx(self, v):
        self._set_velocity(dx=v)

    @property
    def ball_dy(self) -> torch.Tensor:
        return self._velocity()[1]

    @ball_dy.setter
    def ball_dy(self, v):
        self._set_velocity(dy=v)

    def check(self) -> None:
        """Raise if any kernel flagged an error since the last check (one device->host read)."""
        if self._status is None:
            return
        bits = int(self._status.item())
        if bits:
            self._status.zero_()
            raise IndexError(_status_message(bits))

    def gray(self, state: torch.Tensor) -> torch.Tensor:
        """convert_to_grayscale (train_torch.py:334-358) on the device."""
        _lib.require_cuda()
        s = state.to(self._cuda, torch.float32).contiguous()
        out = torch.empty((s.shape[0], 1, 16, 20), dtype=torch.float32, device=self._cuda)
        _lib.check(_lib.lib().bk_gray(s.shape[0], _ptr(s), _ptr(out), self._stream()))
        return out.cpu() if self.device == "cpu" else out

    def render(self, state: torch.Tensor) -> str:
        """ASCII debug view of two states side by side (reference :257-293)."""
        assert state.shape[0] == 2
        s = state.cpu()
        glyph = lambda b, y, x: "█" if s[b, 2, y, x] == 1 else ("●" if s[b, 1, y, x] == 1 else ("=" if s[b, 0, y, x] == 1 else " "))
        lines = []
        for y in range(self.height):
            a = "¦" + "".join(glyph(0, y, x) for x in range(self.width)) + "¦"
            b = "¦" + "".join(glyph(1, y, x) for x in range(self.width)) + "¦"
            lines.append(a + "   " + b)
        return "\n".join(lines)

    # ------------------------------------------------------------------ internals
    def _stream(self):
        return torch.cuda.current_stream(self._cuda).cuda_stream

    def _alloc(self) -> int:
        B = int(self.batch)
        if B != self._B:
            dev = self._cuda
            self._hdr = torch.zeros(B, dtype=torch.int64, device=dev)
            self._bricks = torch.zeros((16, B), dtype=torch.int32, device=dev)
            if self._status is None:
                self._status = torch.zeros(1, dtype=torch.int32, device=dev)
            self._B = B
            self._last_state = None
        return B

    def _hand_out(self, dev_state: torch.Tensor) -> torch.Tensor:
        out = _to_host(dev_state) if self.device == "cpu" else dev_state
        self._last_state, self._last_version = out, out._version
        return out

    def _is_ours(self, state: torch.Tensor) -> bool:
        return state is self._last_state and state._version == self._last_version

    def _ingest(self, state: torch.Tensor, dx=None, dy=None) -> None:
        B = self._B
        if tuple(state.shape) != (B, 3, 16, 20):
            raise ValueError(f"state must have shape {(B, 3, 16, 20)}, got {tuple(state.shape)}")
        s = state.to(self._cuda, torch.float32).contiguous()
        _lib.check(_lib.lib().bk_env_ingest(B, _ptr(s), _ptr(dx), _ptr(dy), _ptr(self._hdr), _ptr(self._bricks),
                                            _ptr(self._status), self._stream()))
        self.check()

    def _velocity(self):
        self._alloc()
        dx = torch.empty(self._B, dtype=torch.int64, device=self._cuda)
        dy = torch.empty(self._B, dtype=torch.float32, device=self._cuda)
        _lib.check(_lib.lib().bk_env_velocity(self._B, _ptr(self._hdr), _ptr(dx), _ptr(dy), self._stream()))
        return (dx.cpu(), dy.cpu()) if self.device == "cpu" else (dx, dy)

    def _set_velocity(self, dx=None, dy=None) -> None:
        """Assigning .ball_dx / .ball_dy (the reference keeps them as plain attributes) rewrites the
        velocity bits of the SoA header."""
        self._alloc()
        cur_dx, cur_dy = self._velocity()
        dx = cur_dx if dx is None else torch.as_tensor(dx)
        dy = cur_dy if dy is None else torch.as_tensor(dy)
        dx = dx.to(self._cuda, torch.int64).contiguous()
        dy = dy.to(self._cuda, torch.float32).contiguous()
        state = torch.empty((self._B, 3, 16, 20), dtype=torch.float32, device=self._cuda)
        L = _lib.lib()
        _lib.check(L.bk_env_render(self._B, _ptr(self._hdr), _ptr(self._bricks), _ptr(state), self._stream()))
        _lib.check(L.bk_env_ingest(self._B, _ptr(state), _ptr(dx), _ptr(dy), _ptr(self._hdr), _ptr(self._bricks),
                                   _ptr(self._status), self._stream()))
