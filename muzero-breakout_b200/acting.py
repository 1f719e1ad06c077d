"""One acting episode entirely on the device: the loop of the reference's RLSystem._acting_stage /
_run_episode / _sample_action (train_torch.py:160-257) with every per-environment Python loop replaced by a
kernel -- the "next" rows 1-3 of SURVEY.md section 8f.

    per move:  observation-history ring -> representation-network input   (mz_rep_input;   :249-253,259-293)
               representation network -> root latents                      (PackedNetworks; :254)
               MCTSSearchVec.search                                        (:256)
               temperature sampling of the action from the visit counts    (mz_sample_actions; :192-198)
               BreakoutEnvironment.step + fused grayscale into the ring    (bk_env_step;   :201-203)
               trajectory record (action, gray frame, reward, visits, value) for envs not yet done (:204-208)

History semantics are the reference's: the trajectory starts with 31 copies of the initial gray frame and 32
zero actions (_pad_initial_state :313-332); the rep-net input is the last 31 appended frames + the current
frame (which from the second move on is also the newest appended one, so it appears twice) + the last 32
actions / 3 as planes.  Environments that are done keep being stepped (as in the reference) but are not
recorded.  Randomness: reset via the env's reset_rng, Dirichlet noise and pUCT tie-breaks via the search,
action sampling via the counter-based stream u32(seed, env, move).
"""
from __future__ import annotations

import torch

from . import _lib
from .src.networks import BF16, F32


def _p(t):
    return None if t is None else t.data_ptr()


class Actor:
    SLOTS = 32

    def __init__(self, env, mcts, temperature: float = 1.0, seed: int = 0, max_moves: int = 261, check_done_every: int = 1,
                 record_frames: bool = True):
        if env.device != "cuda":
            raise ValueError("Actor needs a device-resident environment (cfg['output_device'] = 'cuda')")
        self.env, self.mcts = env, mcts
        self.temperature, self.seed = float(temperature), int(seed)
        self.max_moves, self.check_done_every, self.record_frames = int(max_moves), int(check_done_every), record_frames
        self.keep_rep_inputs = None          # set to a list to collect the rep-net inputs (NCHW fp32) for tests
        self._B = 0
        self._nets = None
        self._episodes = 0

    def _alloc(self, B):
        nets = self.mcts.packed_networks()
        dev = nets.device
        if B == self._B and self._nets is nets:
            return
        self._B, self._nets = B, nets
        self.frames = torch.empty((self.SLOTS, B, 320), dtype=torch.float32, device=dev)
        self.acts = torch.zeros((self.SLOTS, B), dtype=torch.int32, device=dev)
        self.x_cl = torch.empty((B, 320, 64), dtype=nets.dtype, device=dev)
        self.hidden = torch.empty((B, nets.latent_ch, *nets.latent_hw), dtype=torch.float32, device=dev)
        self.rep_prog = nets.representation_program(B, None, self.hidden, x_cl=self.x_cl)
        self.action = torch.zeros(B, dtype=torch.int64, device=dev)

    def run_episode(self):
        """-> dict of device tensors, T = number of moves played:
        action (T,B) i64, reward (T,B) f32, value (T,B) f32, visits (T,B,3) i64, recorded (T,B) bool
        (= not done before the move, train_torch.py:205), done (B,) bool, frames (T,B,1,16,20) f32 (optional),
        initial_gray (B,1,16,20).  replay_buffer.ReplayBuffer.save_episode(record) stores it without a host round trip."""
        _lib.require_cuda()
        env, L = self.env, _lib.lib()
        B = int(env.batch)
        self._alloc(B)
        nets = self._nets
        st = torch.cuda.current_stream(nets.device).cuda_stream
        state, _ = env.reset()
        gray0 = env.gray(state).view(B, 320)
        self.frames.copy_(gray0.unsqueeze(0).expand(self.SLOTS, B, 320))        # 31 copies of the initial frame (:313-332)
        self.acts.zero_()                                                       # 32 zero actions
        head = ahead = self.SLOTS - 1
        done = torch.zeros(B, dtype=torch.bool, device=nets.device)
        cur = gray0
        rec = {k: [] for k in ("action", "reward", "value", "visits", "recorded", "frames")}
        initial_state, initial_dx = state.clone(), env.ball_dx
        dt = nets.dt
        ep_seed = (self.seed * 0x9E3779B97F4A7C15 + self._episodes * 0xC2B2AE3D27D4EB4F + 7) & 0xFFFFFFFFFFFFFFFF
        self._episodes += 1
        for move in range(self.max_moves):
            if move % self.check_done_every == 0 and bool(done.all()):          # while not all(done) (:184)
                break
            _lib.check(L.mz_rep_input(B, self.SLOTS, _p(self.frames), head, _p(cur), _p(self.acts), ahead, _p(self.x_cl), dt, st))
            if self.keep_rep_inputs is not None:
                self.keep_rep_inputs.append(self.x_cl.float().view(B, 16, 20, 64).permute(0, 3, 1, 2).contiguous())
            self.rep_prog.run()
            value, visits = self.mcts.search(self.hidden, None, 0)
            value, visits = value.to(nets.device), visits.to(nets.device)
            ahead = (ahead + 1) % self.SLOTS
            action = torch.empty(B, dtype=torch.int64, device=nets.device)
            _lib.check(L.mz_sample_actions(B, _p(visits), self.temperature, ep_seed, move, _p(action), _p(self.acts[ahead]), None, st))
            prev_done = done.clone()
            head = (head + 1) % self.SLOTS
            gray_slot = self.frames[head].view(B, 1, 16, 20)
            _, reward, done, _valid = self._step_into(action, done, gray_slot)
            cur = self.frames[head]
            rec["action"].append(action); rec["reward"].append(reward); rec["value"].append(value); rec["visits"].append(visits)
            rec["recorded"].append(~prev_done)
            if self.record_frames:
                rec["frames"].append(cur.clone().view(B, 1, 16, 20))
        out = {k: torch.stack(v) for k, v in rec.items() if v}
        out["done"] = done
        out["initial_state"], out["initial_dx"] = initial_state, initial_dx
        out["initial_gray"] = gray0.clone().view(B, 1, 16, 20)                    # the 31 padding frames (:313-332), for ReplayBuffer.save_episode
        return out

    def _step_into(self, action, done, gray_out):
        """env.step with the fused grayscale written straight into a history-ring slot."""
        env, L = self.env, _lib.lib()
        B = env._alloc()
        dev = env._cuda
        reward = torch.empty(B, dtype=torch.float32, device=dev)
        valid = torch.empty((B, 3), dtype=torch.float32, device=dev)
        _lib.check(L.bk_env_step(B, _p(env._hdr), _p(env._bricks), _p(action), _p(done), None, _p(reward), _p(valid), _p(gray_out),
                                 env._rewards, _p(env._status), torch.cuda.current_stream(dev).cuda_stream))
        env._last_state = None
        return None, reward, done, valid
