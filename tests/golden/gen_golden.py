"""Generate the committed golden vectors by RUNNING THE UNMODIFIED REFERENCE (/root/reference).

Build container only (the reference is not on the GPU box):
    cd /root/repo && python tests/golden/gen_golden.py [env|fuzz|play|mcts|net|replay|train|acting|all]

Outputs (np.savez_compressed, all small):
    env_config1.npz  BASELINE.json configs[0]: B=24, torch.manual_seed(42), 10 000 steps, episode
                     protocol of train_torch.py:184-187, actions from Generator(1234)
    env_fuzz.npz     SURVEY.md Appendix B: random synthetic states (bricks rows 0-4, ball anywhere)
    env_play.npz     ball-following play from reset() (wins, losses, row -1 wrap-around)
    mcts_fake.npz    MCTSSearchVec.search with injected RNG + deterministic fake networks
    mcts_real.npz    same with the real fp32 MuZeroAgent (seed 0, perturbed BN), B=4; doubles as the
                     network golden (rep-net output, per-simulation dynamics/prediction outputs)
    replay.npz       replay_buffer.py ObservationTrajectory + ReplayBuffer fed the way train_torch.py:204-208,
                     223-225,313-332 feeds them (synthetic trajectories of 3..261 moves, FIFO eviction at 150 samples)
    train.npz        train_torch.py loss_fn (:33-66) with utils.py ScalarTransforms.supports_representation under autograd
                     (losses + gradients w.r.t. the three logit tensors), and 5 steps of the optimizer networks.py:268 builds
                     (torch.optim.Adam(lr=config.yaml:28, weight_decay=1e-4)) on a 4099-element parameter
    acting.npz       one episode of the reference's own RLSystem._run_episode (train_torch.py:171-233) on the reference environment with
                     injected search outputs and an injected uniform stream for the Categorical draw: rep-net inputs of every move
                     (_pad_initial_state :313-332, _prepare_mcts_input :259-277, _encode_actions :279-293), sampling probabilities
                     (:192-193), actions, and the ObservationTrajectory lists the loop appended (:204-208)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refshim  # noqa: E402
from common import FakeNet, dirichlet_noise, pack_state, perturb_bn  # noqa: E402

refshim.install()
from environment.parallel_breakout import BreakoutEnvironment  # noqa: E402  (the reference)

CFG = refshim.load_cfg()


def _rec_step(env, state, action, done):
    nxt, reward, done_out, valid = env.step(state, action, done)
    assert done_out is done
    return nxt, dict(state=pack_state(nxt), reward=reward.numpy().copy(), done=done.numpy().astype(np.uint8),
                     valid=valid.numpy().astype(np.uint8), dx=env.ball_dx.numpy().astype(np.int8),
                     dy=env.ball_dy.numpy().astype(np.float32).copy())


def _stack(recs):
    return {k: np.stack([r[k] for r in recs]) for k in recs[0]}


def gen_env_config1(steps=10000, B=24):
    torch.manual_seed(42)
    env = BreakoutEnvironment(CFG["environment"])
    assert env.batch == B
    ga = torch.Generator().manual_seed(1234)
    recs, actions, reset_at, reset_state, reset_dx = [], [], [], [], []
    t = 0
    while t < steps:
        state, _ = env.reset()
        reset_at.append(t); reset_state.append(pack_state(state)); reset_dx.append(env.ball_dx.numpy().astype(np.int8))
        done = torch.zeros(B, dtype=torch.bool)
        length = 0
        while not torch.all(done) and length <= 260 and t < steps:      # train_torch.py:184-187
            a = torch.randint(0, 3, (B,), generator=ga)
            state, rec = _rec_step(env, state, a, done)
            recs.append(rec); actions.append(a.numpy().astype(np.int8))
            length += 1; t += 1
    out = _stack(recs)
    out.update(actions=np.stack(actions), reset_at=np.array(reset_at), reset_state=np.stack(reset_state),
               reset_dx=np.stack(reset_dx))
    np.savez_compressed(os.path.join(HERE, "env_config1.npz"), **out)
    print("env_config1: steps", t, "resets", len(reset_at), "dones", int(out["done"][-1].sum()))


def synth_state(B, g):
    """Appendix B synthetic state: paddle x in [0,14], ball anywhere, brick PAIRS with density
    U(0,0.3) in rows 0-4, dx in {-1,+1}, dy in {-1.0,+1.0}."""
    s = torch.zeros(B, 3, 16, 20)
    px = torch.randint(0, 15, (B,), generator=g)
    for b in range(B):
        s[b, 0, 15, px[b]:px[b] + 6] = 1
    by = torch.randint(0, 16, (B,), generator=g)
    bx = torch.randint(0, 20, (B,), generator=g)
    s[torch.arange(B), 1, by, bx] = 1
    dens = torch.rand(B, generator=g) * 0.3
    pairs = (torch.rand(B, 5, 10, generator=g) < dens[:, None, None]).float()
    s[:, 2, :5, :] = pairs.repeat_interleave(2, dim=2)
    dx = torch.randint(0, 2, (B,), generator=g) * 2 - 1
    dy = (torch.randint(0, 2, (B,), generator=g) * 2 - 1).float()
    return s, dx, dy


def gen_env_fuzz(seeds=60, steps=60, B=24):
    env = BreakoutEnvironment(CFG["environment"])
    init_state, init_dx, init_dy, all_actions, outs = [], [], [], [], []
    for seed in range(seeds):
        g = torch.Generator().manual_seed(1000 + seed)
        state, dx, dy = synth_state(B, g)
        env.ball_dx, env.ball_dy = dx.clone(), dy.clone()
        init_state.append(pack_state(state)); init_dx.append(dx.numpy().astype(np.int8)); init_dy.append(dy.numpy().copy())
        done = torch.zeros(B, dtype=torch.bool)
        recs, acts = [], []
        for _ in range(steps):
            a = torch.randint(0, 3, (B,), generator=g)
            state, rec = _rec_step(env, state, a, done)
            recs.append(rec); acts.append(a.numpy().astype(np.int8))
        outs.append(_stack(recs)); all_actions.append(np.stack(acts))
    out = {k: np.stack([o[k] for o in outs]) for k in outs[0]}
    out.update(init_state=np.stack(init_state), init_dx=np.stack(init_dx), init_dy=np.stack(init_dy),
               actions=np.stack(all_actions))
    np.savez_compressed(os.path.join(HERE, "env_fuzz.npz"), **out)
    print("env_fuzz: env-steps", seeds * steps * B, "reward sum", float(out["reward"].sum()))


def gen_env_play(seeds=4, steps=1000, B=24):
    """Real play from reset() with a 90 %-ball-following policy (random play never wins)."""
    env = BreakoutEnvironment(CFG["environment"])
    outs, all_actions, reset_state, reset_dx = [], [], [], []
    wins = losses = wraps = 0
    for seed in range(seeds):
        torch.manual_seed(2000 + seed)
        g = torch.Generator().manual_seed(3000 + seed)
        state, _ = env.reset()
        reset_state.append(pack_state(state)); reset_dx.append(env.ball_dx.numpy().astype(np.int8))
        done = torch.zeros(B, dtype=torch.bool)
        recs, acts = [], []
        for _ in range(steps):
            ball = torch.where(state[:, 1] == 1)
            bx = ball[2]
            px = torch.argmax(state[:, 0, -1, :], dim=1) + 3
            follow = torch.where(bx + env.ball_dx < px, 0, torch.where(bx + env.ball_dx > px, 2, 1))
            rnd = torch.randint(0, 3, (B,), generator=g)
            a = torch.where(torch.rand(B, generator=g) < 0.9, follow, rnd)
            prev_done = done.clone()
            state, rec = _rec_step(env, state, a, done)
            newly = done & ~prev_done
            wins += int((newly & (rec["reward"] >= 5)).sum()); losses += int((newly & (rec["reward"] == -1)).sum())
            wraps += int(((state[:, 1, 15].sum(1) == 1) & (state[:, 2].sum((1, 2)) > 0) & (torch.tensor(rec["dy"]) == -1)).sum())
            recs.append(rec); acts.append(a.numpy().astype(np.int8))
        outs.append(_stack(recs)); all_actions.append(np.stack(acts))
    out = {k: np.stack([o[k] for o in outs]) for k in outs[0]}
    out.update(reset_state=np.stack(reset_state), reset_dx=np.stack(reset_dx), actions=np.stack(all_actions))
    np.savez_compressed(os.path.join(HERE, "env_play.npz"), **out)
    print("env_play: env-steps", seeds * steps * B, "wins", wins, "losses", losses, "row-15 balls moving up", wraps)


# ----------------------------------------------------------------------------------------------- MCTS

def _slots_from_trace(trace, B):
    """Reference node names -> slot numbers in order of first expansion (root 'state_0' = 0; the
    re-expanded sim-0 leaf keeps its name, hence its slot)."""
    names = [{"state_0": 0} for _ in range(B)]
    parent = np.zeros((len(trace), B), np.int32); action = np.zeros_like(parent); leaf = np.zeros_like(parent)
    for s, t in enumerate(trace):
        for b, (prev, a, node) in enumerate(t["last_nodes"]):
            if node not in names[b]:
                names[b][node] = len(names[b])
            parent[s, b], action[s, b], leaf[s, b] = names[b][prev], a, names[b][node]
    return parent, action, leaf


def _run_ref_search(mu_zero, transforms, hidden, seed, noise, S, noise_weight=0.175, c1=None, c2=None):
    from src.mcts import MCTSSearchVec

    cfg = dict(CFG); cfg["num_simulations"] = S
    cfg["search"] = dict(CFG["search"])
    if c1 is not None: cfg["search"]["c1"] = c1
    if c2 is not None: cfg["search"]["c2"] = c2
    m = MCTSSearchVec(cfg, mu_zero, transforms)
    m.noise_weight = noise_weight
    trace = []
    B = hidden.shape[0]
    value, visits, proxy = refshim.injected_search(m, hidden, torch.ones(B, 3), seed, noise, trace)
    parent, action, leaf = _slots_from_trace(trace, B)
    return dict(value=value.numpy().astype(np.float32), visits=visits.numpy().astype(np.int64),
                parent=parent, action=action, leaf=leaf,
                reward=np.stack([t["rewards"].numpy() for t in trace]).astype(np.float32),
                leaf_value=np.stack([t["values"].numpy() for t in trace]).astype(np.float32),
                pi=np.stack([t["policies"].numpy() for t in trace]).astype(np.float32),
                ucb_calls=np.array([proxy.ctr.get(b, 0) for b in range(B)], np.int32))


def gen_mcts_fake():
    from utils import ScalarTransforms

    tr = ScalarTransforms(CFG["model"])
    out = {}
    cases = [("varied", 6, 50, 7, 0.175, None, None), ("deep", 6, 50, 8, 0.175, None, None),
             ("flat", 6, 50, 9, 0.175, None, None), ("varied", 5, 20, 10, 0.1, 2.0, 100.0),
             ("deep", 3, 80, 11, 0.25, None, None), ("mild", 8, 50, 12, 0.175, None, None),
             ("mild", 4, 50, 13, 0.1, None, None)]
    for i, (mode, B, S, seed, w, c1, c2) in enumerate(cases):
        net = FakeNet(mode)
        g = torch.Generator().manual_seed(seed)
        hidden = torch.rand(B, 8, generator=g)
        noise = dirichlet_noise(B, seed)
        with torch.no_grad():
            pol, val = net.evaluate_state(hidden)
        rec = _run_ref_search(net, tr, hidden, seed, noise, S, w, c1, c2)
        rec.update(hidden=hidden.numpy(), noise=noise.numpy(), v_root=tr.inverted_softmax_expectation(val).numpy(),
                   pi_root=torch.softmax(pol, dim=1).numpy(),
                   meta=np.array([B, S, seed, w, c1 or CFG["search"]["c1"], c2 or CFG["search"]["c2"]], np.float64))
        for k, v in rec.items():
            out[f"c{i}_{k}"] = v
        print(f"mcts_fake case {i} {mode}: visits[0]={rec['visits'][0]} value[0]={rec['value'][0]:.6f} "
              f"max depth parent slot={rec['parent'].max()} ucb_calls={rec['ucb_calls'].tolist()}")
    out["modes"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "mcts_fake.npz"), **out)


def gen_mcts_real(B=4, S=50, seed=7):
    from src.networks import MuZeroAgent
    from utils import ScalarTransforms

    torch.manual_seed(0)
    agent = MuZeroAgent(CFG["model"])
    perturb_bn(agent, 1)
    agent.eval_mode()
    tr = ScalarTransforms(CFG["model"])
    g = torch.Generator().manual_seed(5)
    rep_in = torch.rand(B, 64, 16, 20, generator=g)
    with torch.no_grad():
        rep_raw = agent.rep_net(rep_in)
        hidden = agent.create_hidden_state_root(rep_in)
        pol, val = agent.evaluate_state(hidden)
        # one explicit dynamics + prediction call for the pure network golden
        planes = torch.zeros(B, 3, 4, 5); planes[torch.arange(B), torch.tensor([0, 1, 2, 1][:B])] = 1
        h2, rew = agent.hidden_state_transition(hidden, planes)
        dyn_raw, _ = agent.dyn_net(torch.cat([hidden, planes], 1))
        pol2, val2 = agent.evaluate_state(h2)
    noise = dirichlet_noise(B, seed)
    rec = _run_ref_search(agent, tr, hidden, seed, noise, S)
    rec.update(rep_in=rep_in.numpy(), rep_raw=rep_raw.numpy(), hidden=hidden.numpy(), noise=noise.numpy(),
               root_policy_logits=pol.numpy(), root_value_logits=val.numpy(),
               v_root=tr.inverted_softmax_expectation(val).numpy(), pi_root=torch.softmax(pol, 1).numpy(),
               dyn_actions=np.array([0, 1, 2, 1][:B]), dyn_raw=dyn_raw.numpy(), dyn_h=h2.numpy(), dyn_reward_logits=rew.numpy(),
               pred_policy_logits=pol2.numpy(), pred_value_logits=val2.numpy(),
               meta=np.array([B, S, seed], np.float64))
    np.savez_compressed(os.path.join(HERE, "mcts_real.npz"), **rec)
    print("mcts_real: visits", rec["visits"].tolist(), "value", rec["value"].tolist())


def synth_trajectories(lengths, seed=11):
    """Per trajectory: initial gray frame (1,16,20) and per move (action i64, gray frame, reward f32 from the env's
    reward set, visit counts i64 summing to 50, value f32) -- the tuple train_torch.py:205-207 appends."""
    g = torch.Generator().manual_seed(seed)
    levels = torch.tensor([0.0, 0.3, 0.6, 1.0])
    rset = torch.tensor([0.0, 0.0, 0.0, 1.0, 1.0, -1.0, 5.0, 6.0, 4.0])
    out = []
    for T in lengths:
        def frame():
            f = torch.zeros(320)
            idx = torch.randint(0, 320, (12,), generator=g)
            f[idx] = levels[torch.randint(1, 4, (12,), generator=g)]
            return f.view(1, 16, 20)
        init = frame()
        frames = torch.stack([frame() for _ in range(T)]) if T else torch.zeros(0, 1, 16, 20)
        action = torch.randint(0, 3, (T,), generator=g)
        reward = rset[torch.randint(0, len(rset), (T,), generator=g)]
        v0 = torch.randint(0, 51, (T,), generator=g)
        v1 = (torch.rand(T, generator=g) * (51 - v0)).long()
        visits = torch.stack([v0, v1, 50 - v0 - v1], 1)
        value = (torch.rand(T, generator=g) * 8 - 2).float()
        out.append(dict(init=init, frames=frames, action=action, reward=reward, visits=visits, value=value))
    return out


def gen_replay(cap=150, K=5, hist=32, discount=0.985, n_sum=24):
    from replay_buffer import ObservationTrajectory, ReplayBuffer   # the reference

    lengths = [3, 5, 6, 7, 8, 14, 15, 16, 17, 40, 100, 261, 9, 33]
    trajs = synth_trajectories(lengths)
    rb = ReplayBuffer(hist, K, cap, discount, n_sum)
    snaps, snap_after, snap_out = [], [9, len(lengths) - 1], {}

    def snapshot(tag):
        n = rb.length
        idx = torch.arange(n)
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(3))[:40]
        o = {"perm": perm.numpy()}
        for name, fn in (("past_actions", rb.get_batched_past_actions), ("future_actions", rb.get_batched_future_actions),
                         ("states", rb.get_batched_states), ("rewards", rb.get_batched_rewards),
                         ("visit_counts", rb.get_batched_visit_counts), ("values", rb.get_batched_values)):
            full = fn(idx)
            o[name] = full.numpy()
            o[name + "_dtype"] = np.array(str(full.dtype))
            assert torch.equal(fn(perm), full[perm])
        o["value_buffer"] = torch.stack(rb.value_buffer).numpy()
        o["reward_sums"] = np.array(rb.get_reward_sums(), np.float64)
        o["reward_sums_all"] = np.array(rb.reward_sums, np.float64)
        snap_out.update({f"s{tag}_{k}": v for k, v in o.items()})

    for ti, tr in enumerate(trajs):
        ot = ObservationTrajectory(                                   # train_torch.py:313-332 _pad_initial_state
            actions=[0 for _ in range(hist)], states=[tr["init"] for _ in range(hist - 1)],
            rewards=[0 for _ in range(hist)], visit_counts=[torch.zeros(3) for _ in range(hist)],
            values=[0.0 for _ in range(hist)], length=0, reward_sum=0)
        for t in range(len(tr["action"])):                            # train_torch.py:204-208
            ot.add_observation(tr["action"][t], tr["frames"][t], tr["reward"][t], tr["visits"][t], tr["value"][t])
        rb.save_observation_trajectory(ot)                            # every length, not only > K+1 (:224 is the caller's filter)
        snaps.append(rb.length)
        if ti in snap_after:
            snapshot(snap_after.index(ti))
    out = dict(meta=np.array([cap, K, hist, n_sum], np.int64), discount=np.array(discount), lengths=np.array(lengths),
               length_after=np.array(snaps), snap_after=np.array(snap_after))
    for i, tr in enumerate(trajs):
        for k, v in tr.items():
            out[f"t{i}_{k}"] = v.numpy()
    out.update(snap_out)
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)
    print("replay: lengths after each save", snaps, {k: (out["s1_" + k].shape, str(out["s1_" + k + "_dtype"])) for k in
          ("past_actions", "future_actions", "states", "rewards", "visit_counts", "values")})


def gen_train(B=64, K=5, n_params=4099, steps=5):
    """loss_fn + backward, and Adam, from the reference / torch themselves."""
    import train_torch as ref_train                      # the reference (import has no side effects beyond set_seed(42))
    from utils import ScalarTransforms

    tr = ScalarTransforms(CFG["model"])
    g = torch.Generator().manual_seed(5)
    out = {"supports": tr.supports.numpy().copy(), "K": np.array(K)}
    for case, scale in (("a", 1.0), ("b", 4.0)):
        pr = (torch.randn(B, K, 11, generator=g) * scale).requires_grad_()
        pv = (torch.randn(B, K, 11, generator=g) * scale).requires_grad_()
        pp = (torch.randn(B, K, 3, generator=g) * scale).requires_grad_()
        # rewards as Breakout gives them (0, +1, -1, +5, +6), value targets over the range the supports cover, a few exactly on a support
        obs_r = torch.tensor([0.0, 1.0, -1.0, 5.0, 6.0])[torch.randint(0, 5, (B, K), generator=g)]
        val = (torch.rand(B, K, generator=g) - 0.5) * (60.0 if case == "b" else 8.0)
        val[0, :3] = torch.tensor([0.0, 3.0, -3.0])
        vis = torch.multinomial(torch.rand(B * K, 3, generator=g) ** 2 + 1e-3, 50, replacement=True, generator=g)
        vis = torch.stack([(vis == a).sum(1) for a in range(3)], 1).reshape(B, K, 3).to(torch.float32)
        vis[1, 0] = torch.tensor([50.0, 0.0, 0.0])
        loss, rl, vl, pl = ref_train.loss_fn(observed_reward=obs_r, predicted_reward=pr, bootstrapped_reward=val, predicted_value=pv,
                                             visit_counts=vis, predicted_policy=pp, target_transformation=tr.supports_representation, K=K)
        loss.backward()
        out.update({f"{case}_pred_reward": pr.detach().numpy(), f"{case}_pred_value": pv.detach().numpy(), f"{case}_pred_policy": pp.detach().numpy(),
                    f"{case}_obs_reward": obs_r.numpy(), f"{case}_value_target": val.numpy(), f"{case}_visits": vis.numpy(),
                    f"{case}_losses": np.array([loss.item(), rl.item(), vl.item(), pl.item()], np.float32),
                    f"{case}_target_reward": tr.supports_representation(obs_r).numpy(), f"{case}_target_value": tr.supports_representation(val).numpy(),
                    f"{case}_d_reward": pr.grad.numpy(), f"{case}_d_value": pv.grad.numpy(), f"{case}_d_policy": pp.grad.numpy()})
        print("train loss case", case, [float(x) for x in out[f"{case}_losses"]])
    p = torch.nn.Parameter(torch.randn(n_params, generator=g))
    opt = torch.optim.Adam([p], lr=CFG["model"]["learning_rate"], weight_decay=0.0001)     # networks.py:268
    out["adam_p0"] = p.detach().numpy().copy()
    out["adam_lr"] = np.array(CFG["model"]["learning_rate"])
    for s in range(steps):
        p.grad = torch.randn(n_params, generator=g) * (0.1 if s % 2 else 0.003)
        out[f"adam_g{s}"] = p.grad.numpy().copy()
        opt.step()
        st = opt.state[p]
        out[f"adam_p{s + 1}"], out[f"adam_m{s + 1}"], out[f"adam_v{s + 1}"] = p.detach().numpy().copy(), st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **out)
    print("train: adam steps", steps, "params", n_params)


def gen_acting(B=4, seed=3, env_seed=21, temperature=0.5, sims=50):
    """The reference's own acting loop with everything that is not the acting glue injected."""
    import types

    import train_torch as ref_train
    from refshim import rng_u32

    M64 = (1 << 64) - 1
    ep_seed = (seed * 0x9E3779B97F4A7C15 + 0 * 0xC2B2AE3D27D4EB4F + 7) & M64       # acting.Actor's stream key for episode 0 of Actor(seed=seed)
    g = torch.Generator().manual_seed(99)
    rec = dict(rep_inputs=[], probs=[], u=[], visits=[], value=[])
    state_ = dict(move=0, env=0)

    class InjectedCategorical:                      # train_torch.py:196-197: Categorical(probs[i]).sample(), one per env in env order
        def __init__(self, probs):
            self.p = probs

        def sample(self):
            i, mv = state_["env"], state_["move"] - 1
            state_["env"] += 1
            u = np.float32((rng_u32(ep_seed, i, mv) >> 8) / 16777216.0)
            p = self.p.numpy()
            rec["probs"].append(p.copy()); rec["u"].append(u)
            a = 0 if u < p[0] else (1 if u < np.float32(p[0] + p[1]) else 2)
            if p[a] == 0:
                a = 2 if p[2] > 0 else (1 if p[1] > 0 else 0)
            return torch.tensor(a)

    class TorchProxy:                               # the module-global `torch` of train_torch with distributions.Categorical replaced
        distributions = types.SimpleNamespace(Categorical=InjectedCategorical)

        def __getattr__(self, name):
            return getattr(torch, name)

    R = ref_train.RLSystem
    sysm = object.__new__(R)
    sysm.n_parallel, sysm.state_history_length, sysm.actions, sysm.n_actions = B, 32, [0, 1, 2], 3
    sysm.real_resolution, sysm.K, sysm.temperature, sysm.training_iteration = (16, 20), 5, temperature, 0
    sysm.action_stats, sysm.acting_step = torch.tensor([]), 0
    sysm.environment = BreakoutEnvironment(dict(CFG["environment"], n_parallel=B))
    saved = []
    sysm.replay_buffer = types.SimpleNamespace(save_observation_trajectory=saved.append, get_reward_sums=lambda: [0.0])
    sysm.filewriter = types.SimpleNamespace(add_scalar=lambda *a, **k: None)

    def sample_action(self, state, mask):           # _sample_action (:236-257) with the networks + search replaced by preset outputs
        ins = torch.stack([self._prepare_mcts_input(state[i], self.observation_trajectories[i], self.real_resolution) for i in range(B)])
        rec["rep_inputs"].append(ins.numpy().copy())
        c = torch.sort(torch.randint(0, sims + 1, (B, 2), generator=g), dim=1)[0]
        visits = torch.stack([c[:, 0], c[:, 1] - c[:, 0], sims - c[:, 1]], 1).to(torch.int64)
        value = (torch.rand(B, generator=g) * 6 - 1).float()
        rec["visits"].append(visits.numpy().copy()); rec["value"].append(value.numpy().copy())
        state_["move"] += 1; state_["env"] = 0
        return value, visits

    sysm._sample_action = types.MethodType(sample_action, sysm)
    torch.manual_seed(env_seed)
    initial_state, _ = sysm.environment.reset()
    init_np, dx0 = initial_state.numpy().copy(), sysm.environment.ball_dx.numpy().astype(np.int8).copy()
    sysm._pad_initial_state(sysm.convert_to_grayscale(initial_state))
    real_torch = ref_train.torch
    ref_train.torch = TorchProxy()
    try:
        sysm._run_episode(initial_state)
    finally:
        ref_train.torch = real_torch
    T = state_["move"]
    trajs = sysm.observation_trajectories
    out = dict(meta=np.array([B, T, seed, env_seed, sims], np.int64), temperature=np.array(temperature), ep_seed=np.array(ep_seed, np.uint64),
               initial_state=init_np, initial_dx=dx0, rep_inputs=np.stack(rec["rep_inputs"]),
               probs=np.stack(rec["probs"]).reshape(T, B, 3), u=np.array(rec["u"], np.float32).reshape(T, B),
               visits=np.stack(rec["visits"]), value=np.stack(rec["value"]), lengths=np.array([t.length for t in trajs], np.int64),
               saved=np.array([any(t is s_ for s_ in saved) for t in trajs]))
    for i, t in enumerate(trajs):
        out[f"t{i}_actions"] = np.array([int(a) for a in t.actions], np.int64)
        out[f"t{i}_states"] = torch.stack(t.states).numpy()
        out[f"t{i}_rewards"] = np.array([float(r) for r in t.rewards], np.float32)
        out[f"t{i}_visits"] = torch.stack([v.float() for v in t.visit_counts]).numpy()
        out[f"t{i}_values"] = np.array([float(v) for v in t.values], np.float32)
    # the sampling expression alone (train_torch.py:192-193) on many visit-count rows and three temperatures
    c = torch.sort(torch.randint(0, sims + 1, (3000, 2), generator=g), dim=1)[0]
    vis = torch.stack([c[:, 0], c[:, 1] - c[:, 0], sims - c[:, 1]], 1).to(torch.int64)
    out["s_visits"] = vis.numpy()
    for k, temp in enumerate((1.0, 0.5, 0.25)):
        w = vis ** (1 / temp)
        out[f"s_probs{k}"] = (w / w.sum(dim=1, keepdim=True)).numpy()
    out["s_temps"] = np.array([1.0, 0.5, 0.25])
    np.savez_compressed(os.path.join(HERE, "acting.npz"), **out)
    print("acting: moves", T, "lengths", out["lengths"].tolist(), "saved", out["saved"].tolist(), "bytes", os.path.getsize(os.path.join(HERE, "acting.npz")))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.chdir("/tmp")
    if what in ("env", "all"): gen_env_config1()
    if what in ("fuzz", "all"): gen_env_fuzz()
    if what in ("play", "all"): gen_env_play()
    if what in ("mcts", "all"): gen_mcts_fake()
    if what in ("net", "all"): gen_mcts_real()
    if what in ("replay", "all"): gen_replay()
    if what in ("train", "all"): gen_train()
    if what in ("acting", "all"): gen_acting()
