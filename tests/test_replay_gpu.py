"""GPU parity tests of the device replay buffer (muzero-breakout_b200/replay_buffer.py -> rb_append / rb_gather,
csrc/replay.cu) against the reference's golden vectors (tests/golden/replay.npz) and the numpy oracle
(oracle/replay_oracle.py).  Everything bit-exact: integer/index work, frame copies, and the fp32 value targets whose
rounding sequence is the reference's (replay_buffer.py:136-152)."""
import numpy as np
import pytest
import torch

from oracle.replay_oracle import ReplayOracle
from test_oracle_replay import FIELDS, load_replay_golden

pytestmark = pytest.mark.gpu

GETTERS = dict(past_actions="get_batched_past_actions", future_actions="get_batched_future_actions", states="get_batched_states",
               rewards="get_batched_rewards", visit_counts="get_batched_visit_counts", values="get_batched_values")


def _ref_trajectory(tr, hist=32, pad_action=0):
    """An ObservationTrajectory built the way train_torch.py:313-332 + :204-208 builds it (CPU tensors)."""
    from muzero_breakout_b200.replay_buffer import ObservationTrajectory
    init = torch.from_numpy(tr["init"])
    ot = ObservationTrajectory(actions=[pad_action for _ in range(hist)], states=[init for _ in range(hist - 1)],
                               rewards=[0 for _ in range(hist)], visit_counts=[torch.zeros(3) for _ in range(hist)],
                               values=[0.0 for _ in range(hist)], length=0, reward_sum=0)
    a, f, r, n, v = (torch.from_numpy(tr[k]) for k in ("action", "frames", "reward", "visits", "value"))
    for t in range(len(a)):
        ot.add_observation(a[t], f[t], r[t], n[t], v[t])
    return ot


def _check_against(rb, want_of, n, tag=""):
    idx = torch.arange(n)
    for f, getter in GETTERS.items():
        got = getattr(rb, getter)(idx)
        want = want_of(f)
        assert got.is_cuda and tuple(got.shape) == want.shape and str(got.dtype).replace("torch.", "") == str(want.dtype), (tag, f)
        assert np.array_equal(got.cpu().numpy(), want), (tag, f)


def test_save_observation_trajectory_matches_reference_goldens(golden_dir):
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    g, p, trajs = load_replay_golden(golden_dir)
    rb = ReplayBuffer(p["hist"], p["K"], p["cap"], p["discount"], p["n_sum"])
    snap_after = list(g["snap_after"])
    for ti, tr in enumerate(trajs):
        rb.save_observation_trajectory(_ref_trajectory(tr))
        assert rb.length == len(rb) == int(g["length_after"][ti])
        if ti in snap_after:
            tag = f"s{snap_after.index(ti)}_"
            n = rb.length
            _check_against(rb, lambda f: g[tag + f], n, tag)
            perm = torch.from_numpy(g[tag + "perm"])
            mb = rb.minibatch(perm)                                             # one launch, train_torch.py:455-484 order
            for got, f in zip(mb, ("past_actions", "states", "visit_counts", "future_actions", "rewards", "values")):
                assert np.array_equal(got.cpu().numpy(), g[tag + f][perm.numpy()]), f
            # rep-net input of the training step from one launch: cat(states, _encode_actions(past_actions)) of train_torch.py:392,500
            # with :291-293 (actions / n_actions broadcast over the 16 x 20 plane), from the reference's own getter outputs
            st, pa = torch.from_numpy(g[tag + "states"][perm.numpy()]), torch.from_numpy(g[tag + "past_actions"][perm.numpy()])
            planes = torch.ones((len(perm), p["hist"], 16, 20)) * (pa / 3)[:, :, None, None].expand(-1, -1, 16, 20)
            want_in = torch.cat((st.view(len(perm), -1, 16, 20), planes), dim=1)
            got_in = rb.repnet_input(perm, n_actions=3)
            assert got_in.is_cuda and got_in.shape == want_in.shape and torch.equal(got_in.cpu(), want_in)
            assert rb.get_reward_sums() == list(g[tag + "reward_sums"])
            assert rb.reward_sums == list(g[tag + "reward_sums_all"])
            assert np.array_equal(torch.stack(rb.value_buffer).numpy(), g[tag + "value_buffer"])
            assert np.array_equal(torch.stack(rb.bootstrapped_values).numpy(), g[tag + "values"])
    # Python list indexing: negative indices count from the newest sample; out of range raises like the reference
    last = rb.get_batched_rewards(torch.tensor([-1, 0]))
    assert np.array_equal(last.cpu().numpy(), g["s1_rewards"][[-1, 0]])
    with pytest.raises(IndexError):                              # host indices: checked before the launch
        rb.get_batched_rewards(torch.tensor([rb.length]))
    rb.get_batched_rewards(torch.tensor([rb.length]).cuda())     # device indices: the kernel flags it, the next host-facing call raises
    with pytest.raises(IndexError):
        rb.get_reward_sums()
    rb.empty_buffer()
    assert rb.length == 0 and rb.get_reward_sums() == []


def test_nonzero_padding_action_and_cpu_outputs(golden_dir):
    """run_test_simulation pads with action 1 (train_torch.py:547); output_device='cpu' returns host tensors."""
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    g, p, trajs = load_replay_golden(golden_dir)
    rb = ReplayBuffer(p["hist"], p["K"], 500, p["discount"], p["n_sum"], output_device="cpu")
    rb.save_observation_trajectory(_ref_trajectory(trajs[9], pad_action=1))
    pa = rb.get_batched_past_actions(torch.arange(rb.length))
    assert not pa.is_cuda and pa.dtype == torch.int64
    A = np.concatenate([np.ones(32, np.int64), trajs[9]["action"]])
    assert np.array_equal(pa.numpy(), np.stack([A[s:s + 32] for s in range(rb.length)]))
    with pytest.raises(IndexError):
        rb.get_batched_states(torch.tensor([rb.length]))
    bad = _ref_trajectory(trajs[9]); bad.actions[3] = 2
    with pytest.raises(ValueError):
        rb.save_observation_trajectory(bad)


def _episode_record(B, T, seed, min_len=0, device="cuda"):
    """A synthetic (T,B) episode record in acting.Actor.run_episode's layout with random per-env lengths."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, T + 1, (B,), generator=g)
    lens[0] = T
    recorded = torch.arange(T)[:, None] < lens[None, :]
    levels = torch.tensor([0.0, 0.3, 0.6, 1.0])
    rset = torch.tensor([0.0, 0.0, 1.0, -1.0, 5.0, 6.0, 4.0])
    rec = dict(action=torch.randint(0, 3, (T, B), generator=g), reward=rset[torch.randint(0, 7, (T, B), generator=g)],
               value=(torch.rand(T, B, generator=g) * 8 - 2), visits=torch.randint(0, 51, (T, B, 3), generator=g),
               frames=levels[torch.randint(0, 4, (T, B, 1, 16, 20), generator=g)], recorded=recorded,
               initial_gray=levels[torch.randint(0, 4, (B, 1, 16, 20), generator=g)])
    return {k: v.to(device) for k, v in rec.items()}, lens


def _feed_oracle(orc, rec, lens, K):
    c = {k: v.cpu().numpy() for k, v in rec.items()}
    for b in range(len(lens)):
        n = int(lens[b])
        if n > K + 1:                                                          # train_torch.py:223-225
            orc.save(c["initial_gray"][b], c["frames"][:n, b], c["action"][:n, b], c["reward"][:n, b], c["visits"][:n, b], c["value"][:n, b])


@pytest.mark.parametrize("cap", [100000, 700, 90])
def test_save_episode_matches_oracle(cap):
    """Batched append of whole episodes: no eviction (cap 100000), eviction across episodes (700) and a single batch
    larger than the ring (90)."""
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    K, hist, disc = 5, 32, 0.985
    rb = ReplayBuffer(hist, K, cap, disc, 24, max_moves=64)
    orc = ReplayOracle(hist, K, cap, disc, 24)
    for ep, (B, T) in enumerate([(37, 40), (5, 64), (1100, 12), (24, 3)]):
        rec, lens = _episode_record(B, T, seed=ep)
        rb.save_episode(rec)
        _feed_oracle(orc, rec, lens, K)
        n = rb.length
        assert n == len(orc)
        if n:
            idx = np.arange(n) if n <= 400 else np.random.RandomState(ep).choice(n, 400, replace=False)
            for f in FIELDS:
                name = GETTERS.get(f)
                got = getattr(rb, name)(torch.from_numpy(idx)) if name else rb._gather(torch.from_numpy(idx), (f,))[f]
                assert np.array_equal(got.cpu().numpy(), orc.batch(f, idx).astype(got.cpu().numpy().dtype)), (ep, f)
            assert rb.get_reward_sums() == orc.reward_sums()


def test_large_episode_batch_properties():
    """BASELINE-size acting batch: 4096 trajectories of up to 64 moves into a 60 000-sample buffer; sampled windows against
    the oracle, ring-state arithmetic, and a checkpoint round trip."""
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    K, hist, disc, cap = 5, 32, 0.985, 60000
    B, T = 4096, 64
    rb = ReplayBuffer(hist, K, cap, disc, 512, max_moves=261)
    rec, lens = _episode_record(B, T, seed=5, min_len=0)
    rb.save_episode(rec)
    ns = np.where(lens.numpy() > K + 1, lens.numpy() - K + 1, 0)
    total = int(ns.sum())
    assert total > cap and rb.length == cap
    # logical index -> (env, start): samples are appended in env order, the oldest total - cap are evicted
    first = np.cumsum(ns) - ns
    g = np.random.RandomState(0).choice(cap, 300, replace=False)
    glob = g + (total - cap)
    env = np.searchsorted(first, glob, side="right") - 1
    start = glob - first[env]
    c = {k: v.cpu().numpy() for k, v in rec.items()}
    orc = ReplayOracle(hist, K, 10 ** 9, disc, 1)
    pos = {}
    for b in np.unique(env):
        n = int(lens[b]); pos[b] = len(orc)
        orc.save(c["initial_gray"][b], c["frames"][:n, b], c["action"][:n, b], c["reward"][:n, b], c["visits"][:n, b], c["value"][:n, b])
    oidx = np.array([pos[b] + s for b, s in zip(env, start)])
    mb = rb.minibatch(torch.from_numpy(g))
    for got, f in zip(mb, ("past_actions", "states", "visit_counts", "future_actions", "rewards", "values")):
        assert np.array_equal(got.cpu().numpy(), orc.batch(f, oidx).astype(got.cpu().numpy().dtype)), f
    # second episode on top: length stays at cap, newest sample = last start of the last qualifying env
    rec2, lens2 = _episode_record(64, 30, seed=6, min_len=20)
    rb.save_episode(rec2)
    assert rb.length == cap
    b = 63; n = int(lens2[b]); c2 = {k: v.cpu().numpy() for k, v in rec2.items()}
    o2 = ReplayOracle(hist, K, 10 ** 9, disc, 1)
    o2.save(c2["initial_gray"][b], c2["frames"][:n, b], c2["action"][:n, b], c2["reward"][:n, b], c2["visits"][:n, b], c2["value"][:n, b])
    newest = rb.minibatch(torch.tensor([-1]))
    assert np.array_equal(newest[1].cpu().numpy()[0], o2.batch("states", [len(o2) - 1])[0])
    assert np.array_equal(newest[5].cpu().numpy()[0], o2.batch("values", [len(o2) - 1])[0])
    # checkpoint round trip
    sd = rb.state_dict()
    rb2 = ReplayBuffer(hist, K, cap, disc, 512, max_moves=261)
    rb2.load_state_dict(sd)
    assert rb2.length == cap
    for a, b_ in zip(rb.minibatch(torch.from_numpy(g)), rb2.minibatch(torch.from_numpy(g))):
        assert torch.equal(a, b_)


def test_acting_episode_into_replay_buffer():
    """acting.Actor.run_episode -> ReplayBuffer.save_episode without a host round trip == the oracle fed the same record."""
    from test_acting_gpu import make
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    B, sims, K = 6, 4, 5
    actor, _ = make(B, sims, "f32", temperature=1.0, seed=3, max_moves=40)
    torch.manual_seed(5)
    rec = actor.run_episode()
    lens = rec["recorded"].sum(0).cpu()
    rb = ReplayBuffer(32, K, 1000, 0.985, 24, max_moves=64)
    rb.save_episode(rec)
    orc = ReplayOracle(32, K, 1000, 0.985, 24)
    _feed_oracle(orc, rec, lens, K)
    assert rb.length == len(orc) > 0
    idx = np.arange(len(orc))
    for f in FIELDS:
        got = rb._gather(torch.from_numpy(idx), (f,))[f].cpu().numpy()
        assert np.array_equal(got, orc.batch(f, idx).astype(got.dtype)), f
    assert rb.get_reward_sums() == orc.reward_sums()


def test_checkpoint_restore_through_the_reference_list_attributes(golden_dir):
    """train_torch.py:627-636 pickles the reference ReplayBuffer's per-sample lists and :659-668 assigns them back one after the other:
    the same sequence of assignments on a fresh drop-in buffer must reproduce every get_batched_* output (incl. after FIFO eviction),
    and the restored buffer keeps working (further trajectories append and evict as before)."""
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    g, p, trajs = load_replay_golden(golden_dir)
    src = ReplayBuffer(p["hist"], p["K"], p["cap"], p["discount"], p["n_sum"])
    for tr in trajs[:-2]:                                        # past the capacity: the oldest samples are already evicted
        src.save_observation_trajectory(_ref_trajectory(tr))
    n = src.length
    assert n == p["cap"]
    ckpt = {k: getattr(src, k) for k in ("past_actions_buffer", "future_actions_buffer", "state_buffer", "reward_buffer", "visit_counts_buffer",
                                         "value_buffer", "reward_sums", "length", "max_length", "bootstrapped_values")}
    assert isinstance(ckpt["state_buffer"], list) and tuple(ckpt["state_buffer"][0].shape) == (p["hist"], 1, 16, 20)
    dst = ReplayBuffer(p["hist"], p["K"], p["cap"], p["discount"], p["n_sum"])
    for k in ("past_actions_buffer", "future_actions_buffer", "state_buffer", "reward_buffer", "visit_counts_buffer", "value_buffer",
              "reward_sums", "length", "max_length", "bootstrapped_values"):           # the order of train_torch.py:659-668
        setattr(dst, k, ckpt[k])
    assert dst.length == n
    idx = torch.arange(n)
    for f, getter in GETTERS.items():
        assert torch.equal(getattr(dst, getter)(idx), getattr(src, getter)(idx)), f
    assert dst.get_reward_sums() == src.get_reward_sums() and dst.reward_sums == src.reward_sums
    assert all(torch.equal(a, b) for a, b in zip(dst.value_buffer, src.value_buffer))
    for tr in trajs[-2:]:                                        # both keep going identically
        src.save_observation_trajectory(_ref_trajectory(tr)); dst.save_observation_trajectory(_ref_trajectory(tr))
    assert dst.length == src.length == p["cap"]
    for f, getter in GETTERS.items():
        assert torch.equal(getattr(dst, getter)(idx), getattr(src, getter)(idx)), f
    # samples in an arbitrary order (nothing continues anything: one stored trajectory per sample, a larger entry ring)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))[:40]
    shuf = ReplayBuffer(p["hist"], p["K"], p["cap"], p["discount"], p["n_sum"])
    for k in ("past_actions_buffer", "future_actions_buffer", "state_buffer", "reward_buffer", "visit_counts_buffer", "value_buffer", "bootstrapped_values"):
        setattr(shuf, k, [getattr(src, k)[i] for i in perm.tolist()])
    shuf.reward_sums = [src.reward_sums[i] for i in perm.tolist()]
    shuf.length = 40
    for f, getter in GETTERS.items():
        assert torch.equal(getattr(shuf, getter)(torch.arange(40)), getattr(src, getter)(perm)), f
