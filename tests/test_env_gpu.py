"""GPU parity tests of the Breakout kernels (csrc/env.cu through the C ABI / BreakoutEnvironment)
against the committed golden vectors (outputs of the reference) and the CPU oracle.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle
from common import unpack_state

pytestmark = pytest.mark.gpu

ENV_CFG = dict(n_parallel=24, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)


def make_env(B=24, **kw):
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
    cfg = dict(ENV_CFG, n_parallel=B, **kw)
    return BreakoutEnvironment(cfg)


def _eq(a, b, what):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[0].tolist()}: got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")


def _check_rec(rec, t, out, env, where):
    nxt, reward, done, valid = out
    _eq(nxt, unpack_state(rec["state"][t]), f"{where} step {t} next_state")
    _eq(reward, rec["reward"][t], f"{where} step {t} reward")
    _eq(done.to(torch.uint8), rec["done"][t], f"{where} step {t} done")
    _eq(valid, rec["valid"][t].astype(np.float32), f"{where} step {t} valid")
    _eq(env.ball_dx, rec["dx"][t].astype(np.int64), f"{where} step {t} ball_dx")
    _eq(env.ball_dy, rec["dy"][t], f"{where} step {t} ball_dy")


def test_config1_10k_steps_vs_reference_golden(golden_dir):
    """BASELINE.json configs[0]: B=24, torch.manual_seed(42), 10 000 steps, reference episode protocol,
    reference-compatible host tensors (output_device='cpu')."""
    g = np.load(os.path.join(golden_dir, "env_config1.npz"))
    rec = {k: g[k] for k in g.files}
    steps, B = rec["actions"].shape
    env = make_env(B)
    torch.manual_seed(42)
    resets = {int(t): i for i, t in enumerate(rec["reset_at"])}
    state = done = None
    for t in range(steps):
        if t in resets:
            state, zero = env.reset()
            assert zero == 0 and state.device.type == "cpu" and state.dtype == torch.float32
            _eq(state, unpack_state(rec["reset_state"][resets[t]]), f"reset {resets[t]}")
            _eq(env.ball_dx, rec["reset_dx"][resets[t]].astype(np.int64), "reset ball_dx")
            done = torch.zeros(B, dtype=torch.bool)
        action = torch.from_numpy(rec["actions"][t].astype(np.int64))
        out = env.step(state, action, done)
        assert out[2] is done                                  # mutated in place and returned (:204,:247)
        _check_rec(rec, t, out, env, "config1")
        state = out[0]


@pytest.mark.parametrize("name", ["env_fuzz.npz", "env_play.npz"])
def test_corpora_vs_reference_golden(golden_dir, name):
    """Appendix-B fuzz corpus (foreign states -> ingest path, wins, losses, brick rows 0-4) and the
    ball-following play corpus (row -1 wrap-around, +5 after done), device-resident tensors."""
    g = np.load(os.path.join(golden_dir, name))
    seeds, steps, B = g["actions"].shape
    env = make_env(B, output_device="cuda")
    for s in range(seeds):
        rec = {k: g[k][s] for k in ("state", "reward", "done", "valid", "dx", "dy")}
        if "init_state" in g.files:
            state = torch.from_numpy(unpack_state(g["init_state"][s])).cuda()
            env.ball_dx = torch.from_numpy(g["init_dx"][s].astype(np.int64))
            env.ball_dy = torch.from_numpy(g["init_dy"][s].astype(np.float32))
        else:
            torch.manual_seed(2000 + s)
            state, _ = env.reset()
            _eq(state, unpack_state(g["reset_state"][s]), "reset state")
        done = torch.zeros(B, dtype=torch.bool, device="cuda")
        for t in range(steps):
            action = torch.from_numpy(g["actions"][s, t].astype(np.int64)).cuda()
            out = env.step(state, action, done)
            assert out[2] is done and out[0].is_cuda
            _check_rec(rec, t, out, env, f"{name} seed {s}")
            state = out[0]
            if t % 7 == 3:                                     # hand back a copy: forces the ingest path
                state = state.clone()
    env.check()


def test_large_ragged_batch_vs_oracle():
    """B not a multiple of 32 or 128, 400 steps, compared with the CPU oracle every step."""
    B = 4099
    env = make_env(B, output_device="cuda")
    orc = oracle.EnvOracle(B, threads=8)
    torch.manual_seed(7)
    state, _ = env.reset()
    torch.manual_seed(7)
    ostate = orc.reset()
    _eq(state, ostate, "reset")
    done = torch.zeros(B, dtype=torch.bool, device="cuda")
    odone = np.zeros(B, np.uint8)
    g = torch.Generator().manual_seed(11)
    mixed = 0
    for t in range(400):
        # ball-following 70 % of the time so that games last and bricks get hit
        bx = torch.from_numpy(ostate[:, 1].reshape(B, -1).argmax(1) % 20)
        px = torch.from_numpy(ostate[:, 0, 15].argmax(1)) + 3
        follow = torch.where(bx < px, 0, torch.where(bx > px, 2, 1))
        a = torch.where(torch.rand(B, generator=g) < 0.7, follow, torch.randint(0, 3, (B,), generator=g))
        nxt, reward, done, valid, gray = env.step(state, a.cuda(), done, want_gray=True)
        ostate, oreward, odone, ovalid = orc.step(ostate, a, odone)
        _eq(nxt, ostate, f"step {t} next_state")
        _eq(reward, oreward, f"step {t} reward")
        _eq(done.to(torch.uint8), odone, f"step {t} done")
        _eq(valid, ovalid, f"step {t} valid")
        if t % 50 == 0:
            _eq(gray, oracle.gray(ostate), f"step {t} fused gray")
        state = nxt
        mixed += int(0 < odone.sum() < B)
    env.check()
    assert mixed > 20                                          # finished and running games side by side


def test_full_size_properties_65536():
    """BASELINE full size (65 536 envs): size-independent invariants + one oracle spot check."""
    B = 65536
    env = make_env(B, output_device="cuda", reset_rng="device", seed=3)
    state, _ = env.reset()
    s0 = state.clone()
    done = torch.zeros(B, dtype=torch.bool, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    prev_bricks = state[:, 2].sum((1, 2))
    prev_done = done.clone()
    for t in range(120):
        a = torch.randint(0, 3, (B,), generator=g, device="cuda")
        if t == 60:                                            # oracle spot check on the whole batch
            orc = oracle.EnvOracle(B, threads=8)
            dx, dy = env.ball_dx, env.ball_dy
            orc.ball_dx[:], orc.ball_dy[:] = dx.cpu().numpy(), dy.cpu().numpy()
            on, orw, od, ov = orc.step(state.cpu().numpy(), a.cpu(), done.cpu().numpy().astype(np.uint8))
        nxt, reward, done, valid = env.step(state, a, done)
        if t == 60:
            _eq(nxt, on, "65536 next_state"); _eq(reward, orw, "65536 reward"); _eq(done.to(torch.uint8), od, "65536 done"); _eq(valid, ov, "65536 valid")
        assert torch.all((nxt == 0) | (nxt == 1))
        assert torch.all(nxt[:, 1].sum((1, 2)) == 1)                       # exactly one ball pixel (:248)
        prow = nxt[:, 0, 15].sum(1)
        assert torch.all(nxt[:, 0, :15] == 0) and torch.all((prow == 6) | ((prow == 0) & done))
        bricks = nxt[:, 2].sum((1, 2))
        assert torch.all(bricks <= prev_bricks) and torch.all(bricks[done] == 0)
        assert torch.all(done | ~prev_done)                                # done is monotone
        assert torch.all(torch.isin(reward, torch.tensor([-1.0, 0.0, 1.0, 5.0, 6.0], device="cuda")))
        assert torch.all(reward[prev_done] == 5.0)                         # +5 every step after done
        prev_bricks, prev_done, state = bricks, done.clone(), nxt
    env.check()
    # idempotence of the SoA <-> dense round trip: re-ingesting a frame and stepping gives the same
    a = torch.randint(0, 3, (B,), generator=g, device="cuda")
    d1, d2 = done.clone(), done.clone()
    dx, dy = env.ball_dx, env.ball_dy
    hdr, bricks = env._hdr.clone(), env._bricks.clone()
    n1 = env.step(state, a, d1)
    env._hdr.copy_(hdr); env._bricks.copy_(bricks)
    n2 = env.step(state.clone(), a, d2)                                    # foreign tensor -> ingest
    for x, y, w in zip(n1, n2, ("state", "reward", "done", "valid")):
        _eq(x, y, f"round trip {w}")
    # device-RNG reset: reference ranges (parallel_breakout.py:116-137)
    px = s0[:, 0, 15].argmax(1); by = s0[:, 1].sum(2).argmax(1); bx = s0[:, 1].sum(1).argmax(1)
    assert int(px.min()) == 1 and int(px.max()) == 14 and int(bx.min()) == 1 and int(bx.max()) == 18
    assert set(by.unique().tolist()) == {13, 14} and torch.all(s0[:, 2, :3] == 1) and torch.all(s0[:, 2, 3:] == 0)


def test_gray_kernel_matches_oracle():
    env = make_env(5, output_device="cuda")
    g = torch.Generator().manual_seed(0)
    s = (torch.rand(333, 3, 16, 20, generator=g) < 0.3).float()
    _eq(env.gray(s), oracle.gray(s), "gray")


def test_batch_reassignment_and_ownership():
    env = make_env(24)
    torch.manual_seed(1)
    s24, _ = env.reset()
    env.batch = 2                                                          # train_torch.py:448
    s2, _ = env.reset()
    assert s2.shape == (2, 3, 16, 20)
    d = torch.zeros(2, dtype=torch.bool)
    n1, *_ = env.step(s2, torch.tensor([0, 2]), d)
    keep = n1.clone()
    n2, *_ = env.step(n1, torch.tensor([1, 1]), d)
    assert torch.equal(n1, keep) and n2.data_ptr() != n1.data_ptr()       # outputs are never reused
    env.batch = 24
    s, _ = env.reset()
    assert s.shape == (24, 3, 16, 20)


def test_errors_are_loud():
    env = make_env(2)
    # a brick on row 14 under an upward ball on row 15: the reference raises IndexError (:243)
    s = torch.zeros(2, 3, 16, 20)
    s[:, 0, 15, 3:9] = 1; s[:, 1, 15, 10] = 1; s[:, 2, 14, 10:12] = 1; s[:, 2, 0, :] = 1
    env.batch = 2
    env.reset()
    env.ball_dx = torch.tensor([0, 0]); env.ball_dy = torch.tensor([-1.0, -1.0])
    with pytest.raises(IndexError):
        env.step(s, torch.tensor([1, 1]), torch.zeros(2, dtype=torch.bool))
    with pytest.raises(IndexError):
        orc = oracle.EnvOracle(2); orc.ball_dx[:] = 0; orc.ball_dy[:] = -1
        orc.step(s.numpy(), np.array([1, 1]), np.zeros(2, np.uint8))
    bad = torch.zeros(2, 3, 16, 20)                                        # no ball pixel
    with pytest.raises(IndexError):
        env.step(bad, torch.tensor([1, 1]), torch.zeros(2, dtype=torch.bool))


@pytest.mark.parametrize("B", [1, 2, 3, 31, 33])
def test_tiny_batches_vs_oracle(B):
    """batch sizes below one warp task / one block, host tensors (the reference-facing call)."""
    env = make_env(B)
    orc = oracle.EnvOracle(B)
    torch.manual_seed(B)
    state, _ = env.reset()
    torch.manual_seed(B)
    ostate = orc.reset()
    _eq(state, ostate, "reset")
    done, odone = torch.zeros(B, dtype=torch.bool), np.zeros(B, np.uint8)
    g = torch.Generator().manual_seed(B)
    for t in range(70):
        a = torch.randint(0, 3, (B,), generator=g)
        state, reward, done, valid = env.step(state, a, done)
        ostate, oreward, odone, ovalid = orc.step(ostate, a, odone)
        _eq(state, ostate, f"step {t} state"); _eq(reward, oreward, f"step {t} reward")
        _eq(done.to(torch.uint8), odone, f"step {t} done"); _eq(valid, ovalid, f"step {t} valid")
