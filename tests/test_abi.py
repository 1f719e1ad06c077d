"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/mzb200.h declares; the Python hosts mirror the reference signatures and refuse to run
without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import muzero_breakout_b200 as mzb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV_CFG = dict(n_parallel=24, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)


def _declared():
    src = open(os.path.join(ROOT, "include", "mzb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:bk|mz|mzb|rb)_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    so = mzb.build()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mzb200.h but not exported by libmzb200.so"
    assert mzb.lib().mzb_version() == 1


def test_library_is_sm100a_native():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", mzb.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_census_shows_blackwell_tensor_core_and_tma_instructions():
    """profiles/sass_census.py (cuobjdump -sass of the built library): the trunk kernels issue tcgen05 pair MMAs (UTCHMMA.2CTA), TMA
    tensor loads (UTMALDG) and -- the fused trunk's epilogue -- TMA tensor stores (UTMASTG); accumulators come back through LDTM."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import sass_census
    per = sass_census.census(mzb.build())                    # keyed by the mangled kernel names
    def find(sub):
        hits = [v for k, v in per.items() if sub in k]
        assert hits, f"no kernel named *{sub}* in the library"
        return hits
    for c in find("conv_stack_kernel"):
        assert c["UTCHMMA"] >= 4 and c["UTMALDG"] >= 4 and c["UTMASTG"] >= 1 and c["LDTM"] >= 1 and c["full:UTCHMMA.2CTA"] >= 4
    for name in ("conv_tc_kernel", "wgrad_kernel"):
        for c in find(name):
            assert c["UTCHMMA"] >= 4 and c["UTMALDG"] >= 2 and c["LDTM"] >= 1
    for c in find("conv_lat_kernel"):
        assert c["HMMA"] >= 100 and c["UTMALDG"] >= 1          # the latency-mode trunk: warp MMAs fed by TMA weight slices


def test_env_class_mirrors_reference_signature():
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment, MuZeroEnvironment
    env = BreakoutEnvironment(ENV_CFG)
    assert isinstance(env, MuZeroEnvironment)
    assert list(inspect.signature(BreakoutEnvironment.__init__).parameters)[:6] == ["self", "cfg", "width", "height", "paddle_width", "brick_rows"]
    assert list(inspect.signature(env.step).parameters)[:3] == ["state", "action", "done_mask"]
    assert env.state_shape == (24, 3, 16, 20) and env.action_space_size == 3
    env.batch = 2                                   # train_torch.py:448
    assert env.state_shape == (2, 3, 16, 20)
    v = env.get_valid_actions(None, torch.tensor([0, 5, 14]))
    assert v.tolist() == [[0, 1, 1], [1, 1, 1], [1, 1, 0]]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
    env = BreakoutEnvironment(ENV_CFG)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.step(torch.zeros(24, 3, 16, 20), torch.zeros(24, dtype=torch.long), torch.zeros(24, dtype=torch.bool))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "muzero-breakout_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in text


def test_ctypes_mirrors_match_the_c_structs():
    from muzero_breakout_b200 import _lib
    from muzero_breakout_b200.src.networks import MzOp
    L = mzb.lib()
    assert L.mzb_sizeof(0) == ctypes.sizeof(_lib.TreeArgs)
    assert L.mzb_sizeof(1) == ctypes.sizeof(MzOp)
    from muzero_breakout_b200.replay_buffer import RbRing
    assert L.mzb_sizeof(2) == ctypes.sizeof(RbRing)


def test_latency_trunk_host_queries():
    """host-side entry points of the latency-mode trunk (csrc/conv_lat.cu) that need no GPU: sizes, limits, argument checks."""
    from muzero_breakout_b200.src.networks import MzOp, lat_max_samples
    L = mzb.lib()
    assert L.mz_lat_layer_bytes() == 192 and L.mz_lat_max_layers() == 32
    assert L.mz_lat_max_samples() == 81 and lat_max_samples() == 81            # three waves of 27 samples
    for n, rtiles in ((1, 1), (24, 8), (25, 9), (108, 36)):
        assert L.mz_lat_scratch_bytes(n) == 2 * rtiles * 60 * 128 * 8 + 4 * (2 + 2 * rtiles)
    ops = (MzOp * 1)()
    raw = (ctypes.c_uint8 * 512)()
    host = (ctypes.addressof(raw) + 63) & ~63
    assert L.mz_lat_build(ops, 0, host, 192) < 0 and b"bad argument" in L.mzb_last_error()       # no records
    assert L.mz_lat_run(None, 1, 0, 24, None, None, 1, None) < 0                                  # null blob / scratch: refused before any launch


def test_replay_buffer_mirrors_reference_signature():
    """replay_buffer.py:76-94 constructor and the methods train_torch.py calls (:101,147,225,230,377,469-476)."""
    from muzero_breakout_b200.replay_buffer import ObservationTrajectory, ReplayBuffer
    assert list(inspect.signature(ReplayBuffer.__init__).parameters)[:6] == ["self", "seq_len", "K", "max_length", "discount", "num_rewards_to_sum"]
    rb = ReplayBuffer(32, 5, 60000, 0.985, 24)
    for m in ("save_observation_trajectory", "get_batched_past_actions", "get_batched_future_actions", "get_batched_states",
              "get_batched_rewards", "get_batched_visit_counts", "get_batched_values", "get_reward_sums", "empty_buffer"):
        assert callable(getattr(rb, m))
    assert (rb.length, len(rb), rb.max_length, rb.K, rb.hist_seq_len) == (0, 0, 60000, 5, 32)
    ot = ObservationTrajectory(actions=[0] * 32, states=[torch.zeros(1, 16, 20)] * 31, rewards=[0] * 32,
                               visit_counts=[torch.zeros(3)] * 32, values=[0.0] * 32, length=0, reward_sum=0)
    ot.add_observation(torch.tensor(2), torch.ones(1, 16, 20), torch.tensor(1.0), torch.tensor([10, 20, 20]), torch.tensor(0.5))
    assert ot.length == 1 and float(ot.get_reward_sum()) == 1.0 and ot.get_actions().shape == (33,) and ot.get_states().shape == (32, 1, 16, 20)
    L = mzb.lib()
    assert L.rb_entries_for(60000, 5, 512) == (60000 + 512) * 6 + 513 and L.rb_entries_for(0, 5, 512) < 0
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rb.save_observation_trajectory(ot)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rb.get_batched_states(torch.tensor([0]))


def test_training_ends_mirror_reference_signature():
    """loss_fn keeps train_torch.py:33-42's parameter names; mz_loss / mz_adam argument checks run without a GPU."""
    from muzero_breakout_b200 import train
    assert list(inspect.signature(train.loss_fn).parameters) == ["observed_reward", "predicted_reward", "bootstrapped_reward", "predicted_value",
                                                                 "visit_counts", "predicted_policy", "target_transformation", "K"]
    L = mzb.lib()
    assert L.mz_loss_scratch_bytes(2560) == 20 * 3 * 8 + 16 and L.mz_loss_scratch_bytes(0) == 0
    assert L.mz_loss(0, 5, 11, 3, *([None] * 13)) != 0 and b"rows" in L.mzb_last_error()
    assert L.mz_loss(10, 5, 40, 3, *([None] * 13)) != 0 and b"n_supports" in L.mzb_last_error()
    assert L.mz_loss(10, 5, 11, 3, *([None] * 13)) != 0 and b"null" in L.mzb_last_error()
    assert L.mz_adam(16, None, None, None, None, 2e-4, 0.9, 0.999, 1e-8, 1e-4, 1, None) != 0 and b"null" in L.mzb_last_error()
    assert L.mz_adam(16, 64, 64, 64, 64, 2e-4, 0.9, 0.999, 1e-8, 1e-4, 0, None) != 0 and b"step" in L.mzb_last_error()
    assert L.mz_adam(16, 68, 64, 64, 64, 2e-4, 0.9, 0.999, 1e-8, 1e-4, 1, None) != 0 and b"aligned" in L.mzb_last_error()
    # convolution gradients / training-mode BatchNorm: host-side size queries and argument checks
    assert [L.mz_wgrad_padded_samples(n) for n in (0, 1, 64, 65, 512)] == [0, 64, 64, 128, 512]
    assert L.mz_wgrad_partial_bytes(3, 512) == 9 * 8 * 256 * 256 * 4 and L.mz_wgrad_partial_bytes(1, 100) == 2 * 256 * 256 * 4
    assert L.mz_wgrad_partial_bytes(2, 512) == 0
    assert L.mz_conv_wgrad(512, 4, 5, 2, 1, 64, 64, 64, 64, None) != 0 and b"bad argument" in L.mzb_last_error()
    assert L.mz_conv_wgrad(512, 4, 5, 3, 0, 64, 64, 64, 64, None) != 0          # fp32 operands are not built
    assert L.mz_conv_wgrad(512, 4, 5, 3, 1, None, 64, 64, 64, None) != 0 and b"null" in L.mzb_last_error()
    assert L.mz_wgrad_transpose(512, 20, 100, 64, 64, None) != 0                 # channels not a multiple of 64
    assert L.mz_bn_scratch_bytes(10240, 256) == 80 * 2 * 256 * 8 and L.mz_bn_scratch_bytes(10240, 6) == 0      # 128 rows per partial-sum block
    assert L.mz_bn_train_fwd(0, 256, *([None] * 4), 1, 1, 1e-5, 0.1, *([None] * 8)) != 0 and b"M must be positive" in L.mzb_last_error()
    assert L.mz_bn_train_fwd(64, 256, 64, 64, 64, None, 1, 3, 1e-5, 0.1, None, None, 64, 64, 64, None, 64, None) != 0 and b"activation" in L.mzb_last_error()
    assert L.mz_bn_train_bwd(64, 256, 64, 64, 64, 64, None, 0, 1, *([64] * 4), 64, None, None, 64, None) != 0 and b"16-bit" in L.mzb_last_error()
    # the data gradient's weight pack: dx = conv2d(dy, w.transpose(0, 1).flip(2, 3)) (train.ConvDgrad), checked against autograd on the CPU
    w, dy = torch.randn(8, 4, 3, 3), torch.randn(2, 8, 4, 5)
    x = torch.zeros(2, 4, 4, 5, requires_grad=True)
    torch.nn.functional.conv2d(x, w, padding=1).backward(dy)
    assert torch.allclose(torch.nn.functional.conv2d(dy, train.ConvDgrad.dgrad_filter(w), padding=1), x.grad, atol=1e-5)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            train.ConvDgrad(torch.zeros(256, 256, 3, 3))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            train.conv_wgrad(torch.zeros(2, 4, 5, 256), torch.zeros(2, 4, 5, 256), 3)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            train.loss_fn(torch.zeros(2, 5), torch.zeros(2, 5, 11), torch.zeros(2, 5), torch.zeros(2, 5, 11), torch.ones(2, 5, 3),
                          torch.zeros(2, 5, 3), torch.linspace(-5, 5, 11), 5)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            train.Adam([torch.nn.Parameter(torch.zeros(4))])


def test_mcts_class_mirrors_reference_signature():
    from muzero_breakout_b200.src.mcts import MCTSSearchVec
    cfg = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5],
           "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985}}
    m = MCTSSearchVec(cfg, object(), None)
    assert list(inspect.signature(MCTSSearchVec.__init__).parameters)[:4] == ["self", "cfg", "mu_zero", "scalar_transforms"]
    assert list(inspect.signature(m.search).parameters)[:3] == ["hidden_state", "action_mask", "training_iteration"]
    assert (m.noise_weight, m.dirchlet_alpha, m.num_simulations) == (0.175, 0.25, 50)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.search(torch.zeros(2, 256, 4, 5), torch.ones(2, 3), 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout only exists in the build container")
def test_dropin_shadow_modules_resolve_before_the_reference():
    """INTEGRATION.md option A: with dropin/ ahead of the reference on sys.path, the reference's own
    get_class() loads our classes, while src.networks still comes from the reference."""
    import subprocess
    import sys
    code = ("import sys, types\n"
            "for n in ('matplotlib', 'matplotlib.pyplot'): sys.modules.setdefault(n, types.ModuleType(n))\n"
            "from utils import get_class\n"
            "m = get_class('src.mcts', 'MCTSSearchVec'); e = get_class('environment.parallel_breakout', 'BreakoutEnvironment')\n"
            "n = get_class('src.networks', 'MuZeroAgent')\n"
            "import replay_buffer as r\n"
            "print(m.__module__, e.__module__, n.__module__, r.ReplayBuffer.__module__)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, "/root/reference"]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["muzero_breakout_b200.src.mcts", "muzero_breakout_b200.environment.parallel_breakout", "src.networks",
                                  "muzero_breakout_b200.replay_buffer"]
