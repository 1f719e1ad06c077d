import time, torch, sys
sys.path.insert(0, ".")
from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
B = 65536
dev = torch.device("cuda", 0)
ENV_CFG = dict(n_parallel=24, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)
frames = [torch.empty((B, 3, 16, 20), dtype=torch.float32, device=dev) for _ in range(2)]   # like the bench: device buffers allocated first
env2 = BreakoutEnvironment(dict(ENV_CFG, n_parallel=B, output_device="cpu", reset_rng="device", seed=77))
st, _ = env2.reset()
actions = torch.randint(0, 3, (256, B), device=dev)
host_actions = actions.cpu().pin_memory()
hdone = torch.zeros(B, dtype=torch.bool).pin_memory()
ts = []
for i in range(30):
    t0 = time.perf_counter()
    st, r_, hdone, v_ = env2.step(st, host_actions[i % 256], hdone)
    ts.append((time.perf_counter() - t0) * 1e3)
print(" ".join(f"{t:.1f}" for t in ts))
