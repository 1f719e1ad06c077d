#!/usr/bin/env python
"""Benchmark of the acting hot path (BASELINE.json metric: Breakout env-steps/s and latent MCTS
simulations/s).  Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line
(rank 0).  For N>1 it is launched under torch.distributed.run, one rank per GPU.

Workloads (config.workload in the JSON line):
  env   BreakoutEnvironment.step over --envs environments per GPU (default 65 536: the per-step
        frames, 252 MB, are larger than the 126 MB L2), reference-format outputs (fp32 frames,
        reward, done, valid).  A "step" = one step() of the whole batch.
  mcts  MCTSSearchVec.search over --trees roots per GPU x --sims simulations.  A "step" = one search().
`--impl reference` times the CPU port of the reference path (oracle/, all host threads) on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_CFG = dict(n_parallel=24, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)
ENV_BYTES_PER_STEP = 3898       # 3840 frame + 4 reward + 12 valid + 2 done + 8 action + 32 SoA


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier_sync(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------ env
def bench_env(args, rank, local, world):
    import muzero_breakout_b200 as mzb
    from muzero_breakout_b200 import _lib
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment

    B, K, W = args.envs, args.steps, args.warmup
    L = mzb.lib()
    dev = torch.device("cuda", local)
    env = BreakoutEnvironment(dict(ENV_CFG, n_parallel=B, output_device="cuda", reset_rng="device", seed=1234 + rank))
    state, _ = env.reset()
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    actions = torch.randint(0, 3, (K + W, B), generator=g, device=dev)        # resident in HBM before timing
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    frames = [torch.empty((B, 3, 16, 20), dtype=torch.float32, device=dev) for _ in range(2)]
    reward = torch.empty(B, dtype=torch.float32, device=dev)
    valid = torch.empty((B, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: t.data_ptr()

    def step(i):
        if i % args.reset_every == 0:                          # episode protocol: new games, SoA only (device RNG)
            _lib.check(L.bk_env_reset_device_rng(B, p(env._hdr), p(env._bricks), 1234 + rank, i, None, stream))
            done.zero_()
        _lib.check(L.bk_env_step(B, p(env._hdr), p(env._bricks), p(actions[i]), p(done), p(frames[i & 1]), p(reward), p(valid),
                                 None, env._rewards, p(env._status), stream))

    for i in range(W):
        step(i)
    barrier_sync(world)
    launches0 = mzb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(W, W + K):
            step(i)
        e1.record()
        barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = mzb.launch_count() - launches0
    env.check()
    done_frac = float(done.float().mean().item())

    # e2e: the reference-facing call with HOST tensors (pinned action in, frames/reward/done/valid out)
    env2 = BreakoutEnvironment(dict(ENV_CFG, n_parallel=B, output_device="cpu", reset_rng="device", seed=77 + rank))
    st, _ = env2.reset()
    host_actions = actions.cpu().pin_memory()
    hdone = torch.zeros(B, dtype=torch.bool).pin_memory()
    Ke = max(3, min(K, args.e2e_steps))
    for i in range(3):
        st, *_ = env2.step(st, host_actions[i], hdone)
    barrier_sync(world)
    t0 = time.perf_counter()
    for i in range(3, 3 + Ke):
        st, r_, hdone, v_ = env2.step(st, host_actions[i % (K + W)], hdone)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)

    peaks = measured_peaks()
    kernel_ms = ms / K
    achieved = ENV_BYTES_PER_STEP * B / (kernel_ms * 1e-3) / 1e9
    out = {
        "metric": "breakout_env_steps_per_s", "value": world * B * K / (ms * 1e-3), "unit": "env-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32 logic -> f32 frames", "data": "synthetic",
        "config": {"workload": "env: BreakoutEnvironment.step, reference-format outputs (fp32 (B,3,16,20) frames + reward + done + valid)",
                   "envs_per_gpu": B, "actions": "uniform random, resident in HBM", "reset_every": args.reset_every, "l2": "per-step frame output 3840*B bytes exceeds the 126 MB L2",
                   "done_fraction_at_end": done_frac},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                     "traffic": None, "peak_source": peaks["src"], "kernel": "env_step_kernel<frame>", "bytes_per_env_step": ENV_BYTES_PER_STEP,
                     "kernel_ms": kernel_ms},
        "e2e": {"value": world * B * Ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": B * 9, "d2h_bytes_per_step": B * (3840 + 4 + 12 + 1) + 4,
                "steps": Ke, "api": "BreakoutEnvironment.step(state, action, done_mask) with host tensors (output_device='cpu')"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    return out


def cpu_env(args, sample_envs=None, budget_s=12.0):
    """CPU port of the reference env (oracle/breakout_oracle.c), all host threads, bounded sample."""
    import oracle
    cores = os.cpu_count() or 1
    B = sample_envs or min(args.envs, 65536)
    env = oracle.EnvOracle(B, threads=cores)
    torch.manual_seed(0)
    state = env.reset()
    done = np.zeros(B, np.uint8)
    g = torch.Generator().manual_seed(1)
    acts = torch.randint(0, 3, (64, B), generator=g).numpy()
    for i in range(2):
        state, _, done, _ = env.step(state, acts[i], done)
    n, t0 = 0, time.perf_counter()
    while True:
        state, _, done, _ = env.step(state, acts[n % 64], done)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 400:
            break
    return {"value": B * n / el, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} steps of {B} envs, oracle/breakout_oracle.c (C restatement of parallel_breakout.py step) on {cores} threads, {el:.1f} s"}, el / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="env", choices=["env"])
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--reset-every", type=int, default=32, help="env workload: start new games every this many steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        base, ms = cpu_env(args, budget_s=20.0)
        line = {"impl": "reference", "metric": "breakout_env_steps_per_s", "value": base["value"], "unit": "env-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "env: BreakoutEnvironment.step, reference-format outputs (fp32 (B,3,16,20) frames + reward + done + valid)",
                           "envs_per_gpu": args.envs},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    rank, local, world = dist_setup(args.gpus)
    out = bench_env(args, rank, local, world)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], _ = cpu_env(args)
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
