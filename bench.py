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

REFERENCE_ARM = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:]) and any(a.startswith("--impl") for a in sys.argv[1:])
if REFERENCE_ARM:
    # the reference arm times the reference's CPU path on the host cores: its hard-coded "cuda" device strings are redirected to
    # "cpu" (baseline/ref.py) only when no GPU is visible, so hide the GPUs before torch is imported
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):       # torchrun sets OMP_NUM_THREADS=1: the arm uses every host core
        os.environ.pop(_k, None)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_CFG = dict(n_parallel=24, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0)
ENV_BYTES_PER_STEP = 3898       # 3840 frame + 4 reward + 12 valid + 2 done + 8 action + 32 SoA
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/):
ENV_TRAFFIC_PER_ENV = (1.92e6 + 196.2e6) / 65536     # r1_env_step_final_full.txt, 65 536 envs per launch (the L2 keeps part of the frames)
# DRAM bytes per (sample, trunk layer) of conv_stack_kernel: profiles/r1_conv_stack_scout_full.txt, the 28-layer prediction trunk at
# 4096 samples per launch read 801.7 MB and wrote 1078.1 MB
TRUNK_TRAFFIC_PER_SAMPLE_LAYER = (1039.678e6 + 1073.946e6) / (29 * 4096)   # profiles/r2_conv_stack_final_full.txt: prediction trunk, 29 layer records
FLOP_TRUNK_LAYER_VALID = 130 * 256 * 256 * 2           # one 3x3 256->256 conv on the 4x5 latent, in-bounds taps only, per sample


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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's version banner off stdout (ONE JSON line)
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
    NA = 256
    actions = torch.randint(0, 3, (NA, B), generator=g, device=dev)           # resident in HBM before timing, reused every NA steps
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
        _lib.check(L.bk_env_step(B, p(env._hdr), p(env._bricks), p(actions[i % NA]), p(done), p(frames[i & 1]), p(reward), p(valid),
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
    for i in range(4):                                         # warm-up: the pinned output blocks come out of torch's caching host allocator
        st, r_, hdone, v_ = env2.step(st, host_actions[i % NA], hdone)       # (same names as the timed loop: no stale reference pins a block)
    barrier_sync(world)
    step_ms = []
    t0 = time.perf_counter()
    for i in range(4, 4 + Ke):
        t1 = time.perf_counter()
        st, r_, hdone, v_ = env2.step(st, host_actions[i % NA], hdone)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)

    aux24 = None
    if rank == 0 and not getattr(args, "no_aux", False):
        # BASELINE.json configs[0]: config.yaml default env count (24), reference-facing call with host tensors
        env24 = BreakoutEnvironment(dict(ENV_CFG, n_parallel=24))
        torch.manual_seed(42)
        s24, _ = env24.reset()
        d24 = torch.zeros(24, dtype=torch.bool)
        acts24 = torch.randint(0, 3, (1000, 24))
        for i in range(20):
            s24, *_ = env24.step(s24, acts24[i], d24)
        t0 = time.perf_counter()
        for i in range(20, 1000):
            s24, r24, d24, v24 = env24.step(s24, acts24[i], d24)
        el = time.perf_counter() - t0
        aux24 = {"workload": "BreakoutEnvironment.step, 24 envs (config.yaml default), host tensors in and out (BASELINE.json configs[0])",
                 "us_per_step": el / 980 * 1e6, "value": 24 * 980 / el, "unit": "env-steps/s", "note": "launch- and copy-latency-bound"}

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
                     "traffic": ENV_TRAFFIC_PER_ENV * B, "traffic_source": "profiles/r1_env_step_final_full.txt (ncu --set full, scaled by envs per launch)",
                     "peak_source": peaks["src"], "kernel": "env_step_kernel<frame>", "bytes_per_env_step": ENV_BYTES_PER_STEP,
                     "kernel_ms": kernel_ms},
        "e2e": {"value": world * B * Ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": B * 9, "d2h_bytes_per_step": B * (3840 + 4 + 12 + 1) + 4,
                "steps": Ke, "ms_per_step_min_median_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))],
                "api": "BreakoutEnvironment.step(state, action, done_mask) with host tensors (output_device='cpu')"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if aux24:
        out["config_defaults"] = aux24
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



# ------------------------------------------------------------------------------------------ mcts
MCTS_CFG = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5],
            "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "mcts_name": "MCTSSearchVec"}}
# BASELINE.md section 3 (hooks on the reference modules), FLOP = 2 x MAC
FLOP_LEAF_VALID = 984_079_360          # dynamics + prediction, zero-padding taps excluded
FLOP_LEAF_DENSE = 1_360_988_160        # what a dense 3x3 conv executes
FLOP_ROOT_PRED_VALID, FLOP_ROOT_PRED_DENSE = 487_004_160, 673_781_760
FLOP_CONV_VALID = 496_962_560 + 486_932_480      # conv layers only (no Linear heads), per leaf
FLOP_CONV_DENSE = 687_093_760 + 673_710_080


def _time_fn(fn, reps=50):
    """mean duration over `reps` back-to-back runs: long enough (~0.2 s) that the clocks settle under the power cap,
    like inside a search and like the sustained cuBLAS figure the roofline divides by"""
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_mcts(args, rank, local, world):
    from muzero_breakout_b200.src.mcts import MCTSSearchVec
    from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, OP_CONV, PackedNetworks, Program, random_state_dict

    B, S, K, W = args.trees, args.sims, args.steps, args.warmup
    dev = torch.device("cuda", local)
    sd = random_state_dict(DEFAULT_MODEL_CFG, seed=0, bn_jitter=0.2)
    nets = PackedNetworks(sd, DEFAULT_MODEL_CFG, precision=args.precision, device=dev)
    cfg = dict(MCTS_CFG, num_simulations=S, model=DEFAULT_MODEL_CFG)
    cfg["search"] = dict(MCTS_CFG["search"], precision=args.precision, output_device="cuda", cuda_device=str(dev), seed=17 + rank)
    m = MCTSSearchVec(cfg, nets, None)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    hiddens = [torch.rand((B, 256, 4, 5), generator=g, device=dev) for _ in range(2)]    # root latents are in [0,1] (_scale_state)
    mask = torch.ones((B, 3), device=dev)
    # N > 1 (BASELINE.json configs[3]): the algorithm's two exchange steps run INSIDE the timed loop -- the packed target weights are
    # broadcast from rank 0 every 15th search (train_torch.py:361-367 refreshes the target network every 15 iterations) and every search's
    # per-env trajectory record (gray frame, action, reward, visit counts, value: train_torch.py:204-208) is all-gathered
    coll = None
    if world > 1:
        from muzero_breakout_b200 import parallel
        coll = {"bc_ms": [], "ag_ms": [], "bytes_bc": 0}
        rec = torch.zeros((B, parallel.RECORD_FLOATS), device=dev)
        gather = parallel.AsyncTrajectoryGather(B, parallel.RECORD_FLOATS, device=dev, depth=2)

    def one_search(i, timed):
        if coll is not None and i % 15 == 0:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record(); coll["bytes_bc"] = parallel.broadcast_weights(nets, src=0); ev[1].record()
            if timed:
                coll["bc_ms"].append(ev)
        value, visits = m.search(hiddens[i & 1], mask, 0)
        if coll is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            slot = gather.slot()                                          # waits only for the collective of two moves ago
            slot[:, 322:325] = visits; slot[:, 325] = value               # this move's record (the frame / action / reward columns come from the env step)
            gather.submit()                                               # asynchronous: the next search does not wait for the other ranks
            ev[1].record()
            if timed:
                coll["ag_ms"].append(ev)
        return value, visits

    for i in range(W):
        one_search(i, False)
    if coll is not None:
        gather.drain()
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            value, visits = one_search(i, True)
        if coll is not None:
            gathered = gather.drain()                                     # every record of the timed searches has arrived before the clock stops
            assert len(gathered) >= min(K, 2) and all(g_.shape[0] == world * B for g_ in gathered)
            torch.cuda.current_stream(dev).synchronize()
        e1.record()
        barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    assert int(visits.sum().item()) == B * S, "visit counts do not sum to num_simulations"
    plan = next(iter(m._plans.values()))
    launches = K * plan.kernels_per_search

    # dominant kernel: the persistent tcgen05 trunk launches (conv_stack_kernel) of one simulation step, timed alone with CUDA events
    # on the launching stream, back to back for ~0.2 s (clocks settle under the power cap like inside a search)
    prog = plan.sim_prog
    if prog._segs is None:
        prog._build()
    pv_w = nets.pv_conv.w.data_ptr() if nets.pv_conv is not None else None

    def conv_flop(o):
        """in-bounds-tap FLOPs per sample of one convolution record on the 4x5 latent"""
        if pv_w is not None and o.w == pv_w:          # policy 3x3 256->128 + value 1x1 256->128 packed as one 3x3 256->256 layer: count the two ConvBlocks
            return (130 + 20) * 256 * 128 * 2
        return (130 if o.ksize == 3 else 20) * o.cin * o.cout * 2

    stacks, trunk_flop, n_trunk, i_op = [], 0, 0, 0
    for kind, item, cnt in prog._segs:
        if kind == "stack":
            stacks.append(item)
            for o in prog.ops[i_op:i_op + cnt]:
                if o.op == OP_CONV:
                    trunk_flop += conv_flop(o); n_trunk += 1
        i_op += cnt
    st = torch.cuda.current_stream(dev).cuda_stream
    if stacks:
        trunk_ms = _time_fn(lambda: [s_.run(st) for s_ in stacks])
        trunk_launches = len(stacks)
        kernel_name = "conv_stack_kernel (tcgen05 residual trunk + head ConvBlocks)" if not stacks[0].lat else "conv_lat_kernel (latency-mode trunk)"
    else:                                                   # fp32 parity path: no tensor-core trunk, report all convolutions
        conv_prog = Program(B, False)
        conv_prog.ops = [o for o in prog.ops if o.op == OP_CONV]
        conv_prog.keep = prog.keep
        trunk_ms = _time_fn(conv_prog.run)
        trunk_flop, n_trunk, trunk_launches = sum(conv_flop(o) for o in conv_prog.ops), len(conv_prog.ops), conv_prog.n_kernels
        kernel_name = "conv_simt_kernel (fp32 CUDA cores)"
    step_ms = _time_fn(prog.run)

    # e2e: the reference-facing call with a HOST latent tensor in and HOST results out
    cfg2 = dict(cfg); cfg2["search"] = dict(cfg["search"], output_device="cpu")
    m2 = MCTSSearchVec(cfg2, nets, None)
    m2._plans = m._plans                                  # same preallocated plan / graph
    host_hidden = hiddens[0].cpu().pin_memory()
    host_mask = torch.ones(B, 3)
    Ke = max(2, min(K, args.e2e_steps))
    m2.search(host_hidden, host_mask, 0)
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(Ke):
        v_, n_ = m2.search(host_hidden, host_mask, 0)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)

    collectives = None
    if coll is not None:
        # the exchange steps as they ran inside the timed loop above (device events per call, mean over calls, max over ranks)
        torch.cuda.synchronize()
        bc_ms = max_over_ranks(float(np.mean([a_.elapsed_time(b_) for a_, b_ in coll["bc_ms"]])) if coll["bc_ms"] else 0.0, world)
        ag_ms = max_over_ranks(float(np.mean([a_.elapsed_time(b_) for a_, b_ in coll["ag_ms"]])), world)
        rec_bytes = B * parallel.RECORD_FLOATS * 4
        # the same two collectives back to back right after a barrier (no rank skew in the number): what the links deliver
        def alone(fn, reps=5):
            fn(); barrier_sync(world)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(reps):
                fn()
            b_.record(); barrier_sync(world)
            return max_over_ranks(a_.elapsed_time(b_) / reps, world)
        bc_alone = alone(lambda: parallel.broadcast_weights(nets, src=0))
        ag_alone = alone(lambda: parallel.all_gather_trajectory(rec, equal_shards=True))
        collectives = {"in_timed_loop": True, "backend": "nccl", "weights_broadcasts": len(coll["bc_ms"]), "weights_bytes": coll["bytes_bc"],
                       "weights_broadcast_ms_in_loop": bc_ms, "weights_broadcast_ms_alone": bc_alone, "weights_broadcast_GBps": coll["bytes_bc"] / bc_alone / 1e6,
                       "trajectory_allgathers": len(coll["ag_ms"]), "trajectory_bytes_per_rank": rec_bytes,
                       "trajectory_allgather_ms_in_loop": ag_ms, "trajectory_allgather_ms_alone": ag_alone,
                       "trajectory_allgather_GBps_per_rank_in": rec_bytes * (world - 1) / ag_alone / 1e6,
                       "share_of_step": (bc_ms * len(coll["bc_ms"]) + ag_ms * len(coll["ag_ms"])) / ms,
                       "trajectory_allgather_mode": "asynchronous, two staging slots (parallel.AsyncTrajectoryGather): the next search starts without waiting for the other ranks; all records drained before the clock stops",
                       "note": "in_loop = time on the acting stream per call inside the timed searches (broadcast: waits for the slowest rank; all-gather: the slot hand-over only); alone = blocking, back to back after a barrier"}

    peaks = measured_peaks()
    sims = world * B * S * K
    achieved = trunk_flop * B / (trunk_ms * 1e-3) / 1e12
    tol_met = args.precision in ("f16", "f32")
    out = {
        "metric": "latent_mcts_simulations_per_s", "value": sims / (ms * 1e-3), "unit": "simulations/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "tolerance_met": tol_met,
        "config": {"workload": f"mcts: MCTSSearchVec.search, {B} roots x {S} simulations per GPU, random-init MuZero networks (config.yaml sizes), {args.precision}",
                   "trees_per_gpu": B, "num_simulations": S, "step": "one search() call = root prediction + S x (dynamics + prediction + backup/select)"
                   + (" + the asynchronous trajectory all-gather; target-weight broadcast every 15th search" if world > 1 else ""),
                   "l2": f"latent store {B * (S + 2) * 10240 * (4 if args.precision == 'f32' else 2) // 2 / 1e9:.2f} GB per GPU, far larger than the 126 MB L2",
                   "cuda_graph": bool(m.use_graph),
                   "arithmetic": {"f16": "fp16 tensor-core operands, fp32 accumulation and epilogues, residual stream as fp16 + e4m3 correction", "bf16": "bf16 operands, fp32 accumulation",
                                  "f32": "fp32 CUDA cores"}[args.precision],
                   "tolerance": "network outputs within 1e-3 (f16) / 1e-5 (f32) of the fp32 reference, range-relative: tests/test_networks_gpu.py" if tol_met
                   else "bf16 storage: 3-6e-3 of the fp32 reference (misses the north star's 1e-3; use --precision f16)"},
        "e2e": {"value": world * B * S * Ke / e2e_s, "unit": "simulations/s", "h2d_bytes_per_step": B * (5120 * 4 + 12), "d2h_bytes_per_step": B * (4 + 24),
                "steps": Ke, "api": "MCTSSearchVec.search(hidden_state, action_mask, training_iteration) with host tensors in and out"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if collectives:
        out["collectives"] = collectives
    out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                       "frac": achieved / peaks["bf16_sustained"], "traffic": TRUNK_TRAFFIC_PER_SAMPLE_LAYER * n_trunk * B / trunk_launches,
                       "traffic_source": "profiles/r2_conv_stack_final_full.txt (ncu --set full, the 29-record prediction-trunk launch at 4096 samples: 1.04 GB read + 1.07 GB written, scaled by layers x samples per launch)",
                       "peak_source": peaks["src"] + " (sustained cuBLAS bf16, same power cap)",
                       "kernel": f"{kernel_name}: {n_trunk} conv layers of one simulation step in {trunk_launches} launches", "flop_convention": "valid taps only (BASELINE.md section 3)",
                       "flop_per_launch": trunk_flop * B / trunk_launches, "kernel_ms": trunk_ms / trunk_launches,
                       "all_kernels_ms_per_sim_step": step_ms,
                       "search_achieved_valid_tap": (FLOP_LEAF_VALID * S + FLOP_ROOT_PRED_VALID) * B * K / (ms * 1e-3) / 1e12}
    if rank == 0 and world == 1 and not args.no_aux:
        try:                                  # the reference as it is on the same GPU (torch + cuDNN): never let it take the headline line down
            out["gpu_reference"] = gpu_reference(args, sd, dev, B, step_ms)
        except Exception as e:                # noqa: BLE001
            out["gpu_reference"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0 and world == 1 and not args.no_aux:
        try:                                  # one training-loop iteration, drop-in learner vs the reference modules on the same GPU
            out["train_step"] = bench_train_step(dev)
        except Exception as e:                # noqa: BLE001
            out["train_step"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_acting:
        out["acting"] = bench_acting(args, m, dev, rank, world)
    if not args.no_aux:
        out["replay"] = bench_replay(args, dev, rank, world, cpu=not args.no_cpu_baseline)
        try:                                  # auxiliary (SURVEY 8f row 4 pieces): never let it take the headline line down
            out["train_ends"] = bench_train_ends(dev, rank, world, cpu=not args.no_cpu_baseline)
        except Exception as e:                # noqa: BLE001
            out["train_ends"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0 and not args.no_aux:
        # BASELINE.json configs[1]: config.yaml defaults (24 roots x 50 simulations), same weights, one GPU
        cfg24 = dict(cfg); cfg24["search"] = dict(cfg["search"], seed=3)
        m24 = MCTSSearchVec(cfg24, nets, None)
        h24 = hiddens[0][:24].contiguous()
        for _ in range(3):
            m24.search(h24, None, 0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m24.search(h24, None, 0)
        b.record(); torch.cuda.synchronize()
        t24 = a.elapsed_time(b) / 5
        out["config_defaults"] = {"workload": "MCTSSearchVec.search, config.yaml defaults: 24 roots x 50 simulations (BASELINE.json configs[1])",
                                  "ms_per_search": t24, "value": 24 * S / (t24 * 1e-3), "unit": "simulations/s",
                                  "note": "latency-bound (57 dependent trunk layers per simulation): the networks run in latency mode, csrc/conv_lat.cu -- 128 work items "
                                          "of 3 samples x 16 output channels per layer on mma.sync, flag-in-data layer hand-off, head convolutions / heads / _scale_state "
                                          "inside the same launch: 3 launches per simulation step, ~4.8 us per layer against ~20 us on the tcgen05 trunk"}
    return out, sd


def bench_acting(args, m, dev, rank, world):
    """Whole acting moves on the device: history ring -> rep-net input ->
    representation network -> search -> action sampling -> env step (+ fused grayscale) -> record."""
    from muzero_breakout_b200.acting import Actor
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
    B = args.trees
    env = BreakoutEnvironment(dict(ENV_CFG, n_parallel=B, output_device="cuda", reset_rng="device", seed=5 + rank, cuda_device=str(dev)))
    moves = 4
    actor = Actor(env, m, temperature=1.0, seed=rank, max_moves=moves, check_done_every=1 << 30, record_frames=False)
    actor.run_episode()
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = actor.run_episode()
    e1.record()
    barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    return {"metric": "acting_env_moves_per_s", "value": world * B * moves / (ms * 1e-3), "unit": "env-moves/s", "ms_per_move": ms / moves,
            "envs_per_gpu": B, "num_simulations": args.sims, "moves_timed": moves,
            "what": "reset + per move: mz_rep_input, representation net, MCTSSearchVec.search, mz_sample_actions, bk_env_step(+gray), record"}


def bench_replay(args, dev, rank, world, cpu=True):
    """Device replay buffer: one rb_append of a whole acting batch (4096 trajectories x 64
    moves into a 60 000-sample buffer, config.yaml replay_buffer_max) and rb_gather of config.yaml's 512-sample minibatch."""
    from muzero_breakout_b200.replay_buffer import ReplayBuffer
    B, T, K, cap, mb = 4096, 64, 5, 60000, 512
    g = torch.Generator(device=dev).manual_seed(3 + rank)
    lens = torch.randint(40, T + 1, (B,), device=dev, generator=g)
    rec = dict(action=torch.randint(0, 3, (T, B), device=dev, generator=g), reward=torch.randint(-1, 2, (T, B), device=dev, generator=g).float(),
               value=torch.rand((T, B), device=dev, generator=g), visits=torch.randint(0, 51, (T, B, 3), device=dev, generator=g),
               frames=torch.rand((T, B, 1, 16, 20), device=dev, generator=g), recorded=torch.arange(T, device=dev)[:, None] < lens[None, :],
               initial_gray=torch.rand((B, 1, 16, 20), device=dev, generator=g))
    rb = ReplayBuffer(32, K, cap, 0.985, 512, device=dev, max_moves=261)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    app_ms = timed(lambda: rb.save_episode(rec), 10)
    moves = int(lens.sum().item())
    idxs = [torch.randperm(cap, device=dev, generator=g)[:mb] for _ in range(8)]
    it = [0]

    def gather():
        it[0] += 1
        return rb.minibatch(idxs[it[0] % 8])

    ga_ms = timed(gather, 50)
    app_bytes = (moves + B) * 1304 * 2 + moves * 8          # entries read + written (frame + scalars), targets
    ga_bytes = mb * (32 * 1280 * 2 + 32 * 8 + K * 40)       # 32 frames read + written per sample, actions, K-step rows
    peaks = measured_peaks()
    out = {"append": {"ms": app_ms, "trajectories": B, "moves": moves, "moves_per_s": moves / (app_ms * 1e-3), "GBps": app_bytes / app_ms / 1e6,
                      "what": "ReplayBuffer.save_episode: rb_len + rb_scan + rb_store + rb_sample (value targets), no host sync"},
           "minibatch": {"ms": ga_ms, "samples": mb, "samples_per_s": mb / (ga_ms * 1e-3), "GBps": ga_bytes / ga_ms / 1e6,
                         "frac_of_hbm_peak": ga_bytes / ga_ms / 1e6 / peaks["hbm"],
                         "what": "ReplayBuffer.minibatch(512 random indices): one rb_gather launch + 6 output allocations"},
           "buffer": {"samples": rb.length, "max_length": cap, "entry_ring_MB": rb._t["frame"].numel() * 4 / 1e6}}
    if cpu and rank == 0 and world == 1:
        from oracle.replay_oracle import ReplayOracle
        c = {k: v[:, :6].cpu().numpy() for k, v in rec.items() if k != "initial_gray"}
        ig = rec["initial_gray"][:6].cpu().numpy()
        orc = ReplayOracle(32, K, cap, 0.985, 512)
        t0 = time.perf_counter()
        n_moves = 0
        for b in range(6):
            n = int(lens[b])
            orc.save(ig[b], c["frames"][:n, b], c["action"][:n, b], c["reward"][:n, b], c["visits"][:n, b], c["value"][:n, b])
            n_moves += n
        t1 = time.perf_counter()
        ii = np.random.RandomState(0).randint(0, len(orc), mb)
        for f in ("past_actions", "states", "visit_counts", "future_actions", "rewards", "values"):
            orc.batch(f, ii)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"append_moves_per_s": n_moves / (t1 - t0), "minibatch_samples_per_s": mb / (t2 - t1), "cores": 1, "kind": "port",
                               "sample": f"oracle/replay_oracle.py (numpy restatement of replay_buffer.py): 6 trajectories / {n_moves} moves saved, one 512-sample minibatch"}
    return out


def bench_train_ends(dev, rank, world, cpu=True):
    """Loss and optimizer ends of the training step (first slice): mz_adam over the 42 205 081 parameters of
    the three networks (28 B per parameter: HBM-bound) and mz_loss on config.yaml's minibatch (512 x K=5 rows, launch-bound)."""
    from muzero_breakout_b200 import _lib
    from muzero_breakout_b200.train import loss_fn
    L = _lib.lib()
    n = 42_205_081
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    p = torch.randn(n, device=dev, generator=g) * 0.05
    grad = torch.randn(n, device=dev, generator=g) * 1e-3
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    stream = torch.cuda.current_stream(dev).cuda_stream
    step = [0]

    def adam():
        step[0] += 1
        _lib.check(L.mz_adam(n, p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), 2e-4, 0.9, 0.999, 1e-8, 1e-4, step[0], stream))

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    adam_ms = timed(adam, 20)
    B, K = 512, 5
    pr, pv = (torch.randn(B, K, 11, device=dev, generator=g) for _ in range(2))
    pp = torch.randn(B, K, 3, device=dev, generator=g)
    obs = torch.randint(-1, 2, (B, K), device=dev, generator=g).float()
    val = (torch.rand(B, K, device=dev, generator=g) - 0.5) * 20
    vis = torch.randint(1, 30, (B, K, 3), device=dev, generator=g).float()
    sup = torch.linspace(-5, 5, 11, device=dev)
    loss_ms = timed(lambda: loss_fn(obs, pr, val, pv, vis, pp, sup, K), 50)
    # weight gradient of one trunk convolution over the 5 x 512 (dY, X) pairs of a training step (the K unroll steps share their weights)
    nb = B * K
    xa, dya = torch.randn(nb, 4, 5, 256, device=dev, generator=g).bfloat16(), torch.randn(nb, 4, 5, 256, device=dev, generator=g).bfloat16()
    ns = L.mz_wgrad_padded_samples(nb)
    dy_t, x_t = torch.empty(256, 20, ns, dtype=torch.bfloat16, device=dev), torch.empty(256, 20, ns, dtype=torch.bfloat16, device=dev)
    partial = torch.empty(L.mz_wgrad_partial_bytes(3, nb) // 4, dtype=torch.float32, device=dev)
    dw = torch.empty(256, 256, 3, 3, device=dev)

    def wgrad():
        _lib.check(L.mz_wgrad_transpose(nb, 20, 256, dya.data_ptr(), dy_t.data_ptr(), stream))
        _lib.check(L.mz_wgrad_transpose(nb, 20, 256, xa.data_ptr(), x_t.data_ptr(), stream))
        _lib.check(L.mz_conv_wgrad(nb, 4, 5, 3, 1, dy_t.data_ptr(), x_t.data_ptr(), partial.data_ptr(), dw.data_ptr(), stream))

    wg_ms = timed(wgrad, 20)
    wg_flop = 130 * nb * 256 * 256 * 2
    # a 14-block residual trunk (the body of the dynamics / prediction network) as one training step at the minibatch size:
    # forward + backward of this library's kernels only, replayed as one CUDA graph
    from muzero_breakout_b200.train import ResidualBlockTrain, TrunkTrain
    gc = torch.Generator().manual_seed(11)
    mk = lambda *shape, s=1.0: torch.randn(*shape, generator=gc) * s
    trunk = TrunkTrain(ResidualBlockTrain(mk(256, 256, 3, 3, s=0.02), mk(256, s=0.01), torch.rand(256, generator=gc) + 0.5, mk(256, s=0.1),
                                          mk(256, 256, 3, 3, s=0.02), mk(256, s=0.01), torch.rand(256, generator=gc) + 0.5, mk(256, s=0.1), device=dev)
                       for _ in range(14))
    x16 = torch.rand(B, 4, 5, 256, device=dev, generator=g).bfloat16()
    dyf = torch.randn(B, 4, 5, 256, device=dev, generator=g)
    trunk_ms = timed(lambda: trunk.step(x16, dyf, graph=True), 10)
    peaks = measured_peaks()
    out = {"trunk_step": {"ms": trunk_ms, "samples": B, "blocks": 14,
                          "what": "train.TrunkTrain: forward + backward of 14 train-mode ResidualBlocks (tcgen05 conv / dgrad / wgrad + BatchNorm kernels), "
                                  "one CUDA-graph replay incl. the two input copies; torch + cuDNN on the same GPU: 8.0 ms TF32 / 6.6 ms autocast bf16 "
                                  "(profiles/r1_train_ends_timing.txt)"},
           "wgrad": {"ms": wg_ms, "samples": nb, "TFLOPs": wg_flop / wg_ms / 1e9, "frac_of_tensor_peak": wg_flop / wg_ms / 1e9 / peaks["bf16"],
                     "what": "3x3 256->256 weight gradient on tcgen05: 2 transposes + mz_conv_wgrad (9 taps x 8 K-splits of CTA pairs) + fixed-order split reduction; "
                             "in-bounds-tap FLOPs; peak = burst cuBLAS bf16 (a kernel timed alone)"},
           "adam": {"ms": adam_ms, "parameters": n, "GBps": n * 28 / adam_ms / 1e6, "frac_of_hbm_peak": n * 28 / adam_ms / 1e6 / peaks["hbm"],
                    "bytes_per_parameter": 28, "what": "mz_adam: one launch over the flat fp32 parameter / gradient / moment buffers (1.18 GB of traffic, larger than L2)"},
           "loss": {"ms": loss_ms, "rows": B * K, "what": "train.loss_fn: one mz_loss launch (3 KL divergences + total + gradients w.r.t. the logits) + output allocations"}}
    if cpu and rank == 0 and world == 1:
        from oracle import train_oracle as T
        ns = 2_000_000
        cp, cg = p[:ns].cpu().numpy(), grad[:ns].cpu().numpy()
        t0 = time.perf_counter()
        T.adam_step(cp, cg, np.zeros(ns, np.float32), np.zeros(ns, np.float32), 1)
        t1 = time.perf_counter()
        T.loss_fn(obs.cpu().numpy(), pr.cpu().numpy(), val.cpu().numpy(), pv.cpu().numpy(), vis.cpu().numpy(), pp.cpu().numpy(), sup.cpu().numpy(), K)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {"adam_parameters_per_s": ns / (t1 - t0), "loss_ms": (t2 - t1) * 1e3, "cores": 1, "kind": "port",
                               "sample": f"oracle/train_oracle.py (numpy restatement): one Adam update of {ns} parameters, one loss_fn of {B * K} rows"}
    return out


def bench_train_step(dev, minibatch=512, K=5):
    """One iteration of the reference's training loop body (train_torch.py:385-417: zero_grad, _k_step_rollout :487-528 = representation
    network + K x (prediction, dynamics), loss_fn :33-66, loss.backward(), optimizer.step()) at config.yaml's minibatch 512 and K = 5:
    the learner-side drop-in (muzero-breakout_b200/src/agent.py: every layer of the three networks forward + backward on this library's kernels --
    tcgen05 convolution / data / weight gradients, BatchNorm, pools, Linear heads, _scale_state -- plus mz_loss and mz_adam) against the unmodified reference modules + its own loss_fn +
    torch.optim.Adam under torch + cuDNN on the same GPU (TF32 and autocast bf16)."""
    from muzero_breakout_b200.src.agent import MuZeroAgent
    from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG
    from muzero_breakout_b200.train import loss_fn as lib_loss
    import muzero_breakout_b200 as mzb
    cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4, device="cuda")
    g = torch.Generator(device=dev).manual_seed(11)
    frames = torch.rand((minibatch, 64, 16, 20), device=dev, generator=g)
    actions = torch.randint(0, 3, (minibatch, K), device=dev, generator=g)
    obs_r = torch.randint(-1, 2, (minibatch, K), device=dev, generator=g).float()
    val_t = (torch.rand((minibatch, K), device=dev, generator=g) - 0.5) * 8
    visits = torch.randint(1, 30, (minibatch, K, 3), device=dev, generator=g).float()
    supports = torch.linspace(-5, 5, 11, device=dev)

    def rollout(agent):
        h = agent.create_hidden_state_root(frames)
        pol, val, rew = [], [], []
        for k in range(K):
            p_, v_ = agent.evaluate_state(h)
            planes = torch.nn.functional.one_hot(actions[:, k], 3).float().view(-1, 3, 1, 1).expand(-1, -1, 4, 5)
            h, r_ = agent.hidden_state_transition(h, planes)
            pol.append(p_); val.append(v_); rew.append(r_)
        return torch.stack(rew, 1), torch.stack(val, 1), torch.stack(pol, 1)

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            fn()
        b_.record(); torch.cuda.synchronize(dev)
        return a_.elapsed_time(b_) / reps

    torch.manual_seed(0)
    agent = MuZeroAgent(cfg)
    agent.train_mode()

    def lib_step():
        agent.optimizer.zero_grad()
        pr, pv, pp = rollout(agent)
        loss = lib_loss(obs_r, pr, val_t, pv, visits, pp, supports, K)[0]
        loss.backward()
        agent.optimizer.step()

    n0 = mzb.launch_count()
    lib_ms = timed(lib_step)
    launches_per_step = (mzb.launch_count() - n0) // 5
    # the same iteration as ONE CUDA-graph replay (train.GraphedTrainStep; what train.accelerate_training_stage puts under the reference's loop)
    from muzero_breakout_b200.train import GraphedTrainStep
    planes32 = frames[:, 32:]
    gstep = GraphedTrainStep(agent, supports, K)
    graph_ms = timed(lambda: gstep(frames[:, :32], planes32, actions, obs_r, val_t, visits))
    out = {"minibatch": minibatch, "K": K, "library_ms": graph_ms, "library_eager_ms": lib_ms, "library_kernel_launches_per_step": launches_per_step,
           "library_mode": "one CUDA-graph replay per loop iteration (train.GraphedTrainStep / accelerate_training_stage); library_eager_ms = the same kernels launched eagerly",
           "what": "zero_grad + _k_step_rollout + loss_fn + loss.backward() + optimizer.step() (train_torch.py:385-417); library = drop-in MuZeroAgent "
                   "(all layers of the three networks fwd + bwd on library kernels: tcgen05 conv / dgrad / wgrad with fp16 forward and bf16 gradient operands, "
                   "BatchNorm, pools, Linear heads, _scale_state; weight gradients batched over the K unroll steps; mz_loss; mz_adam)"}
    del agent
    torch.cuda.empty_cache()
    R = _reference()
    if R is None:
        out["reference"] = "unavailable: no reference checkout on this box"
        return out
    ref, _, rnets, rutils, _ = R
    import importlib
    rtrain = importlib.import_module("train_torch")            # the reference's loss_fn (import runs its set_seed(42) only)
    rcfg = ref.load_cfg()["model"]
    ragent = rnets.MuZeroAgent(rcfg)
    ragent.train_mode()
    st = rutils.ScalarTransforms(rcfg)

    def ref_step(autocast):
        ragent.optimizer.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            pr, pv, pp = rollout(ragent)
        loss = rtrain.loss_fn(observed_reward=obs_r, predicted_reward=pr.float(), bootstrapped_reward=val_t, predicted_value=pv.float(), visit_counts=visits,
                              predicted_policy=pp.float(), target_transformation=st.supports_representation, K=K)[0]
        loss.backward()
        ragent.optimizer.step()

    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    det = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        out["reference_tf32_ms"] = timed(lambda: ref_step(False))
        out["reference_autocast_bf16_ms"] = timed(lambda: ref_step(True))
        out["reference_cudnn_deterministic"] = bool(torch.backends.cudnn.deterministic)     # train_torch.py:22-23 sets it at import
        out["speedup_vs_best_reference_mode"] = min(out["reference_tf32_ms"], out["reference_autocast_bf16_ms"]) / out["library_ms"]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = det
    return out


def _reference():
    """(ref shim, src.mcts, src.networks, utils, environment.parallel_breakout) of the UNMODIFIED reference checkout (baseline/_ref: copied
    from /root/reference by __graft_entry__.build(), git-ignored, ships to the GPU box), or None when it is not there."""
    from baseline import ref
    if ref.ref_dir() is None:
        return None
    ref.install()
    import importlib
    return (ref,) + tuple(importlib.import_module(n) for n in ("src.mcts", "src.networks", "utils", "environment.parallel_breakout"))


def _use_all_cores():
    cores = os.cpu_count() or 1
    before = torch.get_num_threads()
    torch.set_num_threads(cores)
    return cores, before


def cpu_mcts(args, sd=None, trees=24, budget_s=25.0, max_calls=20, warmup=0):
    """The reference search on the host cores, bounded sample: `trees` roots x args.sims simulations per call (config.yaml's own n_parallel).
    kind "reference": the unmodified src/mcts.py MCTSSearchVec + src/networks.py MuZeroAgent (fp32, eval mode, no_grad) from baseline/_ref on
    every host thread; kind "port" (no checkout on this box): oracle/mcts_oracle.c tree + the fp32 torch restatement of the networks."""
    from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG, random_state_dict
    sd = sd or random_state_dict(DEFAULT_MODEL_CFG, seed=0, bn_jitter=0.2)
    cores, before = _use_all_cores()
    g = torch.Generator().manual_seed(1)
    hidden = torch.rand(trees, 256, 4, 5, generator=g)
    R = _reference()
    if R is not None:
        ref, rmcts, rnets, rutils, _ = R
        cfg = ref.load_cfg()
        cfg["num_simulations"] = args.sims
        agent = rnets.MuZeroAgent(cfg["model"])
        agent.load_state_dict(sd, strict=False)
        agent.eval()
        search_obj = rmcts.MCTSSearchVec(cfg, agent, rutils.ScalarTransforms(cfg["model"]))
        mask = torch.ones(trees, 3)

        def call(i):
            with torch.no_grad():
                v, n = search_obj.search(hidden, mask, 0)
            assert int(n.sum()) == trees * args.sims
        kind, what = "reference", "unmodified reference src/mcts.py MCTSSearchVec.search + src/networks.py MuZeroAgent (baseline/_ref, fp32 CPU, eval, no_grad)"
    else:
        import oracle
        from oracle.networks import OracleAgent
        agent = OracleAgent()
        agent.load_state_dict(sd, strict=False)
        agent.eval_mode()
        noise = torch.distributions.Dirichlet(torch.full((3,), 0.25)).sample((trees,))

        def call(i):
            oracle.search(agent, hidden, noise, seed=i, num_simulations=args.sims)
        kind, what = "port", "oracle/mcts_oracle.c tree + fp32 torch networks (oracle/networks.py); no reference checkout on this box"
    for i in range(warmup):
        call(i)
    n, t0 = 0, time.perf_counter()
    while True:
        call(n)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_calls:
            break
    return {"value": trees * args.sims * n / el, "unit": "simulations/s", "cores": cores, "kind": kind, "torch_threads": torch.get_num_threads(),
            "torch_threads_before_set": before, "calls": n,
            "sample": f"{n} search() calls of {trees} roots x {args.sims} simulations: {what} on {cores} threads, {el:.1f} s"}, el / n * 1e3


def cpu_env_reference(args, envs, budget_s=15.0, max_steps=200):
    """The unmodified reference environment (environment/parallel_breakout.py BreakoutEnvironment.step, torch ops on the host) at `envs`
    environments, random actions; None when there is no checkout on this box."""
    R = _reference()
    if R is None:
        return None
    ref, _, _, _, renv = R
    cores, before = _use_all_cores()
    cfg = ref.load_cfg()["environment"]
    cfg = dict(cfg, n_parallel=envs)
    env = renv.BreakoutEnvironment(cfg)
    torch.manual_seed(0)
    state, _ = env.reset()
    done = torch.zeros(envs, dtype=torch.bool)
    g = torch.Generator().manual_seed(1)
    acts = torch.randint(0, 3, (16, envs), generator=g)
    state, _, done, _ = env.step(state, acts[0], done)
    n, t0 = 0, time.perf_counter()
    while True:
        state, _, done, _ = env.step(state, acts[n % 16], done)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_steps:
            break
    return {"value": envs * n / el, "unit": "env-steps/s", "cores": cores, "kind": "reference", "torch_threads": torch.get_num_threads(),
            "sample": f"{n} steps of {envs} envs: unmodified reference environment/parallel_breakout.py BreakoutEnvironment.step (baseline/_ref, torch CPU ops) "
                      f"on {cores} threads, {el:.1f} s"}, el / n * 1e3


def gpu_reference(args, sd, dev, samples, our_step_ms):
    """The reference AS IT IS on the same GPU (SURVEY.md section 8d's bar for the network kernels): its own nn.Modules under torch + cuDNN,
    (a) one simulation step's networks (hidden_state_transition + evaluate_state) on `samples` leaves, fp32 (TF32 off / on) and autocast bf16, (b) the reference-as-written search (cuda networks + Python dict trees, src/mcts.py:24-71) at config.yaml's 24 roots."""
    R = _reference()
    if R is None:
        return {"unavailable": "no reference checkout on this box (baseline/_ref is copied by __graft_entry__.build() in the build container)"}
    ref, rmcts, rnets, rutils, _ = R
    cfg = ref.load_cfg()
    cfg["num_simulations"] = args.sims
    agent = rnets.MuZeroAgent(cfg["model"])           # hard-coded "cuda" = the current device
    agent.load_state_dict(sd, strict=False)
    agent.eval()
    g = torch.Generator(device=dev).manual_seed(3)
    h = torch.rand((samples, 256, 4, 5), device=dev, generator=g)
    a = torch.zeros((samples, 3, 4, 5), device=dev)
    a[torch.arange(samples, device=dev), torch.randint(0, 3, (samples,), device=dev, generator=g)] = 1.0

    def sim_step():
        with torch.no_grad():
            hs, r = agent.hidden_state_transition(h, a)
            p, v = agent.evaluate_state(hs)
        return p

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"samples": samples, "what": "unmodified reference MuZeroAgent.hidden_state_transition + evaluate_state (one simulation step's networks) under "
                                        "torch + cuDNN on this GPU, eval mode, no_grad; ms per simulation step", "ours_ms": our_step_ms}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        out["fp32_ms"] = timed(sim_step)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        out["tf32_ms"] = timed(sim_step)
        def sim_step_bf16():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                hs, r = agent.hidden_state_transition(h, a)
                p, v = agent.evaluate_state(hs)
            return p
        out["autocast_bf16_ms"] = timed(sim_step_bf16)
        # channels_last (the layout cuDNN's tensor-core kernels want) only for the weights: the reference's own .view() calls
        # (_scale_state, networks.py:318) reject channels_last activations, so the unmodified modules cannot run fully channels_last
        out["speedup_vs_best_reference_mode"] = min(out["fp32_ms"], out["tf32_ms"], out["autocast_bf16_ms"]) / our_step_ms
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    # the reference search as written, on the GPU: 24 roots (config.yaml n_parallel)
    search_obj = rmcts.MCTSSearchVec(cfg, agent, rutils.ScalarTransforms(dict(cfg["model"], device=str(dev))))
    h24, mask = h[:24].contiguous(), torch.ones(24, 3, device=dev)
    with torch.no_grad():
        search_obj.search(h24, mask, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            v, cnt = search_obj.search(h24, mask, 0)
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / n
    out["search_24_roots"] = {"ms_per_search": el * 1e3, "value": 24 * args.sims / el, "unit": "simulations/s",
                              "what": "unmodified reference MCTSSearchVec.search (cuda networks + Python dict trees, src/mcts.py:24-71), 24 roots x 50 simulations"}
    return out


def cpu_baseline_subprocess(primary, args):
    """`bench.py --impl reference` as a child process (it hides the GPUs before importing torch): ~25 s of CPU work for the MCTS sample,
    ~10 s for the env one."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "both" if (primary == "mcts" and args.workload == "both") else primary,
           "--sims", str(args.sims), "--trees", str(args.trees), "--envs", str(args.envs), "--steps", "50" if primary == "env" else "4",
           "--ref-budget", "25"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        return json.loads(lines[-1])
    except Exception as e:                     # noqa: BLE001
        return {"cpu_baseline": {"error": f"{type(e).__name__}: {e}"}}


def main():
    # stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version banner,
    # torch warnings) is sent to stderr by pointing fd 1 at fd 2 and keeping a private handle on the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps of the primary workload (default: mcts 5 searches, env 200 steps)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "mcts", "env"],
                    help="both: MCTS line with the env results nested under 'env' (default)")
    ap.add_argument("--envs", type=int, default=65536, help="env workload: environments per GPU")
    ap.add_argument("--env-steps", type=int, default=12000, help="env workload nested in the default line: timed steps (~0.5 s, so that the clock sampler sees them)")
    ap.add_argument("--trees", type=int, default=None, help="mcts workload: roots per GPU (default 4096 = BASELINE.json configs[2] on one GPU; "
                    "8192 under torchrun = configs[3]'s 65 536 roots over 8 GPUs)")
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--precision", default="f16", choices=["bf16", "f16", "f32"],
                    help="f16 (default): fp16 tensor-core operands, fp32 accumulation, fp16 + e4m3 residual stream -- the 16-bit mode that meets the north star's 1e-3")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--reset-every", type=int, default=32, help="env workload: start new games every this many steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: wall-clock bound of the timed CPU loop, seconds")
    ap.add_argument("--no-acting", action="store_true", help="skip the whole-move (rep net + search + env) aux measurement")
    ap.add_argument("--no-aux", action="store_true", help="skip the config.yaml-default (24 envs / roots) aux measurements")
    args = ap.parse_args()
    if args.trees is None:
        args.trees = 4096 if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) == 1 else 8192
    args.warmup = max(args.warmup, 3)
    primary = "env" if args.workload == "env" else "mcts"
    if args.steps is None:
        args.steps = 12000 if primary == "env" else 5

    if args.impl == "reference":
        # the reference's own CPU implementation of the path on this box's host cores (rank 0 only; the other ranks exit 0 without work)
        if int(os.environ.get("RANK", "0")) != 0:
            return
        budget = args.ref_budget                              # the whole arm ends within a few minutes whatever --steps says
        if primary == "env":
            envs = min(args.envs, 65536)
            res = cpu_env_reference(args, envs, budget_s=budget, max_steps=args.steps)
            if res is None:
                res = cpu_env(args, budget_s=20.0)
            base, ms = res
            metric, unit = "breakout_env_steps_per_s", "env-steps/s"
            wl = f"env: BreakoutEnvironment.step, reference-format outputs (fp32 (B,3,16,20) frames + reward + done + valid), {envs} envs, host cores"
            cfgd = {"workload": wl, "envs_per_gpu": envs}
        else:
            trees = 24
            base, ms = cpu_mcts(args, budget_s=budget, max_calls=args.steps, warmup=1)
            metric, unit = "latent_mcts_simulations_per_s", "simulations/s"
            wl = (f"mcts: MCTSSearchVec.search, {trees} roots x {args.sims} simulations per step (bounded sample: config.yaml's own n_parallel; the B200 arm runs "
                  f"{args.trees} roots per GPU), random-init MuZero networks (config.yaml sizes), fp32 on the host cores")
            cfgd = {"workload": wl, "trees_per_gpu": trees, "num_simulations": args.sims, "b200_arm_trees_per_gpu": args.trees}
        line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": unit,
                "n_gpus": args.gpus, "steps": base.get("calls", args.steps), "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfgd, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if args.workload == "both":
            res = cpu_env_reference(args, 65536, budget_s=30.0, max_steps=50) or cpu_env(args, budget_s=15.0)
            line["env"] = {"metric": "breakout_env_steps_per_s", "value": res[0]["value"], "unit": "env-steps/s", "cpu_baseline": res[0]}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    rank, local, world = dist_setup(args.gpus)
    sd = None
    if primary == "mcts":
        out, sd = bench_mcts(args, rank, local, world)
        if args.workload == "both":
            eargs = argparse.Namespace(**vars(args)); eargs.steps = args.env_steps; eargs.warmup = max(args.warmup, 10)
            out["env"] = bench_env(eargs, rank, local, world)
    else:
        out = bench_env(args, rank, local, world)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            # cpu_baseline = the reference arm itself (same code, GPUs hidden, every host core), run as a child process on a bounded sample
            cb = cpu_baseline_subprocess(primary, args)
            if primary == "mcts":
                out["cpu_baseline"] = cb.get("cpu_baseline", cb)
                if "env" in out and "env" in cb:
                    out["env"]["cpu_baseline"] = cb["env"]["cpu_baseline"]
            else:
                out["cpu_baseline"] = cb.get("cpu_baseline", cb)
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
