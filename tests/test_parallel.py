"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: shard ranges, weight broadcast, trajectory
all-gather.  The GPU path is the same code over the nccl backend (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muzero_breakout_b200 import parallel


def test_shard_ranges_partition_the_batch():
    for total, world in ((65536, 8), (24, 8), (10, 3), (5, 8)):
        r = [parallel.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        ts = [torch.rand(7, 5, generator=g), torch.rand(3, generator=g).bfloat16(), torch.rand(11, generator=g), torch.rand(2, 2, generator=g).bfloat16()]
        want = None
        if rank == 1:
            want = [t.clone() for t in ts]
        nbytes = parallel.broadcast_tensors(ts, src=1)
        ref = [torch.empty_like(t) for t in ts]
        for i, t in enumerate(ts):           # every rank must now hold rank 1's values
            chk = t.clone().float()
            dist.broadcast(chk, src=1)
            assert torch.equal(chk, t.float())
        assert nbytes == (35 + 11) * 4 + (3 + 4) * 2
        # trajectory all-gather: equal and ragged shard sizes, global order = rank-major
        for total in (10, 7):
            lo, hi = parallel.shard_range(total, rank, world)
            B = hi - lo
            idx = torch.arange(lo, hi)
            gray = (idx.view(B, 1, 1, 1) % 2).float().expand(B, 1, 16, 20).contiguous()
            rec = parallel.pack_record(gray, idx % 3, idx.float() * 0.5, torch.stack([idx, idx + 1, 50 - 2 * idx - 1], 1), idx.float() / 7)
            allrec = parallel.all_gather_trajectory(rec)
            assert allrec.shape == (total, parallel.RECORD_FLOATS)
            if total % world == 0:           # the per-move form of an acting loop: no size exchange
                assert torch.equal(parallel.all_gather_trajectory(rec, equal_shards=True), allrec)
            g2, a2, r2, n2, v2 = parallel.unpack_record(allrec)
            full = torch.arange(total)
            assert torch.equal(a2, full % 3) and torch.equal(r2, full.float() * 0.5) and torch.equal(n2[:, 0], full)
            assert torch.equal(g2[:, 0, 0, 0], (full % 2).float()) and torch.allclose(v2, full.float() / 7)
        # asynchronous per-move exchange: 5 moves through 2 staging slots, results in submission order, equal to the blocking form
        ag = parallel.AsyncTrajectoryGather(4, 6, depth=2)
        want_moves = []
        for mv in range(5):
            loc = torch.arange(24, dtype=torch.float32).view(4, 6) + 100 * rank + 1000 * mv
            if mv % 2:
                ag.slot().copy_(loc); ag.submit()             # filled in place
            else:
                ag.submit(loc)
            want_moves.append(torch.cat([torch.arange(24, dtype=torch.float32).view(4, 6) + 100 * r + 1000 * mv for r in range(world)]))
        got = ag.drain()
        assert len(got) == 5 and all(torch.equal(g_, w_) for g_, w_ in zip(got, want_moves))
        assert ag.drain() == []
        # whole-episode exchange: ranks played different numbers of moves with different shard sizes
        lo, hi = parallel.shard_range(7, rank, world)
        B, T = hi - lo, 5 + 3 * rank

        def episode(lo, hi, T):
            e = torch.arange(lo, hi)
            mv = torch.arange(T)
            return dict(action=(mv[:, None] + e[None, :]) % 3, reward=(mv[:, None] * 10 + e[None, :]).float(),
                        value=(mv[:, None] - e[None, :]).float() / 4, visits=torch.stack([mv[:, None] + e[None, :]] * 3, -1),
                        frames=(mv[:, None, None, None, None] + e[None, :, None, None, None]).float().expand(T, hi - lo, 1, 16, 20).contiguous(),
                        recorded=mv[:, None] < (2 + e[None, :]), initial_gray=e.float().view(-1, 1, 1, 1).expand(hi - lo, 1, 16, 20).contiguous())

        merged = parallel.all_gather_episode(episode(lo, hi, T))
        Tm = 5 + 3 * (world - 1)
        for r in range(world):
            a, b = parallel.shard_range(7, r, world)
            want = episode(a, b, 5 + 3 * r)
            for k, v in want.items():
                if k == "initial_gray":
                    assert torch.equal(merged[k][a:b], v)
                else:
                    assert merged[k].shape[0] == Tm and merged[k].dtype == v.dtype
                    assert torch.equal(merged[k][:v.shape[0], a:b], v), k
                    assert not merged[k][v.shape[0]:, a:b].any(), k            # padding moves: zeros, recorded = False
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_allgather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
