"""Multi-GPU plumbing: one process per GPU, environments and trees sharded by batch index.

The acting path needs NO data-path collective: env i and tree i depend only on index i
(reference parallel_breakout.py:158-254 is per-row, src/mcts.py:149,207 loop per sample), so each rank owns
a contiguous index range and throughput scales weakly.  NCCL (torch.distributed, NVLink/NVSwitch) is used
only where the reference has an exchange step:

  broadcast_weights      target-network refresh: the reference copies the learner's state_dict into the
                         target network every 15 iterations (train_torch.py:137-138, 361-367); here the
                         learner rank broadcasts the PACKED weights (84 MB bf16) in one flat bucket per dtype.
  AsyncTrajectoryGather  the same exchange per move, asynchronous and double-buffered (the acting loop never waits for it)
  all_gather_trajectory  per-step trajectory records (gray frame, action, reward, visit counts, value --
                         what train_torch.py:204-208 appends to ObservationTrajectory) gathered from every
                         rank into the replay-buffer owner's tensor, ordered by global env index.

Works with any initialised process group (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist

RECORD_FLOATS = 320 + 1 + 1 + 3 + 1     # gray frame 16x20, action, reward, visit counts, value


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of global indices owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def packed_tensors(nets) -> List[torch.Tensor]:
    """Every device tensor a PackedNetworks holds, in a deterministic order."""
    return [getattr(obj, attr) for obj, attr in nets.holders()]


def broadcast_tensors(tensors: Iterable[torch.Tensor], src: int = 0, group=None) -> int:
    """Broadcast a list of tensors from `src`, coalesced into one flat bucket per dtype (few large
    messages: NVSwitch gives every peer full bandwidth, so bucket for launch latency, not link count).
    In place.  Returns the number of bytes broadcast."""
    tensors = list(tensors)
    total = 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        total += flat.numel() * flat.element_size()
    return total


def broadcast_weights(nets, src: int = 0, group=None) -> int:
    """Target-network refresh: one broadcast per weight arena (PackedNetworks keeps all packed tensors as views
    into one flat buffer per dtype), in place, no staging copies."""
    total = 0
    for arena in nets.arenas.values():
        dist.broadcast(arena, src=src, group=group)
        total += arena.numel() * arena.element_size()
    return total


def pack_record(gray: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, visits: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """(B,1,16,20) f32, (B,) i64, (B,) f32, (B,3) i64, (B,) f32 -> (B, RECORD_FLOATS) f32 (all values are small
    integers or floats exactly representable in fp32)."""
    B = gray.shape[0]
    return torch.cat([gray.reshape(B, -1).float(), action.reshape(B, 1).float(), reward.reshape(B, 1).float(),
                      visits.reshape(B, 3).float(), value.reshape(B, 1).float()], dim=1).contiguous()


def unpack_record(rec: torch.Tensor):
    g = rec[:, :320].reshape(-1, 1, 16, 20)
    return g, rec[:, 320].long(), rec[:, 321], rec[:, 322:325].long(), rec[:, 325]


def all_gather_trajectory(local: torch.Tensor, group=None, equal_shards: bool = False) -> torch.Tensor:
    """(B_local, F) from every rank -> (sum B_local, F), rank-major = global env index order for
    shard_range shards.  Ranks may hold different B_local; equal_shards=True asserts they do not and skips the size
    exchange (one collective and no host synchronisation per call: the per-move form of an acting loop)."""
    world = dist.get_world_size(group)
    if equal_shards:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(sizes)                                        # ragged shards: pad to the largest, gather, trim
    padded = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


class AsyncTrajectoryGather:
    """The per-move trajectory exchange WITHOUT stalling the acting loop.  A blocking all-gather after every search makes the ranks run in
    lock step: every move costs the slowest rank's search plus the collective (measured at 8 GPUs: 4.7 % of a step).  Nothing on the acting
    path reads the gathered records -- only the replay-buffer append at the end of an episode does -- so the collective can run on the
    process group's own stream while the next search is already under way: `submit(local)` copies the record into one of `depth` staging
    slots (waiting only for the collective that used the slot `depth` moves ago) and launches `all_gather_into_tensor(async_op=True)`;
    `drain()` waits for everything in flight and returns the gathered tensors in submission order.  Ranks may drift apart by up to `depth`
    moves; equal shard sizes on all ranks (the acting loop's case)."""

    def __init__(self, rows: int, cols: int = RECORD_FLOATS, device=None, dtype=torch.float32, depth: int = 2, group=None):
        self.group, self.depth = group, int(depth)
        world = dist.get_world_size(group)
        self.local = [torch.zeros((rows, cols), dtype=dtype, device=device) for _ in range(self.depth)]
        self.out = [torch.empty((world * rows, cols), dtype=dtype, device=device) for _ in range(self.depth)]
        self.work = [None] * self.depth
        self.ready = []                 # gathered tensors whose slot was recycled (cloned), in submission order
        self.pending = []               # slot indices in flight, in submission order
        self.n = 0

    def _retire(self, slot, clone):
        self.work[slot].wait()
        self.work[slot] = None
        self.pending.remove(slot)
        self.ready.append(self.out[slot].clone() if clone else self.out[slot])

    def slot(self) -> torch.Tensor:
        """the staging tensor of the next submit(), free to be filled in place (e.g. by slice assignment from the search's outputs)"""
        k = self.n % self.depth
        if self.work[k] is not None:
            self._retire(k, clone=True)
        return self.local[k]

    def submit(self, local: torch.Tensor | None = None):
        """local: this rank's (rows, cols) record, or None when slot() was filled in place"""
        k = self.n % self.depth
        buf = self.slot()
        if local is not None:
            buf.copy_(local)
        self.work[k] = dist.all_gather_into_tensor(self.out[k], buf, group=self.group, async_op=True)
        self.pending.append(k)
        self.n += 1

    def drain(self):
        for k in list(self.pending):
            self._retire(k, clone=False)
        out, self.ready = self.ready, []
        return out


EPISODE_FIELDS = ("action", "reward", "value", "visits", "frames", "recorded")


def all_gather_episode(record: dict, group=None) -> dict:
    """Whole-episode form of the trajectory exchange: every rank's acting.Actor.run_episode record (move-major
    (T_r, B_r, ...) tensors + initial_gray (B_r,1,16,20)) -> one record over all environments in global env order,
    ready for replay_buffer.ReplayBuffer.save_episode on the replay-buffer owner (reference analogue: the per-step
    appends of train_torch.py:204-208 followed by :223-225).  Ranks may have played different numbers of moves
    (their games ended at different times): shorter records are padded with recorded = False moves, which
    save_episode ignores."""
    dev = record["action"].device
    t = torch.tensor([record["action"].shape[0]], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    T = int(t.item())
    out = {}
    for k in EPISODE_FIELDS:
        x = record[k]
        if x.shape[0] < T:
            x = torch.cat([x, torch.zeros((T - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)], dim=0)
        flat = x.transpose(0, 1).contiguous()                                   # env-major so that ranks concatenate along dim 0
        if flat.dtype == torch.bool:
            flat = flat.to(torch.uint8)
        g = all_gather_trajectory(flat, group=group)
        out[k] = (g.bool() if record[k].dtype == torch.bool else g).transpose(0, 1).contiguous()
    out["initial_gray"] = all_gather_trajectory(record["initial_gray"].contiguous(), group=group)
    return out
