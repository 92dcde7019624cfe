"""Data parallelism over games (the only parallelism the path has): one process per GPU, game slots
sharded by contiguous global slot ranges, weight tables replicated, and -- for training only -- one exchange:
every `sync_every` lock-steps the per-rank weight deltas (w - w_sync) are allreduced (NCCL over NVLink/NVSwitch through
torch.distributed) and applied with the per-key mean over contributing ranks (b2048_delta_pack / _apply).
Greedy play needs no collective at all (Philox streams are keyed by GLOBAL game id, so N ranks reproduce the
1-rank games exactly); only the final per-game statistics are gathered.

The arithmetic lives behind a small `ops` object: CudaOps (libb2048.so) is the only production backend and is
what every caller gets by default.  tests/ inject an oracle-backed stand-in to exercise THIS file's sharding,
cadence and reduction logic with world_size-2 gloo on CPU; that stand-in is test code and never ships.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import cabi, engine


def shard(total, world, rank):
    """contiguous split of `total` units over `world` ranks: (first, count); the remainder goes to low ranks"""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class CudaOps:
    """production backend: every call is a kernel launch through the C-ABI"""

    def __init__(self, ctx=None):
        self.ctx = ctx or engine.Context.get()
        self.device = self.ctx.device

    def weights(self, flat):
        return self.ctx.to_device(np.asarray(flat, dtype=np.float32))

    def zeros_like_weights(self, w, mult=1):
        return self.ctx.zeros(w.numel() * mult, torch.float32)

    def trainer(self, n, w, delta, B, alpha, mode, seed, first_id, id_stride):
        games = engine.GameBatch(B, seed=seed, id_stride=id_stride, ctx=self.ctx).init(first_id=first_id)
        return engine.TDTrainer(self.ctx, n, w, games, alpha, mode, delta=delta)

    def run(self, trainer, steps):
        trainer.run(steps)

    def counters(self, trainer):
        return trainer.games.read_counters()

    def delta_pack(self, w, w_sync, packed):
        """packed = [w - w_sync | (w != w_sync)]: what this rank's weights moved by since the last sync"""
        cabi.check(self.ctx.lib.b2048_delta_pack_diff(engine.dptr(w), engine.dptr(w_sync), engine.dptr(packed), w.numel(),
                                                      engine.cur_stream()), "delta_pack_diff")

    def delta_apply(self, w, w_sync, packed):
        n = w.numel()
        cabi.check(self.ctx.lib.b2048_delta_apply(engine.dptr(w), engine.dptr(w_sync), None, engine.dptr(packed),
                                                  engine.dptr(packed[n:]), n, engine.cur_stream()), "delta_apply")

    def greedy(self, n, w, seed, first_id, count, limit_tile=0):
        games = engine.GameBatch(count, seed=seed, ctx=self.ctx).init(first_id=first_id)
        engine.greedy_play(self.ctx, n, w, games, limit_tile=limit_tile)
        h = games.to_host()
        stats = np.stack([h["score"].astype(np.int64), h["moves"].astype(np.int64),
                          engine_max_tile(h["board"])], axis=1)
        return stats, h["board"]


def engine_max_tile(boards):
    b = np.asarray(boards, dtype=np.uint64).reshape(-1, 1)
    sh = (np.uint64(4) * np.arange(16, dtype=np.uint64))
    return ((b >> sh) & np.uint64(15)).max(axis=1).astype(np.int64)


class ShardedTrainer:
    """lock-step TD over `games_per_rank` slots on every rank (weak scaling) with periodic weight-delta sync"""

    def __init__(self, n, weights_flat, games_per_rank, alpha, mode, seed=0, sync_every=64, ops=None, group=None):
        self.ops = ops or CudaOps()
        self.group = group
        self.rank, self.world = rank_world(group)
        self.n, self.sync_every = n, int(sync_every)
        self.B = int(games_per_rank)
        self.w = self.ops.weights(weights_flat)
        multi = self.world > 1
        self.delta = None            # the kernels keep no second accumulator: delta = w - w_sync at sync time
        self.w_sync = self.w.clone() if multi else None
        self.packed = self.ops.zeros_like_weights(self.w, 2) if multi else None
        # global slot s = rank * B + local slot; a finished game's successor is id + world * B
        self.trainer = self.ops.trainer(n, self.w, self.delta, self.B, alpha, mode, seed, first_id=self.rank * self.B,
                                        id_stride=self.world * self.B)
        self.since_sync = 0
        self.syncs = 0

    @property
    def launches(self):
        """kernels of this package enqueued so far on this rank (trainer kernels + 2 per sync)"""
        return getattr(self.trainer, "launches", 0) + 2 * self.syncs

    def sync(self):
        """allreduce(sum) of [delta | touched indicator], then w_sync += sum / contributors on every rank"""
        if self.world == 1 or self.since_sync == 0:
            return
        self.ops.delta_pack(self.w, self.w_sync, self.packed)
        dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.group)
        self.ops.delta_apply(self.w, self.w_sync, self.packed)
        self.since_sync = 0
        self.syncs += 1

    def run(self, lock_steps):
        done = 0
        while done < lock_steps:
            k = lock_steps - done
            if self.world > 1:
                k = min(k, self.sync_every - self.since_sync)
            self.ops.run(self.trainer, k)
            done += k
            self.since_sync += k
            if self.world > 1 and self.since_sync >= self.sync_every:
                self.sync()

    def counters(self):
        """whole-job counters (summed over ranks)"""
        c = self.ops.counters(self.trainer)
        if self.world == 1:
            return c
        keys = sorted(c)
        t = torch.tensor([c[k] for k in keys], dtype=torch.int64, device=self.w.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return {k: int(v) for k, v in zip(keys, t.tolist())}


def greedy_sharded(n, weights_flat, total_games, seed=0, limit_tile=0, ops=None, group=None, gather=True):
    """play global game ids [0, total_games) split over the ranks; returns (stats [total,3] = score, moves, max
    exponent in global id order on every rank when gather=True, else this rank's shard)"""
    ops = ops or CudaOps()
    rank, world = rank_world(group)
    first, count = shard(total_games, world, rank)
    w = ops.weights(weights_flat)
    stats, boards = ops.greedy(n, w, seed, first, count, limit_tile)
    if world == 1 or not gather:
        return stats
    sizes = [shard(total_games, world, r)[1] for r in range(world)]
    pad = max(sizes)
    mine = torch.zeros((pad, 3), dtype=torch.int64, device=w.device)
    mine[:count] = torch.from_numpy(stats).to(w.device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])
