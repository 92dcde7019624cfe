"""Data parallelism over games (the only parallelism the path has): one process per GPU, game slots
sharded by contiguous global slot ranges, weight tables replicated, and -- for training only -- one exchange:
every `sync_every` lock-steps the per-rank weight movements delta_r = w_r - w_sync are combined with the per-key
mean over contributing ranks, w_sync += sum_r delta_r / max(1, #{r: delta_r != 0}), and every replica restarts
from w_sync.  Greedy play needs no collective at all (Philox streams are keyed by GLOBAL game id, so N ranks
reproduce the 1-rank games exactly); only the final per-game statistics are gathered.

Two implementations of the exchange (same formula; each leaves the replicas bit-identical):
  "p2p"   over NVLink / NVSwitch peer memory: the weights and w_sync of every rank are symmetric-memory allocations
          mapped into every process; rank r reduces slice r from remote loads and stores the result into every replica.
          No NCCL call, no message buffers.  One stand-alone kernel per sync (b2048_sync_peers: 40 us at n=4 on
          2 GPUs), or with `fused=True` INSIDE the persistent training launch (b2048_td_run_peers: run(S) is one kernel
          per rank however many syncs fall into it; bit-identical results, but measured 2 % slower over a bench step
          than the stand-alone kernel between launches -- 26.5 vs 26.0 ms per 2,048 lock-steps on 2 GPUs -- so it is
          not the default).
  "nccl"  b2048_delta_pack_bits -> allreduce(sum) of the float32 deltas + allgather of the one-bit-per-weight
          contributor planes (4.125 bytes per weight on the wire instead of the 8 of a float indicator) ->
          b2048_delta_apply_bits.  The portable path, and the one the CPU tier exercises under gloo.
"auto" takes "p2p" when torch's symmetric memory can map the peers, else "nccl".

The arithmetic lives behind a small `ops` object: CudaOps (libb2048.so) is the only production backend and is
what every caller gets by default.  tests/ inject an oracle-backed stand-in to exercise THIS file's sharding,
cadence and reduction logic with world_size-2 gloo on CPU; that stand-in is test code and never ships.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import cabi, engine


def shard(total, world, rank):
    """contiguous split of `total` units over `world` ranks: (first, count); the remainder goes to low ranks"""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def symmetric_memory_ok(device):
    """can torch map a (tiny) buffer of every rank into every process?  A collective: every rank must call it."""
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(64, dtype=torch.int32, device=device)
        h = symm.rendezvous(t, dist.group.WORLD.group_name)
        return len(h.buffer_ptrs) == dist.get_world_size()
    except Exception:                                                 # noqa: BLE001
        return False


def init_distributed(local_rank, peer_exchange=True):
    """torch.distributed over NCCL for this package's N > 1 paths: one process per GPU, NCCL for the plumbing
    (barriers, the initial broadcast, 32-byte reductions of counters and timings).

    Measured on B200 (profiles/r02_nccl_p2p_effect.txt): once NCCL has connected its NVLink P2P transport, every
    GPU-scope fence of the persistent training kernel gets slower (26.8 -> 28.5 ms per 2,048 lock-steps, whether or
    not a single NCCL collective runs in between), while peer mappings made through symmetric memory cost nothing.
    The weight exchange of this package runs over NVLink in its own kernels (b2048_sync_peers), so NCCL's P2P
    transport is switched off (NCCL_P2P_LEVEL=LOC: its few small collectives go through host shared memory) -- unless
    symmetric memory is unavailable, in which case NCCL is re-initialised with P2P on and carries the exchange.
    Returns {"p2p_exchange": bool, "nccl_p2p": bool}."""
    import os
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ours = peer_exchange and "NCCL_P2P_LEVEL" not in os.environ and "NCCL_P2P_DISABLE" not in os.environ
    if ours:
        os.environ["NCCL_P2P_LEVEL"] = "LOC"
    dist.init_process_group("nccl", device_id=dev)
    ok = symmetric_memory_ok(dev) if peer_exchange else False
    if ours and not ok:
        dist.destroy_process_group()
        del os.environ["NCCL_P2P_LEVEL"]
        dist.init_process_group("nccl", device_id=dev)
    return {"p2p_exchange": bool(ok), "nccl_p2p": not (ours and ok)}


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

    def zeros(self, count, dtype=torch.float32):
        return self.ctx.zeros(count, dtype)

    def trainer(self, n, w, delta, B, alpha, mode, seed, first_id, id_stride):
        games = engine.GameBatch(B, seed=seed, id_stride=id_stride, ctx=self.ctx).init(first_id=first_id)
        return engine.TDTrainer(self.ctx, n, w, games, alpha, mode, delta=delta)

    def run(self, trainer, steps):
        trainer.run(steps)

    def run_peers(self, trainer, steps, peers, sync_every, since_sync, epoch):
        return trainer.run_peers(steps, peers, sync_every, since_sync, epoch)

    def counters(self, trainer):
        return trainer.games.read_counters()

    # ---- "nccl" exchange
    def delta_pack_bits(self, w, w_sync, delta, bits):
        cabi.check(self.ctx.lib.b2048_delta_pack_bits(engine.dptr(w), engine.dptr(w_sync), engine.dptr(delta),
                                                      engine.dptr(bits), w.numel(), engine.cur_stream()), "delta_pack_bits")

    def delta_apply_bits(self, w, w_sync, delta_sum, bits_all, world):
        cabi.check(self.ctx.lib.b2048_delta_apply_bits(engine.dptr(w), engine.dptr(w_sync), engine.dptr(delta_sum),
                                                       engine.dptr(bits_all), int(world), w.numel(), engine.cur_stream()),
                   "delta_apply_bits")

    # ---- "p2p" exchange
    def peer_buffers(self, flat, group):
        """(w, w_sync, flags, peers struct) as symmetric-memory tensors mapped on every rank of `group`, or None when
        this torch / driver cannot map peer memory (the caller falls back to the NCCL exchange)"""
        try:
            import torch.distributed._symmetric_memory as symm
            name = (group or dist.group.WORLD).group_name
            nw = int(np.asarray(flat).size)
            bufs, handles = [], []
            for count, dtype in ((nw, torch.float32), (nw, torch.float32), (cabi.PEER_FLAG_WORDS, torch.int32)):
                t = symm.empty(count, dtype=dtype, device=self.device)
                handles.append(symm.rendezvous(t, name))
                bufs.append(t)
            w, w_sync, flags = bufs
            src = torch.from_numpy(np.asarray(flat, dtype=np.float32))
            w.copy_(src)
            w_sync.copy_(src)
            flags.zero_()
            peers = cabi.Peers()
            rank, world = handles[0].rank, handles[0].world_size
            if world > cabi.MAX_PEERS:
                return None
            for q in range(world):
                peers.w[q], peers.w_sync[q], peers.flags[q] = (int(h.buffer_ptrs[q]) for h in handles)
            peers.world, peers.rank = world, rank
            torch.cuda.synchronize()
            handles[0].barrier()                                       # every rank's buffers are initialised
            torch.cuda.synchronize()
            return w, w_sync, flags, peers, handles
        except Exception as e:                                        # noqa: BLE001 - any failure -> NCCL exchange
            self.peer_error = f"{type(e).__name__}: {e}"
            return None

    def sync_peers(self, peers, count, epoch):
        cabi.check(self.ctx.lib.b2048_sync_peers(C.byref(peers), int(count), int(epoch), 0, engine.cur_stream()),
                   "sync_peers")

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
    """lock-step TD over `games_per_rank` slots on every rank with periodic weight sync.  Every rank must pass the
    same arguments; the replicas start from rank 0's `weights_flat` (broadcast), and run() ends on a sync when
    `final_sync` is set, so that the replicas are identical whenever the caller looks at them."""

    def __init__(self, n, weights_flat, games_per_rank, alpha, mode, seed=0, sync_every=64, ops=None, group=None,
                 sync_impl="auto", first_slot=None, total_slots=None, fused=False):
        self.ops = ops or CudaOps()
        self.group = group
        self.rank, self.world = rank_world(group)
        self.n, self.sync_every = n, int(sync_every)
        self.B = int(games_per_rank)
        multi = self.world > 1
        self.peers = None
        self.sync_impl = "none"
        if multi and sync_impl in ("auto", "p2p") and hasattr(self.ops, "peer_buffers"):
            got = self.ops.peer_buffers(weights_flat, group)
            if got is None and sync_impl == "p2p":
                raise cabi.B2048Error("peer memory is not available: " + getattr(self.ops, "peer_error", "?"))
            if got is not None:
                self.w, self.w_sync, self.flags, self.peers, self._handles = got
                self.sync_impl = "p2p"
        if self.peers is None:
            self.w = self.ops.weights(weights_flat)
            self.w_sync = None
            if multi:
                self.sync_impl = "nccl"
        if multi:                    # replicas start from rank 0's tables whatever each rank was constructed with
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast(self.w, src=src, group=group)
            if self.w_sync is None:
                self.w_sync = self.w.clone()
            else:
                self.w_sync.copy_(self.w)
        if self.sync_impl == "nccl":
            nw = self.w.numel()
            self.words = (nw + 31) // 32
            self.delta = self.ops.zeros(nw)
            self.bits = self.ops.zeros(self.words, torch.int32)
            self.bits_all = self.ops.zeros(self.world * self.words, torch.int32)
        # global slot s = first_slot + local slot; a finished game's successor is id + total_slots
        first = self.rank * self.B if first_slot is None else int(first_slot)
        total = self.world * self.B if total_slots is None else int(total_slots)
        self.trainer = self.ops.trainer(n, self.w, None, self.B, alpha, mode, seed, first_id=first, id_stride=total)
        self.since_sync = 0
        self.syncs = 0               # exchanges so far (= the epoch of the last one)
        self.fused_syncs = 0         # ... of which inside a persistent training launch
        self.fused = bool(fused) and self.sync_impl == "p2p"

    @property
    def launches(self):
        """kernels of this package enqueued so far on this rank (trainer kernels + the stand-alone sync kernels)"""
        per_sync = 1 if self.sync_impl == "p2p" else 2
        return getattr(self.trainer, "launches", 0) + per_sync * (self.syncs - self.fused_syncs)

    @property
    def message_bytes(self):
        """bytes one rank contributes to one sync: p2p = remote loads + stores of its slice (2 x 4 B x 2 buffers per
        weight of the slice and peer); nccl = the allreduce + allgather payload"""
        if self.world == 1:
            return 0
        nw = self.w.numel()
        if self.sync_impl == "p2p":
            return 16 * (nw // self.world) * (self.world - 1)
        return 4 * nw + 4 * self.words

    def sync(self):
        """combine what every rank's weights moved by since the last sync (module docstring)"""
        if self.world == 1 or self.since_sync == 0:
            return
        self.syncs += 1
        if self.sync_impl == "p2p":
            self.ops.sync_peers(self.peers, self.w.numel(), self.syncs)
        else:
            self.ops.delta_pack_bits(self.w, self.w_sync, self.delta, self.bits)
            dist.all_reduce(self.delta, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_gather_into_tensor(self.bits_all, self.bits, group=self.group)
            self.ops.delta_apply_bits(self.w, self.w_sync, self.delta, self.bits_all, self.world)
        self.since_sync = 0

    def run(self, lock_steps, final_sync=False):
        done = 0
        if self.sync_impl == "p2p" and self.fused and lock_steps > 0 and not (self.trainer.mode & cabi.RUN_STEPWISE):
            # the whole call as ONE persistent launch per rank, the exchanges inside it
            if self.ops.run_peers(self.trainer, lock_steps, self.peers, self.sync_every, self.since_sync, self.syncs + 1):
                total = self.since_sync + lock_steps
                self.syncs += total // self.sync_every
                self.fused_syncs += total // self.sync_every
                self.since_sync = total % self.sync_every
                done = lock_steps
            else:
                self.fused = False
        while done < lock_steps:
            k = lock_steps - done
            if self.world > 1:
                k = min(k, self.sync_every - self.since_sync)
            self.ops.run(self.trainer, k)
            done += k
            self.since_sync += k
            if self.world > 1 and self.since_sync >= self.sync_every:
                self.sync()
        if final_sync:
            self.sync()

    def check_peer_fault(self):
        if self.sync_impl == "p2p" and int(self.flags[cabi.PEER_FAULT].item()) != 0:
            raise cabi.B2048Error("b2048_sync_peers: a peer never arrived; the replicas are not in sync")

    def replicas_identical(self):
        """True iff every rank holds bit-identical weights (a collective: every rank must call it)"""
        if self.world == 1:
            return True
        self.check_peer_fault()
        bits = self.w.view(torch.int32)
        mine = torch.stack([bits.sum(dtype=torch.int64), (bits.to(torch.int64) * 31 + 7).remainder(1000003).sum()])
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.group)
        return all(bool(torch.equal(e, every[0])) for e in every)

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
