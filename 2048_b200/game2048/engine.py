"""Host side of the batched hot path: torch tensors as device buffers, libb2048.so for every kernel.

PyTorch is plumbing here (allocation, streams, host<->device copies, torch.distributed); all arithmetic
of the path runs in the CUDA kernels behind include/b2048.h.  There is no CPU fallback: constructing
a Context without a CUDA device raises.

Reference functions covered (file:line under /root/reference): game_logic.py:96-183 (board ops, moves,
spawns, greedy loop), r_learning.py:17-69 (features), :202-252 (evaluate / update / episode).
"""
import ctypes as C

import numpy as np
import torch

from . import cabi
from .cabi import B2048Error, check

I64, I32, U8, F32 = torch.int64, torch.int32, torch.uint8, torch.float32


def dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """Per-device immutable state: the 65,536-entry row LUT (create_table, game_logic.py:18-39)."""
    _cache = {}

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise B2048Error("no CUDA device visible: the b2048 hot path runs only on the GPU (no CPU fallback)")
        self.lib = cabi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            self.lut = torch.empty(cabi.LUT_ENTRIES, dtype=I32, device=self.device)
            check(self.lib.b2048_lut_build(dptr(self.lut), cur_stream()), "lut_build")

    @classmethod
    def get(cls, device=None):
        key = str(torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)) \
            if torch.cuda.is_available() else "none"
        ctx = cls._cache.get(key)
        if ctx is None:
            ctx = cls._cache[key] = cls(device)
        return ctx

    # ------------------------------------------------------------------ buffers
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def to_device(self, a, dtype=None, pinned=False):
        """numpy -> device tensor (uint64 travels as int64 bits, uint32 as int32, ...)."""
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        elif a.dtype == np.uint32:
            a = a.view(np.int32)
        elif a.dtype == np.uint16:
            a = a.view(np.int16)
        elif a.dtype == np.int8:
            a = a.view(np.uint8)
        t = torch.from_numpy(a)
        if pinned:
            t = t.pin_memory()
        t = t.to(self.device, non_blocking=pinned)
        return t if dtype is None else t.to(dtype)

    # ------------------------------------------------------------------ (1) pack / unpack
    def pack(self, rows):
        """int32 rows [m,4,4] (device tensor or numpy) -> int64-bit boards tensor [m] on device."""
        r = rows if torch.is_tensor(rows) else self.to_device(np.asarray(rows, dtype=np.int32))
        r = r.reshape(-1, 16).contiguous()
        out = self.empty(r.shape[0], I64)
        check(self.lib.b2048_pack(dptr(r), dptr(out), r.shape[0], cur_stream()), "pack")
        return out

    def unpack(self, boards):
        out = self.empty((boards.shape[0], 4, 4), I32)
        check(self.lib.b2048_unpack(dptr(boards), dptr(out), boards.shape[0], cur_stream()), "unpack")
        return out

    # ------------------------------------------------------------------ (2) moves
    def move4(self, boards, want_over=True):
        m = boards.shape[0]
        after, gain = self.empty((m, 4), I64), self.empty((m, 4), I32)
        flags = self.empty(m, U8)
        over = self.empty(m, U8) if want_over else None
        check(self.lib.b2048_move4(dptr(self.lut), dptr(boards), m, dptr(after), dptr(gain), dptr(flags), dptr(over),
                                   cur_stream()), "move4")
        return after, gain, flags, over

    def board_stats(self, boards, want_mask=True):
        m = boards.shape[0]
        stats = self.empty((m, 4), U8)
        mask = self.empty(m, torch.int16) if want_mask else None
        check(self.lib.b2048_board_stats(dptr(boards), m, dptr(stats), dptr(mask), cur_stream()), "board_stats")
        return stats, mask

    def sweep(self, boards, seed=0, first_index=0, spawn=True, out=None):
        m = boards.shape[0]
        if out is None:
            out = (self.empty((m, 4), I64), self.empty((m, 4), I32), self.empty(m, U8),
                   self.empty((m, 4), I64) if spawn else None)
        after, gain, flags, spawned = out
        check(self.lib.b2048_sweep(dptr(self.lut), dptr(boards), m, seed, first_index, dptr(after), dptr(gain),
                                   dptr(flags), dptr(spawned), cur_stream()), "sweep")
        return out

    # ------------------------------------------------------------------ (3) spawns
    def spawn_philox(self, boards, seed, game_id, move_no, want_spawn=False):
        m = boards.shape[0]
        sp = self.empty(m, torch.int16) if want_spawn else None
        check(self.lib.b2048_spawn_philox(dptr(boards), m, seed, dptr(game_id), dptr(move_no), dptr(sp), cur_stream()),
              "spawn_philox")
        return sp

    def spawn_initial(self, m, seed, first_id=0, id_step=1):
        boards = self.empty(m, I64)
        check(self.lib.b2048_spawn_initial(dptr(boards), m, seed, first_id, id_step, cur_stream()), "spawn_initial")
        return boards

    def spawn_replay(self, boards, tile, pos):
        check(self.lib.b2048_spawn_replay(dptr(boards), boards.shape[0], dptr(tile), dptr(pos), cur_stream()),
              "spawn_replay")

    # ------------------------------------------------------------------ (4) agent
    def features(self, n, boards):
        m = boards.shape[0]
        out = self.empty((m, cabi.num_feat(n)), I32)
        check(self.lib.b2048_features(n, dptr(boards), m, dptr(out), cur_stream()), "features")
        return out

    def evaluate(self, n, weights, boards):
        m = boards.shape[0]
        out = self.empty(m, F32)
        check(self.lib.b2048_evaluate(n, dptr(weights), dptr(boards), m, dptr(out), cur_stream()), "evaluate")
        return out

    def look_forward(self, n, weights, boards, game_id, move_no, root_dir, depth, width=1, since_empty=6, seed=0):
        """Game.look_forward (game_logic.py:214-243) values of m afterstates (device tensors: boards int64 bits,
        game_id int64, move_no int32, root_dir uint8) -> float32 [m]"""
        m = boards.shape[0]
        out = self.empty(m, F32)
        check(self.lib.b2048_look_forward(n, dptr(weights), dptr(self.lut), dptr(boards), dptr(game_id), dptr(move_no),
                                          dptr(root_dir), m, int(depth), int(width), int(since_empty), int(seed), dptr(out),
                                          cur_stream()), "look_forward")
        return out

    def update_workspace(self, n, m, mode):
        nbytes = self.lib.b2048_td_update_workspace(n, m, mode)
        return self.zeros(max(nbytes, 16), U8)

    def td_update(self, n, weights, boards, dw, mode=cabi.UPD_ATOMIC, work=None, delta=None):
        m = boards.shape[0]
        if work is None:
            work = self.update_workspace(n, m, mode)
        check(self.lib.b2048_td_update(n, dptr(weights), dptr(delta), dptr(boards), dptr(dw), m, mode, dptr(work),
                                       work.numel(), cur_stream()), "td_update")
        return work


def boards_to_numpy(t):
    return t.detach().cpu().numpy().view(np.uint64)


# ---------------------------------------------------------------------------------------------------
# weights: reference file layout (list of float32 arrays per signature group, r_learning.py:151-164)
# <-> one flat float32 device buffer in table order
# ---------------------------------------------------------------------------------------------------
SIGNATURE = {2: (24,), 3: (52,), 4: (17,), 5: (17, 4), 6: (17, 4, 12)}
GROUP_SIZE = {2: (256,), 3: (4096,), 4: (65536,), 5: (65536, 1048576), 6: (65536, 1048576, 14 ** 6)}


def flat_from_arrays(arrays):
    return np.concatenate([np.asarray(a, dtype=np.float32).reshape(-1) for a in arrays])


def arrays_from_flat(n, flat):
    out, o = [], 0
    for d, s in zip(SIGNATURE[n], GROUP_SIZE[n]):
        out.append(np.asarray(flat[o:o + d * s], dtype=np.float32).reshape(d, s).copy())
        o += d * s
    return out


# ---------------------------------------------------------------------------------------------------
# device-resident batch of game slots (b2048_games_t)
# ---------------------------------------------------------------------------------------------------
class GameBatch:
    def __init__(self, B, seed=0, id_stride=None, ctx=None, fin_cap=0):
        self.ctx = ctx or Context.get()
        self.B, self.seed = int(B), int(seed)
        self.id_stride = int(self.B if id_stride is None else id_stride)
        z = self.ctx.zeros
        self.board, self.score, self.moves = z(B, I64), z(B, I32), z(B, I32)
        self.game_id, self.state, self.old_label = z(B, I64), z(B, I64), z(B, F32)
        self.flags = z(B, U8)
        self.counters, self.tile_hist = z(cabi.CTR_COUNT, I64), z(17, I32)
        self.fin_cap = int(fin_cap)
        self.fin_log = z((self.fin_cap, 8), I32) if self.fin_cap else None
        self.c = cabi.Games(self.B, self.board.data_ptr(), self.score.data_ptr(), self.moves.data_ptr(),
                            self.game_id.data_ptr(), self.state.data_ptr(), self.old_label.data_ptr(),
                            self.flags.data_ptr(), self.counters.data_ptr(), self.tile_hist.data_ptr(),
                            self.seed, self.id_stride, self.fin_log.data_ptr() if self.fin_cap else None,
                            self.fin_cap)

    def init(self, first_id=0, reset_counters=True):
        """Game.__init__ for every slot (game_logic.py:55-70), Philox ids first_id + slot."""
        check(self.ctx.lib.b2048_games_init(C.byref(self.c), int(first_id), int(reset_counters), cur_stream()),
              "games_init")
        return self

    def set_positions(self, boards, scores=None):
        """adopt given boards (Game(row=...), game_logic.py:67-69); numpy uint64 or device tensor"""
        b = boards if torch.is_tensor(boards) else self.ctx.to_device(np.asarray(boards, dtype=np.uint64))
        self.board.copy_(b)
        self.score.zero_() if scores is None else self.score.copy_(self.ctx.to_device(np.asarray(scores, np.int32)))
        self.moves.zero_(); self.state.zero_(); self.old_label.zero_(); self.flags.zero_()
        return self

    def read_counters(self):
        c = self.counters.cpu().numpy()
        if c[cabi.CTR_FAULT]:
            raise B2048Error("a persistent training launch gave up at a grid barrier (B2048_CTR_FAULT): its results "
                             "and the update workspace are invalid")
        names = ["moves", "evals", "updates", "finished", "score_sum", "moves_sum", "overflow", "active"]
        return {k: int(v) for k, v in zip(names, c)}

    def drain_finished(self, strict=True):
        """finished-game records since the last drain, completion order: uint32 [k,8] =
        (id lo, id hi, score, moves, max exponent, board lo, board hi, 0).  The log holds fin_cap records; games
        that finished beyond that are counted by the kernels but not stored: with strict=True that raises (the
        caller's statistics would silently miss episodes), otherwise the number dropped is kept in `self.dropped`."""
        if not self.fin_cap:
            return np.zeros((0, 8), np.uint32)
        head = int(self.counters[cabi.CTR_LOG].item())
        k = min(head, self.fin_cap)
        rec = self.fin_log[:k].cpu().numpy().view(np.uint32).copy()
        self.counters[cabi.CTR_LOG] = 0
        self.dropped = head - k
        if self.dropped and strict:
            raise B2048Error(f"finished-game log overflow: {head} games finished since the last drain, fin_cap = "
                             f"{self.fin_cap}; drain more often or enlarge fin_cap")
        return rec

    def to_host(self):
        return dict(board=boards_to_numpy(self.board), score=self.score.cpu().numpy().view(np.uint32),
                    moves=self.moves.cpu().numpy().view(np.uint32), game_id=boards_to_numpy(self.game_id),
                    flags=self.flags.cpu().numpy(), state=boards_to_numpy(self.state),
                    old_label=self.old_label.cpu().numpy(), tile_hist=self.tile_hist.cpu().numpy())


class ReplayBuffers:
    """recorded spawns (Game.tiles, game_logic.py:121) as device arrays [B, len]; tile 0 = exhausted"""

    def __init__(self, ctx, tiles_per_game):
        B = len(tiles_per_game)
        L = max(1, max(len(t) for t in tiles_per_game) + 1)
        tile, pos = np.zeros((B, L), np.uint8), np.zeros((B, L), np.uint8)
        for j, t in enumerate(tiles_per_game):
            t = np.asarray(t, dtype=np.int64).reshape(-1, 3)
            tile[j, :len(t)] = t[:, 0]
            pos[j, :len(t)] = 4 * t[:, 1] + t[:, 2]
        self.tile, self.pos = ctx.to_device(tile), ctx.to_device(pos)
        self.c = cabi.Replay(self.tile.data_ptr(), self.pos.data_ptr(), L)
        self.len = L


def greedy_play(ctx, n, weights, games, limit_tile=0, step_limit=100000, chunk=2048, replay=None, trace_len=0,
                max_launches=None):
    """Game.trial_run (depth 0) for every slot until all are done.
    Returns (trace_dir, trace_value, trace_spawn) device tensors [B, trace_len] (None when trace_len == 0)."""
    tdir = ctx.empty((games.B, trace_len), torch.int8).fill_(-2) if trace_len else None
    tval = ctx.zeros((games.B, trace_len), F32) if trace_len else None
    tsp = ctx.zeros((games.B, trace_len), torch.int16) if trace_len else None
    rp = C.byref(replay.c) if replay is not None else None
    launches = 0
    while True:
        check(ctx.lib.b2048_greedy_play(n, dptr(weights), dptr(ctx.lut), C.byref(games.c), chunk, limit_tile, step_limit,
                                        rp, dptr(tdir), dptr(tval), dptr(tsp), trace_len, cur_stream()), "greedy_play")
        launches += 1
        active = int(games.counters[cabi.CTR_ACTIVE].item())
        if active == 0 or replay is not None or (max_launches and launches >= max_launches):
            break
    return tdir, tval, tsp


def expectimax_play(ctx, n, weights, games, depth, width=1, since_empty=6, limit_tile=0, step_limit=100000, chunk=256,
                    max_launches=None, trace_len=0):
    """Game.trial_run with look-ahead (depth / width / since_empty, game_logic.py:150-183, 214-243) for every slot
    until all are done; one warp per game, whole games on the device.
    Returns (trace_dir, None, trace_spawn) like greedy_play (None when trace_len == 0)."""
    tdir = ctx.empty((games.B, trace_len), torch.int8).fill_(-2) if trace_len else None
    tsp = ctx.zeros((games.B, trace_len), torch.int16) if trace_len else None
    launches = 0
    while True:
        check(ctx.lib.b2048_expectimax_play(n, dptr(weights), dptr(ctx.lut), C.byref(games.c), chunk, limit_tile, step_limit,
                                            int(depth), int(width), int(since_empty), dptr(tdir), dptr(tsp), trace_len,
                                            cur_stream()), "expectimax_play")
        launches += 1
        if int(games.counters[cabi.CTR_ACTIVE].item()) == 0 or (max_launches and launches >= max_launches):
            break
    return tdir, None, tsp


class TDTrainer:
    """Lock-step batched TD(0) on afterstates (QAgent.episode, r_learning.py:224-252) over a GameBatch."""

    def __init__(self, ctx, n, weights, games, alpha, mode, delta=None):
        self.ctx, self.n, self.w, self.games, self.alpha, self.mode = ctx, n, weights, games, float(alpha), int(mode)
        self.delta = delta
        self.upd_board, self.upd_dw = ctx.zeros(games.B, I64), ctx.zeros(games.B, F32)
        self.work = ctx.update_workspace(n, games.B, self.mode & 7)
        self.launches = 0                                        # kernels enqueued so far (bench: gpu_launches)

    def step(self, replay=None, trace=None):
        rp = C.byref(replay.c) if replay is not None else None
        td, tv, tw, ts, tl = (None, None, None, None, 0) if trace is None else trace
        self.launches += 2 if (self.mode & 7) == 0 else 3
        check(self.ctx.lib.b2048_td_step(self.n, dptr(self.w), dptr(self.delta), dptr(self.ctx.lut),
                                         C.byref(self.games.c), self.alpha, self.mode & 7, dptr(self.upd_board),
                                         dptr(self.upd_dw), dptr(self.work), self.work.numel(), rp, dptr(td), dptr(tv),
                                         dptr(tw), dptr(ts), tl, cur_stream()), "td_step")

    def phase_a(self):
        """the gather half alone (evaluate / argmax / commit / spawn); phase_b() applies the updates"""
        check(self.ctx.lib.b2048_td_phase_a(self.n, dptr(self.w), dptr(self.ctx.lut), C.byref(self.games.c), self.alpha,
                                            dptr(self.upd_board), dptr(self.upd_dw), None, None, None, None, None, 0,
                                            cur_stream()), "td_phase_a")

    def phase_b(self):
        check(self.ctx.lib.b2048_td_update(self.n, dptr(self.w), dptr(self.delta), dptr(self.upd_board),
                                           dptr(self.upd_dw), self.games.B, self.mode & 7, dptr(self.work),
                                           self.work.numel(), cur_stream()), "td_update")

    def run_peers(self, steps, peers, sync_every, since_sync, epoch):
        """`steps` lock-steps in ONE persistent launch with the multi-GPU weight exchange inside it (b2048_td_run_peers):
        a sync after every lock-step that completes a period of `sync_every`.  Returns False when this device / batch
        cannot take the persistent kernel (the caller then alternates run() and the stand-alone sync kernel)."""
        rc = self.ctx.lib.b2048_td_run_peers(self.n, dptr(self.w), dptr(self.ctx.lut), C.byref(self.games.c), self.alpha,
                                             self.mode & ~cabi.RUN_STEPWISE, int(steps), dptr(self.upd_board),
                                             dptr(self.upd_dw), dptr(self.work), self.work.numel(), C.byref(peers),
                                             int(sync_every), int(since_sync), int(epoch), cur_stream())
        if rc == -2:                                             # B2048_ENOTSUP
            return False
        check(rc, "td_run_peers")
        self.launches += 1
        return True

    def run(self, steps):
        """`steps` lock-steps without host work in between: one persistent cooperative kernel (default), or
        3 launches per lock-step with mode | RUN_STEPWISE"""
        self.launches += int(self.ctx.lib.b2048_td_run_launches(self.n, self.games.B, self.mode, int(steps)))
        check(self.ctx.lib.b2048_td_run(self.n, dptr(self.w), dptr(self.delta), dptr(self.ctx.lut),
                                        C.byref(self.games.c), self.alpha, self.mode, int(steps), dptr(self.upd_board),
                                        dptr(self.upd_dw), dptr(self.work), self.work.numel(), cur_stream()), "td_run")
