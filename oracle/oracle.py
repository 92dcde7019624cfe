"""ctypes front-end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (2048_b200/) never does.  See oracle.c for the pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MAX_FEAT = 52
NUM_FEAT = {2: 24, 3: 52, 4: 17, 5: 21, 6: 33}


def build(force=False):
    """Compile oracle.c -> liboracle.so (gcc, OpenMP).  Idempotent."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_impl.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _declare(L):
    i32p, i64p, u8p, u32p, u64p = (C.POINTER(t) for t in (C.c_int32, C.c_int64, C.c_uint8, C.c_uint32, C.c_uint64))
    i8p = C.POINTER(C.c_int8)
    f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
    L.orc_create_table.argtypes = [u8p, u32p, u8p]
    L.orc_pre_move_batch.argtypes = [i32p, i64p, C.c_int64, i32p, i64p, i8p]
    L.orc_empty.argtypes = [i32p, i32p]
    L.orc_empty_count.argtypes = [i32p]
    L.orc_adjacent_pair_count.argtypes = [i32p]
    L.orc_game_over.argtypes = [i32p]
    L.orc_pack.argtypes = [i32p]
    L.orc_pack.restype = C.c_uint64
    L.orc_unpack.argtypes = [C.c_uint64, i32p]
    L.orc_features_batch.argtypes = [C.c_int, i32p, C.c_int64, i32p]
    L.orc_table_offset.argtypes = [C.c_int, C.c_int]
    L.orc_table_offset.restype = C.c_int64
    L.orc_table_size.argtypes = [C.c_int, C.c_int]
    L.orc_table_size.restype = C.c_int64
    L.orc_num_weights.argtypes = [C.c_int]
    L.orc_num_weights.restype = C.c_int64
    L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.orc_spawn_initial.argtypes = [C.c_uint64, C.c_uint64, i32p]
    L.orc_spawn_move.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, i32p]
    L.orc_spawn_sweep.argtypes = [C.c_uint64, C.c_uint64, C.c_int, i32p]
    L.orc_sweep.argtypes = [u64p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, u64p, u32p, u8p, u64p]
    L.orc_rot90.argtypes = [i32p, C.c_int, i32p]
    L.orc_transpose.argtypes = [i32p, i32p]
    for suf, rp, rt in (("f64", f64p, C.c_double), ("f32", f32p, C.c_float)):
        f = getattr(L, "orc_evaluate_" + suf)
        f.argtypes = [C.c_int, rp, i32p]
        f.restype = rt
        f = getattr(L, "orc_update_" + suf)
        f.argtypes = [C.c_int, rp, i32p, rt]
        f.restype = None
        f = getattr(L, "orc_update_keys_" + suf)
        f.argtypes = [C.c_int, i32p, i64p]
        f = getattr(L, "orc_update_batch_" + suf)
        f.argtypes = [C.c_int, rp, u64p, rp, C.c_int64, C.c_int]
        f.restype = C.c_int64
        f = getattr(L, "orc_episode_replay_" + suf)
        f.argtypes = [C.c_int, rp, rt, i32p, i32p, C.c_int, i32p, rp, rp, i32p, i64p, C.c_int]
        f = getattr(L, "orc_trial_replay_" + suf)
        f.argtypes = [C.c_int, rp, i32p, i32p, C.c_int, C.c_int, C.c_int, i32p, rp, i32p, i64p]
        f = getattr(L, "orc_play_philox_" + suf)
        f.argtypes = [C.c_int, rp, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int,
                      i64p, i32p, i32p, u64p, i64p]
        f.restype = C.c_int64
        f = getattr(L, "orc_look_forward_" + suf)
        f.argtypes = [C.c_int, rp, i32p, i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int64, i32p,
                      C.c_int64, C.c_uint64, u64p, u32p, i32p, rp]
        f = getattr(L, "orc_play_expectimax_" + suf)
        f.argtypes = [C.c_int, rp, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      i64p, i32p, u64p]
        f.restype = C.c_int64
        f = getattr(L, "orc_td_lockstep_" + suf)
        f.argtypes = [C.c_int, rp, rt, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_int, u64p, i64p, i32p, u64p, u64p, rp, u8p, i64p, i64p, i64p, i32p, i64p]
        f.restype = C.c_int64


def _rows(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a.reshape(-1, 16)


def _real(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(dtype)


# ---------------------------------------------------------------- game_logic.py restated
def create_table():
    lines = np.zeros((65536, 4), np.uint8)
    score = np.zeros(65536, np.uint32)
    changed = np.zeros(65536, np.uint8)
    lib().orc_create_table(_p(lines, C.c_uint8), _p(score, C.c_uint32), _p(changed, C.c_uint8))
    return lines, score, changed


def pre_move_batch(rows, scores=None):
    """rows [M,4,4] -> (new_rows [M,4,4,4] (direction axis 1), new_scores [M,4], change [M,4])."""
    r = _rows(rows)
    m = r.shape[0]
    sc = np.zeros(m, np.int64) if scores is None else np.ascontiguousarray(scores, dtype=np.int64)
    out = np.zeros((m, 4, 4, 4), np.int32)
    ns = np.zeros((m, 4), np.int64)
    ch = np.zeros((m, 4), np.int8)
    lib().orc_pre_move_batch(_p(r, C.c_int32), _p(sc, C.c_int64), m, _p(out, C.c_int32), _p(ns, C.c_int64),
                             _p(ch, C.c_int8))
    return out, ns, ch


def empty(row):
    r = _rows(row)[0].copy()
    pos = np.zeros(16, np.int32)
    m = lib().orc_empty(_p(r, C.c_int32), _p(pos, C.c_int32))
    return [(int(p) // 4, int(p) % 4) for p in pos[:m]]


def empty_count(row):
    r = _rows(row)[0].copy()
    return lib().orc_empty_count(_p(r, C.c_int32))


def adjacent_pair_count(row):
    r = _rows(row)[0].copy()
    return lib().orc_adjacent_pair_count(_p(r, C.c_int32))


def game_over(row):
    r = _rows(row)[0].copy()
    return bool(lib().orc_game_over(_p(r, C.c_int32)))


def pack(rows):
    r = _rows(rows)
    return np.array([lib().orc_pack(_p(r[i:i + 1].copy(), C.c_int32)) for i in range(r.shape[0])], dtype=np.uint64)


def pack_np(rows):
    """Vectorised pack (numpy) -- same layout as orc_pack; used for large test inputs."""
    r = _rows(rows).astype(np.uint64) & np.uint64(15)
    sh = (np.uint64(4) * (np.uint64(15) - np.arange(16, dtype=np.uint64)))
    return np.bitwise_or.reduce(r << sh, axis=1)


def unpack_np(boards):
    b = np.asarray(boards, dtype=np.uint64).reshape(-1, 1)
    sh = (np.uint64(4) * (np.uint64(15) - np.arange(16, dtype=np.uint64)))
    return ((b >> sh) & np.uint64(15)).astype(np.int32).reshape(-1, 4, 4)


def rot90(row, k):
    r = _rows(row)[0].copy()
    out = np.zeros(16, np.int32)
    lib().orc_rot90(_p(r, C.c_int32), k, _p(out, C.c_int32))
    return out.reshape(4, 4)


# ---------------------------------------------------------------- r_learning.py restated
def features_batch(n, rows):
    r = _rows(rows)
    out = np.zeros((r.shape[0], NUM_FEAT[n]), np.int32)
    lib().orc_features_batch(n, _p(r, C.c_int32), r.shape[0], _p(out, C.c_int32))
    return out


def table_offsets(n):
    return np.array([lib().orc_table_offset(n, i) for i in range(NUM_FEAT[n] + 1)], dtype=np.int64)


def num_weights(n):
    return int(lib().orc_num_weights(n))


def evaluate(n, w, row):
    suf, _ = _real(w.dtype)
    r = _rows(row)[0].copy()
    return getattr(lib(), "orc_evaluate_" + suf)(n, _p(w, _real(w.dtype)[1]), _p(r, C.c_int32))


def evaluate_batch(n, w, rows):
    r = _rows(rows)
    return np.array([evaluate(n, w, r[i]) for i in range(r.shape[0])], dtype=w.dtype)


def update(n, w, row, dw):
    suf, ct = _real(w.dtype)
    r = _rows(row)[0].copy()
    getattr(lib(), "orc_update_" + suf)(n, _p(w, ct), _p(r, C.c_int32), ct(dw))


def update_keys(n, row):
    r = _rows(row)[0].copy()
    keys = np.zeros(8 * NUM_FEAT[n], np.int64)
    m = lib().orc_update_keys_f64(n, _p(r, C.c_int32), _p(keys, C.c_int64))
    return keys[:m]


def update_batch(n, w, boards, dw, rule):
    """b2048_td_update restated: rule 0 sequential, 1 per-key sum then add, 2 per-key mean over entries."""
    suf, ct = _real(w.dtype)
    b = np.ascontiguousarray(boards, dtype=np.uint64)
    d = np.ascontiguousarray(dw, dtype=w.dtype)
    return getattr(lib(), "orc_update_batch_" + suf)(n, _p(w, ct), _p(b, C.c_uint64), _p(d, ct), len(b), rule)


def _tiles_array(tiles):
    """reference Game.tiles = [(tile, (i, j)), ...] -> int32 [K,3]"""
    if isinstance(tiles, np.ndarray):
        return np.ascontiguousarray(tiles, dtype=np.int32).reshape(-1, 3)
    return np.array([[t, p[0], p[1]] for t, p in tiles], dtype=np.int32).reshape(-1, 3)


def episode_replay(n, w, alpha, start, tiles, rule=0):
    """QAgent.episode teacher-forced on recorded spawns.  Mutates w.  Returns dict.
    rule 0 = the reference's sequential update; 1..4 = the batch rules of update_batch with m = 1."""
    suf, ct = _real(w.dtype)
    t = _tiles_array(tiles)
    k = t.shape[0]
    st = _rows(start)[0].copy()
    moves = np.zeros(k + 2, np.int32)
    values = np.zeros(k + 2, w.dtype)
    dws = np.zeros(k + 2, w.dtype)
    frow = np.zeros(16, np.int32)
    fscore = np.zeros(1, np.int64)
    odo = getattr(lib(), "orc_episode_replay_" + suf)(
        n, _p(w, ct), ct(alpha), _p(st, C.c_int32), _p(t, C.c_int32), k, _p(moves, C.c_int32),
        _p(values, ct), _p(dws, ct), _p(frow, C.c_int32), _p(fscore, C.c_int64), rule)
    if odo < 0:
        raise RuntimeError(f"orc_episode_replay failed: {odo}")
    return dict(odometer=odo, moves=moves[:odo + 1], values=values[:odo + 1], dws=dws[:odo + 1],
                row=frow.reshape(4, 4), score=int(fscore[0]))


def trial_replay(n, w, start, tiles, limit_tile=0, step_limit=100000):
    suf, ct = _real(w.dtype)
    t = _tiles_array(tiles)
    k = t.shape[0]
    st = _rows(start)[0].copy()
    moves = np.zeros(k + 2, np.int32)
    values = np.zeros(k + 2, w.dtype)
    frow = np.zeros(16, np.int32)
    fscore = np.zeros(1, np.int64)
    odo = getattr(lib(), "orc_trial_replay_" + suf)(
        n, _p(w, ct), _p(st, C.c_int32), _p(t, C.c_int32), k, limit_tile, step_limit, _p(moves, C.c_int32),
        _p(values, ct), _p(frow, C.c_int32), _p(fscore, C.c_int64))
    if odo < 0:
        raise RuntimeError(f"orc_trial_replay failed: {odo}")
    return dict(odometer=odo, moves=moves[:odo], values=values[:odo], row=frow.reshape(4, 4), score=int(fscore[0]))


def play_philox(n, w, seed, first_id, num, limit_tile=0, step_limit=100000, threads=0):
    suf, ct = _real(w.dtype)
    scores = np.zeros(num, np.int64)
    nm = np.zeros(num, np.int32)
    mt = np.zeros(num, np.int32)
    fb = np.zeros(num, np.uint64)
    ne = np.zeros(1, np.int64)
    total = getattr(lib(), "orc_play_philox_" + suf)(
        n, _p(w, ct), seed, first_id, num, limit_tile, step_limit, threads, _p(scores, C.c_int64),
        _p(nm, C.c_int32), _p(mt, C.c_int32), _p(fb, C.c_uint64), _p(ne, C.c_int64))
    return dict(total_moves=int(total), scores=scores, moves=nm, max_tile=mt, boards=fb, n_eval=int(ne[0]))


def look_forward(n, w, rows, scores, depth, width, since_empty, log=None, seed=0, ids=None, move_no=None,
                 root_dir=None):
    """Game.look_forward (game_logic.py:214-243) with estimator = evaluate, for m afterstates.
    log = (positions, tiles): replay the reference's logged random.sample / randrange results (depth-first order);
    otherwise the Philox node-keyed spec with per-afterstate ids / move_no / root_dir."""
    suf, ct = _real(w.dtype)
    r = _rows(rows)
    m = r.shape[0]
    sc = np.ascontiguousarray(scores, dtype=np.int64) if scores is not None else np.zeros(m, np.int64)
    out = np.zeros(m, w.dtype)
    if log is not None:
        lp, lt = (np.ascontiguousarray(x, dtype=np.int32) for x in log)
        rc = getattr(lib(), "orc_look_forward_" + suf)(n, _p(w, ct), _p(r, C.c_int32), _p(sc, C.c_int64), m, depth, width,
                                                       since_empty, 1, _p(lp, C.c_int32), len(lp), _p(lt, C.c_int32),
                                                       len(lt), 0, None, None, None, _p(out, ct))
    else:
        idv = np.ascontiguousarray(ids, dtype=np.uint64)
        mv = np.ascontiguousarray(move_no, dtype=np.uint32)
        rd = np.ascontiguousarray(root_dir, dtype=np.int32)
        rc = getattr(lib(), "orc_look_forward_" + suf)(n, _p(w, ct), _p(r, C.c_int32), _p(sc, C.c_int64), m, depth, width,
                                                       since_empty, 0, None, 0, None, 0, seed, _p(idv, C.c_uint64),
                                                       _p(mv, C.c_uint32), _p(rd, C.c_int32), _p(out, ct))
    if rc:
        raise RuntimeError(f"orc_look_forward failed: {rc}")
    return out


def play_expectimax(n, w, seed, first_id, num, depth, width, since_empty, limit_tile=0, step_limit=100000, threads=0):
    """trial_run with look-ahead (game_logic.py:150-183, 214-243) on the Philox streams, games [first_id, +num)"""
    suf, ct = _real(w.dtype)
    scores, nm, fb = np.zeros(num, np.int64), np.zeros(num, np.int32), np.zeros(num, np.uint64)
    total = getattr(lib(), "orc_play_expectimax_" + suf)(n, _p(w, ct), seed, first_id, num, depth, width, since_empty,
                                                         limit_tile, step_limit, threads, _p(scores, C.c_int64),
                                                         _p(nm, C.c_int32), _p(fb, C.c_uint64))
    return dict(total_moves=int(total), scores=scores, moves=nm, boards=fb)


class LockStep:
    """Host mirror of the lock-step batched TD state (see orc_td_lockstep)."""

    def __init__(self, n, w, alpha, seed, B, first_id=0, id_stride=None, segmented=0, threads=1):
        self.n, self.w, self.alpha, self.seed, self.B = n, w, alpha, seed, B
        self.first_id = first_id
        self.id_stride = B if id_stride is None else id_stride
        self.segmented, self.threads = segmented, threads
        self.board = np.zeros(B, np.uint64)
        self.score = np.zeros(B, np.int64)
        self.odo = np.zeros(B, np.int32)
        self.game_id = np.zeros(B, np.uint64)
        self.state = np.zeros(B, np.uint64)
        self.old_label = np.zeros(B, w.dtype)
        self.have_state = np.zeros(B, np.uint8)
        self.fin = np.zeros(3, np.int64)       # count, score sum, moves sum
        self.hist = np.zeros(17, np.int32)
        self.n_moves = 0
        self.n_updates = 0
        self._init = 1

    def run(self, steps):
        suf, ct = _real(self.w.dtype)
        nm = np.zeros(1, np.int64)
        fin = self.fin
        u = getattr(lib(), "orc_td_lockstep_" + suf)(
            self.n, _p(self.w, ct), ct(self.alpha), self.seed, self.first_id, self.id_stride, self.B, steps,
            self.segmented, self._init, self.threads, _p(self.board, C.c_uint64), _p(self.score, C.c_int64),
            _p(self.odo, C.c_int32), _p(self.game_id, C.c_uint64), _p(self.state, C.c_uint64),
            _p(self.old_label, ct), _p(self.have_state, C.c_uint8),
            _p(fin[0:1], C.c_int64), _p(fin[1:2], C.c_int64), _p(fin[2:3], C.c_int64), _p(self.hist, C.c_int32),
            _p(nm, C.c_int64))
        self._init = 0
        self.n_updates += int(u)
        self.n_moves += int(nm[0])
        return int(u)


# ---------------------------------------------------------------- Philox spawn stream (new spec)
def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def spawn_initial(seed, gid):
    row = np.zeros(16, np.int32)
    lib().orc_spawn_initial(seed, gid, _p(row, C.c_int32))
    return row.reshape(4, 4)


def spawn_move(seed, gid, move_no, row):
    r = _rows(row)[0].copy()
    res = lib().orc_spawn_move(seed, gid, move_no, _p(r, C.c_int32))
    return r.reshape(4, 4), res


def sweep(boards, seed=0, first_index=0, threads=0, spawn=True):
    b = np.ascontiguousarray(boards, dtype=np.uint64)
    m = b.shape[0]
    after = np.zeros((m, 4), np.uint64)
    gain = np.zeros((m, 4), np.uint32)
    flags = np.zeros(m, np.uint8)
    spawned = np.zeros((m, 4), np.uint64) if spawn else None
    lib().orc_sweep(_p(b, C.c_uint64), m, seed, first_index, threads, _p(after, C.c_uint64), _p(gain, C.c_uint32),
                    _p(flags, C.c_uint8), _p(spawned, C.c_uint64) if spawn else None)
    return after, gain, flags, spawned


def max_threads():
    return lib().orc_max_threads()
