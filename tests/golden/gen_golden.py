"""Generate tests/golden/*.npz from the REAL reference (abachurin/2048) imported in this container.

    python tests/golden/gen_golden.py            # needs /root/reference; a few minutes, one core

Everything written here is produced by executing the unmodified reference code
(game2048/game_logic.py, game2048/r_learning.py) through oracle/ref_shim.py; the oracle and the CUDA
path are then checked against these files on machines where the reference is absent.
The reference never seeds its RNGs; this script seeds `random` and `np.random` explicitly.
"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures as fx  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (only for the Philox spawn spec in the lock-step fixture)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
gl, rl = ref_shim.load()
Game, QAgent = gl.Game, rl.QAgent
FEATS = {2: rl.f_2, 3: rl.f_3, 4: rl.f_4, 5: rl.f_5, 6: rl.f_6}


def save(name, **kw):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", flush=True)


def pack(row):
    return orc.pack_np(np.asarray(row).reshape(1, 4, 4))[0]


# ------------------------------------------------------------------ 1. move table (game_logic.py:18-39, :51)
def gen_table():
    lines = np.zeros((65536, 4), np.uint8)
    score = np.zeros(65536, np.uint32)
    changed = np.zeros(65536, np.uint8)
    for (a, b, c, d), (line, s, ch) in Game.table.items():
        k = (a << 12) | (b << 8) | (c << 4) | d
        lines[k] = line
        score[k] = s
        changed[k] = ch
    save("move_table.npz", lines=lines, score=score, changed=changed)


# ------------------------------------------------------------------ 2. board ops on seeded boards
def harvest_boards(num_games, seed):
    """boards met in seeded reference games under random play (realistic merge density)."""
    ref_shim.seed_all(seed)
    out = []
    for _ in range(num_games):
        g = Game()
        while not g.game_over(g.row):
            out.append(g.row.copy())
            d = random.randrange(4)
            if g.make_move(d):
                g.new_tile()
            g.moves.pop()
    return np.array(out, dtype=np.int32)


def gen_boards():
    boards = np.concatenate([fx.edge_boards(), fx.random_boards(12000, seed=5), harvest_boards(40, seed=6)])
    m = boards.shape[0]
    g = Game()
    after = np.zeros((m, 4, 4, 4), np.int8)
    gain = np.zeros((m, 4), np.int32)
    change = np.zeros((m, 4), np.uint8)
    over = np.zeros(m, np.uint8)
    n_empty = np.zeros(m, np.int8)
    n_pairs = np.zeros(m, np.int8)
    empties = -np.ones((m, 16), np.int8)
    for q in range(m):
        row = boards[q]
        for d in range(4):
            nr, ns, ch = g.pre_move(row, 1000, d)          # game_logic.py:136
            after[q, d] = nr
            gain[q, d] = ns - 1000
            change[q, d] = ch
        over[q] = g.game_over(row)                          # :109
        n_empty[q] = g.empty_count(row)                     # :101
        n_pairs[q] = g.adjacent_pair_count(row)             # :105
        em = g.empty(row)                                   # :96
        empties[q, :len(em)] = [4 * i + j for i, j in em]
    feats = {}
    mf = 3000
    for n, f in FEATS.items():
        feats[f"f_{n}"] = np.array([f(boards[q]) for q in range(mf)], dtype=np.int32)   # r_learning.py:17-69
    save("boards.npz", boards=boards.astype(np.int8), after=after, gain=gain, change=change, over=over,
         n_empty=n_empty, n_pairs=n_pairs, empties=empties, **feats)


# ------------------------------------------------------------------ 3. D4 order of update (r_learning.py:207-214)
def gen_d4():
    class Rec:
        def __init__(self):
            self.imgs = []

        def __call__(self, row):
            self.imgs.append(np.array(row).ravel().copy())
            return []

    agent = ref_shim.make_agent(rl, 2, weights32=[np.zeros((24, 256), np.float32)])
    rec = Rec()
    agent.features = rec
    agent.update(np.arange(16).reshape(4, 4), 1.0)
    imgs = np.array(rec.imgs, dtype=np.int8)
    # multiplicity of every (table, index) key of one update(), dw = 1, on a few boards, per n
    out = dict(images=imgs)
    sample = np.concatenate([fx.edge_boards(), fx.random_boards(20, seed=9)])
    out["sample"] = sample.astype(np.int8)
    for n in (2, 3, 4, 5):
        zeros = [np.zeros_like(a) for a in fx.init_weights32(n, 0)]
        keys, counts, offs = [], [], [0]
        for b in sample:
            ag = ref_shim.make_agent(rl, n, weights32=zeros)
            ag.update(b.copy(), 1.0)
            w = fx.flat_from_ref_lists(ag.weights)
            k = np.nonzero(w)[0]
            keys.append(k)
            counts.append(w[k].astype(np.int32))
            offs.append(offs[-1] + len(k))
        out[f"keys_{n}"] = np.concatenate(keys).astype(np.int64)
        out[f"counts_{n}"] = np.concatenate(counts)
        out[f"offs_{n}"] = np.array(offs, dtype=np.int64)
    save("d4.npz", **out)


# ------------------------------------------------------------------ 4./5. episodes and greedy games
class EventLog:
    """wraps agent.evaluate / agent.update on the INSTANCE: the reference's own methods still run."""

    def __init__(self, agent):
        self.kind, self.board, self.val = [], [], []
        ev, up = agent.evaluate, agent.update

        def evaluate(row, score=None):
            v = ev(row, score)
            self.kind.append(0); self.board.append(pack(row)); self.val.append(v)
            return v

        def update(row, dw):
            self.kind.append(1); self.board.append(pack(row)); self.val.append(dw)
            return up(row, dw)

        agent.evaluate, agent.update = evaluate, update

    def arrays(self):
        return (np.array(self.kind, np.uint8), np.array(self.board, np.uint64), np.array(self.val, np.float64))


def game_record(games):
    """concatenate reference Game records: starting_position, moves, tiles (game_logic.py:57-66)."""
    start = np.array([g.starting_position for g in games], dtype=np.int8)
    moves = np.concatenate([np.array(g.moves, dtype=np.int8) for g in games])
    tiles = np.concatenate([np.array([[t, p[0], p[1]] for t, p in g.tiles], dtype=np.int8).reshape(-1, 3)
                            for g in games])
    m_off = np.cumsum([0] + [len(g.moves) for g in games]).astype(np.int64)
    t_off = np.cumsum([0] + [len(g.tiles) for g in games]).astype(np.int64)
    final = np.array([g.row for g in games], dtype=np.int8)
    score = np.array([g.score for g in games], dtype=np.int64)
    odo = np.array([g.odometer for g in games], dtype=np.int32)
    return dict(start=start, moves=moves, tiles=tiles, m_off=m_off, t_off=t_off, final=final, score=score, odo=odo)


def gen_episodes(n, seed, episodes, alpha=0.25, log_events_for=2):
    w32 = fx.init_weights32(n, seed)
    base = fx.flat(w32).astype(np.float64)
    agent = ref_shim.make_agent(rl, n, weights32=w32, alpha=alpha)
    ref_shim.seed_all(seed + 1000)
    log = EventLog(agent)
    games, ev_counts, w_after_first = [], [], None
    t0 = time.time()
    for e in range(episodes):
        games.append(agent.episode())                       # r_learning.py:224
        ev_counts.append(len(log.kind))
        if e == 0:
            w_after_first = fx.flat_from_ref_lists(agent.weights)
    w_final = fx.flat_from_ref_lists(agent.weights)
    kind, board, val = log.arrays()
    n_ev = ev_counts[min(log_events_for, episodes) - 1]
    # per-step dw for every episode (the update events), to pin episode_replay over all episodes
    i1, v1 = fx.sparse_diff(base, w_after_first)
    i2, v2 = fx.sparse_diff(base, w_final)
    print(f"n={n}: {episodes} episodes, {sum(g.odometer for g in games)} moves, {time.time() - t0:.1f}s, "
          f"{len(i2)} weights touched", flush=True)
    save(f"episodes_n{n}.npz", n=n, seed=seed, alpha=alpha, ev_kind=kind[:n_ev], ev_board=board[:n_ev],
         ev_val=val[:n_ev], upd_dw=val[kind == 1], upd_board=board[kind == 1],
         w1_idx=i1, w1_val=v1, w_idx=i2, w_val=v2, **game_record(games))
    return agent


def gen_greedy(n, agent_trained, seed, num):
    """greedy trial_run games (game_logic.py:170-183) from float32-rounded trained weights."""
    base32 = fx.flat(fx.init_weights32(n, seed))
    w32 = fx.flat_from_ref_lists(agent_trained.weights).astype(np.float32)      # what save_agent would write (:156)
    idx, val = fx.sparse_diff(base32, w32)
    agent = ref_shim.make_agent(rl, n, weights32=fx.unflat(n, w32))
    ref_shim.seed_all(seed + 2000)
    games, values = [], []
    for _ in range(num):
        g = Game()
        vals = []

        def est(row, score, _vals=vals):
            v = agent.evaluate(row, score)
            _vals.append(v)
            return v

        g.trial_run(est)
        games.append(g)
        values.append(np.array(vals))
    print(f"greedy n={n}: {num} games, scores {[g.score for g in games]}", flush=True)
    v_off = np.cumsum([0] + [len(v) for v in values]).astype(np.int64)
    save(f"greedy_n{n}.npz", n=n, seed=seed, w_idx=idx, w_val=val, eval_values=np.concatenate(values), v_off=v_off,
         **game_record(games))


# ------------------------------------------------------------------ 6. lock-step TD built from reference objects
def gen_lockstep(n, seed, B, steps, alpha=0.25):
    """N reference Games + one reference QAgent; every lock-step all slots evaluate with
    W_t (Game.pre_move + QAgent.evaluate), then all updates are applied in slot order (QAgent.update).
    Spawns follow the Philox spawn spec (oracle.spawn_*), applied to the reference boards."""
    w32 = fx.init_weights32(n, seed)
    base = fx.flat(w32).astype(np.float64)
    agent = ref_shim.make_agent(rl, n, weights32=w32, alpha=alpha)
    pseed = 77
    games = [Game(row=orc.spawn_initial(pseed, j)) for j in range(B)]
    ids = list(range(B))
    state = [None] * B
    label = [0.0] * B
    fin = []
    n_upd = n_mv = 0
    for _ in range(steps):
        todo = []
        for j, g in enumerate(games):
            if g.game_over(g.row):
                if state[j] is not None:
                    todo.append((state[j], -label[j] * agent.alpha / agent.num_feat))     # r_learning.py:248
                fin.append((ids[j], g.score, g.odometer, int(np.max(g.row))))
                ids[j] += B
                games[j] = Game(row=orc.spawn_initial(pseed, ids[j]))
                state[j], label[j] = None, 0.0
                continue
            action, best_value, best_row, best_score = 0, -np.inf, None, None
            for d in range(4):                                                           # :231-237
                nr, ns, ch = g.pre_move(g.row, g.score, d)
                if ch:
                    v = agent.evaluate(nr)
                    if v > best_value:
                        action, best_value, best_row, best_score = d, v, nr, ns
            if state[j] is not None:
                todo.append((state[j], (best_score - g.score + best_value - label[j]) * agent.alpha / agent.num_feat))
            g.row, g.score = best_row, best_score
            g.odometer += 1
            state[j], label[j] = g.row.copy(), best_value
            g.row, _ = orc.spawn_move(pseed, ids[j], g.odometer, g.row)
            n_mv += 1
        for st, dw in todo:
            agent.update(st, dw)
            n_upd += 1
    w = fx.flat_from_ref_lists(agent.weights)
    idx, val = fx.sparse_diff(base, w)
    save(f"lockstep_n{n}.npz", n=n, seed=seed, pseed=pseed, B=B, steps=steps, alpha=alpha, w_idx=idx, w_val=val,
         boards=np.array([pack(g.row) for g in games], dtype=np.uint64),
         scores=np.array([g.score for g in games], dtype=np.int64), ids=np.array(ids, dtype=np.int64),
         labels=np.array(label), fin=np.array(fin, dtype=np.int64).reshape(-1, 4), n_upd=n_upd, n_mv=n_mv)


if __name__ == "__main__":
    t0 = time.time()
    if "--only-lockstep" in sys.argv:
        gen_lockstep(4, seed=24, B=16, steps=400, alpha=0.01)
        sys.exit(0)
    gen_table()
    gen_boards()
    gen_d4()
    gen_episodes(2, seed=12, episodes=6)
    gen_episodes(3, seed=13, episodes=6)
    trained = gen_episodes(4, seed=14, episodes=60)
    gen_greedy(4, trained, seed=14, num=8)
    gen_episodes(5, seed=15, episodes=4)
    # sum-rule lock-step is only stable for alpha << 1/B; keep the fixture in that regime
    gen_lockstep(4, seed=24, B=16, steps=400, alpha=0.01)
    if "--n6" in sys.argv or True:
        gen_episodes(6, seed=16, episodes=2)
    print(f"done in {time.time() - t0:.0f}s")
