"""Generate tests/golden/lookforward.npz from the REAL reference: Game.look_forward (game_logic.py:214-243) on
crowded boards harvested from reference games, estimator = QAgent.evaluate, with every random.sample /
random.randrange result of the call logged in call (depth-first) order.

    python tests/golden/gen_lookforward.py        # needs /root/reference

The oracle replays the logged draws (oracle.look_forward(..., log=...)) and must reproduce the values; the whole
_find_best_move decision (direction chosen above the four look_forward values) is recorded as well.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures as fx  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
gl, rl = ref_shim.load()


class Logger:
    """wraps random.sample / random.randrange of the reference's module namespace"""

    def __init__(self):
        self.pos, self.tile = [], []

    def sample(self, population, k):
        out = random.sample(population, k)
        self.pos.extend(4 * i + j for (i, j) in out)
        return out

    def randrange(self, n):
        v = random.randrange(n)
        self.tile.append(1 if v else 2)
        return v


class RandomProxy:
    def __init__(self, log):
        self.log = log

    def __getattr__(self, name):
        if name == "sample":
            return self.log.sample
        if name == "randrange":
            return self.log.randrange
        return getattr(random, name)


def main():
    n = 4
    ref_shim.seed_all(123)
    w32 = fx.init_weights32(n, 77)
    # a slightly informative estimator: a few reference training episodes on top of the seeded init
    agent = ref_shim.make_agent(rl, n, [a.copy() for a in w32])
    for _ in range(30):
        agent.episode()
    weights = agent.list_to_np() if hasattr(agent, "list_to_np") else None
    w_flat = np.concatenate([np.asarray(t, dtype=np.float64) for t in agent.weights])
    # crowded boards from reference greedy games
    boards, scores = [], []
    while len(boards) < 40:
        g = gl.Game()
        for state, _ in g.generate_run(estimator=agent.evaluate, depth=0, width=1, since_empty=16):
            if g.empty_count(g.row) <= 5 and random.random() < 0.2:
                boards.append(g.row.copy()); scores.append(g.score)
            if len(boards) >= 40:
                break
    cases = []
    params = [(1, 1, 6), (1, 4, 8), (2, 2, 6), (2, 4, 8), (3, 3, 7), (3, 4, 6), (4, 2, 8), (2, 3, 3)]
    real_random = gl.random
    for ci, (row, score) in enumerate(zip(boards, scores)):
        depth, width, since_empty = params[ci % len(params)]
        g = gl.Game(row=row.copy(), score=int(score))
        for d in range(4):
            new_row, new_score, change = g.pre_move(g.row, g.score, d)
            if not change:
                continue
            log = Logger()
            gl.random = RandomProxy(log)
            try:
                v = g.look_forward(agent.evaluate, new_row, new_score, depth=depth, width=width, since_empty=since_empty)
            finally:
                gl.random = real_random
            cases.append(dict(row=np.asarray(new_row, np.int32).reshape(16), score=int(new_score), depth=depth, width=width,
                              since_empty=since_empty, value=float(v), pos=np.array(log.pos, np.int32),
                              tile=np.array(log.tile, np.int32)))
    rows = np.stack([c["row"] for c in cases])
    meta = np.array([[c["score"], c["depth"], c["width"], c["since_empty"]] for c in cases], np.int64)
    values = np.array([c["value"] for c in cases], np.float64)
    pos_off = np.cumsum([0] + [len(c["pos"]) for c in cases]).astype(np.int64)
    tile_off = np.cumsum([0] + [len(c["tile"]) for c in cases]).astype(np.int64)
    # weights = the seeded float32 init (fx.init_weights32(n, 77)) + the entries the 30 episodes changed (float64)
    init = fx.flat(w32).astype(np.float64)
    changed = np.flatnonzero(w_flat != init)
    np.savez_compressed(os.path.join(OUT, "lookforward.npz"), n=n, w_seed=77, w_idx=changed.astype(np.int64),
                        w_val=w_flat[changed], rows=rows, meta=meta,
                        values=values, pos=np.concatenate([c["pos"] for c in cases]), pos_off=pos_off,
                        tile=np.concatenate([c["tile"] for c in cases]), tile_off=tile_off)
    print(f"lookforward.npz: {len(cases)} look_forward calls, {pos_off[-1]} sampled positions, depth up to 4; "
          f"{os.path.getsize(os.path.join(OUT, 'lookforward.npz')) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
