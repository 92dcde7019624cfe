"""Quality check of the batched trainer through the drop-in API: train n=4 for --episodes
episodes with --batch lock-step games, then QAgent.trial on --games greedy games; prints the ma_100 history and the
reach table the reference's README reports (README.md:79-121)."""
import argparse
import importlib
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=4)
p.add_argument("--episodes", type=int, default=100000)
p.add_argument("--batch", type=int, default=4096)
p.add_argument("--games", type=int, default=1000)
p.add_argument("--mode", default="atomic")
p.add_argument("--depth", type=int, default=0)
p.add_argument("--width", type=int, default=1)
p.add_argument("--since-empty", type=int, default=6)
p.add_argument("--look-games", type=int, default=100)
a = p.parse_args()
importlib.import_module("2048_b200")
from game2048 import r_learning as rl
random.seed(0)
np.random.seed(0)
lines = []
agent = rl.QAgent(name="q", storage="local", console="local", n=a.n, batch=a.batch, update_mode=a.mode)
agent.print = lines.append
t0 = time.time()
agent.train_run(num_eps=a.episodes, saving=False, chunk=256)
dt = time.time() - t0
h = agent.train_history
print(f"n={a.n} batch={a.batch} mode={a.mode}: {agent.step} episodes in {dt:.1f} s; ma_100 every 10k episodes: "
      f"{[h[i] for i in range(99, len(h), 100)]}; final alpha {agent.alpha}")
res = rl.QAgent.trial(estimator=agent.evaluate, num=a.games, storage="local", seed=1)
if a.depth > 0:
    t0 = time.time()
    res = rl.QAgent.trial(estimator=agent.evaluate, num=a.look_games, depth=a.depth, width=a.width, since_empty=a.since_empty,
                          storage="local", seed=2)
    print(f"look-ahead trial depth={a.depth} width={a.width} since_empty={a.since_empty}: {a.look_games} games in {time.time() - t0:.1f} s")
