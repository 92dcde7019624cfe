"""Small driver for profiling the greedy kernel: python profiles/run_greedy.py --n 6 --games 131072 [--pretrain K]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=6)
p.add_argument("--games", type=int, default=131072)
p.add_argument("--pretrain", type=int, default=0)
p.add_argument("--chunk", type=int, default=4096)
p.add_argument("--reps", type=int, default=2)
p.add_argument("--so", default="", help="load this libb2048.so instead of the in-tree one (lab builds)")
a = p.parse_args()
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
if a.so:
    cabi.SO_PATH = os.path.abspath(a.so)
import bench
ctx = engine.Context.get()
wd = ctx.to_device(bench.seeded_weights(a.n))
if a.pretrain:
    g0 = engine.GameBatch(4096, seed=5, ctx=ctx).init()
    engine.TDTrainer(ctx, a.n, wd, g0, 0.25, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN).run(a.pretrain)
games = engine.GameBatch(a.games, seed=0, ctx=ctx)
for rep in range(a.reps):
    games.init()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    e[0].record()
    engine.greedy_play(ctx, a.n, wd, games, chunk=1 << 20)
    e[1].record()
    torch.cuda.synchronize()
    c = games.read_counters()
    ms = e[0].elapsed_time(e[1])
    print(f"n={a.n} games={a.games} pretrain={a.pretrain} {os.path.basename(os.path.dirname(a.so)) or 'in-tree'} rep {rep}: {c['moves']} moves in {ms:.2f} ms = {c['moves'] / ms / 1e6:.3f} G moves/s, "
          f"{c['evals'] / max(c['moves'], 1):.2f} evals/move, avg score {c['score_sum'] / max(c['finished'], 1):.0f}")
