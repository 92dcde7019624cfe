"""Small driver for profiling the look-ahead kernel: python profiles/run_expectimax.py --games 1024 --depth 3 --width 4"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=4)
p.add_argument("--games", type=int, default=1024)
p.add_argument("--depth", type=int, default=3)
p.add_argument("--width", type=int, default=4)
p.add_argument("--since-empty", type=int, default=6)
p.add_argument("--pretrain", type=int, default=3000)
p.add_argument("--chunk", type=int, default=64)
p.add_argument("--launches", type=int, default=4)
a = p.parse_args()
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
import bench
ctx = engine.Context.get()
wd = ctx.to_device(bench.seeded_weights(a.n))
g0 = engine.GameBatch(4096, seed=5, ctx=ctx).init()
engine.TDTrainer(ctx, a.n, wd, g0, 0.25, cabi.UPD_ATOMIC | cabi.UPD_MEAN).run(a.pretrain)
games = engine.GameBatch(a.games, seed=0, ctx=ctx).init()
engine.greedy_play(ctx, a.n, wd, games, chunk=600, max_launches=1)          # mid-game boards (crowded enough to look ahead)
games.counters.zero_()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
e[0].record()
engine.expectimax_play(ctx, a.n, wd, games, a.depth, a.width, a.since_empty, chunk=a.chunk, max_launches=a.launches)
e[1].record()
torch.cuda.synchronize()
c = games.read_counters()
ms = e[0].elapsed_time(e[1])
print(f"{c['moves']} moves, {c['evals']} leaf evaluations in {ms:.2f} ms = {c['moves'] / ms / 1e3:.3f} M moves/s, "
      f"{c['evals'] / ms / 1e6:.2f} G evaluations/s")
