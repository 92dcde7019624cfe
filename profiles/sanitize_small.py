"""Smallest cases that exercise every kernel family once (for compute-sanitizer): python profiles/sanitize_small.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
import bench
ctx = engine.Context.get()
for n, B in ((4, 200), (5, 4500), (2, 64)):
    w = ctx.to_device(bench.seeded_weights(n))
    for mode in (cabi.UPD_ATOMIC | cabi.UPD_MEAN, cabi.UPD_ATOMIC | cabi.UPD_SUM, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN,
                 cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN | cabi.RUN_STEPWISE):
        g = engine.GameBatch(B, seed=1, ctx=ctx).init()
        tr = engine.TDTrainer(ctx, n, w, g, 0.25 if mode & 2 else 0.25 / B, mode)
        tr.run(12)
        tr.step()
    g = engine.GameBatch(min(B, 512), seed=2, ctx=ctx).init()
    engine.greedy_play(ctx, n, w, g, chunk=64, max_launches=2)
b = ctx.spawn_initial(4096, 3)
ctx.sweep(b, seed=0)
ctx.move4(b)
ctx.board_stats(b)
torch.cuda.synchronize()
print("ok")
