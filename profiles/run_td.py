"""Small driver for profiling: python profiles/run_td.py --n 4 --games 4096 --mode atomic --rule mean --steps 200
Runs `--warm` lock-steps (so that games are desynchronised like in steady state), then `--steps` lock-steps."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=4)
p.add_argument("--games", type=int, default=4096)
p.add_argument("--mode", default="atomic")
p.add_argument("--rule", default="mean")
p.add_argument("--sorted", action="store_true")
p.add_argument("--warm", type=int, default=600)
p.add_argument("--steps", type=int, default=100)
p.add_argument("--split", action="store_true", help="time phase A / phase B separately with events")
p.add_argument("--layout", default="", help="force persistent-kernel variants: generic, scan, lists ('+'-joined)")
p.add_argument("--so", default="", help="load this libb2048.so instead of the in-tree one (lab builds)")
a = p.parse_args()
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
if a.so:
    cabi.SO_PATH = os.path.abspath(a.so)
import bench
ctx = engine.Context.get()
mode = (cabi.UPD_DETERMINISTIC if a.mode == "deterministic" else 0) | (cabi.UPD_MEAN if a.rule == "mean" else 0) | \
       (cabi.UPD_SORTED if a.sorted else 0)
for word in a.layout.split("+"):
    mode |= {"": 0, "generic": cabi.RUN_GENERIC, "scan": cabi.RUN_SCAN, "lists": cabi.RUN_LISTS, "even": cabi.RUN_EVEN}[word]
wd = ctx.to_device(bench.seeded_weights(a.n))
games = engine.GameBatch(a.games, seed=0, ctx=ctx).init()
tr = engine.TDTrainer(ctx, a.n, wd, games, 0.25 if a.rule == "mean" else 0.25 / a.games, mode)
tr.run(a.warm)
torch.cuda.synchronize()
c0 = games.read_counters()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
if a.split:
    ta = tb = 0.0
    for _ in range(a.steps):
        e[0].record(); tr.phase_a(); e[1].record(); tr.phase_b(); e[2].record()
        torch.cuda.synchronize()
        ta += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f"phase A {ta / a.steps * 1e3:.2f} us, phase B {tb / a.steps * 1e3:.2f} us per lock-step")
else:
    e[0].record(); tr.run(a.steps); e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1])
    c1 = games.read_counters()
    print(f"n={a.n} games={a.games} {a.mode} {a.rule} {a.layout or 'default'}: {a.steps} lock-steps: {ms / a.steps * 1e3:.2f} us each, "
          f"{(c1['updates'] - c0['updates']) / ms / 1e3:.1f} M updates/s")
