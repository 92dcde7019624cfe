"""Small driver for profiling the board sweep: python profiles/run_sweep.py --boards 16777216"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--boards", type=int, default=1 << 24)
p.add_argument("--reps", type=int, default=3)
a = p.parse_args()
import torch
importlib.import_module("2048_b200")
from game2048 import engine
ctx = engine.Context.get()
gen = torch.Generator(device=ctx.device).manual_seed(0)
parts = []
for i in range(0, a.boards, 1 << 22):
    k = min(1 << 22, a.boards - i)
    cells = torch.randint(1, 12, (k, 16), dtype=torch.int32, device=ctx.device, generator=gen)
    cells.mul_((torch.rand((k, 16), device=ctx.device, generator=gen) >= 0.3).to(torch.int32))
    parts.append(ctx.pack(cells))
boards = torch.cat(parts)
bufs = ctx.sweep(boards, seed=0)
for rep in range(a.reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    e[0].record()
    ctx.sweep(boards, seed=0, out=bufs)
    e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1])
    print(f"rep {rep}: {a.boards} boards in {ms:.3f} ms = {a.boards / ms / 1e6:.2f} G boards/s = {a.boards * 89 / ms / 1e6:.0f} GB/s algorithmic")
