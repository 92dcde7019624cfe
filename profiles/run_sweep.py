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
p.add_argument("--so", default="", help="load this libb2048.so instead of the in-tree one (lab builds)")
p.add_argument("--games", action="store_true", help="boards harvested from greedy n=4 games instead of iid cells")
a = p.parse_args()
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
if a.so:
    cabi.SO_PATH = os.path.abspath(a.so)
ctx = engine.Context.get()
gen = torch.Generator(device=ctx.device).manual_seed(0)
parts = []
for i in range(0, a.boards, 1 << 22):
    k = min(1 << 22, a.boards - i)
    cells = torch.randint(1, 12, (k, 16), dtype=torch.int32, device=ctx.device, generator=gen)
    cells.mul_((torch.rand((k, 16), device=ctx.device, generator=gen) >= 0.3).to(torch.int32))
    parts.append(ctx.pack(cells))
boards = torch.cat(parts)
if a.games:
    import bench
    wd = ctx.to_device(bench.seeded_weights(4))
    gh = engine.GameBatch(131072, seed=5, ctx=ctx).init()
    snaps = []
    for _ in range(max(1, a.boards // 131072)):
        engine.greedy_play(ctx, 4, wd, gh, chunk=8, max_launches=1)
        snaps.append(gh.board.clone())
    boards = torch.cat(snaps)
bufs = ctx.sweep(boards, seed=0)
import hashlib
print("output digest", hashlib.sha1(b"".join(t.cpu().numpy().tobytes() for t in bufs[:3])).hexdigest()[:16], a.so or "in-tree")
for rep in range(a.reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    e[0].record()
    ctx.sweep(boards, seed=0, out=bufs)
    e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1])
    print(f"rep {rep}: {a.boards} boards in {ms:.3f} ms = {a.boards / ms / 1e6:.2f} G boards/s = {a.boards * 89 / ms / 1e6:.0f} GB/s algorithmic")
