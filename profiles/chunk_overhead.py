import importlib, os, sys, time
sys.path.insert(0, '/root/repo')
import torch
importlib.import_module("2048_b200")
from game2048 import cabi, engine
import bench
ctx = engine.Context.get()
wd = ctx.to_device(bench.seeded_weights(4))
games = engine.GameBatch(4096, seed=0, ctx=ctx).init()
delta = ctx.zeros(wd.numel(), torch.float32)
tr = engine.TDTrainer(ctx, 4, wd, games, 0.25, cabi.UPD_ATOMIC | cabi.UPD_MEAN, delta=delta)
tr.run(600)
def timed(fn):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); fn(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b), (t1 - t0) * 1e3
print("1 x 2048:", timed(lambda: tr.run(2048)))
print("32 x 64 :", timed(lambda: [tr.run(64) for _ in range(32)]))
packed = ctx.zeros(2 * wd.numel(), torch.float32); ws = wd.clone()
def chunk():
    tr.run(64)
    cabi.check(ctx.lib.b2048_delta_pack(engine.dptr(delta), engine.dptr(packed), delta.numel(), engine.cur_stream()))
    cabi.check(ctx.lib.b2048_delta_apply(engine.dptr(wd), engine.dptr(ws), engine.dptr(delta), engine.dptr(packed), engine.dptr(packed[wd.numel():]), wd.numel(), engine.cur_stream()))
print("32 x (64 + pack + apply):", timed(lambda: [chunk() for _ in range(32)]))
print("(gpu ms, cpu enqueue ms)")
