"""torchrun --nproc-per-node N profiles/sync_overhead.py : where the per-sync time of the multi-GPU trainer goes"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
importlib.import_module("2048_b200")
from game2048 import cabi, engine, parallel
import bench
st = parallel.ShardedTrainer(4, bench.seeded_weights(4), 4096, 0.25, cabi.UPD_ATOMIC | cabi.UPD_MEAN, seed=0, sync_every=64)
st.run(640)
def timed(fn, reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = {}
res["allreduce 8.9 MB fp32"] = timed(lambda: dist.all_reduce(st.packed), 50)
half = st.packed[:st.w.numel()]
res["allreduce 4.5 MB fp32"] = timed(lambda: dist.all_reduce(half), 50)
res["64 lock-steps (persistent launch)"] = timed(lambda: st.ops.run(st.trainer, 64), 20)
res["pack + apply kernels"] = timed(lambda: (st.ops.delta_pack(st.w, st.w_sync, st.packed), st.ops.delta_apply(st.w, st.w_sync, st.packed)), 20)
res["64 lock-steps + full sync"] = timed(lambda: st.run(64), 20)
if rank == 0:
    for k, v in res.items(): print(f"{k:40s} {v:9.1f} us")
dist.destroy_process_group()
