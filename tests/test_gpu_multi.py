"""GPU tier, needs >= 2 GPUs on the box (skipped otherwise; run with `gpurun --gpus 2|4|8`): the real multi-process
weight exchange of 2048_b200/game2048/parallel.py -- one process per GPU, NCCL for the plumbing -- in both forms:
  "p2p"   the fused peer-memory kernel (b2048_sync_peers over symmetric memory / NVLink)
  "nccl"  pack_bits -> allreduce + allgather -> apply_bits
Deterministic update mode, so that the result is comparable bit for bit: every replica must equal the schedule emulated
in one process on the oracle with the rank-order reduction formula (tests/test_gpu_sync.py::oracle_schedule).  The p2p
kernel sums in rank order by construction; NCCL's order is its own, so for "nccl" the exact comparison is made at
world size 2 (a + b is commutative) and replicas-identical + 1e-6 closeness beyond."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, impl, n, B, periods):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    importlib.import_module("2048_b200")
    from game2048 import cabi, parallel
    from oracle import fixtures as fx
    w0 = fx.flat(fx.init_weights32(n, 17)).astype(np.float32)
    tr = parallel.ShardedTrainer(n, w0, B, alpha=0.25, mode=cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN, seed=21,
                                 sync_every=periods[0], sync_impl=impl)
    identical = []
    for steps in periods:
        tr.run(steps, final_sync=True)
        identical.append(tr.replicas_identical())
    c = tr.counters()
    np.savez(os.path.join(out_dir, f"{impl}_rank{rank}.npz"), w=tr.w.cpu().numpy(), w_sync=tr.w_sync.cpu().numpy(),
             impl=tr.sync_impl, identical=np.array(identical), updates=c["updates"], syncs=tr.syncs,
             why=str(getattr(tr.ops, "peer_error", "")))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("impl", ["p2p", "nccl"])
def test_sharded_trainer_on_real_gpus(tmp_path, orc, fx, impl):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_sync import oracle_schedule
    n, B, periods = 4, 48, (5, 5, 2)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), impl, n, B, periods), nprocs=world, join=True)
    r = [np.load(tmp_path / f"{impl}_rank{k}.npz") for k in range(world)]
    assert all(str(x["impl"]) == impl for x in r), [str(x["why"]) for x in r]
    assert all(x["identical"].all() for x in r) and all(int(x["syncs"]) == 3 for x in r)
    for x in r[1:]:
        assert np.array_equal(x["w"], r[0]["w"]) and np.array_equal(x["w_sync"], r[0]["w"])
    w0 = fx.flat(fx.init_weights32(n, 17)).astype(np.float32)
    ref_w, ref_updates = oracle_schedule(orc, n, w0, world, B, 0.25, 21, periods)
    assert int(r[0]["updates"]) == ref_updates
    if impl == "p2p" or world == 2:
        assert np.array_equal(r[0]["w"], ref_w)
    else:
        assert np.abs(r[0]["w"] - ref_w).max() <= 1e-6 * max(1.0, np.abs(ref_w).max())
