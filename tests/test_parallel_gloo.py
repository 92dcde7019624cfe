"""CPU tier: the N>1 host logic of 2048_b200/game2048/parallel.py under torch.distributed (gloo, world_size 2,
127.0.0.1): sharding by global game id, sync cadence, the [delta | contributors] allreduce and the per-key-mean
apply.  The CUDA kernels cannot run here, so an oracle-backed ops object (TEST CODE, below) stands in for
CudaOps; the same ShardedTrainer / greedy_sharded code runs on the GPUs with CudaOps."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleOps:
    """test double for parallel.CudaOps built on the CPU oracle (float32 rules == the device's)"""

    def __init__(self):
        from oracle import oracle as orc
        self.orc = orc
        self.device = torch.device("cpu")

    def weights(self, flat):
        return torch.from_numpy(np.array(flat, dtype=np.float32))

    def zeros(self, count, dtype=torch.float32):
        return torch.zeros(count, dtype=dtype)

    def trainer(self, n, w, delta, B, alpha, mode, seed, first_id, id_stride):
        rule = 4 if mode & 2 else 3                         # DETERMINISTIC | MEAN / SUM
        ls = self.orc.LockStep(n, w.numpy(), alpha, seed, B, first_id=first_id, id_stride=id_stride, segmented=rule,
                               threads=1)
        assert delta is None                                 # delta = w - w_sync at sync time, no second accumulator
        return ls

    def run(self, ls, steps):
        ls.run(steps)

    def counters(self, ls):
        return {"updates": ls.n_updates, "moves": ls.n_moves, "finished": int(ls.fin[0])}

    def delta_pack_bits(self, w, w_sync, delta, bits):
        """b2048_delta_pack_bits restated: float32 difference + one bit per weight, 32 per int32 word"""
        n = w.numel()
        delta.copy_(w - w_sync)
        moved = np.zeros(bits.numel() * 32, np.uint8)
        moved[:n] = (w != w_sync).numpy()
        bits.copy_(torch.from_numpy(np.packbits(moved, bitorder="little").view(np.int32)))

    def delta_apply_bits(self, w, w_sync, delta_sum, bits_all, world):
        n = w.numel()
        planes = np.unpackbits(bits_all.numpy().view(np.uint8).reshape(world, -1), axis=1, bitorder="little")[:, :n]
        c = torch.from_numpy(planes.sum(axis=0).astype(np.float32)).clamp(min=1.0)
        w_sync += delta_sum / c
        w.copy_(w_sync)

    def greedy(self, n, w, seed, first_id, count, limit_tile=0):
        r = self.orc.play_philox(n, w.numpy(), seed, first_id, count, limit_tile=limit_tile, threads=1)
        return np.stack([r["scores"], r["moves"].astype(np.int64), r["max_tile"].astype(np.int64)], axis=1), r["boards"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    importlib.import_module("2048_b200")
    from game2048 import parallel
    from oracle import fixtures as fx
    from test_parallel_gloo import OracleOps
    n, B = 2, 6
    w0 = fx.flat(fx.init_weights32(n, 4))
    # rank 1 is handed different tables on purpose: the constructor broadcasts rank 0's
    tr = parallel.ShardedTrainer(n, w0 if rank == 0 else w0[::-1].copy(), B, alpha=0.25, mode=3, seed=9, sync_every=5,
                                 ops=OracleOps())
    assert tr.sync_impl == "nccl" and tr.replicas_identical()
    tr.run(12)                                               # syncs after lock-steps 5 and 10, 2 steps pending
    mid = tr.w.clone()
    assert not tr.replicas_identical()
    tr.run(0, final_sync=True)                               # flush the pending delta
    assert tr.replicas_identical()
    c = tr.counters()
    stats = parallel.greedy_sharded(n, w0, 11, seed=3, ops=OracleOps())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=tr.w.numpy(), mid=mid.numpy(), syncs=tr.syncs,
             updates=c["updates"], stats=stats, first=tr.trainer.first_id, stride=tr.trainer.id_stride)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_training_and_greedy_world2(tmp_path, orc, fx):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    n, B = 2, 6
    # (1) replicas are bit-identical after every sync; 3 syncs happened (5, 10, final flush)
    assert np.array_equal(r[0]["w"], r[1]["w"]) and int(r[0]["syncs"]) == 3
    assert not np.array_equal(r[0]["mid"], r[1]["mid"])     # between syncs the replicas drift apart
    assert [int(x["first"]) for x in r] == [0, B] and all(int(x["stride"]) == world * B for x in r)
    # (2) the same schedule emulated in ONE process: two shard trainers + the same reduction formula
    w0 = fx.flat(fx.init_weights32(n, 4))
    ws = [w0.copy() for _ in range(world)]
    ls = [orc.LockStep(n, ws[k], 0.25, 9, B, first_id=k * B, id_stride=world * B, segmented=4, threads=1)
          for k in range(world)]
    w_sync = w0.copy()
    for steps in (5, 5, 2):
        deltas = []
        for k in range(world):
            before = ws[k].copy()
            ls[k].run(steps)
            deltas.append(torch.from_numpy(ws[k]) - torch.from_numpy(before))
        tot = deltas[0] + deltas[1]
        cont = ((deltas[0] != 0).float() + (deltas[1] != 0).float()).clamp(min=1.0)
        w_sync = (torch.from_numpy(w_sync) + tot / cont).numpy()
        for k in range(world):
            ws[k][:] = w_sync
    assert np.array_equal(r[0]["w"], w_sync)
    assert int(r[0]["updates"]) == ls[0].n_updates + ls[1].n_updates
    # (3) greedy play: the gathered result equals the unsharded run, in global id order, on every rank
    ref = orc.play_philox(n, w0, 3, 0, 11, threads=1)
    for x in r:
        assert np.array_equal(x["stats"][:, 0], ref["scores"]) and np.array_equal(x["stats"][:, 1], ref["moves"])


def test_shard_ranges():
    importlib.import_module("2048_b200")
    from game2048 import parallel
    for total, world in ((1000, 8), (1048576, 8), (7, 4), (3, 8)):
        parts = [parallel.shard(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
