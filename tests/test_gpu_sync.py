"""GPU tier: the multi-GPU weight exchange on hardware, with R ranks emulated on ONE GPU (per-rank weight buffers
trained separately in the deterministic mode on their own game-id ranges, then combined).  All three forms of the
exchange must give, bit for bit, the formula the CPU tier checks under gloo (tests/test_parallel_gloo.py) evaluated
on the oracle's lock-step trainers:  w_sync += (sum over ranks of delta_r, rank order) / max(1, #{r: delta_r != 0}).
  (a) b2048_delta_pack_diff  -> sum -> b2048_delta_apply with the contributors half (float indicator, ABI v1)
  (b) b2048_delta_pack_bits  -> sum + stacked bit planes -> b2048_delta_apply_bits   (what the NCCL path runs)
  (c) b2048_sync_peers: R instances of the fused peer-memory kernel on R streams, every "peer" pointer local
      (what the p2p path runs; on the box the same pointers are NVLink mappings of the other GPUs' buffers)."""
import ctypes as C
import importlib
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, ROOT)
    importlib.import_module("2048_b200")
    from game2048 import cabi, engine
    return engine.Context.get(), engine, cabi


def oracle_schedule(orc, n, w0, R, B, alpha, seed, periods):
    """R shard trainers + the reduction formula, in one process (float32, exact per-key mean inside a rank)"""
    import torch
    ws = [w0.copy() for _ in range(R)]
    ls = [orc.LockStep(n, ws[k], alpha, seed, B, first_id=k * B, id_stride=R * B, segmented=4, threads=2)
          for k in range(R)]
    w_sync = torch.from_numpy(w0.copy())
    for steps in periods:
        deltas = []
        for k in range(R):
            ls[k].run(steps)
            deltas.append(torch.from_numpy(ws[k]) - w_sync)
        tot, cont = deltas[0].clone(), (deltas[0] != 0).float()
        for d in deltas[1:]:
            tot = tot + d
            cont = cont + (d != 0).float()
        w_sync = w_sync + tot / cont.clamp(min=1.0)
        for k in range(R):
            ws[k][:] = w_sync.numpy()
    return w_sync.numpy(), sum(l.n_updates for l in ls)


@pytest.mark.parametrize("R,n,B", [(2, 4, 64), (3, 4, 40), (4, 5, 24), (8, 2, 16)])
def test_sync_forms_emulated_ranks_vs_oracle(eng, orc, fx, R, n, B):
    import torch
    ctx, engine, cabi = eng
    lib = ctx.lib
    alpha, seed, periods = 0.25, 21, (5, 5, 2)
    w0 = fx.flat(fx.init_weights32(n, 17)).astype(np.float32)
    nw = len(w0)
    ref_w, ref_updates = oracle_schedule(orc, n, w0, R, B, alpha, seed, periods)
    mode = cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN
    words = (nw + 31) // 32

    def fresh():
        w = [ctx.to_device(w0) for _ in range(R)]
        ws = [ctx.to_device(w0) for _ in range(R)]
        tr = []
        for k in range(R):
            g = engine.GameBatch(B, seed=seed, id_stride=R * B, ctx=ctx).init(first_id=k * B)
            tr.append(engine.TDTrainer(ctx, n, w[k], g, alpha, mode))
        return w, ws, tr

    def finish(w, ws, tr, what):
        torch.cuda.synchronize()
        for k in range(R):
            assert np.array_equal(w[k].cpu().numpy(), ref_w), f"{what}: replica {k} differs from the oracle formula"
            assert np.array_equal(ws[k].cpu().numpy(), ref_w), f"{what}: w_sync of replica {k}"
        assert sum(t.games.read_counters()["updates"] for t in tr) == ref_updates

    # (a) float indicator
    w, ws, tr = fresh()
    packed = [ctx.zeros(2 * nw, torch.float32) for _ in range(R)]
    for steps in periods:
        for k in range(R):
            tr[k].run(steps)
            cabi.check(lib.b2048_delta_pack_diff(engine.dptr(w[k]), engine.dptr(ws[k]), engine.dptr(packed[k]), nw,
                                                 engine.cur_stream()))
        tot = packed[0].clone()
        for k in range(1, R):
            tot += packed[k]                                           # rank order, as the oracle formula
        for k in range(R):
            cabi.check(lib.b2048_delta_apply(engine.dptr(w[k]), engine.dptr(ws[k]), None, engine.dptr(tot),
                                             engine.dptr(tot[nw:]), nw, engine.cur_stream()))
    finish(w, ws, tr, "pack_diff/apply")

    # (b) one bit per weight
    w, ws, tr = fresh()
    delta = [ctx.zeros(nw, torch.float32) for _ in range(R)]
    bits_all = ctx.zeros(R * words, torch.int32)
    for steps in periods:
        for k in range(R):
            tr[k].run(steps)
            cabi.check(lib.b2048_delta_pack_bits(engine.dptr(w[k]), engine.dptr(ws[k]), engine.dptr(delta[k]),
                                                 engine.dptr(bits_all[k * words:]), nw, engine.cur_stream()))
        moved = torch.stack([(w[k] != ws[k]) for k in range(R)]).cpu().numpy()
        planes = np.unpackbits(bits_all.cpu().numpy().view(np.uint8).reshape(R, -1), axis=1, bitorder="little")
        assert np.array_equal(planes[:, :nw].astype(bool), moved) and not planes[:, nw:].any()
        tot = delta[0].clone()
        for k in range(1, R):
            tot += delta[k]
        for k in range(R):
            cabi.check(lib.b2048_delta_apply_bits(engine.dptr(w[k]), engine.dptr(ws[k]), engine.dptr(tot),
                                                  engine.dptr(bits_all), R, nw, engine.cur_stream()))
    finish(w, ws, tr, "pack_bits/apply_bits")

    # (c) fused peer-memory kernel: R instances in flight at once (they rendezvous through the flags)
    w, ws, tr = fresh()
    flags = [ctx.zeros(cabi.PEER_FLAG_WORDS, torch.int32) for _ in range(R)]
    peers = []
    for k in range(R):
        p = cabi.Peers()
        for q in range(R):
            p.w[q], p.w_sync[q], p.flags[q] = w[q].data_ptr(), ws[q].data_ptr(), flags[q].data_ptr()
        p.world, p.rank = R, k
        peers.append(p)
    streams = [torch.cuda.Stream() for _ in range(R)]
    for epoch, steps in enumerate(periods, start=1):
        for k in range(R):
            tr[k].run(steps)
        torch.cuda.synchronize()
        for k in range(R):
            with torch.cuda.stream(streams[k]):
                cabi.check(lib.b2048_sync_peers(C.byref(peers[k]), nw, epoch, 16, engine.cur_stream()))
        torch.cuda.synchronize()
        for k in range(R):
            assert int(flags[k][cabi.PEER_FAULT].item()) == 0
    finish(w, ws, tr, "sync_peers")

    # (d) the exchange inside the persistent training launch (b2048_td_run_peers): the 12 lock-steps of the schedule as
    #     ONE launch per rank (syncs after lock-steps 5 and 10 inside it), then the stand-alone kernel for the last 2
    w, ws, tr = fresh()
    flags = [ctx.zeros(cabi.PEER_FLAG_WORDS, torch.int32) for _ in range(R)]
    peers = []
    for k in range(R):
        p = cabi.Peers()
        for q in range(R):
            p.w[q], p.w_sync[q], p.flags[q] = w[q].data_ptr(), ws[q].data_ptr(), flags[q].data_ptr()
        p.world, p.rank = R, k
        peers.append(p)
    torch.cuda.synchronize()
    for k in range(R):
        with torch.cuda.stream(streams[k]):
            assert tr[k].run_peers(sum(periods), peers[k], periods[0], 0, 1)
    torch.cuda.synchronize()
    for k in range(R):
        with torch.cuda.stream(streams[k]):
            cabi.check(lib.b2048_sync_peers(C.byref(peers[k]), nw, 3, 16, engine.cur_stream()))
    torch.cuda.synchronize()
    for k in range(R):
        assert int(flags[k][cabi.PEER_FAULT].item()) == 0
    finish(w, ws, tr, "td_run_peers")


def test_sync_peers_argument_errors_and_ragged_count(eng):
    """count not a multiple of 4 (scalar tail on the last rank), world 1 (a no-op exchange), bad arguments"""
    import torch
    ctx, engine, cabi = eng
    lib = ctx.lib
    R, count = 3, 1003
    gen = torch.Generator(device=ctx.device).manual_seed(3)
    base = torch.rand(count + 1, device=ctx.device, generator=gen)[1:]     # odd offset: exercises the alignment check
    assert base.data_ptr() % 16 != 0
    ws = [torch.rand(count, device=ctx.device, generator=gen)] * 1
    ws = [ws[0].clone() for _ in range(R)]
    w = [ws[0].clone() for _ in range(R)]
    for k in range(R):
        w[k][k::5] += 0.25 * (k + 1)                                   # overlapping and exclusive movers
    want = ws[0].clone()
    tot = (w[0] - ws[0])
    cont = (w[0] != ws[0]).float()
    for k in range(1, R):
        tot = tot + (w[k] - ws[k])
        cont = cont + (w[k] != ws[k]).float()
    want = want + tot / cont.clamp(min=1.0)
    flags = [ctx.zeros(cabi.PEER_FLAG_WORDS, torch.int32) for _ in range(R)]
    peers = []
    for k in range(R):
        p = cabi.Peers()
        for q in range(R):
            p.w[q], p.w_sync[q], p.flags[q] = w[q].data_ptr(), ws[q].data_ptr(), flags[q].data_ptr()
        p.world, p.rank = R, k
        peers.append(p)
    streams = [torch.cuda.Stream() for _ in range(R)]
    torch.cuda.synchronize()
    for k in range(R):
        with torch.cuda.stream(streams[k]):
            cabi.check(lib.b2048_sync_peers(C.byref(peers[k]), count, 1, 8, engine.cur_stream()))
    torch.cuda.synchronize()
    for k in range(R):
        assert torch.equal(w[k], want) and torch.equal(ws[k], want)
    # world 1: w_sync catches up with w
    one = cabi.Peers()
    a, b, f = torch.rand(64, device=ctx.device), torch.zeros(64, device=ctx.device), ctx.zeros(cabi.PEER_FLAG_WORDS, torch.int32)
    one.w[0], one.w_sync[0], one.flags[0], one.world, one.rank = a.data_ptr(), b.data_ptr(), f.data_ptr(), 1, 0
    keep = a.clone()
    cabi.check(lib.b2048_sync_peers(C.byref(one), 64, 1, 0, engine.cur_stream()))
    torch.cuda.synchronize()
    assert torch.equal(a, keep) and torch.equal(b, keep)
    # errors
    bad = cabi.Peers()
    bad.world, bad.rank = 2, 0
    assert lib.b2048_sync_peers(C.byref(bad), 64, 1, 0, engine.cur_stream()) == -1          # NULL peers
    assert lib.b2048_sync_peers(C.byref(one), 64, 0, 0, engine.cur_stream()) == -1          # epoch 0
    one.w[0] = base.data_ptr()
    assert lib.b2048_sync_peers(C.byref(one), 64, 2, 0, engine.cur_stream()) == -1          # misaligned
