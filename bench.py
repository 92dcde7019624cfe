#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native 2048 / n-tuple hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload td|greedy|sweep]

Workload `td` (default) is BASELINE.json configs[1]: Q_agent n=4 TD(0) training, 4,096 parallel seeded games
per GPU (weak scaling), per-key-mean lock-step rule, atomic update mode; one bench "step" = `--lock-steps`
lock-steps of all games.  metric = TD updates/s (one update = one QAgent.update() equivalent = 8*F weight
RMWs).  One JSON line on stdout (rank 0); see DESIGN.md "Measurement" for every field.

--impl reference times the CPU restatement of the reference algorithm (oracle/, C + OpenMP, all host
threads) on a bounded sample of the same workload; the Python reference itself cannot travel to the GPU box.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_OF_N = {2: 24, 3: 52, 4: 17, 5: 21, 6: 33}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="td", choices=["td", "greedy", "sweep", "expectimax"])
    p.add_argument("--depth", type=int, default=3, help="expectimax: look-ahead depth (the reference's best: 3)")
    p.add_argument("--width", type=int, default=4, help="expectimax: sampled tiles per node")
    p.add_argument("--since-empty", type=int, default=6, help="expectimax: look ahead only below this many empty cells")
    p.add_argument("--n", "--tuple", dest="n", type=int, default=4,
                   help="tuple size 2..6 (use --tuple under torchrun, whose own parser finds --n ambiguous)")
    p.add_argument("--games", type=int, default=4096, help="game slots per GPU")
    p.add_argument("--lock-steps", type=int, default=2048, help="lock-steps per bench step (td)")
    p.add_argument("--mode", default="atomic", choices=["atomic", "deterministic"])
    p.add_argument("--stepwise", action="store_true", help="3 launches per lock-step instead of the persistent kernel")
    p.add_argument("--rule", default="mean", choices=["mean", "sum"])
    p.add_argument("--alpha", type=float, default=0.25)
    p.add_argument("--sync-every", type=int, default=128, help="lock-steps between weight-delta allreduces (N>1)")
    p.add_argument("--boards", type=int, default=1 << 24, help="boards per GPU (sweep)")
    p.add_argument("--no-extras", action="store_true")
    p.add_argument("--chunk", type=int, default=4096, help="moves per greedy launch")
    p.add_argument("--pretrain", type=int, default=0, help="greedy: TD lock-steps (4096 games) before playing")
    p.add_argument("--cpu-games", type=int, default=4096, help="greedy: games of the cpu_baseline sample")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    return p.parse_args()


# --------------------------------------------------------------------------------------------- helpers
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 10 ms
    from a thread (nvidia-smi -lms cannot sample a region of a few hundred ms reliably)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.t_mark, self.err = index, [], False, 0.0, None
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.replace(",", "").isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                   # noqa: BLE001 - any NVML failure: report, don't die
            self.err = f"nvml unavailable: {e}"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:                                # noqa: BLE001 - older NVML name
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.time(), sm, r))
            except Exception as e:                               # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(0.01)

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.thread and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """start of the timed region: only samples taken after this call are reported"""
        self.t_mark = time.time()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"], "samples": 0}
        timed = [x for x in self.samples if x[0] >= self.t_mark] or self.samples[-3:]
        reasons = sorted({name for _, _, r in timed for name, bit in self.REASONS if r & bit})
        return {"sm_mhz": float(np.median([x[1] for x in timed])), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(timed)}


def seeded_weights(n, seed=0):
    """the reference's init (np.random.random/100, r_learning.py:136-149) from a seeded legacy stream, float32"""
    sig = {2: (24,), 3: (52,), 4: (17,), 5: (17, 4), 6: (17, 4, 12)}[n]
    size = {2: (256,), 3: (4096,), 4: (65536,), 5: (65536, 1048576), 6: (65536, 1048576, 14 ** 6)}[n]
    rs = np.random.RandomState(seed)
    return np.concatenate([(rs.random_sample((d, s)) / 100).astype(np.float32).reshape(-1) for d, s in zip(sig, size)])


def mode_bits(cabi, args):
    return (cabi.UPD_DETERMINISTIC if args.mode == "deterministic" else cabi.UPD_ATOMIC) | \
           (cabi.UPD_MEAN if args.rule == "mean" else cabi.UPD_SUM) | (cabi.RUN_STEPWISE if args.stepwise else 0)


def ncu_traffic(key, field):
    """bytes per unit of work measured by ncu (profiles/ncu_traffic.json), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)[key][field])
    except (OSError, KeyError, ValueError):
        return None


def launches_per_step_of(st, S):
    """kernel launches one bench step (S lock-steps) costs on this rank, sync kernels excluded"""
    from game2048 import cabi
    return int(cabi.lib().b2048_td_run_launches(st.n, st.B, st.trainer.mode, S))


def bytes_per_update(n, evals_per_move):
    F = F_OF_N[n]
    return 16 + 4 * F * evals_per_move + 64 * F          # board in/out + gathers + 8F RMWs x (4+4) B


# --------------------------------------------------------------------------------------------- reference arm
def cpu_td_sample(args, seconds, threads=0):
    """oracle lock-step TD (float32, per-key-mean rule == the GPU's semantics) on the host cores"""
    from oracle import oracle as orc
    orc.build()
    threads = threads or orc.max_threads()
    rule = 2 if args.rule == "mean" else 1
    w = seeded_weights(args.n)
    ls = orc.LockStep(args.n, w, args.alpha, 0, args.games, segmented=rule, threads=threads)
    ls.run(2)                                            # touch memory
    u0, t0, steps = ls.n_updates, time.perf_counter(), 0
    chunk = 4
    while True:
        ls.run(chunk)
        steps += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
        chunk = max(1, min(64, int(chunk * seconds / max(dt, 1e-3) / 4)))
    return (ls.n_updates - u0) / dt, threads, f"{steps} lock-steps of {args.games} games, n={args.n} ({dt:.1f} s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    rule = 2 if args.rule == "mean" else 1
    w = seeded_weights(args.n)
    ls = orc.LockStep(args.n, w, args.alpha, 0, args.games, segmented=rule, threads=threads)
    per_step = max(1, min(args.lock_steps, 8))           # bounded sample of the GPU step
    for _ in range(args.warmup):
        ls.run(per_step)
    u0, t0 = ls.n_updates, time.perf_counter()
    for _ in range(args.steps):
        ls.run(per_step)
    dt = time.perf_counter() - t0
    val = (ls.n_updates - u0) / dt
    sample = f"{per_step} lock-steps of {args.games} games per step (GPU step = {args.lock_steps})"
    line = {"impl": "reference", "metric": "td_updates_per_sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE configs[1]: Q_agent n={args.n} TD(0), {args.games} parallel seeded games per GPU, "
                        f"{args.lock_steps} lock-steps per step, rule={args.rule}, update mode={args.mode}",
            "n": args.n, "games_per_gpu": args.games, "lock_steps_per_step": args.lock_steps, "alpha": args.alpha,
            "update_mode": args.mode, "update_rule": args.rule, "sync_every": args.sync_every,
            "l2": "weight tables (4.46 MB at n=4) are L2-resident by construction; a 256 MiB buffer is written "
                  "between timed steps to flush L2"}


# --------------------------------------------------------------------------------------------- GPU arm
def run_td(args):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    importlib.import_module("2048_b200")
    from game2048 import cabi, engine
    from game2048 import parallel
    ctx = engine.Context.get()
    n, B, S = args.n, args.games, args.lock_steps
    mode = mode_bits(cabi, args)
    w_host = torch.from_numpy(seeded_weights(n)).pin_memory()
    # games sharded by global slot id (rank r owns slots [r*B, (r+1)*B)), weights replicated, per-rank deltas
    # allreduced over NCCL every --sync-every lock-steps (2048_b200/game2048/parallel.py)
    st = parallel.ShardedTrainer(n, w_host.numpy(), B, args.alpha, mode, seed=0, sync_every=args.sync_every)
    tr, wd, games = st.trainer, st.w, st.trainer.games
    flush = ctx.zeros(64 << 20, torch.int32)                     # 256 MiB > 126 MB L2

    def step():
        st.run(S)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(args.warmup):
        step()
    barrier()
    c0 = games.read_counters()
    l0 = st.launches
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        step()
        b.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches_timed = st.launches - l0                            # this rank's kernels inside the timed region
    ms = sum(a.elapsed_time(b) for a, b in ev)
    c1 = games.read_counters()
    upd = c1["updates"] - c0["updates"]
    mv = c1["moves"] - c0["moves"]
    evals = c1["evals"] - c0["evals"]
    t = torch.tensor([ms, float(upd), float(mv), float(evals)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, upd, mv, evals = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tsum[3])
    value = upd / (ms * 1e-3)

    # ---- e2e: the same step through host buffers: weights H2D (pinned) -> S lock-steps (with the NCCL delta syncs
    # at N > 1) -> weights + counters D2H, all inside the timed region; max over ranks, updates summed
    e2e_ms, e2e_upd = 0.0, 0
    for i in range(min(args.steps, 5) + 1):
        cA = games.read_counters()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        wd.copy_(w_host, non_blocking=True)
        if st.w_sync is not None:
            st.w_sync.copy_(wd)                                  # replicas restart from the same host weights
        st.run(S)
        w_host.copy_(wd, non_blocking=True)
        cnt = games.counters.cpu()                               # D2H read of the step's result
        b.record()
        torch.cuda.synchronize()
        if i:                                                    # first pass = warm-up
            e2e_ms += a.elapsed_time(b)
            e2e_upd += int(cnt[cabi.CTR_UPDATES]) - cA["updates"]
    h2d, d2h = wd.numel() * 4, wd.numel() * 4 + cabi.CTR_COUNT * 8
    if world > 1:
        t2 = torch.tensor([e2e_ms, float(e2e_upd)], dtype=torch.float64, device=ctx.device)
        tm, ts = t2.clone(), t2.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        e2e_ms, e2e_upd = float(tm[0]), float(ts[1])
    e2e = {"value": e2e_upd / (e2e_ms * 1e-3) if e2e_ms else None, "unit": "updates/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    # ---- roofline of the dominant kernel: the persistent lock-step kernel IS the timed step (one launch per step at
    # N=1), so its average launch duration is ms / steps, measured by the CUDA events above on the launching stream.
    # Algorithmic bytes per launch = updates per launch x (16 + 4 F E + 64 F)  (DESIGN 3).
    roof = None
    if rank == 0:
        F = F_OF_N[n]
        peak, how = peaks()
        e_per_move = evals / max(mv, 1)
        # N > 1: a step is S / sync_every persistent launches with a delta sync after each; the average below then
        # includes the sync kernels and the allreduce (whole-step view)
        n_persist = 1 if world == 1 else max(1, -(-S // args.sync_every))
        per_gpu_updates_per_launch = upd / world / args.steps / n_persist
        launch_s = ms * 1e-3 / args.steps / n_persist
        ach = per_gpu_updates_per_launch * bytes_per_update(n, e_per_move) / launch_s / 1e9
        persistent = launches_per_step_of(st, S) == 1
        roof = {"bound": "hbm", "kernel": "td_persist_kernel (phase A gather + argmax + spawn, phase B 8F-way scatter, "
                                          "apply; one cooperative launch per bench step)" if persistent else
                                          "td_phase_a + td_accum + td_apply (stepwise path)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": (lambda t: t * per_gpu_updates_per_launch if t and n == 4 and B == 4096 else None)(
                    ncu_traffic("td_persist_n4_B4096_atomic_mean", "bytes_per_update")),
                "traffic_source": "profiles/ncu_traffic.json: DRAM bytes per update of one ncu --set full capture of this kernel "
                                  "(n=4, 4096 games) x updates per launch; the tables never leave L2",
                "peak_source": how, "launch_us": launch_s * 1e6, "updates_per_launch": per_gpu_updates_per_launch,
                "bytes_per_update": bytes_per_update(n, e_per_move), "evals_per_move": e_per_move,
                "atomics_per_sec": value / world * 8 * F,
                "atomic_issue_peak_per_sec": 126e9,
                "note": "tables are L2-resident at n<=5, so HBM is the judged but not the physical bound: the kernel is "
                        "bound by the SM-side issue rate of returning L2 atomics (126 G/s measured chip-wide, "
                        "profiles/microbench/atomics2.cu) and by two grid barriers per lock-step"}

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = td_extras(args, ctx, engine, cabi, wd)
    cpu = None
    if rank == 0 and world == 1:
        v, cores, sample = cpu_td_sample(args, args.cpu_seconds)
        cpu = {"value": v, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        line = {"metric": "td_updates_per_sec", "value": value, "unit": "updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args), "moves_per_sec": mv / (ms * 1e-3), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches_timed),
                "roofline": roof, "cpu_baseline": cpu, "extras": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def td_extras(args, ctx, engine, cabi, wd):
    """secondary numbers on the same box: deterministic-mode updates/s, greedy moves/s, board sweep"""
    import torch
    out = {}
    n, B = args.n, args.games

    def timed(fn, reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3

    for name, mode in (("deterministic_mean", cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN),
                       ("deterministic_mean_sorted_stepwise", cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN | cabi.UPD_SORTED),
                       ("atomic_sum", cabi.UPD_ATOMIC | cabi.UPD_SUM),
                       ("atomic_mean_stepwise", cabi.UPD_ATOMIC | cabi.UPD_MEAN | cabi.RUN_STEPWISE)):
        w2 = wd.clone()
        g2 = engine.GameBatch(B, seed=1, ctx=ctx).init()
        alpha = args.alpha if mode & cabi.UPD_MEAN else args.alpha / B
        t2 = engine.TDTrainer(ctx, n, w2, g2, alpha, mode)
        t2.run(600)                                              # desynchronise the games (steady state)
        c0 = g2.read_counters()
        dt = timed(lambda: t2.run(256), 2)
        c1 = g2.read_counters()
        out[f"td_updates_per_sec_{name}"] = (c1["updates"] - c0["updates"]) / dt
    # deterministic mode: the weights after 300 lock-steps from seeded weights, twice -- the checksum (xor and wrapping
    # sum of the float bit patterns) is identical run to run (and equals the CPU oracle's, tests/test_gpu_full_size.py)
    sums = []
    for rep in range(2):
        w3 = ctx.to_device(seeded_weights(n))
        g4 = engine.GameBatch(B, seed=3, ctx=ctx).init()
        engine.TDTrainer(ctx, n, w3, g4, args.alpha, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN).run(300)
        bits = w3.view(torch.int32).to(torch.int64)
        x = int(np.bitwise_xor.reduce(bits.cpu().numpy()))
        sums.append(f"{x & 0xFFFFFFFF:08x}-{int(bits.sum().item()) & 0xFFFFFFFFFFFFFFFF:016x}")
    out["deterministic_weights_checksum_300_locksteps"] = sums[0]
    out["deterministic_checksum_repeat_equal"] = sums[0] == sums[1]
    # greedy play, BASELINE configs[0] shape: 1,000 seeded games to completion from the trained-so-far weights
    g3 = engine.GameBatch(1000, seed=2, ctx=ctx).init()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    engine.greedy_play(ctx, n, wd, g3)
    b.record()
    torch.cuda.synchronize()
    c = g3.read_counters()
    out["greedy_1000_games_moves_per_sec"] = c["moves"] / (a.elapsed_time(b) * 1e-3)
    out["greedy_1000_games_avg_score"] = c["score_sum"] / max(c["finished"], 1)
    # config 5 sweep: 16M boards
    m = args.boards
    gen = torch.Generator(device=ctx.device).manual_seed(0)      # cell iid: empty p=0.3 else exponent 1..11
    cells = torch.randint(1, 12, (m, 16), dtype=torch.int32, device=ctx.device, generator=gen)
    cells.mul_((torch.rand((m, 16), device=ctx.device, generator=gen) >= 0.3).to(torch.int32))
    boards = ctx.pack(cells)
    del cells
    bufs = ctx.sweep(boards, seed=0)
    dt = timed(lambda: ctx.sweep(boards, seed=0, out=bufs), 3) / 3
    out["sweep_boards_per_sec"] = m / dt
    out["sweep_GBps_algorithmic"] = m * 89 / dt / 1e9
    # the same sweep on boards met in games (SURVEY 8d config 5's second set: realistic merge density): snapshots of
    # 131,072 greedy games every 8 moves, played with the weights trained so far
    del boards, bufs
    per = min(131072, m)
    gh = engine.GameBatch(per, seed=5, ctx=ctx).init()
    snaps = []
    for _ in range(max(1, m // per)):
        engine.greedy_play(ctx, n, wd, gh, chunk=8, max_launches=1)
        snaps.append(gh.board.clone())
    boards = torch.cat(snaps)
    del snaps
    bufs = ctx.sweep(boards, seed=0)
    dt = timed(lambda: ctx.sweep(boards, seed=0, out=bufs), 3) / 3
    out["sweep_game_boards_per_sec"] = boards.numel() / dt
    return out


# --------------------------------------------------------------------------------------------- secondary workloads
def dist_setup():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def reduce_max_sum(vals, ctx, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=ctx.device)
    if world == 1:
        return list(vals), list(vals)
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return tmax.tolist(), tsum.tolist()


def run_greedy(args):
    """BASELINE configs[3] shape: greedy n-tuple play of `--games` seeded games per GPU to completion (default n=6,
    131,072 games per GPU = 1M games on 8 GPUs), weights fixed.  One bench step = all games of the rank, played by
    b2048_greedy_play (whole games per launch, Philox spawns keyed by GLOBAL game id, no collective)."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_setup()
    importlib.import_module("2048_b200")
    from game2048 import cabi, engine
    ctx = engine.Context.get()
    n, B = args.n, args.games
    w_host = torch.from_numpy(seeded_weights(n)).pin_memory()
    wd = ctx.empty(w_host.numel(), torch.float32)
    wd.copy_(w_host)
    if args.pretrain:                                            # a briefly trained agent plays longer games
        g0 = engine.GameBatch(4096, seed=5, ctx=ctx).init()
        engine.TDTrainer(ctx, n, wd, g0, args.alpha, cabi.UPD_ATOMIC | cabi.UPD_MEAN).run(args.pretrain)
        w_host.copy_(wd)
    games = engine.GameBatch(B, seed=0, ctx=ctx)
    flush = ctx.zeros(64 << 20, torch.int32)
    score_host, moves_host = torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()

    look = args.workload == "expectimax"

    def step(e2e=False):
        if e2e:
            wd.copy_(w_host, non_blocking=True)
        games.init(first_id=rank * B)
        if look:
            engine.expectimax_play(ctx, n, wd, games, args.depth, args.width, args.since_empty, chunk=args.chunk)
        else:
            engine.greedy_play(ctx, n, wd, games, chunk=args.chunk)
        if e2e:                                                  # per-game result: score, moves (+ counters)
            score_host.copy_(games.score, non_blocking=True)
            moves_host.copy_(games.moves, non_blocking=True)
            return games.read_counters()
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); sampler.wait_first()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark()
    ms, moves, evals, launches = 0.0, 0, 0, 0
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        c = games.read_counters()
        moves += c["moves"]; evals += c["evals"]
        score_avg = c["score_sum"] / max(c["finished"], 1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e_ms, e_moves = 0.0, 0
    for i in range(args.steps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        c = step(e2e=True)
        b.record()
        torch.cuda.synchronize()
        if i:
            e_ms += a.elapsed_time(b); e_moves += c["moves"]
    (mx, sm) = reduce_max_sum([ms, float(moves), float(evals), e_ms, float(e_moves)], ctx, world)
    ms, e_ms = mx[0], mx[3]
    moves, evals, e_moves = sm[1], sm[2], sm[4]
    if rank == 0:
        F = F_OF_N[n]
        peak, how = peaks()
        E = evals / max(moves, 1)
        bpm = 16 + 4 * F * E
        value = moves / (ms * 1e-3)
        ach = value / world * bpm / 1e9
        sector = value / world * (16 + 32 * F * E) / 1e9
        line = {"metric": "expectimax_moves_per_sec" if look else "greedy_moves_per_sec", "value": value, "unit": "moves/s",
                "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": (f"SURVEY 8(f) rank 1: Q_agent n={n} play with look-ahead depth={args.depth} width={args.width} "
                                        f"since_empty={args.since_empty} (game_logic.py:214-243), {B} seeded games per GPU to completion"
                                        if look else
                                        f"BASELINE configs[3] shape: Q_agent n={n} greedy play of {B} seeded games per GPU to "
                                        f"completion") + ", seeded random-init weights" +
                                       (f" + {args.pretrain} TD lock-steps" if args.pretrain else ""),
                           "n": n, "games_per_gpu": B, "moves_per_game": moves / world / args.steps / B,
                           "avg_score_rank0": score_avg,
                           "l2": "a 256 MiB buffer is written between timed steps to flush L2"},
                "clocks": clocks,
                "e2e": {"value": e_moves / (e_ms * 1e-3), "unit": "moves/s", "h2d_bytes_per_step": wd.numel() * 4,
                        "d2h_bytes_per_step": B * 8 + cabi.CTR_COUNT * 8},
                "gpu_launches": None,
                "roofline": {"bound": "hbm", "kernel": "expectimax_play_kernel (one warp per game, 16 lanes per root afterstate, "
                                                       "depth-first subtrees, F-table gather per leaf)" if look else
                                                       "greedy_play_kernel (4 LUT moves, F-table gather per valid afterstate, "
                                                       "argmax, Philox spawn; whole games per launch)",
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": (lambda t: t * moves / world / args.steps if t and n == 6 and not args.pretrain and not look else None)(
                                 ncu_traffic("greedy_n6_B131072_random_init", "bytes_per_move")),
                             "peak_source": how, "bytes_per_move": bpm, "evals_per_move": E,
                             "sector_granular_GBps": sector,
                             "note": "algorithmic bytes = 16 + 4 F E per move; a random 4-byte gather moves a 32-byte sector, "
                                     "sector_granular_GBps counts those"},
                "cpu_baseline": None}
        if world == 1:
            from oracle import oracle as orc
            orc.build()
            t0 = time.perf_counter()
            if look:
                r = orc.play_expectimax(n, w_host.numpy(), 0, 0, min(B, max(orc.max_threads(), 8)), args.depth, args.width,
                                        args.since_empty, threads=orc.max_threads())
            else:
                r = orc.play_philox(n, w_host.numpy(), seed=0, first_id=0, num=min(B, args.cpu_games), threads=orc.max_threads())
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": r["total_moves"] / dt, "unit": "moves/s", "cores": orc.max_threads(), "kind": "port",
                                    "sample": f"{len(r['scores'])} of the same games ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE configs[4]: `--boards` packed boards per GPU x 4 directions: afterstates, merge scores, changed /
    overflow flags and a Philox spawn on every changed afterstate (b2048_sweep), inputs larger than L2."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_setup()
    importlib.import_module("2048_b200")
    from game2048 import engine
    ctx = engine.Context.get()
    m = args.boards
    gen = torch.Generator(device=ctx.device).manual_seed(rank)       # cell iid: empty p=0.3 else exponent 1..11
    boards = None
    chunk = 1 << 22
    parts = []
    for i in range(0, m, chunk):
        k = min(chunk, m - i)
        cells = torch.randint(1, 12, (k, 16), dtype=torch.int32, device=ctx.device, generator=gen)
        cells.mul_((torch.rand((k, 16), device=ctx.device, generator=gen) >= 0.3).to(torch.int32))
        parts.append(ctx.pack(cells))
    boards = torch.cat(parts)
    del parts
    bufs = ctx.sweep(boards, seed=0)
    host_in = torch.empty(m, dtype=torch.int64).pin_memory()
    host_in.copy_(boards)
    host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in bufs]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); sampler.wait_first()
    for _ in range(args.warmup):
        ctx.sweep(boards, seed=0, out=bufs)
    barrier()
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:                                              # 128 MiB in + 1.3 GB out per step: larger than L2
        a.record()
        ctx.sweep(boards, seed=0, first_index=rank * m, out=bufs)
        b.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    e_ms = 0.0
    for i in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        boards.copy_(host_in, non_blocking=True)
        ctx.sweep(boards, seed=0, first_index=rank * m, out=bufs)
        for h, t in zip(host_out, bufs):
            h.copy_(t, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        if i:
            e_ms += a.elapsed_time(b) / 2
    (mx, _) = reduce_max_sum([ms, e_ms], ctx, world)
    if rank == 0:
        peak, how = peaks()
        value = world * m * args.steps / (mx[0] * 1e-3)
        ach = value / world * 89 / 1e9
        line = {"metric": "sweep_boards_per_sec", "value": value, "unit": "boards/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": mx[0] / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[4]: {m} packed boards per GPU x 4 directions move/merge/score + spawn",
                           "boards_per_gpu": m, "l2": "inputs + outputs (1.4 GB per step) are larger than L2"},
                "clocks": clocks,
                "e2e": {"value": world * m / (mx[1] * 1e-3), "unit": "boards/s", "h2d_bytes_per_step": m * 8,
                        "d2h_bytes_per_step": m * 81},
                "gpu_launches": args.steps,
                "roofline": {"bound": "hbm", "kernel": "sweep_kernel (row LUT in shared memory, persistent grid)", "achieved": ach,
                             "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": (lambda t: t * m if t else None)(ncu_traffic("sweep_16M", "bytes_per_board")),
                             "peak_source": how, "bytes_per_board": 89},
                "cpu_baseline": None}
        if world == 1:
            from oracle import oracle as orc
            orc.build()
            k = min(m, 1 << 22)
            hb = host_in.numpy().view(np.uint64)[:k]
            t0 = time.perf_counter()
            orc.sweep(hb, seed=0, threads=orc.max_threads())
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": k / dt, "unit": "boards/s", "cores": orc.max_threads(), "kind": "port",
                                    "sample": f"the first {k} boards ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "greedy" and args.n == 4 and "--n" not in sys.argv and "--tuple" not in sys.argv:
        args.n = 6
    if args.workload == "greedy" and "--games" not in sys.argv:
        args.games = 131072
    if args.workload == "expectimax":
        if "--games" not in sys.argv:
            args.games = 1024
        if "--pretrain" not in sys.argv:
            args.pretrain = 3000
        if "--chunk" not in sys.argv:
            args.chunk = 256
    if args.impl == "reference":
        if args.workload != "td":
            raise SystemExit("--impl reference is the td workload (the headline); greedy / sweep lines carry cpu_baseline")
        run_reference(args)
        return
    {"td": run_td, "greedy": run_greedy, "sweep": run_sweep, "expectimax": run_greedy}[args.workload](args)


if __name__ == "__main__":
    main()
