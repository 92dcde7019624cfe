#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native 2048 / n-tuple hot path over all five BASELINE.json configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload td|greedy|sweep|expectimax]

Default (`--workload td`): ONE JSON line on stdout (rank 0).  Its top-level keys are the headline, BASELINE
configs[1]: Q_agent n=4 TD(0) training, 4,096 parallel seeded games per GPU (weak scaling), per-key-mean lock-step
rule, atomic update mode; one bench "step" = `--lock-steps` lock-steps of all games; metric = TD updates/s (one
update = one QAgent.update() equivalent = 8*F weight RMWs).  The `configs` object of the same line carries the other
four configs, each with value / e2e / roofline / cpu_baseline measured in the same invocation:
    configs[0]  greedy n=4, 1,000 seeded games in total from fixed (pre-trained) weights          moves/s, strong
    configs[2]  n=5 TD(0), 65,536 games IN TOTAL split over the N GPUs, weight sync every 64        updates/s, strong
                (one step = 512 lock-steps = 8 sync periods)
    configs[3]  greedy n=6, 131,072 games per GPU, random-init AND pre-trained weights              moves/s, weak
    configs[4]  board sweep, 16M packed boards per GPU x 4 directions + spawn                       boards/s, weak
See DESIGN.md "Measurement" for every field.

--impl reference times the reference algorithm on the host CPU for the headline config: the C + OpenMP restatement in
oracle/ (1 thread and all threads, the better one is the line's value) and, where the unmodified Python reference is
importable (/root/reference or $B2048_REFERENCE: in the build container, not on the GPU box), its own
QAgent.episode on one core beside it.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_OF_N = {2: 24, 3: 52, 4: 17, 5: 21, 6: 33}
CPU_CACHE = os.path.join(ROOT, "gpurun_out", ".cpu_baseline_td.json")


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
    p.add_argument("--sync-every", type=int, default=128, help="lock-steps between weight syncs (N>1)")
    p.add_argument("--sync-impl", default="auto", choices=["auto", "p2p", "nccl"],
                   help="N>1 weight exchange: fused peer-memory kernel or NCCL allreduce + allgather")
    p.add_argument("--fused", action="store_true",
                   help="p2p exchange inside the persistent launches (b2048_td_run_peers) instead of a kernel between them")
    p.add_argument("--boards", type=int, default=1 << 24, help="boards per GPU (sweep)")
    p.add_argument("--no-extras", action="store_true")
    p.add_argument("--no-configs", action="store_true", help="headline only: skip the `configs` object")
    p.add_argument("--only-config", type=int, default=None, help="with --workload td: run just this config of `configs`")
    p.add_argument("--config-steps", type=int, default=5, help="timed steps of each entry of `configs`")
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
        """start of a timed region: only samples taken after this call are reported by the next snapshot()"""
        self.t_mark = time.time()

    def snapshot(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"], "samples": 0}
        timed = [x for x in self.samples if x[0] >= self.t_mark] or self.samples[-3:]
        reasons = sorted({name for _, _, r in timed for name, bit in self.REASONS if r & bit})
        return {"sm_mhz": float(np.median([x[1] for x in timed])), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(timed)}

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        return self.snapshot()


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
    except (OSError, KeyError, ValueError, TypeError):
        return None


def launches_per_step_of(st, S):
    """kernel launches one bench step (S lock-steps) costs on this rank, sync kernels excluded"""
    from game2048 import cabi
    return int(cabi.lib().b2048_td_run_launches(st.n, st.B, st.trainer.mode, S))


def bytes_per_update(n, evals_per_move):
    F = F_OF_N[n]
    return 16 + 4 * F * evals_per_move + 64 * F          # board in/out + gathers + 8F RMWs x (4+4) B


class Dist:
    """rank / world of this process and max / sum reductions over the ranks (device tensors, NCCL)"""

    def __init__(self, sync_impl="auto"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.comm = None
        if self.world > 1 and not dist.is_initialized():
            importlib.import_module("2048_b200")
            from game2048 import parallel
            self.comm = parallel.init_distributed(self.local, peer_exchange=sync_impl != "nccl")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_sum(self, vals):
        if self.world == 1:
            return list(vals), list(vals)
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        tmax, tsum = t.clone(), t.clone()
        self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(tsum, op=self.dist.ReduceOp.SUM)
        return tmax.tolist(), tsum.tolist()

    def finish(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


def timed_region(D, flush, step, steps, warmup):
    """`warmup` untimed steps, then `steps` steps timed one by one with CUDA events on the launching stream; an L2
    flush (a write larger than L2) before each; barrier + synchronize on both sides.  Returns this rank's total ms."""
    torch = D.torch
    for _ in range(warmup):
        step()
    D.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.fill_(1)
        a.record()
        step()
        b.record()
    D.barrier()
    return sum(a.elapsed_time(b) for a, b in ev)


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_td_port(n, games, alpha, rule, seconds, threads):
    """oracle lock-step TD (float32, same rule as the GPU's) on `threads` host threads for about `seconds`"""
    from oracle import oracle as orc
    orc.build()
    w = seeded_weights(n)
    ls = orc.LockStep(n, w, alpha, 0, games, segmented=2 if rule == "mean" else 1, threads=threads)
    ls.run(2)                                            # touch memory
    u0, t0, steps = ls.n_updates, time.perf_counter(), 0
    chunk = 1 if games > 8192 else 4
    while True:
        ls.run(chunk)
        steps += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
        chunk = max(1, min(64, int(chunk * seconds / max(dt, 1e-3) / 4)))
    return (ls.n_updates - u0) / dt, f"{steps} lock-steps of {games} games, n={n}, {threads} thread(s) ({dt:.1f} s)"


def cpu_td_baseline(n, games, alpha, rule, seconds):
    """the port at 1 thread and at all host threads (only its evaluate phase is parallel: the scaling is what it is);
    value = the better of the two, cores = the thread count that produced it"""
    from oracle import oracle as orc
    orc.build()
    tmax = max(1, orc.max_threads())
    res = {}
    for t in sorted({1, tmax}):
        res[t] = cpu_td_port(n, games, alpha, rule, seconds / (2 if tmax > 1 else 1), t)
    best = max(res, key=lambda t: res[t][0])
    return {"value": res[best][0], "unit": "updates/s", "cores": best, "kind": "port", "sample": res[best][1],
            "thread_scaling": {str(t): res[t][0] for t in res}, "host_threads": tmax}


def python_reference_td(n, alpha, seconds):
    """the UNMODIFIED Python reference (QAgent.episode, r_learning.py:224-252) on one core, where it is importable
    (oracle/ref_shim.py: the build container; not the GPU box).  None otherwise."""
    try:
        from oracle import ref_shim
        if not ref_shim.available():
            return None
        gl, rl = ref_shim.load()
        ref_shim.seed_all(0)
        agent = ref_shim.make_agent(rl, n, alpha=alpha)
        updates, episodes, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            game = agent.episode()
            updates += game.odometer                 # one update() per move from the 2nd on, + the terminal one
            episodes += 1
        dt = time.perf_counter() - t0
        return {"value": updates / dt, "unit": "updates/s", "cores": 1, "kind": "reference",
                "sample": f"{episodes} QAgent.episode() calls, n={n}, random-init weights, 1 core ({dt:.1f} s)"}
    except Exception as e:                               # noqa: BLE001 - the reference arm must still print its line
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    orc.build()
    tmax = max(1, orc.max_threads())
    rule = 2 if args.rule == "mean" else 1
    per_step = max(1, min(args.lock_steps, 8))           # bounded sample of the GPU step
    runs = {}
    for threads in sorted({1, tmax}):
        ls = orc.LockStep(args.n, seeded_weights(args.n), args.alpha, 0, args.games, segmented=rule, threads=threads)
        for _ in range(args.warmup):
            ls.run(per_step)
        u0, t0 = ls.n_updates, time.perf_counter()
        for _ in range(args.steps):
            ls.run(per_step)
        dt = time.perf_counter() - t0
        runs[threads] = ((ls.n_updates - u0) / dt, dt)
    best = max(runs, key=lambda t: runs[t][0])
    val, dt = runs[best]
    sample = f"{per_step} lock-steps of {args.games} games per step (GPU step = {args.lock_steps}), {best} thread(s)"
    cpu = {"value": val, "unit": "updates/s", "cores": best, "kind": "port", "sample": sample,
           "thread_scaling": {str(t): runs[t][0] for t in runs}, "host_threads": tmax}
    line = {"impl": "reference", "metric": "td_updates_per_sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    py = python_reference_td(args.n, args.alpha, min(args.cpu_seconds, 10.0))
    if py is not None:
        line["python_reference"] = py
    try:                                                 # the b200 arm of the same box reports THIS measurement
        os.makedirs(os.path.dirname(CPU_CACHE), exist_ok=True)
        with open(CPU_CACHE, "w") as f:
            json.dump({"when": time.time(), "n": args.n, "games": args.games, "rule": args.rule, "cpu_baseline": cpu,
                       "python_reference": py}, f)
    except OSError:
        pass
    print(json.dumps(line), flush=True)


def cached_cpu_td(args):
    """the --impl reference measurement of this box (the driver runs that arm first), if it is the same config and
    less than an hour old: both arms then report ONE cpu measurement"""
    try:
        with open(CPU_CACHE) as f:
            c = json.load(f)
        if c["n"] == args.n and c["games"] == args.games and c["rule"] == args.rule and time.time() - c["when"] < 3600:
            cpu = dict(c["cpu_baseline"])
            cpu["source"] = "the --impl reference run on this box (one measurement for both arms)"
            return cpu, c.get("python_reference")
    except (OSError, KeyError, ValueError):
        pass
    return None, None


def workload_config(args):
    return {"workload": f"BASELINE configs[1]: Q_agent n={args.n} TD(0), {args.games} parallel seeded games per GPU, "
                        f"{args.lock_steps} lock-steps per step, rule={args.rule}, update mode={args.mode}",
            "n": args.n, "games_per_gpu": args.games, "lock_steps_per_step": args.lock_steps, "alpha": args.alpha,
            "update_mode": args.mode, "update_rule": args.rule, "sync_every": args.sync_every,
            "l2": "weight tables (4.46 MB at n=4) are L2-resident by construction; a 256 MiB buffer is written "
                  "between timed steps to flush L2"}


# --------------------------------------------------------------------------------------------- TD (configs[1], [2])
def td_bench(D, ctx, args, n, B, S, steps, warmup, mode, sync_every, first_slot, total_slots, sampler=None,
             traffic_key=None):
    """Lock-step TD on this rank's B slots (global slots first_slot .. of total_slots), S lock-steps per bench step.
    Returns (result dict for rank 0 | None, trainer): value, e2e, roofline, sync figures, sync_check."""
    torch = D.torch
    from game2048 import cabi, parallel
    w_host = torch.from_numpy(seeded_weights(n)).pin_memory()
    st = parallel.ShardedTrainer(n, w_host.numpy(), B, args.alpha, mode, seed=0, sync_every=sync_every,
                                 sync_impl=args.sync_impl, first_slot=first_slot, total_slots=total_slots,
                                 fused=args.fused)
    wd, games = st.w, st.trainer.games
    flush = ctx.zeros(64 << 20, torch.int32)                     # 256 MiB > 126 MB L2
    for _ in range(warmup):
        st.run(S)
    D.barrier()
    c0, l0 = games.read_counters(), st.launches
    if sampler:
        sampler.mark()
    ms = timed_region(D, flush, lambda: st.run(S), steps, 0)
    clocks = sampler.snapshot() if sampler else None
    launches_timed = st.launches - l0
    c1 = games.read_counters()
    upd, mv, evals = c1["updates"] - c0["updates"], c1["moves"] - c0["moves"], c1["evals"] - c0["evals"]
    print(f"[rank {D.rank}] n={n} B={B} S={S}: {ms / steps:.3f} ms/step, {upd} updates", file=sys.stderr, flush=True)
    mx, sm = D.max_sum([ms, float(upd), float(mv), float(evals)])
    ms, upd, mv, evals = mx[0], sm[1], sm[2], sm[3]
    value = upd / (ms * 1e-3)
    in_sync = st.replicas_identical() if S % max(sync_every, 1) == 0 or D.world == 1 else None

    # ---- e2e: the same step through host buffers: weights H2D (pinned) -> S lock-steps (with the weight syncs at
    # N > 1) -> weights + counters D2H, all inside the timed region; max over ranks, updates summed
    e2e_ms, e2e_upd = 0.0, 0
    for i in range(min(steps, 5) + 1):
        cA = games.read_counters()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
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
    mx2, sm2 = D.max_sum([e2e_ms, float(e2e_upd)])
    e2e = {"value": sm2[1] / (mx2[0] * 1e-3) if mx2[0] else None, "unit": "updates/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    # ---- one weight sync in isolation (N > 1): CUDA events around sync() after one lock-step, max over ranks
    sync_us = None
    if D.world > 1:
        tot = 0.0
        for i in range(6):
            st.run(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            D.barrier()
            a.record()
            st.sync()
            b.record()
            torch.cuda.synchronize()
            if i:
                tot += a.elapsed_time(b) / 5
        sync_us = D.max_sum([tot * 1e3])[0][0]
        ok2 = st.replicas_identical()
        in_sync = ok2 if in_sync is None else (in_sync and ok2)
    if D.rank != 0:
        return None, st
    # ---- roofline of the dominant kernel: the persistent lock-step kernel IS the timed step (one launch per step at
    # N=1), so its average launch duration is ms / steps, measured by the CUDA events above on the launching stream.
    # Algorithmic bytes per launch = updates per launch x (16 + 4 F E + 64 F)  (DESIGN 3).
    F = F_OF_N[n]
    peak, how = peaks()
    e_per_move = evals / max(mv, 1)
    # N > 1: a step is S / sync_every persistent launches with a weight sync after each; the average below then
    # includes the sync (whole-step view)
    n_persist = 1 if D.world == 1 else max(1, -(-S // sync_every))
    upl = upd / D.world / steps / n_persist
    launch_s = ms * 1e-3 / steps / n_persist
    ach = upl * bytes_per_update(n, e_per_move) / launch_s / 1e9
    persistent = launches_per_step_of(st, S) == 1
    tr_b = ncu_traffic(traffic_key, "bytes_per_update") if traffic_key else None
    roof = {"bound": "hbm", "kernel": "td_persist_kernel (phase A gather + argmax + spawn, phase B 8F-way scatter, "
                                      "apply; one cooperative launch per sync period)" if persistent else
                                      "td_phase_a + td_accum + td_apply (stepwise path)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": tr_b * upl if tr_b else None,
            "traffic_source": f"profiles/ncu_traffic.json[{traffic_key}]: DRAM bytes per update of one ncu --set full "
                              "capture of this kernel x updates per launch" if tr_b else None,
            "peak_source": how, "launch_us": launch_s * 1e6, "updates_per_launch": upl,
            "bytes_per_update": bytes_per_update(n, e_per_move), "evals_per_move": e_per_move,
            "l2": {"achieved_atomics_per_sec": value / D.world * 8 * F, "peak_atomics_per_sec": 126e9,
                   "frac": value / D.world * 8 * F / 126e9,
                   "peak_source": "profiles/r01b_microbench_atomics_steady.txt: returning L2 atomics, chip-wide"},
            "note": "tables are L2-resident at n<=5, so HBM is the judged but not the physical bound: the kernel is "
                    "bound by the SM-side issue rate of returning L2 atomics and by two grid barriers per lock-step"}
    res = {"metric": "td_updates_per_sec", "value": value, "unit": "updates/s", "ms_per_step": ms / steps,
           "steps": steps, "warmup": warmup, "moves_per_sec": mv / (ms * 1e-3), "clocks": clocks, "e2e": e2e,
           "gpu_launches": int(launches_timed), "roofline": roof}
    if D.world > 1:
        res.update(sync_check=bool(in_sync), sync_us=sync_us,
                   sync_impl=st.sync_impl + (" (inside the persistent launch)" if st.fused else ""),
                   sync_message_bytes_per_rank=st.message_bytes, comm=D.comm,
                   sync_fallback_reason=getattr(st.ops, "peer_error", None))
    return res, st


def run_td(args):
    D = Dist(args.sync_impl)
    importlib.import_module("2048_b200")
    from game2048 import cabi, engine
    ctx = engine.Context.get()
    n, B, S = args.n, args.games, args.lock_steps
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
        sampler.wait_first()
    # games sharded by global slot id (rank r owns slots [r*B, (r+1)*B)), weights replicated, per-rank weight
    # movements combined every --sync-every lock-steps (2048_b200/game2048/parallel.py)
    res, st = td_bench(D, ctx, args, n, B, S, args.steps, args.warmup, mode_bits(cabi, args), args.sync_every,
                       first_slot=None, total_slots=None, sampler=sampler if D.rank == 0 else None,
                       traffic_key="td_persist_n4_B4096_atomic_mean" if (n, B) == (4, 4096) else None)
    if D.world > 1 and res is not None and res.get("sync_check") is False:
        raise SystemExit("replicas differ after the last weight sync: no value printed")
    line = None
    if D.rank == 0:
        extras = td_extras(args, ctx, engine, cabi, st.w) if D.world == 1 and not args.no_extras else {}
        cpu, py = (None, None)
        if D.world == 1:
            cpu, py = cached_cpu_td(args)
            if cpu is None:
                cpu = cpu_td_baseline(n, B, args.alpha, args.rule, args.cpu_seconds)
        line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": D.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args)}
        for k in ("moves_per_sec", "clocks", "e2e", "gpu_launches", "roofline", "sync_check", "sync_us", "sync_impl",
                  "sync_message_bytes_per_rank", "sync_fallback_reason", "comm"):
            if k in res:
                line[k] = res[k]
        line["cpu_baseline"] = cpu
        if py:
            line["python_reference"] = py
        line["extras"] = extras
    del st
    D.torch.cuda.empty_cache()
    if not args.no_configs:
        cfg = run_configs(D, ctx, args, sampler if D.rank == 0 else None)
        if D.rank == 0:
            line["configs"] = cfg
    if D.rank == 0:
        sampler.stop()
        print(json.dumps(line), flush=True)
    D.finish()


def td_extras(args, ctx, engine, cabi, wd):
    """secondary numbers on the same box: the other update modes of the headline shape"""
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
    return out


# --------------------------------------------------------------------------------------------- configs[0], [3]
def pretrained_weights(ctx, n, lock_steps, alpha=0.25):
    """seeded random-init weights + `lock_steps` deterministic per-key-mean TD lock-steps of 4,096 games: the same
    bits on every rank and every run (fixed weights for the greedy configs), returned as a device tensor"""
    from game2048 import cabi, engine
    wd = ctx.to_device(seeded_weights(n))
    if lock_steps:
        g0 = engine.GameBatch(4096, seed=5, ctx=ctx).init()
        engine.TDTrainer(ctx, n, wd, g0, alpha, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN).run(lock_steps)
    return wd


def greedy_bench(D, ctx, args, n, first_id, B, wd, steps, warmup, label, sampler=None, traffic_key=None, cpu_games=0,
                 look=None):
    """greedy (or look-ahead) play of this rank's B games (global ids first_id ..) to completion per bench step"""
    torch = D.torch
    from game2048 import cabi, engine
    w_host = torch.empty(wd.numel(), dtype=torch.float32).pin_memory()
    w_host.copy_(wd)
    games = engine.GameBatch(max(B, 1), seed=0, ctx=ctx)
    flush = ctx.zeros(64 << 20, torch.int32)
    score_host = torch.empty(max(B, 1), dtype=torch.int32).pin_memory()
    moves_host = torch.empty(max(B, 1), dtype=torch.int32).pin_memory()
    launches = [0]

    def play():
        games.init(first_id=first_id)
        if look:
            engine.expectimax_play(ctx, n, wd, games, *look, chunk=args.chunk)
        else:
            engine.greedy_play(ctx, n, wd, games, chunk=1 << 20)
        launches[0] += 2

    tot = {"moves": 0, "evals": 0, "score": 0, "fin": 0}

    def step():
        play()
        c = games.read_counters()
        tot["moves"] += c["moves"]; tot["evals"] += c["evals"]; tot["score"] += c["score_sum"]; tot["fin"] += c["finished"]

    for _ in range(warmup):
        play()
    for k in tot:
        tot[k] = 0
    launches[0] = 0
    if sampler:
        sampler.mark()
    # per step: flush, event, all games to completion (one persistent launch + the init kernel), event
    D.barrier()
    ms = 0.0
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        play()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
        c = games.read_counters()
        tot["moves"] += c["moves"]; tot["evals"] += c["evals"]; tot["score"] += c["score_sum"]; tot["fin"] += c["finished"]
    D.barrier()
    clocks = sampler.snapshot() if sampler else None
    e_ms, e_moves = 0.0, 0
    for i in range(min(steps, 3) + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        a.record()
        wd.copy_(w_host, non_blocking=True)                          # weights H2D from pinned memory
        play()
        score_host.copy_(games.score, non_blocking=True)             # per-game result D2H
        moves_host.copy_(games.moves, non_blocking=True)
        c = games.read_counters()
        b.record()
        torch.cuda.synchronize()
        if i:
            e_ms += a.elapsed_time(b); e_moves += c["moves"]
    longest = float(games.moves.max().item()) if B else 0.0          # the longest game bounds a latency-bound step
    mx, sm = D.max_sum([ms, float(tot["moves"]), float(tot["evals"]), e_ms, float(e_moves), float(tot["score"]),
                        float(tot["fin"]), longest])
    if D.rank != 0:
        return None
    ms, e_ms = mx[0], mx[3]
    moves, evals, e_moves = sm[1], sm[2], sm[4]
    F = F_OF_N[n]
    peak, how = peaks()
    E = evals / max(moves, 1)
    bpm = 16 + 4 * F * E
    value = moves / (ms * 1e-3)
    ach = value / D.world * bpm / 1e9
    tr_b = ncu_traffic(traffic_key, "bytes_per_move") if traffic_key else None
    nw_bytes = wd.numel() * 4
    res = {"metric": "expectimax_moves_per_sec" if look else "greedy_moves_per_sec", "value": value, "unit": "moves/s",
           "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "config": {"workload": label, "n": n, "games_this_rank": B, "moves_per_game": moves / max(sm[6], 1),
                      "avg_score": sm[5] / max(sm[6], 1), "longest_game_moves": mx[7],
                      "us_per_move_of_longest_game": ms / steps * 1e3 / max(mx[7], 1),
                      "l2": "a 256 MiB buffer is written between timed steps to flush L2"},
           "clocks": clocks,
           "e2e": {"value": e_moves / (e_ms * 1e-3) if e_ms else None, "unit": "moves/s", "h2d_bytes_per_step": nw_bytes,
                   "d2h_bytes_per_step": B * 8 + cabi.CTR_COUNT * 8},
           "gpu_launches": launches[0],
           "roofline": {"bound": "hbm", "kernel": "expectimax_play_kernel" if look else "greedy_play_kernel (4 LUT moves, "
                        "F-table gather per valid afterstate, shuffle argmax, Philox spawn; whole games per launch)",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": tr_b * moves / D.world / steps if tr_b else None,
                        "peak_source": how, "bytes_per_move": bpm, "evals_per_move": E,
                        "sector_granular_GBps": value / D.world * (16 + 32 * F * E) / 1e9,
                        "l2_gather": {"achieved_gathers_per_sec": value / D.world * F * E, "peak_gathers_per_sec": 293e9,
                                      "frac": value / D.world * F * E / 293e9,
                                      "peak_source": "profiles/r02_microbench_gather.txt: random 4-byte gathers over an "
                                                     "L2-resident footprint, 9.4 TB/s of 32-byte sectors (57 G/s = 1.8 TB/s "
                                                     "when the footprint is the whole 383 MB of n=6 in HBM); a fraction "
                                                     "above 1 is L1 hits"},
                        "note": "algorithmic bytes = 16 + 4 F E per move; a random 4-byte gather moves a 32-byte sector, "
                                "sector_granular_GBps counts those; tables of n <= 5 are L2-resident"},
           "cpu_baseline": None}
    if D.world == 1 and cpu_games:
        from oracle import oracle as orc
        orc.build()
        k = min(B, cpu_games)
        wh = w_host.numpy()
        t0, cpu_moves, reps, first_scores = time.perf_counter(), 0, 0, None
        while True:                                      # the sample, repeated on the following id ranges, for >= 3 s
            if look:
                r = orc.play_expectimax(n, wh, 0, first_id + reps * k, k, *look, threads=orc.max_threads())
            else:
                r = orc.play_philox(n, wh, seed=0, first_id=first_id + reps * k, num=k, threads=orc.max_threads())
            if first_scores is None:
                first_scores = r["scores"]
            cpu_moves += r["total_moves"]
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= 3.0 or look:
                break
        res["cpu_baseline"] = {"value": cpu_moves / dt, "unit": "moves/s", "cores": orc.max_threads(), "kind": "port",
                               "sample": f"{reps} x {k} games (the first {k} are the GPU's own), all host threads ({dt:.1f} s)"}
        # the oracle doubles as the checker: the games it played must be the GPU's, score for score
        res["oracle_match"] = bool(np.array_equal(score_host.numpy()[:k].astype(np.int64), first_scores))
    return res


# --------------------------------------------------------------------------------------------- configs[4]
def sweep_bench(D, ctx, args, m, steps, warmup, sampler=None):
    """BASELINE configs[4]: m packed boards on this rank x 4 directions: afterstates, merge scores, changed /
    overflow flags and a Philox spawn on every changed afterstate (b2048_sweep), inputs larger than L2."""
    torch = D.torch
    gen = torch.Generator(device=ctx.device).manual_seed(D.rank)     # cell iid: empty p=0.3 else exponent 1..11
    parts = []
    for i in range(0, m, 1 << 22):
        k = min(1 << 22, m - i)
        cells = torch.randint(1, 12, (k, 16), dtype=torch.int32, device=ctx.device, generator=gen)
        cells.mul_((torch.rand((k, 16), device=ctx.device, generator=gen) >= 0.3).to(torch.int32))
        parts.append(ctx.pack(cells))
        del cells
    boards = torch.cat(parts)
    del parts
    bufs = ctx.sweep(boards, seed=0)
    host_in = torch.empty(m, dtype=torch.int64).pin_memory()
    host_in.copy_(boards)
    host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in bufs]
    if sampler:
        sampler.mark()
    # 128 MiB in + 1.3 GB out per step: larger than L2, no flush needed
    ms = timed_region(D, None, lambda: ctx.sweep(boards, seed=0, first_index=D.rank * m, out=bufs), steps, warmup)
    clocks = sampler.snapshot() if sampler else None
    e_ms = 0.0
    for i in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        a.record()
        boards.copy_(host_in, non_blocking=True)
        ctx.sweep(boards, seed=0, first_index=D.rank * m, out=bufs)
        for h, t in zip(host_out, bufs):
            h.copy_(t, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        if i:
            e_ms += a.elapsed_time(b) / 2
    mx, _ = D.max_sum([ms, e_ms])
    if D.rank != 0:
        return None
    peak, how = peaks()
    value = D.world * m * steps / (mx[0] * 1e-3)
    ach = value / D.world * 89 / 1e9
    tr_b = ncu_traffic("sweep_16M", "bytes_per_board")
    res = {"metric": "sweep_boards_per_sec", "value": value, "unit": "boards/s", "ms_per_step": mx[0] / steps,
           "steps": steps, "warmup": warmup, "dtype": "u64",
           "config": {"workload": f"BASELINE configs[4]: {m} packed boards per GPU x 4 directions move/merge/score + spawn",
                      "boards_per_gpu": m, "l2": "inputs + outputs (1.4 GB per step) are larger than L2"},
           "clocks": clocks,
           "e2e": {"value": D.world * m / (mx[1] * 1e-3), "unit": "boards/s", "h2d_bytes_per_step": m * 8,
                   "d2h_bytes_per_step": m * 81},
           "gpu_launches": steps,
           "roofline": {"bound": "hbm", "kernel": "sweep_kernel (row LUT in shared memory, persistent grid)", "achieved": ach,
                        "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": tr_b * m if tr_b else None,
                        "peak_source": how, "bytes_per_board": 89},
           "cpu_baseline": None}
    if D.world == 1:
        from oracle import oracle as orc
        orc.build()
        k = min(m, 1 << 22)
        hb = host_in.numpy().view(np.uint64)[:k]
        t0 = time.perf_counter()
        ra = orc.sweep(hb, seed=0, threads=orc.max_threads())
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": k / dt, "unit": "boards/s", "cores": orc.max_threads(), "kind": "port",
                               "sample": f"the first {k} boards, all host threads ({dt:.1f} s)"}
        res["oracle_match"] = bool(np.array_equal(host_out[0].numpy().view(np.uint64)[:k], ra[0]) and
                                   np.array_equal(host_out[3].numpy().view(np.uint64)[:k], ra[3]))
    return res


# --------------------------------------------------------------------------------------------- the configs object
def run_configs(D, ctx, args, sampler):
    """BASELINE configs[0], [2], [3], [4] in the same invocation (configs[1] is the headline)"""
    torch = D.torch
    from game2048 import cabi, parallel
    out = {}
    ks, kw = args.config_steps, 3
    want = (lambda i: args.only_config is None or args.only_config == i)

    def free():
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    if want(0):       # ---- configs[0]: 1,000 games IN TOTAL, n=4, fixed pre-trained weights (strong scaling)
        total = 1000
        first, count = parallel.shard(total, D.world, D.rank)
        wd = pretrained_weights(ctx, 4, 3000)
        r = greedy_bench(D, ctx, args, 4, first, count, wd, ks, kw,
                         f"BASELINE configs[0]: Q_agent n=4 greedy play of {total} seeded games in total from fixed weights "
                         "(seeded init + 3000 deterministic TD lock-steps of 4096 games), to completion", sampler,
                         traffic_key="greedy_n4_B1000_pretrained", cpu_games=total)
        if r:
            r.update(scaling="strong", total_games=total)
            out["configs[0]"] = r
        del wd
        free()
    if want(2):       # ---- configs[2]: n=5, 65,536 games IN TOTAL over the N GPUs, weight sync every 64 lock-steps
        total, K, S = 65536, 64, 512                      # one bench step = 512 lock-steps = 8 sync periods
        first, count = parallel.shard(total, D.world, D.rank)
        r, st = td_bench(D, ctx, args, 5, count, S, ks, kw, cabi.UPD_ATOMIC | cabi.UPD_MEAN, K, first_slot=first,
                         total_slots=total, sampler=sampler, traffic_key=f"td_persist_n5_B{count}_atomic_mean")
        if r:
            r.update(scaling="strong", total_games=total,
                     config={"workload": f"BASELINE configs[2]: Q_agent n=5 TD(0), {total} games in total sharded over "
                                         f"{D.world} GPU(s) ({count} per GPU), weight sync every {K} lock-steps, "
                                         f"rule=mean, update mode=atomic; one step = {S} lock-steps",
                             "n": 5, "games_per_gpu": count, "lock_steps_per_step": S, "sync_every": K,
                             "l2": "21.2 MB of tables are L2-resident; 256 MiB flush between timed steps"})
            if D.world == 1:
                r["cpu_baseline"] = cpu_td_baseline(5, total, args.alpha, "mean", min(args.cpu_seconds, 6.0))
            out["configs[2]"] = r
        del st
        free()
    if want(3):       # ---- configs[3]: n=6 greedy, 131,072 games per GPU (1M on 8), random-init and pre-trained
        per = 131072
        for name, pre, key in (("random_init", 0, "greedy_n6_B131072_random_init"),
                               ("pretrained", 3000, "greedy_n6_B131072_pretrained")):
            wd = pretrained_weights(ctx, 6, pre)
            r = greedy_bench(D, ctx, args, 6, D.rank * per, per, wd, min(ks, 3), kw,
                             f"BASELINE configs[3]: Q_agent n=6 greedy play of {per} seeded games per GPU to completion, "
                             + ("seeded random-init weights" if not pre else
                                f"seeded init + {pre} deterministic TD lock-steps of 4096 games"), sampler,
                             traffic_key=key, cpu_games=4096 if not pre else 256)
            if r:
                r.update(scaling="weak", games_per_gpu=per)
                out.setdefault("configs[3]", {})[name] = r
            del wd
            free()
    if want(4):       # ---- configs[4]: 16M boards per GPU
        r = sweep_bench(D, ctx, args, args.boards, ks, kw, sampler)
        if r:
            r.update(scaling="weak")
            out["configs[4]"] = r
        free()
    return out


# --------------------------------------------------------------------------------------------- secondary workloads
def run_greedy(args):
    """stand-alone line for one greedy / look-ahead shape (`--workload greedy|expectimax`), same keys as the headline"""
    D = Dist()
    importlib.import_module("2048_b200")
    from game2048 import engine
    ctx = engine.Context.get()
    n, B = args.n, args.games
    look = (args.depth, args.width, args.since_empty) if args.workload == "expectimax" else None
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start(); sampler.wait_first()
    wd = pretrained_weights(ctx, n, args.pretrain, args.alpha)
    label = (f"SURVEY 8(f) rank 1: Q_agent n={n} play with look-ahead depth={args.depth} width={args.width} "
             f"since_empty={args.since_empty} (game_logic.py:214-243)" if look else
             f"BASELINE configs[3] shape: Q_agent n={n} greedy play") + \
            f", {B} seeded games per GPU to completion, seeded init + {args.pretrain} TD lock-steps"
    r = greedy_bench(D, ctx, args, n, D.rank * B, B, wd, args.steps, args.warmup, label, sampler if D.rank == 0 else None,
                     cpu_games=min(B, args.cpu_games if not look else 16), look=look)
    if D.rank == 0:
        sampler.stop()
        r.update(n_gpus=D.world, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic")
        print(json.dumps(r), flush=True)
    D.finish()


def run_sweep(args):
    D = Dist()
    importlib.import_module("2048_b200")
    from game2048 import engine
    ctx = engine.Context.get()
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start(); sampler.wait_first()
    r = sweep_bench(D, ctx, args, args.boards, args.steps, args.warmup, sampler if D.rank == 0 else None)
    if D.rank == 0:
        sampler.stop()
        r.update(n_gpus=D.world, higher_is_better=True, scaling="weak", vs_baseline=None, data="synthetic")
        print(json.dumps(r), flush=True)
    D.finish()


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
