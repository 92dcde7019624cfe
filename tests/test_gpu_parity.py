"""GPU tier (-m gpu): every kernel group called through the C-ABI (libb2048.so via 2048_b200/game2048/cabi.py)
and compared with the CPU oracle / the reference-generated golden fixtures on identical seeded inputs.
Integer, byte and index work: bit-exact.  Float32 values/weights: bit-exact against the oracle's float32
restatement where the summation order is defined (evaluate, deterministic updates, B=1), and within a stated
tolerance against the float64 reference fixtures."""
import ctypes as C
import importlib

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    pkg = importlib.import_module("2048_b200")
    from game2048 import cabi, engine
    ctx = engine.Context.get()
    return ctx, engine, cabi


def dev_boards(ctx, rows, orc):
    return ctx.to_device(orc.pack_np(np.asarray(rows, dtype=np.int32)))


def u64(t):
    return t.cpu().numpy().view(np.uint64)


def w_dev(ctx, fx, n, seed):
    w = fx.flat(fx.init_weights32(n, seed)).astype(np.float32)
    return w, ctx.to_device(w)


# ------------------------------------------------------------------------------------------- (1) (2)
def test_lut_matches_reference_table(eng):
    ctx, engine, cabi = eng
    g = load_golden("move_table.npz")
    lut = ctx.lut.cpu().numpy().view(np.uint32)
    lines = np.stack([(lut >> 12) & 15, (lut >> 8) & 15, (lut >> 4) & 15, lut & 15], axis=1).astype(np.uint8)
    a, b = (lut >> 16) & 15, (lut >> 20) & 15
    score = np.where(a > 0, 2 << a.astype(np.int64), 0) + np.where(b > 0, 2 << b.astype(np.int64), 0)
    assert np.array_equal(lines, np.minimum(g["lines"], 15))
    assert np.array_equal(score, g["score"].astype(np.int64))
    assert np.array_equal((lut >> 24) & 1, g["changed"])
    assert np.array_equal(((lut >> 25) & 1).astype(bool), (g["lines"] == 16).any(axis=1))


def test_pack_unpack_roundtrip(eng, orc, fx):
    ctx, engine, cabi = eng
    rows = np.concatenate([fx.edge_boards(), fx.random_boards(5000, seed=1, max_exp=15)])
    b = ctx.pack(rows)
    assert np.array_equal(u64(b), orc.pack_np(rows))
    assert np.array_equal(ctx.unpack(b).cpu().numpy(), rows)
    assert ctx.pack(np.zeros((0, 4, 4), np.int32)).shape[0] == 0          # empty input


def test_move4_predicates_vs_reference_fixture(eng, orc):
    ctx, engine, cabi = eng
    g = load_golden("boards.npz")
    rows = g["boards"].astype(np.int32)
    m = len(rows)
    b = dev_boards(ctx, rows, orc)
    after, gain, flags, over = ctx.move4(b)
    ref_after = g["after"].astype(np.int32)
    ovf = (ref_after > 15).any(axis=(2, 3))
    assert np.array_equal(orc.unpack_np(u64(after).reshape(-1)).reshape(m, 4, 4, 4), np.minimum(ref_after, 15))
    assert np.array_equal(gain.cpu().numpy().astype(np.int64), g["gain"].astype(np.int64))
    fl = flags.cpu().numpy()
    for d in range(4):
        assert np.array_equal((fl >> d) & 1, g["change"][:, d])
        assert np.array_equal(((fl >> (4 + d)) & 1).astype(bool), ovf[:, d])
    assert np.array_equal(over.cpu().numpy(), g["over"])
    stats, mask = ctx.board_stats(b)
    st = stats.cpu().numpy()
    assert np.array_equal(st[:, 0], g["n_empty"].astype(np.uint8))
    assert np.array_equal(st[:, 1], g["n_pairs"].astype(np.uint8))
    assert np.array_equal(st[:, 2], g["over"])
    assert np.array_equal(st[:, 3], rows.reshape(m, -1).max(axis=1).astype(np.uint8))
    mk = mask.cpu().numpy().view(np.uint16)
    for q in range(0, m, 5):                      # Game.empty order = ascending flat cell = ascending bit
        assert [p for p in range(16) if (mk[q] >> p) & 1] == [int(p) for p in g["empties"][q] if p >= 0]


def test_move4_exhaustive_lines_all_directions(eng, orc):
    """all 65,536 lines as a row (left/right) and as a column (up/down) vs the oracle"""
    ctx, engine, cabi = eng
    lines = np.arange(65536, dtype=np.uint64)
    rows = np.stack([(lines >> 12) & 15, (lines >> 8) & 15, (lines >> 4) & 15, lines & 15], axis=1).astype(np.int32)
    filler = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], dtype=np.int32)
    b_row = np.concatenate([np.repeat(filler[None, :1], 65536, 0), rows[:, None, :],
                            np.repeat(filler[None, 1:], 65536, 0)], axis=1)
    for rows4 in (b_row, np.transpose(b_row, (0, 2, 1)).copy()):
        after, gain, flags, over = ctx.move4(dev_boards(ctx, rows4, orc))
        ra, rs, rc = orc.pre_move_batch(rows4)
        assert np.array_equal(orc.unpack_np(u64(after).reshape(-1)).reshape(-1, 4, 4, 4), np.minimum(ra, 15))
        assert np.array_equal(gain.cpu().numpy().astype(np.int64), rs)
        fl = flags.cpu().numpy()
        assert np.array_equal(np.stack([(fl >> d) & 1 for d in range(4)], 1), rc.astype(np.uint8))


# ------------------------------------------------------------------------------------------- (3)
def test_spawns_vs_oracle(eng, orc, fx):
    ctx, engine, cabi = eng
    ids = np.arange(3000, dtype=np.uint64) * np.uint64(7) + np.uint64(2 ** 33)
    b0 = ctx.spawn_initial(3000, seed=99, first_id=int(ids[0]), id_step=7)
    ref0 = np.array([orc.pack_np(orc.spawn_initial(99, int(i))[None])[0] for i in ids], dtype=np.uint64)
    assert np.array_equal(u64(b0), ref0)
    rows = fx.random_boards(3000, seed=3)
    rows[:5] = 1                                             # full boards -> 0xFFFF, unchanged
    b = dev_boards(ctx, rows, orc)
    mv = (np.arange(3000) % 300 + 1).astype(np.uint32)
    sp = ctx.spawn_philox(b, 5, ctx.to_device(ids), ctx.to_device(mv), want_spawn=True)
    sp = sp.cpu().numpy().view(np.uint16)
    got = u64(b)
    for i in range(3000):
        new, res = orc.spawn_move(5, int(ids[i]), int(mv[i]), rows[i])
        assert got[i] == orc.pack_np(new[None])[0]
        assert sp[i] == (0xFFFF if res < 0 else res)
    # replay mode (Game.replay, game_logic.py:259-260); tile 0 = skip
    rows = fx.random_boards(1000, seed=4)
    tile = (np.arange(1000) % 3).astype(np.uint8)
    pos = (np.arange(1000) * 5 % 16).astype(np.uint8)
    b = dev_boards(ctx, rows, orc)
    ctx.spawn_replay(b, ctx.to_device(tile), ctx.to_device(pos))
    exp = rows.copy().reshape(-1, 16)
    sel = tile > 0
    exp[np.nonzero(sel)[0], pos[sel]] = tile[sel]
    assert np.array_equal(u64(b), orc.pack_np(exp))


def test_sweep_vs_oracle(eng, orc, fx):
    """config-5 kernel (shared-memory LUT, persistent grid) on 200k synthetic + harvested boards"""
    ctx, engine, cabi = eng
    g = load_golden("boards.npz")
    rows = np.concatenate([fx.random_boards(200_000, seed=0), g["boards"].astype(np.int32)])
    boards = orc.pack_np(rows)
    after, gain, flags, spawned = ctx.sweep(ctx.to_device(boards), seed=42, first_index=1 << 35)
    ra, rg, rf, rs = orc.sweep(boards, seed=42, first_index=1 << 35)
    assert np.array_equal(u64(after), ra)
    assert np.array_equal(gain.cpu().numpy().view(np.uint32), rg)
    assert np.array_equal(flags.cpu().numpy(), rf)
    assert np.array_equal(u64(spawned), rs)
    a2, g2, f2, _ = ctx.sweep(ctx.to_device(boards[:1000]), spawn=False)       # no-spawn variant, small grid
    assert np.array_equal(u64(a2), ra[:1000])


# ------------------------------------------------------------------------------------------- (4)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_features_and_evaluate(eng, orc, fx, n):
    ctx, engine, cabi = eng
    g = load_golden("boards.npz")
    ref = g[f"f_{n}"]
    rows = g["boards"][:len(ref)].astype(np.int32)
    b = dev_boards(ctx, rows, orc)
    assert np.array_equal(ctx.features(n, b).cpu().numpy(), ref)
    assert cabi.table_offsets(n) == orc.table_offsets(n).tolist()
    w, wd = w_dev(ctx, fx, n, 100 + n)
    v = ctx.evaluate(n, wd, b).cpu().numpy()
    sub = slice(0, 400)
    v32 = orc.evaluate_batch(n, w, rows[sub])                    # float32, same order: bit-exact
    assert np.array_equal(v[sub], v32)
    v64 = orc.evaluate_batch(n, w.astype(np.float64), rows[sub])  # reference arithmetic
    assert np.allclose(v[sub], v64, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_update_single_entry_is_reference_update(eng, orc, fx, n):
    """m = 1: every mode equals QAgent.update (r_learning.py:207-214); key multiplicities from the fixture"""
    ctx, engine, cabi = eng
    g = load_golden("d4.npz")
    sample = g["sample"].astype(np.int32)
    offs = g[f"offs_{n}"]
    nw = cabi.num_weights(n)
    for mode in (0, 1, 2, 3, 5, 7):
        for q in (0, 3, 6, 7, 12, 20):
            wd = ctx.zeros(nw, torch.float32)
            b = dev_boards(ctx, sample[q:q + 1], orc)
            ctx.td_update(n, wd, b, ctx.to_device(np.array([1.0], np.float32)), mode=mode)
            w = wd.cpu().numpy()
            k = np.nonzero(w)[0]
            assert np.array_equal(k, g[f"keys_{n}"][offs[q]:offs[q + 1]])
            assert np.array_equal(w[k].astype(np.int32), g[f"counts_{n}"][offs[q]:offs[q + 1]])


@pytest.mark.parametrize("n", [4, 6])
def test_update_batch_modes_vs_oracle(eng, orc, fx, n):
    """m = 3000 entries with heavy key sharing (early-game boards) and NaN holes:
    deterministic sum / mean (exact fixed-point reduction by key), direct and SORTED implementations: bit-exact vs
    the oracle's exact rules and run-to-run identical; atomic sum / mean within 1e-5 (relative to the largest
    change) of the float64 sequential oracle."""
    ctx, engine, cabi = eng
    rs = np.random.RandomState(7)
    rows = fx.random_boards(3000, seed=8, p_empty=0.7, max_exp=3)
    rows[:50] = 0
    rows[50:60] = fx.edge_boards()[:10]
    boards = orc.pack_np(rows)
    dw = (rs.standard_normal(3000) * 0.3).astype(np.float32)
    dw[::17] = np.nan
    w0, _ = w_dev(ctx, fx, n, 5)
    bd, dd = ctx.to_device(boards), ctx.to_device(dw)
    exact, seq64 = {}, {}
    for rule in (1, 2):
        w32 = w0.copy()
        n_upd = orc.update_batch(n, w32, boards, dw, rule + 2)
        assert n_upd == np.isfinite(dw).sum()
        exact[rule] = w32
        w64 = w0.astype(np.float64)
        orc.update_batch(n, w64, boards, dw.astype(np.float64), rule)
        seq64[rule] = w64
        assert np.abs(w32 - w64).max() <= 1e-5 * max(np.abs(w64 - w0).max(), 1.0)     # exact rule ~ sequential rule
    D, M, S = cabi.UPD_DETERMINISTIC, cabi.UPD_MEAN, cabi.UPD_SORTED
    for rule, mode in ((1, D), (2, D | M), (1, D | S), (2, D | M | S)):
        outs, work = [], None
        for rep in range(3):                                              # rep 1, 2 reuse the workspace of rep 0
            wd = ctx.to_device(w0)
            work = ctx.td_update(n, wd, bd, dd, mode=mode, work=work)
            outs.append(wd.cpu().numpy())
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])    # deterministic
        assert np.array_equal(outs[0], exact[rule])                      # bit-exact vs the oracle
    for rule, mode in ((1, cabi.UPD_ATOMIC | cabi.UPD_SUM), (2, cabi.UPD_ATOMIC | M)):
        wd = ctx.to_device(w0)
        work = ctx.td_update(n, wd, bd, dd, mode=mode)
        got = wd.cpu().numpy()
        scale = np.abs(seq64[rule] - w0).max()
        assert np.abs(got - seq64[rule]).max() <= 1e-5 * max(scale, 1.0)  # stated tolerance
        wd2 = ctx.to_device(w0)
        ctx.td_update(n, wd2, bd, dd, mode=mode, work=work)               # workspace is reusable
        assert np.abs(wd2.cpu().numpy() - seq64[rule]).max() <= 1e-5 * max(scale, 1.0)
    # delta buffer receives the same increments
    wd, delta = ctx.to_device(w0), ctx.zeros(len(w0), torch.float32)
    ctx.td_update(n, wd, bd, dd, mode=D | M, delta=delta)
    assert np.allclose(w0 + delta.cpu().numpy(), wd.cpu().numpy(), rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------------- fused loops
def test_greedy_replay_of_reference_games(eng, orc, fx):
    """Game.trial_run games recorded from the real reference (trained float32 weights), teacher-forced on
    the recorded spawns: same moves, boards and scores (bit-exact); evaluate() values within 2e-6 rel."""
    ctx, engine, cabi = eng
    g = load_golden("greedy_n4.npz")
    n = 4
    w = fx.apply_sparse(fx.flat(fx.init_weights32(n, int(g["seed"]))), g["w_idx"], g["w_val"]).astype(np.float32)
    wd = ctx.to_device(w)
    G = len(g["odo"])
    tiles = [g["tiles"][g["t_off"][i]:g["t_off"][i + 1]] for i in range(G)]
    games = engine.GameBatch(G, ctx=ctx)
    games.set_positions(orc.pack_np(g["start"].astype(np.int32)))
    rp = engine.ReplayBuffers(ctx, tiles)
    tdir, tval, tsp = engine.greedy_play(ctx, n, wd, games, replay=rp, trace_len=rp.len, chunk=rp.len + 2)
    h = games.to_host()
    tdir, tval = tdir.cpu().numpy(), tval.cpu().numpy()
    assert np.array_equal(h["moves"], g["odo"].astype(np.uint32))
    assert np.array_equal(h["score"].astype(np.int64), g["score"])
    assert np.array_equal(h["board"], orc.pack_np(g["final"].astype(np.int32)))
    assert (h["flags"] & cabi.F_DONE).all()
    for i in range(G):
        mv = g["moves"][g["m_off"][i]:g["m_off"][i + 1]]
        assert np.array_equal(tdir[i, :len(mv)], mv)
        assert (tdir[i, len(mv):] == -2).all()                            # no sentinel after trial_run
        # the reference evaluated every valid direction; the kernel's trace holds the max per move
        r = orc.trial_replay(n, w.astype(np.float64), g["start"][i].astype(np.int32), tiles[i])
        assert np.allclose(tval[i, :len(mv)], r["values"], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("n", [4, 5, 6])
def test_greedy_philox_vs_oracle(eng, orc, fx, n):
    """Philox-driven greedy games: identical scores / move counts / final boards as the float32 oracle;
    splitting the id range over two batches (sharding) changes nothing."""
    ctx, engine, cabi = eng
    w, wd = w_dev(ctx, fx, n, 9)
    num = 96 if n < 6 else 32
    ref = orc.play_philox(n, w, seed=11, first_id=1000, num=num)
    games = engine.GameBatch(num, seed=11, ctx=ctx).init(first_id=1000)
    engine.greedy_play(ctx, n, wd, games, chunk=64)                       # several launches per game
    h = games.to_host()
    assert np.array_equal(h["score"].astype(np.int64), ref["scores"])
    assert np.array_equal(h["moves"].astype(np.int32), ref["moves"])
    assert np.array_equal(h["board"], ref["boards"])
    c = games.read_counters()
    assert c["moves"] == ref["total_moves"] and c["evals"] == ref["n_eval"] and c["finished"] == num
    assert c["score_sum"] == ref["scores"].sum() and c["active"] == 0
    hist = np.bincount(ref["max_tile"], minlength=17)
    assert np.array_equal(h["tile_hist"], hist)
    a = engine.GameBatch(num // 3, seed=11, ctx=ctx).init(first_id=1000)
    b = engine.GameBatch(num - num // 3, seed=11, ctx=ctx).init(first_id=1000 + num // 3)
    engine.greedy_play(ctx, n, wd, a)
    engine.greedy_play(ctx, n, wd, b)
    assert np.array_equal(np.concatenate([a.to_host()["board"], b.to_host()["board"]]), ref["boards"])
    # limit_tile / step_limit (trial_run arguments)
    lim = engine.GameBatch(16, seed=11, ctx=ctx).init(first_id=1000)
    engine.greedy_play(ctx, n, wd, lim, limit_tile=5, step_limit=40)
    ref_l = orc.play_philox(n, w, seed=11, first_id=1000, num=16, limit_tile=5, step_limit=40)
    assert np.array_equal(lim.to_host()["board"], ref_l["boards"])
    # odd launch budgets and limits (the look-ahead kernel commits up to two moves per iteration and must stop exactly)
    odd = engine.GameBatch(24, seed=11, ctx=ctx).init(first_id=1000)
    engine.greedy_play(ctx, n, wd, odd, step_limit=41, chunk=7)
    ref_o = orc.play_philox(n, w, seed=11, first_id=1000, num=24, step_limit=41)
    ho = odd.to_host()
    assert np.array_equal(ho["board"], ref_o["boards"]) and np.array_equal(ho["moves"].astype(np.int32), ref_o["moves"])
    assert np.array_equal(ho["score"].astype(np.int64), ref_o["scores"])


@pytest.mark.parametrize("n", [3, 4, 6])
def test_gpu_philox_games_replay_through_reference_rules(eng, orc, fx, n):
    """Reverse direction of the replay protocol: games the GPU played on its own Philox spawns,
    recorded as (starting_position, moves, tiles) like a reference Game, are re-played by the oracle's
    teacher-forced trial_run (pinned to the reference's recorded games in test_oracle_golden): it must choose
    the same move at every step and end on the same board and score."""
    ctx, engine, cabi = eng
    w, wd = w_dev(ctx, fx, n, 13)
    num, L = 20, 6000
    games = engine.GameBatch(num, seed=77, ctx=ctx).init(first_id=500)
    starts = games.to_host()["board"].copy()
    tdir, _, tsp = engine.greedy_play(ctx, n, wd, games, trace_len=L)
    h = games.to_host()
    tdir, tsp = tdir.cpu().numpy(), tsp.cpu().numpy().view(np.uint16)
    for j in range(num):
        odo = int(h["moves"][j])
        assert 0 < odo < L
        tiles = [(int(s >> 8), (int(s & 15) // 4, int(s & 15) % 4)) for s in tsp[j, :odo]]
        ref = orc.trial_replay(n, w, orc.unpack_np(starts[j:j + 1])[0], tiles)
        assert ref["odometer"] == odo and ref["score"] == int(h["score"][j])
        assert np.array_equal(ref["moves"], tdir[j, :odo].astype(np.int32))
        assert np.array_equal(orc.pack_np(ref["row"][None].astype(np.int32))[0], h["board"][j])


def _run_replay_episodes(ctx, engine, cabi, orc, g, n, wd, mode, episodes):
    out = []
    for i in range(episodes):
        tiles = g["tiles"][g["t_off"][i]:g["t_off"][i + 1]]
        games = engine.GameBatch(1, ctx=ctx)
        games.set_positions(orc.pack_np(g["start"][i:i + 1].astype(np.int32)))
        rp = engine.ReplayBuffers(ctx, [tiles])
        L = rp.len + 1
        td, tv, tw = ctx.empty((1, L), torch.int8).fill_(-2), ctx.zeros((1, L), torch.float32), ctx.zeros((1, L), torch.float32)
        ts = ctx.zeros((1, L), torch.int16)
        tr = engine.TDTrainer(ctx, n, wd, games, float(g["alpha"]), mode)
        for _ in range(len(tiles) + 2):
            tr.step(replay=rp, trace=(td, tv, tw, ts, L))
        out.append((games.to_host(), td.cpu().numpy()[0], tv.cpu().numpy()[0], tw.cpu().numpy()[0]))
    return out


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_td_episode_teacher_forced_vs_reference(eng, orc, fx, n):
    """QAgent.episode recorded from the real reference, replayed on the GPU with B = 1 (deterministic, sum rule):
    moves incl. the -1 sentinel, boards, scores bit-exact; per-step dw, values and final weights bit-exact vs the
    float32 oracle (same exact-sum rule) and within tolerance of the float64 reference
    (|dw err| <= 1e-4 (1 + |dw|); weights <= 2e-3 absolute after all episodes).  The atomic mode (float adds,
    duplicates merged before the add) stays within 1e-5 of the deterministic weights."""
    ctx, engine, cabi = eng
    g = load_golden(f"episodes_n{n}.npz")
    E = min(len(g["odo"]), 12)
    w0 = fx.flat(fx.init_weights32(n, int(g["seed"]))).astype(np.float32)
    wd = ctx.to_device(w0)
    res = _run_replay_episodes(ctx, engine, cabi, orc, g, n, wd, cabi.UPD_DETERMINISTIC | cabi.UPD_SUM, E)
    w32 = w0.copy()
    u = 0
    for i, (h, td, tv, tw) in enumerate(res):
        mv = g["moves"][g["m_off"][i]:g["m_off"][i + 1]]
        tiles = g["tiles"][g["t_off"][i]:g["t_off"][i + 1]]
        r32 = orc.episode_replay(n, w32, float(g["alpha"]), g["start"][i].astype(np.int32), tiles, rule=3)
        odo = r32["odometer"]
        assert np.array_equal(td[:odo + 1], r32["moves"]) and td[odo] == -1
        assert np.array_equal(tw[1:odo + 1], r32["dws"][1:]) and np.isnan(tw[0])
        assert np.array_equal(tv[:odo], r32["values"][:odo])
        assert h["flags"][0] & cabi.F_DONE
        if np.array_equal(r32["moves"], mv):                              # float32 followed the reference game
            assert h["moves"][0] == g["odo"][i] and h["score"][0] == g["score"][i]
            assert h["board"][0] == orc.pack_np(g["final"][i:i + 1].astype(np.int32))[0]
            ref_dw = g["upd_dw"][u:u + odo]
            assert np.all(np.abs(tw[1:odo + 1] - ref_dw) <= 1e-4 * (1 + np.abs(ref_dw)))
        u += len(mv) - 1
    assert np.array_equal(wd.cpu().numpy(), w32)                          # bit-exact vs float32 oracle
    if E == len(g["odo"]):
        ref_w = fx.apply_sparse(w0.astype(np.float64), g["w_idx"], g["w_val"])
        assert np.abs(wd.cpu().numpy() - ref_w).max() <= 2e-3
    wa = ctx.to_device(w0)
    _run_replay_episodes(ctx, engine, cabi, orc, g, n, wa, cabi.UPD_ATOMIC | cabi.UPD_SUM, min(E, 3))
    w3 = w0.copy()
    for i in range(min(E, 3)):
        orc.episode_replay(n, w3, float(g["alpha"]), g["start"][i].astype(np.int32),
                           g["tiles"][g["t_off"][i]:g["t_off"][i + 1]], rule=0)       # the reference's sequential adds
    assert np.abs(wa.cpu().numpy() - w3).max() <= 1e-4


@pytest.mark.parametrize("n,B", [(4, 1), (4, 64), (5, 48), (3, 33)])
def test_td_lockstep_deterministic_bit_exact(eng, orc, fx, n, B):
    """Philox lock-step TD with in-place restart: deterministic modes are bit-exact vs the float32 oracle
    (weights, boards, scores, ids, labels, counters) and run-to-run identical."""
    ctx, engine, cabi = eng
    steps = 300
    for rule, mode, alpha in ((3, cabi.UPD_DETERMINISTIC | cabi.UPD_SUM, 0.25 / max(B, 4)),
                              (4, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN, 0.25),
                              (4, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN | cabi.UPD_SORTED, 0.25)):
        w0 = fx.flat(fx.init_weights32(n, 31)).astype(np.float32)
        ref_w = w0.copy()
        ls = orc.LockStep(n, ref_w, alpha, 77, B, first_id=5, id_stride=B, segmented=rule, threads=2)
        ls.run(steps)
        outs = []
        for rep in range(2):
            wd = ctx.to_device(w0)
            games = engine.GameBatch(B, seed=77, id_stride=B, ctx=ctx).init(first_id=5)
            tr = engine.TDTrainer(ctx, n, wd, games, alpha, mode)
            tr.run(steps // 2)
            for _ in range(steps - steps // 2):
                tr.step()
            outs.append((wd.cpu().numpy(), games.to_host(), games.read_counters()))
        (wa, ha, ca), (wb, hb, cb) = outs
        assert np.array_equal(wa, wb) and np.array_equal(ha["board"], hb["board"])
        assert np.array_equal(ha["board"], ls.board) and np.array_equal(ha["game_id"], ls.game_id)
        assert np.array_equal(ha["score"].astype(np.int64), ls.score)
        assert np.array_equal(ha["moves"].astype(np.int32), ls.odo)
        assert np.array_equal(ha["old_label"], ls.old_label)
        assert np.array_equal(wa, ref_w)
        assert ca["updates"] == ls.n_updates and ca["moves"] == ls.n_moves
        assert ca["finished"] == ls.fin[0] and ca["score_sum"] == ls.fin[1] and ca["moves_sum"] == ls.fin[2]
        assert np.array_equal(ha["tile_hist"], ls.hist)


@pytest.mark.parametrize("n,B,steps,force", [(4, 64, 200, "generic"), (5, 5000, 60, None), (6, 700, 40, None),
                                             (2, 4096, 50, None), (3, 4096, 40, None), (5, 17000, 30, None),
                                             (4, 3000, 80, None), (3, 1000, 60, None), (2, 2000, 50, None),
                                             (4, 8, 300, None),
                                             # BASELINE configs[2] shapes: n=5 with 65,536 games on 1 / 2 GPUs
                                             (5, 32768, 12, None), (5, 65536, 8, None),
                                             # n=6 in the generic layout (more slots per CTA than one round; forced)
                                             (6, 4096, 12, None), (6, 300, 40, "generic"),
                                             # both apply phases forced where the launcher would pick the other one
                                             (4, 4096, 40, "lists"), (4, 1000, 80, "scan"), (4, 64, 100, "generic+scan"),
                                             (5, 6000, 30, "scan"), (3, 2000, 40, "scan"), (2, 500, 60, "scan"),
                                             (4, 20000, 24, "even"), (4, 20000, 24, None)])
def test_td_persistent_paths_bit_exact(eng, orc, fx, monkeypatch, n, B, steps, force):
    """b2048_td_run's persistent kernel gives the oracle's bits in the deterministic modes in both slot layouts:
    one phase-B round per CTA with the state in registers ((6, 700), (4, 3000), (3, 1000), (2, 2000), (4, 8)), and
    the generic layout (rounds over global staging and key lists: forced with B2048_RUN_GENERIC, or the larger
    shapes, among them n=5 at 32,768 / 65,536 games and n=6 at 4,096)."""
    ctx, engine, cabi = eng
    layout = 0
    for word in (force or "").split("+"):
        layout |= {"": 0, "generic": cabi.RUN_GENERIC, "scan": cabi.RUN_SCAN, "lists": cabi.RUN_LISTS,
                   "even": cabi.RUN_EVEN}[word]
    for rule, mode, alpha in ((4, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN, 0.25),
                              (3, cabi.UPD_DETERMINISTIC | cabi.UPD_SUM, 0.25 / B)):
        w0 = fx.flat(fx.init_weights32(n, 41)).astype(np.float32)
        ref_w = w0.copy()
        ls = orc.LockStep(n, ref_w, alpha, 91, B, segmented=rule, threads=8)
        ls.run(steps)
        wd = ctx.to_device(w0)
        games = engine.GameBatch(B, seed=91, ctx=ctx).init()
        tr = engine.TDTrainer(ctx, n, wd, games, alpha, mode | layout)
        tr.run(steps // 3)
        tr.run(steps - steps // 3)                                        # a second launch resumes from memory
        h, c = games.to_host(), games.read_counters()
        assert np.array_equal(h["board"], ls.board) and np.array_equal(h["game_id"], ls.game_id)
        assert np.array_equal(h["old_label"], ls.old_label)
        assert np.array_equal(wd.cpu().numpy(), ref_w)
        assert c["updates"] == ls.n_updates and c["moves"] == ls.n_moves and c["finished"] == ls.fin[0]


def test_td_lockstep_atomic_within_tolerance(eng, orc, fx):
    """atomic modes: same sums, unordered -> weights within 1e-5 (relative to the largest change) of the
    float64 oracle over a short stable run"""
    ctx, engine, cabi = eng
    n, B, steps = 4, 256, 40
    for rule, mode, alpha in ((1, cabi.UPD_ATOMIC | cabi.UPD_SUM, 0.25 / B), (2, cabi.UPD_ATOMIC | cabi.UPD_MEAN, 0.25)):
        w0 = fx.flat(fx.init_weights32(n, 32)).astype(np.float32)
        w64 = w0.astype(np.float64)
        ls = orc.LockStep(n, w64, alpha, 78, B, segmented=rule, threads=4)
        ls.run(steps)
        wd = ctx.to_device(w0)
        games = engine.GameBatch(B, seed=78, ctx=ctx).init()
        engine.TDTrainer(ctx, n, wd, games, alpha, mode).run(steps)
        got = wd.cpu().numpy()
        scale = np.abs(w64 - w0).max()
        assert np.abs(got - w64).max() <= 1e-4 * max(scale, 1.0)
        assert games.read_counters()["updates"] == ls.n_updates


def test_finished_game_log_and_delta_apply(eng, orc, fx):
    ctx, engine, cabi = eng
    n, B = 4, 64
    w0 = fx.flat(fx.init_weights32(n, 33)).astype(np.float32)
    wd = ctx.to_device(w0)
    games = engine.GameBatch(B, seed=3, ctx=ctx, fin_cap=4096).init()
    delta = ctx.zeros(len(w0), torch.float32)
    tr = engine.TDTrainer(ctx, n, wd, games, 0.25, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN, delta=delta)
    tr.run(600)
    rec = games.drain_finished()
    c = games.read_counters()
    assert len(rec) == c["finished"] > 0 and rec[:, 2].sum() == c["score_sum"] and rec[:, 3].sum() == c["moves_sum"]
    assert len(games.drain_finished()) == 0
    # multi-GPU sync kernel: w_sync += delta_sum; w = w_sync; delta = 0
    w_sync = ctx.to_device(w0)
    dsum = delta.clone() * 2
    check = w0 + 2 * delta.cpu().numpy()
    cabi.check(ctx.lib.b2048_delta_apply(engine.dptr(wd), engine.dptr(w_sync), engine.dptr(delta), engine.dptr(dsum),
                                         None, len(w0), engine.cur_stream()))
    assert np.array_equal(wd.cpu().numpy(), check) and np.array_equal(w_sync.cpu().numpy(), check)
    assert not delta.any().item()
    # per-key mean over contributing ranks: pack -> (sum over 3 fake ranks) -> apply
    d1 = torch.zeros(1000, device=ctx.device); d1[::3] = 0.5
    packed = ctx.zeros(2000, torch.float32)
    cabi.check(ctx.lib.b2048_delta_pack(engine.dptr(d1), engine.dptr(packed), 1000, engine.cur_stream()))
    assert torch.equal(packed[:1000], d1) and torch.equal(packed[1000:], (d1 != 0).float())
    tot = packed * 3                                              # three identical ranks
    w1, ws = ctx.zeros(1000, torch.float32), ctx.zeros(1000, torch.float32)
    cabi.check(ctx.lib.b2048_delta_apply(engine.dptr(w1), engine.dptr(ws), engine.dptr(d1), engine.dptr(tot),
                                         engine.dptr(tot[1000:]), 1000, engine.cur_stream()))
    assert torch.equal(w1[::3], torch.full_like(w1[::3], 0.5)) and not w1[1::3].any().item() and not d1.any().item()


@pytest.mark.parametrize("n", [4, 6])
def test_look_forward_and_expectimax_play_vs_oracle(eng, orc, fx, n):
    """Game.look_forward / trial_run with look-ahead (game_logic.py:150-183, 214-243): device values and whole games
    == the float32 oracle, whose recursion is pinned to the reference by tests/golden/lookforward.npz."""
    ctx, engine, cabi = eng
    w, wd = w_dev(ctx, fx, n, 12)
    rng = np.random.default_rng(5)
    # crowded afterstates: 0-6 empty cells, exponents 1..9
    rows = rng.integers(1, 10, size=(300, 16)).astype(np.int32)
    rows[rng.random((300, 16)) < rng.random((300, 1)) * 0.4] = 0
    rows[np.arange(300), rng.integers(0, 16, size=300)] = 0             # an afterstate always has an empty cell
    boards = orc.pack_np(rows)
    ids = rng.integers(0, 1 << 40, size=300).astype(np.uint64)
    mv = rng.integers(0, 5000, size=300).astype(np.uint32)
    rd = rng.integers(0, 4, size=300).astype(np.int32)
    for depth, width, since_empty in ((0, 1, 6), (1, 1, 6), (1, 4, 16), (2, 2, 6), (2, 4, 8), (3, 3, 7), (4, 2, 5)):
        ref = orc.look_forward(n, w, rows, None, depth, width, since_empty, seed=21, ids=ids, move_no=mv, root_dir=rd)
        got = ctx.look_forward(n, wd, ctx.to_device(boards), ctx.to_device(ids), ctx.to_device(mv),
                               ctx.to_device(rd.astype(np.uint8)), depth, width, since_empty, seed=21).cpu().numpy()
        assert np.array_equal(got, ref), (depth, width, since_empty, np.abs(got - ref).max())
    num = 24
    for depth, width, since_empty in ((1, 2, 8), (2, 2, 6), (3, 4, 5)):
        ref = orc.play_expectimax(n, w, 31, 100, num, depth, width, since_empty, threads=8)
        games = engine.GameBatch(num, seed=31, ctx=ctx).init(first_id=100)
        tdir, _, tsp = engine.expectimax_play(ctx, n, wd, games, depth, width, since_empty, chunk=50, trace_len=4096)
        h, c = games.to_host(), games.read_counters()
        assert np.array_equal(h["board"], ref["boards"]) and np.array_equal(h["score"].astype(np.int64), ref["scores"])
        assert np.array_equal(h["moves"].astype(np.int32), ref["moves"])
        assert c["moves"] == ref["total_moves"] and c["finished"] == num and c["score_sum"] == ref["scores"].sum()
        td = tdir.cpu().numpy()
        assert all((td[j, :h["moves"][j]] >= 0).all() and (td[j, h["moves"][j]:] == -2).all() for j in range(num))
    # more than 2,048 games: the launcher's high-occupancy variant, which at depth 3 also lays the third level out and
    # gathers on compacted leaves; 10 moves per game keep the oracle's share short
    if n == 4:
        big = 2304
        ref = orc.play_expectimax(n, w, 32, 7, big, 3, 4, 9, step_limit=10, threads=orc.max_threads())
        games = engine.GameBatch(big, seed=32, ctx=ctx).init(first_id=7)
        engine.expectimax_play(ctx, n, wd, games, 3, 4, 9, step_limit=10)
        h, c = games.to_host(), games.read_counters()
        assert np.array_equal(h["board"], ref["boards"]) and np.array_equal(h["score"].astype(np.int64), ref["scores"])
        assert c["moves"] == ref["total_moves"] == 10 * big
    # depth 0 == greedy play
    g0 = engine.GameBatch(num, seed=31, ctx=ctx).init(first_id=100)
    engine.expectimax_play(ctx, n, wd, g0, 0)
    g1 = engine.GameBatch(num, seed=31, ctx=ctx).init(first_id=100)
    engine.greedy_play(ctx, n, wd, g1)
    assert np.array_equal(g0.to_host()["board"], g1.to_host()["board"])


def test_argument_errors(eng):
    ctx, engine, cabi = eng
    L = ctx.lib
    assert L.b2048_num_feat(7) == -1 and L.b2048_num_weights(1) == -1
    assert L.b2048_features(7, None, 1, None, None) == -1
    assert L.b2048_move4(None, None, 1, None, None, None, None, None) == -1
    assert L.b2048_evaluate(4, None, None, 0, None, None) == -1
    b = ctx.zeros(4, torch.int64)
    assert L.b2048_td_update(4, engine.dptr(ctx.zeros(cabi.num_weights(4), torch.float32)), None, engine.dptr(b),
                             engine.dptr(ctx.zeros(4, torch.float32)), 4, 3, None, 0, None) == -3     # EWORK
    assert L.b2048_td_update(4, engine.dptr(b), None, engine.dptr(b), engine.dptr(b), 4, 4, None, 0, None) == -1  # SORTED needs DETERMINISTIC
    assert b"workspace" in L.b2048_strerror(-3)
    # b2048_td_run: launch plan and argument checks
    assert L.b2048_td_run_launches(4, 4096, cabi.UPD_ATOMIC | cabi.UPD_MEAN, 100) == 1            # persistent kernel
    assert L.b2048_td_run_launches(4, 4096, cabi.UPD_ATOMIC | cabi.UPD_MEAN | cabi.RUN_STEPWISE, 100) == 300
    assert L.b2048_td_run_launches(4, 4096, cabi.UPD_ATOMIC | cabi.UPD_SUM | cabi.RUN_STEPWISE, 100) == 200
    assert L.b2048_td_run_launches(4, 0, 0, 100) == 0 and L.b2048_td_run_launches(7, 16, 0, 1) == -1
    games = engine.GameBatch(8, seed=1, ctx=ctx).init()
    w = ctx.zeros(cabi.num_weights(4), torch.float32)
    tr = engine.TDTrainer(ctx, 4, w, games, 0.25, cabi.UPD_ATOMIC | cabi.UPD_MEAN)
    assert L.b2048_td_run(4, engine.dptr(w), None, engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                          engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), None, 0, None) == -3   # no workspace
    assert L.b2048_td_run(4, engine.dptr(w), None, engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 256, 5,
                          engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                          None) == -1                                                               # unknown mode bit
    wmis = ctx.zeros(cabi.num_weights(4) + 4, torch.float32)[1:]                                    # 4-byte aligned only
    assert L.b2048_td_run(4, engine.dptr(wmis), None, engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                          engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                          None) == -1                                                               # 16-byte alignment
    # multi-GPU exchange entry points
    z = ctx.zeros(64, torch.float32)
    assert L.b2048_delta_pack_bits(engine.dptr(z), engine.dptr(z), None, engine.dptr(z), 64, None) == -1
    assert L.b2048_delta_pack_bits(None, None, None, None, 0, None) == 0
    assert L.b2048_delta_apply_bits(engine.dptr(z), engine.dptr(z), engine.dptr(z), engine.dptr(z), 0, 64, None) == -1   # world 0
    pz = cabi.Peers()
    assert L.b2048_td_run_peers(4, engine.dptr(w), engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                                engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                                C.byref(pz), 4, 0, 1, None) == -1                                     # world 0
    fl = ctx.zeros(cabi.PEER_FLAG_WORDS, torch.int32)
    ws = w.clone()
    pz.w[0], pz.w_sync[0], pz.flags[0], pz.world, pz.rank = w.data_ptr(), ws.data_ptr(), fl.data_ptr(), 1, 0
    assert L.b2048_td_run_peers(4, engine.dptr(w), engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                                engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                                C.byref(pz), 4, 4, 1, None) == -1                                     # since_sync >= sync_every
    assert L.b2048_td_run_peers(4, engine.dptr(ws), engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                                engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                                C.byref(pz), 4, 0, 1, None) == -1                                     # weights != peers.w[rank]
    # world 1 through the persistent launch: 5 lock-steps, one exchange after the 4th (w_sync catches up with w there)
    assert L.b2048_td_run_peers(4, engine.dptr(w), engine.dptr(ctx.lut), C.byref(games.c), C.c_float(0.25), 2, 5,
                                engine.dptr(tr.upd_board), engine.dptr(tr.upd_dw), engine.dptr(tr.work), tr.work.numel(),
                                C.byref(pz), 4, 0, 1, engine.cur_stream()) == 0
    torch.cuda.synchronize()
    assert int(fl[cabi.PEER_FAULT].item()) == 0 and bool((w != ws).any()) and games.read_counters()["updates"] > 0
    # look-ahead entry points
    assert L.b2048_look_forward(4, engine.dptr(w), engine.dptr(ctx.lut), None, None, None, None, 0, 5, 1, 6, 0, None, None) == -1
    assert L.b2048_look_forward(4, engine.dptr(w), engine.dptr(ctx.lut), None, None, None, None, 0, 2, 5, 6, 0, None, None) == -1
    assert L.b2048_look_forward(4, engine.dptr(w), engine.dptr(ctx.lut), None, None, None, None, 0, 2, 2, 6, 0, None, None) == 0
    assert L.b2048_expectimax_play(4, engine.dptr(w), engine.dptr(ctx.lut), C.byref(games.c), 4, 0, 100, 9, 1, 6, None, None, 0,
                                   None) == -1


@pytest.mark.parametrize("B", [3, 3000])
def test_greedy_stops_at_the_2_16_escape(eng, orc, fx, B):
    """A position whose only legal moves would create a 2^16 tile (the reference raises KeyError one move later,
    game_logic.py:129): both greedy kernels (look-ahead kernel for 3 games, 4-lane kernel for 3,000) stop the game
    untouched, flag it DONE | OVERFLOW and count it."""
    ctx, engine, cabi = eng
    n = 4
    w, wd = w_dev(ctx, fx, n, 9)
    row = np.array([[15, 15, 1, 2], [3, 4, 5, 6], [1, 2, 3, 4], [5, 6, 7, 8]], np.int32)
    start = orc.pack_np(row[None])[0]
    games = engine.GameBatch(B, seed=1, ctx=ctx).init()
    games.set_positions(np.full(B, start, dtype=np.uint64))
    engine.greedy_play(ctx, n, wd, games)
    h, c = games.to_host(), games.read_counters()
    assert (h["board"] == start).all() and (h["score"] == 0).all() and (h["moves"] == 0).all()
    assert (h["flags"] == (cabi.F_DONE | cabi.F_OVERFLOW)).all()
    assert c["finished"] == B and c["overflow"] == B and c["moves"] == 0 and c["active"] == 0
    assert h["tile_hist"][16] == B
    # one move before the escape: [14 14 . .] merges to 2^15 legally, then the game goes on
    row2 = row.copy()
    row2[0, :2] = 14
    g2 = engine.GameBatch(B, seed=1, ctx=ctx).init()
    g2.set_positions(np.full(B, orc.pack_np(row2[None])[0], dtype=np.uint64))
    engine.greedy_play(ctx, n, wd, g2, step_limit=1)
    h2 = g2.to_host()
    assert (h2["moves"] == 1).all() and (h2["score"] == 1 << 15).all() and not (h2["flags"] & cabi.F_OVERFLOW).any()


@pytest.mark.parametrize("stepwise", [False, True])
def test_td_treats_the_2_16_escape_as_the_end_of_the_episode(eng, orc, fx, stepwise):
    """lock-step TD on a position whose only legal moves would create 2^16: the slot ends its episode there (no update on
    the first move of an episode: there is no previous afterstate), is counted as overflowed and restarts in place"""
    ctx, engine, cabi = eng
    n, B = 4, 8
    w, wd = w_dev(ctx, fx, n, 9)
    w_before = wd.clone()
    row = np.array([[15, 15, 1, 2], [3, 4, 5, 6], [1, 2, 3, 4], [5, 6, 7, 8]], np.int32)
    games = engine.GameBatch(B, seed=4, ctx=ctx).init()
    ids0 = games.to_host()["game_id"].copy()
    games.set_positions(np.full(B, orc.pack_np(row[None])[0], dtype=np.uint64))
    mode = cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN | (cabi.RUN_STEPWISE if stepwise else 0)
    engine.TDTrainer(ctx, n, wd, games, 0.25, mode).run(1)
    h, c = games.to_host(), games.read_counters()
    assert c["finished"] == B and c["overflow"] == B and c["moves"] == 0 and c["updates"] == 0
    assert np.array_equal(h["game_id"], ids0 + B) and (h["moves"] == 0).all() and (h["score"] == 0).all()
    assert (h["flags"] == 0).all()                                      # fresh games
    assert np.array_equal(h["board"], np.array([orc.pack_np(orc.spawn_initial(4, int(i))[None])[0] for i in h["game_id"]]))
    assert torch.equal(wd, w_before) and h["tile_hist"][16] == B


def test_empty_batches_are_no_ops(eng):
    """m = 0 / B = 0 / steps = 0 / count = 0: every batch entry point returns 0 and touches nothing"""
    ctx, engine, cabi = eng
    L = ctx.lib
    lut, st = engine.dptr(ctx.lut), engine.cur_stream()
    w = ctx.zeros(cabi.num_weights(2), torch.float32)
    assert L.b2048_pack(None, None, 0, st) == 0 and L.b2048_unpack(None, None, 0, st) == 0
    assert L.b2048_move4(lut, None, 0, None, None, None, None, st) == 0
    assert L.b2048_board_stats(None, 0, None, None, st) == 0
    assert L.b2048_spawn_philox(None, 0, 1, None, None, None, st) == 0
    assert L.b2048_spawn_initial(None, 0, 1, 0, 1, st) == 0 and L.b2048_spawn_replay(None, 0, None, None, st) == 0
    assert L.b2048_sweep(lut, None, 0, 0, 0, None, None, None, None, st) == 0
    assert L.b2048_features(2, None, 0, None, st) == 0 and L.b2048_evaluate(2, engine.dptr(w), None, 0, None, st) == 0
    assert L.b2048_td_update(2, engine.dptr(w), None, None, None, 0, 3, None, 0, st) == 0
    assert L.b2048_look_forward(2, engine.dptr(w), lut, None, None, None, None, 0, 2, 2, 6, 0, None, st) == 0
    assert L.b2048_delta_pack(None, None, 0, st) == 0 and L.b2048_delta_pack_diff(None, None, None, 0, st) == 0
    assert L.b2048_delta_apply(None, None, None, None, None, 0, st) == 0
    assert L.b2048_delta_pack_bits(None, None, None, None, 0, st) == 0
    assert L.b2048_delta_apply_bits(None, None, None, None, 1, 0, st) == 0
    empty = engine.GameBatch(0, seed=1, ctx=ctx)
    assert L.b2048_games_init(C.byref(empty.c), 0, 1, st) == 0
    assert L.b2048_greedy_play(2, engine.dptr(w), lut, C.byref(empty.c), 8, 0, 100, None, None, None, None, 0, st) == 0
    assert L.b2048_expectimax_play(2, engine.dptr(w), lut, C.byref(empty.c), 8, 0, 100, 2, 2, 6, None, None, 0, st) == 0
    dummy = ctx.zeros(1, torch.int64)
    assert L.b2048_td_run(2, engine.dptr(w), None, lut, C.byref(empty.c), C.c_float(0.25), 2, 5, engine.dptr(dummy),
                          engine.dptr(dummy), None, 0, st) == 0
    games = engine.GameBatch(4, seed=1, ctx=ctx).init()
    before = games.to_host()
    tr = engine.TDTrainer(ctx, 2, w, games, 0.25, cabi.UPD_DETERMINISTIC | cabi.UPD_MEAN)
    tr.run(0)
    engine.greedy_play(ctx, 2, w, engine.GameBatch(4, seed=1, ctx=ctx).init(), chunk=0, max_launches=1)
    torch.cuda.synchronize()
    after = games.to_host()
    assert np.array_equal(before["board"], after["board"]) and not w.any().item()
    assert games.read_counters()["moves"] == 0
