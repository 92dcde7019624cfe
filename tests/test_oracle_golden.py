"""Pin the CPU oracle (oracle/oracle.c) against fixtures produced by the real reference
(tests/golden/gen_golden.py).  CPU only; runs on every machine."""
import numpy as np
import pytest

from conftest import load_golden


def test_move_table_exhaustive(orc):
    """all 65,536 entries of Game.table (game_logic.py:18-39, :51)"""
    g = load_golden("move_table.npz")
    lines, score, changed = orc.create_table()
    assert np.array_equal(lines, g["lines"])
    assert np.array_equal(score, g["score"])
    assert np.array_equal(changed, g["changed"])
    assert int(changed.sum()) == 21210 and int(score.max()) == 131072
    assert int((lines == 16).any(axis=1).sum()) == 767


def test_survey_golden_vectors(orc):
    """the hand-listed vectors of SURVEY.md 8c"""
    lines, score, changed = orc.create_table()
    key = lambda a, b, c, d: (a << 12) | (b << 8) | (c << 4) | d
    for line, (out, s, ch) in {
        (1, 1, 1, 1): ((2, 2, 0, 0), 8, True), (1, 1, 1, 0): ((2, 1, 0, 0), 4, True),
        (2, 1, 1, 0): ((2, 2, 0, 0), 4, True), (1, 0, 0, 1): ((2, 0, 0, 0), 4, True),
        (0, 1, 2, 3): ((1, 2, 3, 0), 0, True), (1, 2, 3, 4): ((1, 2, 3, 4), 0, False),
        (0, 0, 0, 0): ((0, 0, 0, 0), 0, False), (14, 14, 14, 14): ((15, 15, 0, 0), 65536, True),
        (15, 15, 0, 0): ((16, 0, 0, 0), 65536, True), (15, 15, 15, 15): ((16, 16, 0, 0), 131072, True),
    }.items():
        k = key(*line)
        assert tuple(lines[k]) == out and score[k] == s and bool(changed[k]) == ch
    b = np.array([[1, 1, 2, 2], [0, 0, 0, 0], [3, 0, 3, 0], [1, 2, 3, 4]])
    after, ns, ch = orc.pre_move_batch(b[None])
    assert after[0, 0].tolist() == [[2, 3, 0, 0], [0, 0, 0, 0], [4, 0, 0, 0], [1, 2, 3, 4]] and ns[0, 0] == 28
    assert after[0, 1].tolist() == [[1, 1, 2, 2], [3, 2, 4, 4], [1, 0, 0, 0], [0, 0, 0, 0]] and ns[0, 1] == 16
    assert after[0, 2].tolist() == [[0, 0, 2, 3], [0, 0, 0, 0], [0, 0, 0, 4], [1, 2, 3, 4]] and ns[0, 2] == 28
    assert after[0, 3].tolist() == [[0, 0, 0, 0], [1, 0, 0, 0], [3, 1, 2, 2], [1, 2, 4, 4]] and ns[0, 3] == 16
    assert ch[0].all()
    x = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 0]])
    f6 = orc.features_batch(6, x[None])[0]
    assert f6[:17].tolist() == [5533, 9902, 14271, 18624, 4660, 22136, 39612, 57072, 5414, 9783, 14152, 22890,
                                27259, 31628, 40366, 44735, 49088]
    assert f6[17:21].tolist() == [402855, 472760, 682475, 752380]
    assert f6[21:].tolist() == [755086, 1334281, 1913476, 3071865, 3648315, 4224752, 623959, 1203154, 2940739,
                                3519934, 5257503, 5836474]
    assert orc.features_batch(2, x[None])[0][:4].tolist() == [21, 38, 55, 72]
    assert orc.features_batch(3, x[None])[0][:4].tolist() == [345, 618, 891, 1164]
    assert orc.pack_np(x[None])[0] == 0x123456789ABCDEF0


def test_board_ops(orc):
    """pre_move x4, game_over, empty, empty_count, adjacent_pair_count on 13k seeded boards"""
    g = load_golden("boards.npz")
    boards = g["boards"].astype(np.int32)
    after, ns, ch = orc.pre_move_batch(boards)
    assert np.array_equal(after, g["after"].astype(np.int32))
    assert np.array_equal(ns, g["gain"].astype(np.int64))
    assert np.array_equal(ch.astype(np.uint8), g["change"])
    for q in range(0, boards.shape[0], 7):
        assert orc.game_over(boards[q]) == bool(g["over"][q])
        assert orc.empty_count(boards[q]) == g["n_empty"][q]
        assert orc.adjacent_pair_count(boards[q]) == g["n_pairs"][q]
        em = [4 * i + j for i, j in orc.empty(boards[q])]
        assert em == [int(p) for p in g["empties"][q] if p >= 0]
    assert g["over"].sum() > 0 and (g["over"] == 0).sum() > 0


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_features(orc, n):
    g = load_golden("boards.npz")
    ref = g[f"f_{n}"]
    got = orc.features_batch(n, g["boards"][:ref.shape[0]].astype(np.int32))
    assert np.array_equal(got, ref)
    offs = orc.table_offsets(n)
    sizes = np.diff(offs)
    assert (got < sizes[None, :]).all() and offs[-1] == orc.num_weights(n)


def test_weight_counts(orc):
    """6,144 / 212,992 / 1,114,112 / 5,308,416 / 95,662,848 weights"""
    assert [orc.num_weights(n) for n in (2, 3, 4, 5, 6)] == [6144, 212992, 1114112, 5308416, 95662848]


def test_d4_order_and_update_keys(orc):
    """update() visits r, r^T, Rr, (Rr)^T, ... (r_learning.py:207-214); key multiplicities per update"""
    g = load_golden("d4.npz")
    row = np.arange(16, dtype=np.int32).reshape(4, 4)
    imgs = []
    for _ in range(4):
        imgs.append(row.ravel().copy())
        row = row.T
        imgs.append(row.ravel().copy())
        row = orc.rot90(row.T, 1)
    assert np.array_equal(np.array(imgs), g["images"].astype(np.int32))
    assert np.array_equal(row, np.arange(16).reshape(4, 4))
    assert len({tuple(i) for i in imgs}) == 8
    sample = g["sample"].astype(np.int32)
    for n in (2, 3, 4, 5):
        offs = g[f"offs_{n}"]
        for q, b in enumerate(sample):
            keys = orc.update_keys(n, b)
            assert len(keys) == 8 * orc.NUM_FEAT[n]
            k, c = np.unique(keys, return_counts=True)
            assert np.array_equal(k, g[f"keys_{n}"][offs[q]:offs[q + 1]])
            assert np.array_equal(c, g[f"counts_{n}"][offs[q]:offs[q + 1]])


def _weights(fx, n, seed, dtype):
    return fx.flat(fx.init_weights32(n, seed)).astype(dtype)


def _games(g):
    for i in range(len(g["odo"])):
        yield (i, g["start"][i].astype(np.int32), g["moves"][g["m_off"][i]:g["m_off"][i + 1]].astype(np.int32),
               g["tiles"][g["t_off"][i]:g["t_off"][i + 1]].astype(np.int32), g["final"][i].astype(np.int32),
               int(g["score"][i]), int(g["odo"][i]))


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_evaluate_update_event_stream(orc, fx, n):
    """every evaluate()/update() call the reference made in its first episodes, replayed in order:
    values equal to 1e-12 relative (float64), final weights of episode 1 equal."""
    g = load_golden(f"episodes_n{n}.npz")
    w = _weights(fx, n, int(g["seed"]), np.float64)
    base = w.copy()
    rows = orc.unpack_np(g["ev_board"])
    first_len = int(g["m_off"][1])          # moves of episode 0 incl. sentinel == its update count
    n_upd = 0
    checked_first = False
    for kind, row, val in zip(g["ev_kind"], rows, g["ev_val"]):
        if kind == 0:
            v = orc.evaluate(n, w, row)
            assert abs(v - val) <= 1e-12 * max(1.0, abs(val))
        else:
            orc.update(n, w, row, float(val))
            n_upd += 1
            if n_upd == first_len - 1 and not checked_first:     # episode 0: odometer updates (odo-1 + terminal)
                ref = fx.apply_sparse(base, g["w1_idx"], g["w1_val"])
                assert np.allclose(w, ref, rtol=0, atol=1e-13)
                checked_first = True
    assert checked_first


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_episode_replay_f64(orc, fx, n):
    """QAgent.episode teacher-forced on the recorded spawns: same moves (incl. -1 sentinel), same
    dw sequence, same final board/score, same final weights (float64, 1e-12)."""
    g = load_golden(f"episodes_n{n}.npz")
    w = _weights(fx, n, int(g["seed"]), np.float64)
    base = w.copy()
    upd_dw, u = g["upd_dw"], 0
    for i, start, moves, tiles, final, score, odo in _games(g):
        r = orc.episode_replay(n, w, float(g["alpha"]), start, tiles)
        assert r["odometer"] == odo and r["score"] == score
        assert np.array_equal(r["moves"], moves) and moves[-1] == -1
        assert np.array_equal(r["row"], final)
        dws = r["dws"][1:]                                  # step 0 has no update
        ref = upd_dw[u:u + len(dws)]
        assert np.allclose(dws, ref, rtol=1e-11, atol=1e-15)
        u += len(dws)
    assert u == len(upd_dw)
    ref_w = fx.apply_sparse(base, g["w_idx"], g["w_val"])
    assert np.allclose(w, ref_w, rtol=1e-11, atol=1e-14)
    assert len(g["w_idx"]) > 1000


def test_episode_replay_f32_tracks_f64(orc, fx):
    """the float32 restatement (the device's arithmetic) stays within tolerance of the float64 one:
    |dw32 - dw64| <= 2e-5 * (1 + |dw|), weights within 1e-4 after 60 episodes (teacher-forced)."""
    g = load_golden("episodes_n4.npz")
    n = 4
    w = _weights(fx, n, int(g["seed"]), np.float32)
    base = w.astype(np.float64)
    agree = total = 0
    for i, start, moves, tiles, final, score, odo in _games(g):
        r = orc.episode_replay(n, w, float(g["alpha"]), start, tiles)
        m = min(len(moves), len(r["moves"]))
        agree += int((r["moves"][:m] == moves[:m]).sum())
        total += len(moves)
    ref_w = fx.apply_sparse(base, g["w_idx"], g["w_val"])
    assert agree / total > 0.995          # near-ties may flip an argmax in float32
    assert np.abs(w - ref_w).max() < 5e-2


def test_greedy_replay(orc, fx):
    """Game.trial_run (depth 0) teacher-forced on recorded spawns, from float32-rounded trained
    weights: same moves, no sentinel, same final board and score, same evaluate() values."""
    g = load_golden("greedy_n4.npz")
    n = 4
    w32 = fx.apply_sparse(fx.flat(fx.init_weights32(n, int(g["seed"]))), g["w_idx"], g["w_val"])
    w = w32.astype(np.float64)
    for i, start, moves, tiles, final, score, odo in _games(g):
        r = orc.trial_replay(n, w, start, tiles)
        assert r["odometer"] == odo and r["score"] == score
        assert np.array_equal(r["moves"], moves) and (moves >= 0).all()
        assert np.array_equal(r["row"], final)
        # f32 arithmetic follows the same trajectory on these games (no near-tie flips)
        r32 = orc.trial_replay(n, w32.astype(np.float32), start, tiles)
        assert np.array_equal(r32["moves"], moves)
        assert np.allclose(r32["values"], r["values"], rtol=2e-6, atol=1e-6)


def test_lockstep_matches_reference_objects(orc, fx):
    """orc_td_lockstep (float64, sequential updates) == the lock-step loop built from reference
    Game/QAgent objects (gen_golden.gen_lockstep)."""
    g = load_golden("lockstep_n4.npz")
    n, B, steps = int(g["n"]), int(g["B"]), int(g["steps"])
    w = _weights(fx, n, int(g["seed"]), np.float64)
    base = w.copy()
    ls = orc.LockStep(n, w, float(g["alpha"]), int(g["pseed"]), B, segmented=0, threads=1)
    ls.run(steps // 2)
    ls.run(steps - steps // 2)           # chained calls == one call
    assert ls.n_updates == int(g["n_upd"]) and ls.n_moves == int(g["n_mv"])
    assert np.array_equal(ls.board, g["boards"])
    assert np.array_equal(ls.score, g["scores"])
    assert np.array_equal(ls.game_id.astype(np.int64), g["ids"])
    assert ls.fin[0] == len(g["fin"]) and ls.fin[1] == g["fin"][:, 1].sum() and ls.fin[2] == g["fin"][:, 2].sum()
    assert np.allclose(ls.old_label, g["labels"], rtol=1e-11)
    ref_w = fx.apply_sparse(base, g["w_idx"], g["w_val"])
    assert np.allclose(w, ref_w, rtol=1e-11, atol=1e-14)
    # segmented (delta-then-apply) float64 variant: same to rounding
    w2 = base.copy()
    ls2 = orc.LockStep(n, w2, float(g["alpha"]), int(g["pseed"]), B, segmented=1, threads=2)
    ls2.run(steps)
    assert np.array_equal(ls2.board, g["boards"])
    assert np.allclose(w2, ref_w, rtol=1e-9, atol=1e-12)


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10"""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, out in kat:
        assert orc.philox4x32_10(ctr, key).tolist() == out


def test_spawn_statistics(orc):
    """Philox spawn: P(tile '4') = 0.1, uniform over empties, two tiles at start (game_logic.py:61-66,112-116)"""
    fours = cells = 0
    pos_hist = np.zeros(16)
    for gid in range(4000):
        row = orc.spawn_initial(123, gid)
        assert (row > 0).sum() == 2 and set(np.unique(row)) <= {0, 1, 2}
        fours += int((row == 2).sum())
        cells += 2
        pos_hist += (row.ravel() > 0)
    assert abs(fours / cells - 0.1) < 0.012
    assert pos_hist.min() > 0.8 * pos_hist.mean() and pos_hist.max() < 1.2 * pos_hist.mean()
    row = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [1, 2, 3, 4], [5, 6, 0, 7]])
    new, res = orc.spawn_move(5, 9, 3, row)
    assert res >= 0 and (res & 0xff) == 14 and new[3, 2] == (res >> 8)
    full, res = orc.spawn_move(5, 9, 3, np.ones((4, 4)))
    assert res == -1


def test_play_philox_threads_agree(orc, fx):
    """greedy Philox play: results independent of thread count and of how the id range is split"""
    n = 4
    w = _weights(fx, n, 3, np.float32)
    a = orc.play_philox(n, w, seed=11, first_id=0, num=24, threads=1)
    b = orc.play_philox(n, w, seed=11, first_id=0, num=24, threads=4)
    c1 = orc.play_philox(n, w, seed=11, first_id=0, num=10, threads=2)
    c2 = orc.play_philox(n, w, seed=11, first_id=10, num=14, threads=2)
    assert np.array_equal(a["scores"], b["scores"]) and np.array_equal(a["boards"], b["boards"])
    assert np.array_equal(a["scores"], np.concatenate([c1["scores"], c2["scores"]]))
    assert a["total_moves"] == a["moves"].sum() > 24 * 20


def test_look_forward_replays_reference(orc, fx):
    """Game.look_forward (game_logic.py:214-243): the oracle, fed the reference's logged random.sample /
    random.randrange results in call order, reproduces the reference's values (float64, 141 calls, depth 1-4)."""
    g = load_golden("lookforward.npz")
    n = int(g["n"])
    w = fx.flat(fx.init_weights32(n, int(g["w_seed"]))).astype(np.float64)
    w[g["w_idx"]] = g["w_val"]
    for q in range(len(g["rows"])):
        score, depth, width, since_empty = (int(x) for x in g["meta"][q])
        log = (g["pos"][g["pos_off"][q]:g["pos_off"][q + 1]], g["tile"][g["tile_off"][q]:g["tile_off"][q + 1]])
        v = orc.look_forward(n, w, g["rows"][q:q + 1], [score], depth, width, since_empty, log=log)[0]
        assert abs(v - g["values"][q]) <= 1e-9 * max(1.0, abs(g["values"][q])), (q, v, g["values"][q])
