"""GPU tier: the reference-facing Python surface (game2048.game_logic.Game, game2048.r_learning.QAgent) on top
of the C-ABI.  These read like tests of the reference itself: same calls, same seeds, same expected results."""
import importlib
import os
import pickle
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
P = os.path.join(GOLDEN, "pickles")


@pytest.fixture(scope="module")
def mods():
    importlib.import_module("2048_b200")
    from game2048 import game_logic, r_learning
    return game_logic, r_learning


def agent_from(rl, fx, n, arrays):
    a = rl.QAgent(name="t", storage="local", console="local", n=n, with_weights=False)
    a.weights = arrays
    a.np_to_list()
    return a


def test_game_single_board_api_vs_reference_fixture(mods, orc):
    gl, rl = mods
    g = load_golden("boards.npz")
    game = gl.Game(row=np.zeros((4, 4), dtype=np.int32))
    c0 = gl.Game.counter
    for q in list(range(11)) + list(range(11, 3000, 97)):
        row = g["boards"][q].astype(np.int32)
        for d in range(4):
            if (g["after"][q, d] > 15).any():
                with pytest.raises(OverflowError):
                    game.pre_move(row, 1000, d)
                continue
            nr, ns, ch = game.pre_move(row, 1000, d)
            assert nr.dtype == np.int32 and nr.shape == (4, 4)
            assert np.array_equal(nr, g["after"][q, d]) and ns == 1000 + g["gain"][q, d] and ch == bool(g["change"][q, d])
        assert game.game_over(row) == bool(g["over"][q])
        assert gl.Game.empty_count(row) == g["n_empty"][q] and gl.Game.adjacent_pair_count(row) == g["n_pairs"][q]
        assert [4 * i + j for i, j in gl.Game.empty(row)] == [int(p) for p in g["empties"][q] if p >= 0]
    assert gl.Game.counter > c0
    t = gl.Game.table
    assert t[(1, 1, 1, 1)] == ((2, 2, 0, 0), 8, True) and t[(1, 2, 3, 4)] == ((1, 2, 3, 4), 0, False)
    assert t[(14, 14, 14, 14)] == ((15, 15, 0, 0), 65536, True) and len(t) == 65536
    with pytest.raises(KeyError):
        game.pre_move(np.full((4, 4), 16), 0, 0)              # the reference raises KeyError on a 2^16 tile


def test_seeded_game_reproduces_reference_rng_stream(mods, fx):
    """random.seed(s); Game(); trial_run(estimator): the single-game API consumes Python's `random` exactly like
    the reference (randrange(10) then choice(empties)), so the recorded reference game is reproduced move by move"""
    gl, rl = mods
    g = load_golden("greedy_n4.npz")
    n, seed = 4, int(g["seed"])
    w = fx.apply_sparse(fx.flat(fx.init_weights32(n, seed)), g["w_idx"], g["w_val"]).astype(np.float32)
    agent = agent_from(rl, fx, n, fx.unflat(n, w))
    random.seed(seed + 2000)
    np.random.seed(seed + 2000)
    for i in range(2):
        game = gl.Game()
        assert np.array_equal(game.starting_position, g["start"][i])
        game.trial_run(lambda row, score: agent.evaluate(row, score))     # generic estimator -> reference loop
        assert game.moves == g["moves"][g["m_off"][i]:g["m_off"][i + 1]].tolist()
        assert [(t, p[0], p[1]) for t, p in game.tiles] == [tuple(t) for t in g["tiles"][g["t_off"][i]:g["t_off"][i + 1]]]
        assert game.score == g["score"][i] and game.odometer == g["odo"][i] and np.array_equal(game.row, g["final"][i])
        chain = game.replay(verbose=False)
        assert np.array_equal(chain[game.odometer][0], game.row) and chain[game.odometer + 1] == (None, None, -1)


def test_agent_evaluate_update_match_reference_pickle(mods):
    """a reference-written agent file: evaluate() on the float32 file contents (<= 1e-6 rel), update() like the
    reference (8 images x F tables, r_learning.py:207-214)"""
    gl, rl = mods
    agent = rl.QAgent.load_agent_local(os.path.join(P, "ref_agent_n2.pkl"))
    probe = np.load(os.path.join(P, "probe.npz"))
    for row, v in zip(probe["rows"], probe["values32"]):
        assert abs(agent.evaluate(row) - v) <= 1e-6 * max(1.0, abs(v))
    assert np.allclose(agent.evaluate_batch(probe["rows"]), probe["values32"], rtol=1e-6)
    before = np.concatenate([a.reshape(-1) for a in agent.list_to_np()])
    row = probe["rows"][0]
    agent.update(row, 0.5)
    after = np.concatenate([a.reshape(-1) for a in agent.list_to_np()])
    assert abs((after - before).sum() - 0.5 * 8 * 24) < 1e-3 and 0 < np.count_nonzero(after != before) <= 8 * 24
    assert abs(agent.evaluate(row) - (probe["values32"][0] + (after - before)[rl.f_2(row) + 256 * np.arange(24)].sum())) < 1e-4
    # the reference's `agent.weights[i][f]` (list of lists while the agent is live, r_learning.py:136-164) keeps working as
    # a read-only per-table view of the device buffer
    view = agent.weights
    assert len(view) == 24 and view[3].shape == (256,) and view[-1].dtype == np.float32
    f = rl.f_2(row)
    assert abs(sum(float(view[i][f[i]]) for i in range(24)) - agent.evaluate(row)) < 1e-4
    assert np.array_equal(np.concatenate(list(view)), after) and np.array_equal(view[2:4][1], view[3])
    with pytest.raises(ValueError):
        view[0][0] = 1.0


def test_episode_returns_a_replayable_game(mods, fx):
    gl, rl = mods
    random.seed(5)
    agent = agent_from(rl, fx, 4, fx.init_weights32(4, 2))
    w0 = agent._device_weights().clone()
    game = agent.episode()
    assert agent.step == 1 and game.moves[-1] == -1 and len(game.moves) == game.odometer + 1
    assert len(game.tiles) == game.odometer and game.game_over(game.row)
    chain = game.replay(verbose=False)
    assert np.array_equal(chain[game.odometer][0], game.row) and chain[game.odometer][1] == game.score
    assert (agent._device_weights() != w0).sum().item() > 100


def test_trial_and_train_run_drivers(mods, fx, tmp_path, monkeypatch, capsys):
    gl, rl = mods
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("B2048_STORAGE", str(tmp_path / "store"))
    np.random.seed(0)
    random.seed(0)
    agent = rl.QAgent(name="drv", storage="local", console="local", n=4, batch=256)
    info = agent.train_run(num_eps=1200, saving=True)
    out = capsys.readouterr().out
    assert agent.step == 1201 and len(agent.train_history) == 12           # num_eps + 1 episodes, one entry per 100
    assert "average over last 1000 episodes" in out and "agent saved in drv.pkl" in out and "new best game" in out
    assert info["updates"] > 100_000 and os.path.exists("drv.pkl") and os.path.exists("best_of_drv.pkl")
    assert agent.train_history[-1] > agent.train_history[0]                # it learns
    results = rl.QAgent.trial(estimator=agent.evaluate, num=40, storage="local", game_file="best.pkl", seed=3)
    out = capsys.readouterr().out
    assert len(results) == 40 and results[0].score >= results[-1].score
    assert "average score of 40 runs" in out and "time per shuffle" in out
    best = gl.Game.load_game("best.pkl")
    assert best.score == results[0].score and len(best.moves) == best.odometer == len(best.tiles)
    chain = best.replay(verbose=False)
    assert np.array_equal(chain[best.odometer][0], best.row)
    # file round trips: local whole-object pickle and the two-object layout
    again = rl.QAgent.load_agent_local("drv.pkl")
    assert again.step == agent.step and torch.equal(again._device_weights(), agent._device_weights())
    agent.s3 = True
    agent.save_agent()
    third = rl.QAgent.load_agent("a/drv.pkl")
    assert third.alpha == agent.alpha and torch.equal(third._device_weights(), agent._device_weights())
    row = results[0].row
    assert third.evaluate(row) == agent.evaluate(row)


def test_look_ahead_drivers(mods, fx, capsys):
    """trial / trial_run with depth > 0 (game_logic.py:150-183, 214-243): the batched device expectimax plays
    better than depth 0 from the same weights, its best game replays, and a single Game.trial_run(depth=2) with the
    agent's estimator runs its look_forward trees on the device"""
    gl, rl = mods
    np.random.seed(1)
    random.seed(1)
    agent = rl.QAgent(name="la", storage="local", console="local", n=4, batch=256)
    agent.train_run(num_eps=600, saving=False)
    capsys.readouterr()
    r0 = rl.QAgent.trial(estimator=agent.evaluate, num=48, depth=0, storage="local", seed=4)
    r2 = rl.QAgent.trial(estimator=agent.evaluate, num=48, depth=2, width=3, since_empty=8, storage="local", seed=4)
    out = capsys.readouterr().out
    assert len(r2) == 48 and "average score of 48 runs" in out
    assert np.mean([g.score for g in r2]) > np.mean([g.score for g in r0])
    best = r2[0]
    assert len(best.moves) == best.odometer == len(best.tiles)
    chain = best.replay(verbose=False)
    assert np.array_equal(chain[best.odometer][0], best.row) and chain[best.odometer][1] == best.score
    game = gl.Game()
    game.trial_run(agent.evaluate, depth=2, width=2, since_empty=8, step_limit=60)
    assert game.odometer == 60 or game.game_over(game.row)
    assert len(game.moves) == game.odometer


def test_thread_mode_hooks(mods, fx, capsys):
    """the Dash call sites run the path from daemon threads with cooperative stop flags
    (application.py:427-468, 566-621; game_logic.py:186-200; r_learning.py:285-290, 363-368)."""
    import time as _time
    gl, rl = mods
    from game2048 import start
    random.seed(2)
    agent = agent_from(rl, fx, 4, fx.init_weights32(4, 5))
    # 'Agent Play': thread_trial records history until the pane's id changes
    start.GAME_PANE["u1"] = {"id": 7}
    game = gl.Game()
    game.thread_trial(agent.evaluate, depth=0, width=1, since_empty=6, stopper={"parent": "u1", "n": 7})
    t0 = _time.time()
    while len(game.history) < 20 and _time.time() - t0 < 60:
        _time.sleep(0.01)
    start.GAME_PANE["u1"]["id"] = 8                                       # another job took the pane: the thread must stop
    _time.sleep(0.3)
    n_hist = len(game.history)
    _time.sleep(0.3)
    assert n_hist >= 20 and len(game.history) == n_hist
    row0, score0, dir0 = game.history[0]
    assert row0.shape == (4, 4) and score0 == 0 and dir0 in (0, 1, 2, 3)
    # 'Train' from a thread with a stopper: stops at the next chunk boundary once the pane changes hands
    from threading import Thread
    start.AGENT_PANE["u1"] = {"id": 3}
    start.RUNNING["u1"] = 1
    worker = Thread(target=agent.train_run, kwargs={"num_eps": 10 ** 7, "saving": False, "batch": 64,
                                                    "stopper": {"parent": "u1", "a": 3}}, daemon=True)
    worker.start()
    t0 = _time.time()
    while agent.step < 50 and _time.time() - t0 < 60:
        _time.sleep(0.01)
    start.AGENT_PANE["u1"]["id"] = 4
    worker.join(timeout=30)
    assert not worker.is_alive() and 50 <= agent.step < 10 ** 6
