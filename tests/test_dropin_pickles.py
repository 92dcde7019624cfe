"""CPU tier: files written by the REAL reference (tests/golden/pickles, produced by gen_pickles.py) unpickle
through the drop-in package -- same module / class / function names, same attributes, float32 weight arrays in
the reference's file layout (r_learning.py:151-200, game_logic.py:77-86).  No kernel is launched here."""
import importlib
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN

P = os.path.join(GOLDEN, "pickles")
ATTRS = ["name", "file", "game_file", "s3", "log_file", "print", "n", "alpha", "decay", "decay_step",
         "low_alpha_limit", "num_feat", "size_feat", "features", "step", "top_game", "top_score", "train_history",
         "next_decay", "top_tile", "weights", "weight_signature"]


@pytest.fixture(scope="module")
def mods():
    importlib.import_module("2048_b200")
    from game2048 import game_logic, r_learning
    return game_logic, r_learning


def test_local_mode_agent_pickle_loads(mods):
    gl, rl = mods
    with open(os.path.join(P, "ref_agent_n2.pkl"), "rb") as f:
        agent = pickle.load(f)
    probe = np.load(os.path.join(P, "probe.npz"))
    assert type(agent) is rl.QAgent and type(agent).__module__ == "game2048.r_learning"
    for a in ATTRS:
        assert hasattr(agent, a), a
    assert agent.n == 2 and agent.num_feat == 24 and agent.step == int(probe["step"])
    assert agent.features is rl.f_2 and agent.print is print
    assert agent.weight_signature == (24,) and agent.train_history == [123, 456]
    assert isinstance(agent.weights, list) and agent.weights[0].dtype == np.float32
    assert agent.weights[0].shape == (24, 256)
    assert np.array_equal(np.concatenate([w.reshape(-1) for w in agent.weights]), probe["w32"])
    assert type(agent.top_game) is gl.Game and agent.top_game.score == int(probe["last_score"])
    assert agent._w is None                                   # nothing uploaded until first use


def test_s3_mode_pair_and_game_pickle(mods, tmp_path, monkeypatch):
    gl, rl = mods
    monkeypatch.setenv("B2048_STORAGE", P)
    from game2048 import start
    params = start.load_s3("a/ref_agent_n2.pkl")
    weights = start.load_s3("weights/ref_agent_n2.pkl")
    assert params.weights is None and params.n == 2
    assert [w.shape for w in weights] == [(24, 256)] and weights[0].dtype == np.float32
    assert start.is_data_there("a/ref_agent_n2.pkl") and "weights/ref_agent_n2.pkl" in start.list_names_s3()
    game = gl.Game.load_game(os.path.join(P, "ref_game.pkl"))
    probe = np.load(os.path.join(P, "probe.npz"))
    assert game.score == int(probe["last_score"]) and game.odometer == int(probe["last_odo"])
    assert game.moves == probe["last_moves"].tolist() and game.moves[-1] == -1
    assert np.array_equal(game.row, probe["last_row"]) and len(game.tiles) == game.odometer
    assert game.tiles[0][0] in (1, 2) and len(game.tiles[0][1]) == 2


def test_weightless_roundtrip_keeps_reference_layout(mods):
    gl, rl = mods
    a = rl.QAgent(name="x", storage="local", console="local", n=5, with_weights=False)
    b = pickle.loads(pickle.dumps(a, -1))
    assert b.n == 5 and b.weights is None and b.num_feat == 21 and b.features is rl.f_5
    assert rl.Q_agent is rl.QAgent
    assert rl.QAgent.parameter_shape == {2: (24, 256), 3: (52, 4096), 4: (17, 65536), 5: (21, 1048576), 6: (33, 0)}
