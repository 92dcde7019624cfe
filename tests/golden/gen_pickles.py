"""Generate reference-format pickles with the REAL reference imported under its own package name
(`game2048`), so that the pickles name game2048.r_learning.QAgent / f_2 and game2048.game_logic.Game exactly as
files written by the reference do.  Run in a fresh interpreter:  python tests/golden/gen_pickles.py
Outputs (tests/golden/pickles/): local-mode agent (whole QAgent, weights as float32 arrays, r_learning.py:177-180),
S3-mode pair a/<name>.pkl + weights/<name>.pkl (:168-174), a Game (game_logic.py:77-80), and probe values."""
import contextlib
import io
import os
import pickle
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "pickles")
os.makedirs(os.path.join(OUT, "a"), exist_ok=True)
os.makedirs(os.path.join(OUT, "weights"), exist_ok=True)
sys.modules.setdefault("boto3", types.ModuleType("boto3"))
os.environ["S3_URL"] = "none"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
with contextlib.redirect_stdout(io.StringIO()):
    from game2048.r_learning import QAgent, Game   # noqa: E402

random.seed(21)
np.random.seed(21)
os.chdir(OUT)
agent = QAgent(name="ref_agent_n2", storage="local", console="local", n=2)
with contextlib.redirect_stdout(io.StringIO()):
    games = [agent.episode() for _ in range(5)]
agent.top_game, agent.top_score = games[-1], games[-1].score
agent.train_history = [123, 456]
agent.save_agent()                                         # -> ref_agent_n2.pkl (local mode)
nps = agent.list_to_np()
params = QAgent(name=agent.name, with_weights=False, storage="local", console="local", n=2)
for key in agent.__dict__:
    if key != "weights":
        setattr(params, key, getattr(agent, key))
with open(os.path.join("a", "ref_agent_n2.pkl"), "wb") as f:   # what save_s3(agent_params, 'a/...') uploads
    pickle.dump(params, f, -1)
with open(os.path.join("weights", "ref_agent_n2.pkl"), "wb") as f:
    pickle.dump(nps, f, -1)
games[-1].save_game("ref_game.pkl")
probe = np.array([g.row for g in games], dtype=np.int32)
values = np.array([agent.evaluate(r) for r in probe])
w32 = np.concatenate([a.reshape(-1) for a in nps])
values32 = []
for r in probe:                                            # the same sum over the float32 file contents
    values32.append(sum(float(nps[0][i][f]) for i, f in enumerate(agent.features(r))))
np.savez(os.path.join(OUT, "probe.npz"), rows=probe, values=values, values32=np.array(values32), w32=w32,
         step=agent.step, alpha=agent.alpha, last_score=games[-1].score, last_odo=games[-1].odometer,
         last_moves=np.array(games[-1].moves), last_row=games[-1].row)
print("wrote", sorted(os.listdir(OUT)))
