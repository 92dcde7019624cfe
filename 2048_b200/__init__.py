"""2048_b200 -- B200-native batched 2048 simulation + n-tuple TD agent (drop-in for abachurin/2048's
game2048.game_logic.Game / game2048.r_learning.QAgent hot path).

The directory name is not a Python identifier; import it with
    importlib.import_module("2048_b200")
which puts the drop-in `game2048` package (same module names as the reference, so that pickled
agents/games resolve `game2048.r_learning.QAgent`) on sys.path and imports it.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import game2048  # noqa: E402,F401  (the drop-in package living in this directory)
from game2048 import cabi, engine  # noqa: E402,F401

__all__ = ["game2048", "cabi", "engine"]
