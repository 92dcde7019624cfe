"""Drop-in `game2048` package: same module names as abachurin/2048 (start, game_logic, r_learning) so
that `from game2048.r_learning import QAgent` and pickles that reference these modules keep working;
the hot path underneath runs in libb2048.so (CUDA, sm_100a) through cabi.py / engine.py."""
