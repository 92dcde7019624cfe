"""Import the unmodified Python reference (abachurin/2048) in THIS container.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, so nothing that runs there may call this module
(tests that use it skip when the directory is absent; golden fixtures made with it are committed under
tests/golden/ together with gen_golden.py).

The reference needs two nudges to import without network/S3:
  * game2048/start.py:12 does `import boto3`                     -> an empty stub module
  * start.py:35-45 opens a hard-coded credential file unless S3_URL is set to something that is
    neither 'local' nor 'AWS'                                   -> S3_URL=none
It is loaded under the alias package name `ref_game2048` (its modules only use relative imports), so
that the product's own drop-in `game2048` package can live in the same interpreter.
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

REF_ROOT = os.environ.get("B2048_REFERENCE", "/root/reference")
ALIAS = "ref_game2048"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "game2048", "r_learning.py"))


def load(alias=ALIAS):
    """Returns (game_logic_module, r_learning_module) of the reference."""
    if alias + ".r_learning" in sys.modules:
        return sys.modules[alias + ".game_logic"], sys.modules[alias + ".r_learning"]
    if not available():
        raise ImportError(f"reference not found under {REF_ROOT}")
    sys.modules.setdefault("boto3", types.ModuleType("boto3"))
    os.environ["S3_URL"] = "none"
    pkg_dir = os.path.join(REF_ROOT, "game2048")
    spec = importlib.util.spec_from_file_location(alias, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[alias] = pkg
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True      # /root/reference is read-only
    try:
        with contextlib.redirect_stdout(io.StringIO()):   # 'Unknown environment', 'table of moves created'
            spec.loader.exec_module(pkg)
            gl = importlib.import_module(alias + ".game_logic")
            rl = importlib.import_module(alias + ".r_learning")
    finally:
        sys.dont_write_bytecode = dont
    return gl, rl


def seed_all(seed):
    """The reference never seeds its RNGs; the harness does."""
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)


def make_agent(rl, n, weights32=None, alpha=0.25):
    """QAgent with storage/console local (r_learning.py:97-99); optional float32 weight arrays in the
    reference's file layout (r_learning.py:151-164) installed through its own np_to_list()."""
    agent = rl.QAgent(name="oracle", storage="local", console="local", n=n, alpha=alpha,
                      with_weights=weights32 is None)
    if weights32 is not None:
        sig = {2: (24,), 3: (52,), 4: (17,), 5: (17, 4), 6: (17, 4, 12)}[n]
        agent.weights = list(weights32)
        agent.weight_signature = sig
        agent.np_to_list()
    return agent
