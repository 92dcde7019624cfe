"""Drop-in for the reference's game2048/game_logic.py: class Game with the same attributes, methods and
semantics (file:line citations are into /root/reference/game2048/game_logic.py), but every board
operation -- moves, merge scores, game-over test, empty cells -- runs in the CUDA kernels of libb2048.so.

Two ways to use it:
  * the reference's single-game API (Game(), pre_move, make_move, new_tile, trial_run, replay ...), kept call
    for call so existing callers work; each call is one tiny kernel launch through the C-ABI (correct, not fast);
  * GameBatch / engine.greedy_play for many games at once (what QAgent.trial and the benchmarks use).
Tile spawning in the single-game API consumes Python's `random` exactly like the reference (randrange(10)
then choice(empties), :112-116), so a seeded script reproduces the reference's games; the batched kernels use
the counter-based Philox stream instead (include/b2048.h).
"""
from .start import *  # noqa: F401,F403  (the reference's star-import chain, game_logic.py:1)
from . import cabi, engine

import numpy as np
import random
import pickle
from threading import Thread

_SHIFTS = (4 * (15 - np.arange(16))).astype(np.uint64)


def pack_row(row):
    """4x4 exponents -> packed uint64 (cell (r,c) at nibble 15-4r-c); host-side, one board"""
    r = np.asarray(row, dtype=np.int64).reshape(16)
    if ((r < 0) | (r > 15)).any():
        raise KeyError(tuple(int(v) for v in r))       # the reference fails the same way on a 2^16 tile (:129)
    return int(np.bitwise_or.reduce(r.astype(np.uint64) << _SHIFTS))


def unpack_board(b):
    return ((np.uint64(b) >> _SHIFTS) & np.uint64(15)).astype(np.int32).reshape(4, 4)


def random_eval(row, score):
    """:5-6"""
    return np.random.random()


def score_eval(row, score):
    """:9-10"""
    return score


class _MoveTable:
    """Game.table (:18-39, :51) materialised lazily from the device LUT.  Lines that would hold a 2^16 tile are
    stored with that cell saturated at 15 (the packed format cannot hold 16; see include/b2048.h)."""

    def __init__(self):
        self._d = None

    def _build(self):
        lut = engine.Context.get().lut.cpu().numpy().view(np.uint32)
        a, b = (lut >> 16) & 15, (lut >> 20) & 15
        score = np.where(a > 0, 2 << a.astype(np.int64), 0) + np.where(b > 0, 2 << b.astype(np.int64), 0)
        d = {}
        for key in range(65536):
            e = int(lut[key])
            line = ((key >> 12) & 15, (key >> 8) & 15, (key >> 4) & 15, key & 15)
            new = ((e >> 12) & 15, (e >> 8) & 15, (e >> 4) & 15, e & 15)
            d[line] = (new, int(score[key]), bool((e >> 24) & 1))
        self._d = d

    def __get__(self, obj, owner):
        if self._d is None:
            self._build()
        return self._d


def create_table():
    """:18-39 -- returns the same dict the reference builds, read back from the GPU LUT"""
    t = _MoveTable()
    t._build()
    print('table of moves created')
    return t._d


class Game:
    """:48-269"""

    actions = {0: 'left', 1: 'up', 2: 'right', 3: 'down'}
    table = _MoveTable()
    counter = 0
    save_file = 'saved_game.pkl'
    _cache = (None, None)          # (packed board, its move4 result): pre_move is called 4x per position

    def __init__(self, score=0, row=None, file=None):
        self.score = score
        self.odometer = 0
        self.moves = []
        self.tiles = []
        self.history = {}
        if row is None:
            self.row = np.zeros((4, 4), dtype=np.int32)
            self.new_tile()
            self.new_tile()
            self.tiles = []                                  # the two initial spawns are not recorded (:65)
            self.starting_position = self.row.copy()
        else:
            self.row = np.array(row, dtype=np.int32)
            self.starting_position = row
        self.file = file or Game.save_file

    def copy(self):
        return Game(self.score, self.row)

    def save_game(self, file=None):
        with open(file or self.file, 'wb') as f:
            pickle.dump(self, f, -1)

    @staticmethod
    def load_game(file=save_file):
        with open(file, 'rb') as f:
            return pickle.load(f)

    def __eq__(self, other):
        return np.array_equal(self.row, other.row)

    def __str__(self):
        lines = []
        for r in self.row:
            lines.append(''.join(str(1 << v if v else 0) + '\t' * (4 if (1 << v) < 1000 else 3) for v in r))
        return '\n'.join(lines) + f'\n score = {str(self.score)} moves = {str(self.odometer)} ' \
                                  f'reached {1 << np.max(self.row)}'

    # ------------------------------------------------------------------ device helpers
    @staticmethod
    def _stats(row):
        ctx = engine.Context.get()
        b = ctx.to_device(np.array([pack_row(row)], dtype=np.uint64))
        stats, mask = ctx.board_stats(b)
        return stats.cpu().numpy()[0], int(mask.cpu().numpy().view(np.uint16)[0])

    @staticmethod
    def _move4(row):
        packed = pack_row(row)
        if Game._cache[0] == packed:
            return Game._cache[1]
        ctx = engine.Context.get()
        after, gain, flags, _ = ctx.move4(ctx.to_device(np.array([packed], dtype=np.uint64)), want_over=False)
        res = (after.cpu().numpy().view(np.uint64)[0], gain.cpu().numpy().view(np.uint32)[0], int(flags.cpu()[0]))
        Game._cache = (packed, res)
        return res

    # ------------------------------------------------------------------ board predicates (:96-110)
    @staticmethod
    def empty(row):
        _, mask = Game._stats(row)
        return [(p // 4, p % 4) for p in range(16) if (mask >> p) & 1]

    @staticmethod
    def empty_count(row):
        return int(Game._stats(row)[0][0])

    @staticmethod
    def adjacent_pair_count(row):
        return int(Game._stats(row)[0][1])

    def game_over(self, row):
        return bool(Game._stats(row)[0][2])

    # ------------------------------------------------------------------ spawns (:112-121)
    def create_new_tile(self, row):
        em = self.empty(row)
        tile = 1 if random.randrange(10) else 2
        position = random.choice(em)
        return tile, position

    def new_tile(self):
        tile, position = self.create_new_tile(self.row)
        self.row[position] = tile
        self.tiles.append((tile, position))

    # ------------------------------------------------------------------ moves (:123-148)
    @staticmethod
    def _left(row, score):
        after, gain, flags = Game._move4(row)
        change = bool(flags & 1)
        if flags & 16:
            raise OverflowError('a 2^16 tile cannot be represented (the reference raises KeyError on its next move)')
        return unpack_board(after[0]), score + int(gain[0]), change

    def pre_move(self, row, score, direction):
        Game.counter += 1
        after, gain, flags = Game._move4(row)
        if (flags >> (4 + direction)) & 1:
            raise OverflowError('a 2^16 tile cannot be represented (the reference raises KeyError on its next move)')
        change = bool((flags >> direction) & 1)
        # the reference adds the line scores only for changed lines, which is all lines with a merge
        return unpack_board(after[direction]), score + int(gain[direction]), change

    def make_move(self, direction):
        self.row, self.score, change = self.pre_move(self.row, self.score, direction)
        self.odometer += 1
        self.moves.append(direction)
        return change

    # ------------------------------------------------------------------ greedy play (:150-211)
    def _find_best_move(self, estimator, depth, width, since_empty):
        best_dir, best_value = 0, -np.inf
        best_row, best_score = None, None
        agent = self._agent_of(estimator) if depth > 0 else None
        cand = [(d,) + tuple(self.pre_move(self.row, self.score, d)) for d in range(4)]
        device_values = {}
        if agent is not None and 1 <= width <= 4 and depth <= 4:
            # all look_forward trees of the move in one kernel (b2048_look_forward, node-keyed Philox draws);
            # other estimators / parameters keep the reference's host recursion on Python's `random`
            ctx = engine.Context.get()
            valid = [c for c in cand if c[3]]
            if valid:
                if not hasattr(self, '_lf_seed'):
                    self._lf_seed = random.getrandbits(63)
                boards = ctx.to_device(np.array([pack_row(c[1]) for c in valid], dtype=np.uint64))
                k = len(valid)
                vals = ctx.look_forward(agent.n, agent._device_weights(), boards,
                                        ctx.to_device(np.zeros(k, dtype=np.uint64)),
                                        ctx.to_device(np.full(k, self.odometer, dtype=np.uint32)),
                                        ctx.to_device(np.array([c[0] for c in valid], dtype=np.uint8)), depth, width,
                                        since_empty, seed=self._lf_seed).cpu().numpy()
                device_values = {c[0]: float(v) for c, v in zip(valid, vals)}
        for direction, new_row, new_score, change in cand:
            if not change:
                continue
            value = device_values[direction] if direction in device_values else \
                self.look_forward(estimator, new_row, new_score, depth=depth, width=width, since_empty=since_empty)
            if value > best_value:                           # strict: the lowest direction wins ties
                best_dir, best_value = direction, value
                best_row, best_score = new_row, new_score
        return best_dir, best_row, best_score

    def _move_on(self, best_dir, best_row, best_score):
        self.moves.append(best_dir)
        self.odometer += 1
        self.row, self.score = best_row, best_score
        self.new_tile()

    def _agent_of(self, estimator):
        """the QAgent behind `estimator` if it is that agent's bound evaluate(), else None"""
        owner = getattr(estimator, '__self__', None)
        if owner is not None and getattr(estimator, '__func__', None) is getattr(type(owner), 'evaluate', None) \
                and hasattr(owner, '_device_weights'):
            return owner
        return None

    def trial_run(self, estimator, limit_tile=0, step_limit=100000, depth=0, width=1, since_empty=0, verbose=False):
        """:170-183.  With an agent's evaluate() and depth 0 the whole game runs in one kernel
        (b2048_greedy_play, Philox spawns keyed by a draw from `random`); otherwise the reference's loop."""
        agent = self._agent_of(estimator)
        if agent is not None and depth == 0 and not verbose:
            return self._trial_run_device(agent, limit_tile, step_limit)
        if verbose:
            print('Starting position:')
            print(self)
        while self.odometer < step_limit:
            if self.game_over(self.row):
                return
            if limit_tile and np.max(self.row) >= limit_tile:
                break
            best_dir, best_row, best_score = self._find_best_move(estimator, depth, width, since_empty)
            self._move_on(best_dir, best_row, best_score)
            if verbose:
                print(f'On {self.odometer} we moved {Game.actions[best_dir]}')
                print(self)

    def _trial_run_device(self, agent, limit_tile, step_limit, trace_len=1 << 15):
        ctx = engine.Context.get()
        games = engine.GameBatch(1, seed=random.getrandbits(63), ctx=ctx)
        games.set_positions(np.array([pack_row(self.row)], dtype=np.uint64), scores=[self.score])
        games.moves.fill_(self.odometer)
        start = self.odometer
        tdir, _, tsp = engine.greedy_play(ctx, agent.n, agent._device_weights(), games, limit_tile=limit_tile,
                                          step_limit=step_limit, trace_len=min(trace_len + start, 1 << 20))
        self.adopt_device_result(games.to_host(), 0, tdir, tsp, start)

    def adopt_device_result(self, host, slot, tdir, tsp, start=0):
        """fill row / score / odometer / moves / tiles from a finished device slot and its traces"""
        self.row = unpack_board(host['board'][slot])
        self.score = int(host['score'][slot])
        self.odometer = int(host['moves'][slot])
        if tdir is not None:
            k = min(self.odometer, tdir.shape[1])
            d = tdir[slot, start:k].cpu().numpy()
            sp = tsp[slot, start:k].cpu().numpy().view(np.uint16)
            self.moves += [int(v) for v in d]
            self.tiles += [(int(s >> 8), (int(s & 15) // 4, int(s & 15) % 4)) for s in sp]

    def trial_run_for_thread(self, estimator, depth=0, width=1, since_empty=0, stopper=None):
        """:186-197 (Dash 'Agent Play'): same loop, records history, polls the stop flag"""
        parent, this_thread = stopper['parent'], stopper['n']
        while True:
            if GAME_PANE[parent]['id'] != this_thread:  # noqa: F405
                return
            if self.game_over(self.row):
                self.history[self.odometer] = (self.row.copy(), self.score, -1)
                self.moves.append(-1)
                return
            best_dir, best_row, best_score = self._find_best_move(estimator, depth, width, since_empty)
            self.history[self.odometer] = (self.row.copy(), self.score, best_dir)
            self._move_on(best_dir, best_row, best_score)

    def thread_trial(self, *args, **kwargs):
        Thread(target=self.trial_run_for_thread, args=args, kwargs=kwargs, daemon=True).start()

    def generate_run(self, estimator, limit_tile=0, depth=0, width=1, since_empty=16):
        """:203-211 (show.py)"""
        while True:
            if self.game_over(self.row):
                return
            if limit_tile and np.max(self.row) >= limit_tile:
                break
            best_dir, best_row, best_score = self._find_best_move(estimator, depth, width, since_empty)
            yield self, best_dir
            self._move_on(best_dir, best_row, best_score)

    # ------------------------------------------------------------------ look-ahead (:214-243)
    def look_forward(self, estimator, row, score, depth, width, since_empty):
        """depth 0 is the hot path (estimator call); depth > 0 is the reference's sampled expectimax, kept on
        the host on top of the GPU board primitives ('next', not yet batched)"""
        if depth == 0:
            return estimator(row, score)
        empty = self.empty_count(row)
        if empty >= since_empty:
            return estimator(row, score)
        num_tiles = min(width, empty)
        tile_positions = random.sample(self.empty(row), num_tiles)
        average = 0
        for position in tile_positions:
            new_tile = 1 if random.randrange(10) else 2
            new_row = row.copy()
            new_row[position] = new_tile
            if self.game_over(new_row):
                best_value = -100
            else:
                best_value = -np.inf
                for direction in range(4):
                    test_row, test_score, change = self.pre_move(new_row, score, direction)
                    if change:
                        value = self.look_forward(estimator, test_row, test_score, depth=depth - 1, width=width,
                                                  since_empty=since_empty)
                        best_value = max(best_value, value)
            average += max(best_value, 0)
        return average / num_tiles

    # ------------------------------------------------------------------ replay (:245-269)
    def replay(self, verbose=True):
        """rebuild the chain of boards from starting_position, moves and tiles.  Unlike the reference this does
        not need a trailing -1 in moves (trial_run does not append one, :267 would raise IndexError there)."""
        chain = {}
        replay_game = Game(row=self.starting_position)
        if verbose:
            print('Starting position:')
            print(replay_game)
        for i in range(self.odometer):
            move = self.moves[i]
            chain[i] = (replay_game.row.copy(), replay_game.score, move)
            new_tile, position = self.tiles[i]
            if verbose:
                print(i, new_tile, position)
            replay_game.make_move(move)
            replay_game.row[position] = new_tile
            if verbose:
                print(f'On {replay_game.odometer} we move = {Game.actions[move]}, '
                      f'new tile = {new_tile} at position = {position}')
                print(replay_game)
        if verbose:
            print('no more moves possible, final position')
        last = self.moves[self.odometer] if len(self.moves) > self.odometer else -1
        chain[self.odometer] = (self.row.copy(), self.score, last)
        chain[self.odometer + 1] = (None, None, -1)
        return chain
