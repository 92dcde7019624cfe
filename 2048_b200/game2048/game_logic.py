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

    actions = dict(enumerate(('left', 'up', 'right', 'down')))
    table = _MoveTable()
    counter = 0
    save_file = 'saved_game.pkl'
    _cache = (None, None)          # (packed board, its move4 result): pre_move is called 4x per position

    def __init__(self, score=0, row=None, file=None):
        self.score, self.odometer = score, 0
        self.moves, self.tiles, self.history = [], [], {}
        self.file = file if file else Game.save_file
        if row is not None:
            self.row, self.starting_position = np.array(row, dtype=np.int32), row
            return
        self.row = np.zeros((4, 4), dtype=np.int32)
        for _ in range(2):
            self.new_tile()
        del self.tiles[:]                                    # the two initial spawns are not recorded (:65)
        self.starting_position = self.row.copy()

    def copy(self):
        return Game(score=self.score, row=self.row)

    def save_game(self, file=None):
        with open(file or self.file, 'wb') as out:
            pickle.dump(self, out, pickle.HIGHEST_PROTOCOL)

    @staticmethod
    def load_game(file=save_file):
        with open(file, 'rb') as src:
            return pickle.load(src)

    def __eq__(self, other):
        return bool((np.asarray(self.row) == np.asarray(other.row)).all())

    def __str__(self):
        def cell(v):
            shown = (1 << int(v)) if v else 0
            return f'{shown}' + '\t' * (3 if shown >= 1000 else 4)     # the reference's tab layout (:89-93)

        body = [''.join(cell(v) for v in line) for line in self.row]
        body.append(f' score = {self.score} moves = {self.odometer} reached {1 << np.max(self.row)}')
        return '\n'.join(body)

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
        cached = Game._cache                                 # ONE read: worker threads replace the tuple concurrently
        if cached[0] == packed:
            return cached[1]
        ctx = engine.Context.get()
        after, gain, flags, _ = ctx.move4(ctx.to_device(np.array([packed], dtype=np.uint64)), want_over=False)
        res = (after.cpu().numpy().view(np.uint64)[0], gain.cpu().numpy().view(np.uint32)[0], int(flags.cpu()[0]))
        Game._cache = (packed, res)
        return res

    # ------------------------------------------------------------------ board predicates (:96-110)
    @staticmethod
    def empty(row):
        mask = Game._stats(row)[1]
        return [divmod(p, 4) for p in range(16) if mask & (1 << p)]

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
        """consumes `random` in the reference's order: randrange(10) for the tile, then choice over the empties"""
        cells = self.empty(row)
        four = random.randrange(10) == 0
        return (2 if four else 1), random.choice(cells)

    def new_tile(self):
        spawn = self.create_new_tile(self.row)
        self.row[spawn[1]] = spawn[0]
        self.tiles.append(spawn)

    # ------------------------------------------------------------------ moves (:123-148)
    _OVERFLOW = 'a 2^16 tile cannot be represented (the reference raises KeyError on its next move)'

    @staticmethod
    def _left(row, score):
        after, gain, flags = Game._move4(row)
        if flags & 16:
            raise OverflowError(Game._OVERFLOW)
        return unpack_board(after[0]), score + int(gain[0]), bool(flags & 1)

    def pre_move(self, row, score, direction):
        Game.counter += 1
        after, gain, flags = Game._move4(row)
        if flags & (16 << direction):
            raise OverflowError(Game._OVERFLOW)
        # the reference adds the line scores only for changed lines, which is all lines with a merge
        return unpack_board(after[direction]), score + int(gain[direction]), bool(flags & (1 << direction))

    def make_move(self, direction):
        moved = self.pre_move(self.row, self.score, direction)
        self.row, self.score = moved[0], moved[1]
        self._count_move(direction)
        return moved[2]

    def _count_move(self, direction):
        self.moves.append(direction)
        self.odometer += 1

    # ------------------------------------------------------------------ greedy play (:150-211)
    def _device_look_forward(self, agent, valid, depth, width, since_empty):
        """all look_forward trees of one move in one kernel (b2048_look_forward, node-keyed Philox draws):
        {direction: value} for the changed directions in `valid`"""
        if not valid:
            return {}
        ctx = engine.Context.get()
        if not hasattr(self, '_lf_seed'):
            self._lf_seed = random.getrandbits(63)
        k = len(valid)
        boards = ctx.to_device(np.array([pack_row(c[1]) for c in valid], dtype=np.uint64))
        vals = ctx.look_forward(agent.n, agent._device_weights(), boards,
                                ctx.to_device(np.zeros(k, dtype=np.uint64)),
                                ctx.to_device(np.full(k, self.odometer, dtype=np.uint32)),
                                ctx.to_device(np.array([c[0] for c in valid], dtype=np.uint8)), depth, width,
                                since_empty, seed=self._lf_seed).cpu().numpy()
        return {c[0]: float(v) for c, v in zip(valid, vals)}

    def _find_best_move(self, estimator, depth, width, since_empty):
        """(direction, afterstate, score) of the best changed direction; strict '>' so the lowest direction wins
        ties, direction 0 with no afterstate when nothing moves (:150-161)"""
        cand = [(d,) + tuple(self.pre_move(self.row, self.score, d)) for d in range(4)]
        valid = [c for c in cand if c[3]]
        agent = self._agent_of(estimator) if depth > 0 else None
        on_device = agent is not None and 1 <= width <= 4 and depth <= 4
        # other estimators / parameters keep the reference's host recursion on Python's `random`
        values = self._device_look_forward(agent, valid, depth, width, since_empty) if on_device else {}
        pick, top = (0, None, None), -np.inf
        for d, after, gained, _ in valid:
            v = values[d] if d in values else self.look_forward(estimator, after, gained, depth=depth, width=width,
                                                                since_empty=since_empty)
            if v > top:
                pick, top = (d, after, gained), v
        return pick

    def _move_on(self, best_dir, best_row, best_score):
        self._count_move(best_dir)
        self.row, self.score = best_row, best_score
        self.new_tile()

    def _agent_of(self, estimator):
        """the QAgent behind `estimator` if it is that agent's bound evaluate(), else None"""
        owner = getattr(estimator, '__self__', None)
        if owner is not None and getattr(estimator, '__func__', None) is getattr(type(owner), 'evaluate', None) \
                and hasattr(owner, '_device_weights'):
            return owner
        return None

    def _decisions(self, estimator, limit_tile, depth, width, since_empty, step_limit=None):
        """the reference's three play loops (:170-211) share this generator: yields the chosen
        (direction, afterstate, score) while the game can go on; the caller applies it with _move_on.
        Returns True if the game ended because no move was left."""
        while step_limit is None or self.odometer < step_limit:
            if self.game_over(self.row):
                return True
            if limit_tile and self.row.max() >= limit_tile:
                return False
            yield self._find_best_move(estimator, depth, width, since_empty)
        return False

    def trial_run(self, estimator, limit_tile=0, step_limit=100000, depth=0, width=1, since_empty=0, verbose=False):
        """:170-183.  With an agent's evaluate() and depth 0 the whole game runs in one kernel
        (b2048_greedy_play, Philox spawns keyed by a draw from `random`); otherwise the reference's loop."""
        agent = self._agent_of(estimator)
        if agent is not None and depth == 0 and not verbose:
            return self._trial_run_device(agent, limit_tile, step_limit)
        say = print if verbose else (lambda *a: None)
        say('Starting position:')
        say(self)
        for choice in self._decisions(estimator, limit_tile, depth, width, since_empty, step_limit):
            self._move_on(*choice)
            say(f'On {self.odometer} we moved {Game.actions[choice[0]]}')
            say(self)

    def _trial_run_device(self, agent, limit_tile, step_limit, trace_len=1 << 15):
        ctx = engine.Context.get()
        seed, start = random.getrandbits(63), self.odometer
        board, score = np.array([pack_row(self.row)], dtype=np.uint64), [self.score]
        length = trace_len + start
        while True:
            games = engine.GameBatch(1, seed=seed, ctx=ctx)
            games.set_positions(board, scores=score)
            games.moves.fill_(start)
            tdir, _, tsp = engine.greedy_play(ctx, agent.n, agent._device_weights(), games, limit_tile=limit_tile,
                                              step_limit=step_limit, trace_len=length)
            host = games.to_host()
            if int(host['moves'][0]) <= length:
                break
            length = int(host['moves'][0]) + 1               # the record was too short: the same game again, traced fully
        self.adopt_device_result(host, 0, tdir, tsp, start)

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
        owner, ticket = stopper['parent'], stopper['n']
        plan = self._decisions(estimator, 0, depth, width, since_empty)
        while GAME_PANE[owner]['id'] == ticket:  # noqa: F405
            choice = next(plan, None)
            snapshot = (self.row.copy(), self.score, -1 if choice is None else choice[0])
            self.history[self.odometer] = snapshot
            if choice is None:
                self.moves.append(-1)
                break
            self._move_on(*choice)

    def thread_trial(self, *args, **kwargs):
        Thread(target=self.trial_run_for_thread, args=args, kwargs=kwargs, daemon=True).start()

    def generate_run(self, estimator, limit_tile=0, depth=0, width=1, since_empty=16):
        """:203-211 (show.py)"""
        for choice in self._decisions(estimator, limit_tile, depth, width, since_empty):
            yield self, choice[0]
            self._move_on(*choice)

    # ------------------------------------------------------------------ look-ahead (:214-243)
    def look_forward(self, estimator, row, score, depth, width, since_empty):
        """the reference's sampled expectimax on the host, on top of the GPU board primitives; used for
        estimators that are not a QAgent (agents go through _device_look_forward).  Draw order on `random`
        as in the reference: one sample() of the positions, then one randrange(10) per position."""
        room = self.empty_count(row) if depth else 0
        if depth == 0 or room >= since_empty:
            return estimator(row, score)
        spots = random.sample(self.empty(row), min(width, room))
        total = 0
        for spot in spots:
            child = row.copy()
            child[spot] = 1 if random.randrange(10) else 2
            if self.game_over(child):
                continue                                     # max(-100, 0) adds nothing (:231, :242)
            replies = (self.pre_move(child, score, d) for d in range(4))
            below = [self.look_forward(estimator, r, s, depth=depth - 1, width=width, since_empty=since_empty)
                     for r, s, ok in replies if ok]
            total += max(max(below), 0)
        return total / len(spots)

    # ------------------------------------------------------------------ replay (:245-269)
    def replay(self, verbose=True):
        """rebuild the chain of boards from starting_position, moves and tiles.  Unlike the reference this does
        not need a trailing -1 in moves (trial_run does not append one, :267 would raise IndexError there)."""
        say = print if verbose else (lambda *a: None)
        ghost = Game(row=self.starting_position)
        say('Starting position:')
        say(ghost)
        chain = {}
        n = self.odometer
        for i, (move, (tile, where)) in enumerate(zip(self.moves[:n], self.tiles[:n])):
            chain[i] = (ghost.row.copy(), ghost.score, move)
            say(i, tile, where)
            ghost.make_move(move)
            ghost.row[where] = tile
            say(f'On {ghost.odometer} we move = {Game.actions[move]}, new tile = {tile} at position = {where}')
            say(ghost)
        say('no more moves possible, final position')
        chain[n] = (self.row.copy(), self.score, self.moves[n] if len(self.moves) > n else -1)
        chain[n + 1] = (None, None, -1)
        return chain
