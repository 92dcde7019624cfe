"""Drop-in for the reference's game2048/r_learning.py: f_2..f_6 and class QAgent (alias Q_agent) with the same
constructor, attributes, methods, report text and weight-file format (citations: file:line in
/root/reference/game2048/r_learning.py), with the n-tuple gather (evaluate), the D4-symmetric TD scatter
(update) and the whole episode / trial loops running in libb2048.so on the GPU.

Weights live in ONE float32 device buffer (table i at cabi.table_offsets(n)[i]); `weights` is materialised as
the reference's list of float32 arrays per signature group only for pickling (list_to_np, :151-158).

Batched semantics (new; DESIGN.md "Semantics of batched TD"): train_run plays `batch` games in lock-step; all
games evaluate with the weights of the start of the lock-step, then the updates are combined per weight with
the per-key-mean rule (sum of the contributions / number of distinct games contributing), which equals the
reference's update at batch = 1 and stays stable at the reference's alpha for thousands of games.
"""
import os

from .game_logic import *  # noqa: F401,F403  (the reference's star-import chain, r_learning.py:3)
from . import cabi, engine
from .game_logic import Game, pack_row, unpack_board
from .start import AGENT_PANE, RUNNING, Logger, dash_intervals, load_s3, save_s3

import numpy as np
import pickle
import random
import time
from collections import deque


def check_thread(parent, benchmark):
    """:6-13 heartbeat of the Dash session that owns a worker thread: once two check intervals have passed, the
    session must have raised its RUNNING flag again, else 0 tells the worker to stop; the flag is cleared and
    the clock restarted otherwise"""
    now = time.time()
    if now - benchmark <= 2 * dash_intervals['check_run']:
        return benchmark
    alive, RUNNING[parent] = RUNNING[parent], 0
    return now if alive else 0


class _Owner:
    """the `stopper` protocol of the Dash front end (:280-290, :362-369): {'parent': session, 'a': ticket}.
    go_on() is False once another job took the agent pane; lost() is True when the session stopped pinging."""

    def __init__(self, stopper):
        self.session = stopper['parent'] if stopper else None
        self.ticket = stopper['a'] if stopper else None
        self.clock = time.time()

    def go_on(self):
        return self.session is None or AGENT_PANE[self.session]['id'] == self.ticket

    def lost(self):
        if self.session is not None:
            self.clock = check_thread(self.session, self.clock)
        return not self.clock


def _features(n, x):
    ctx = engine.Context.get()
    b = ctx.to_device(np.array([pack_row(x)], dtype=np.uint64))
    return ctx.features(n, b).cpu().numpy()[0]


def f_2(x):
    """:17-20 all adjacent pairs (24 indices)"""
    return _features(2, x)


def f_3(x):
    """:24-31 lines of three and L-shaped triples (52 indices)"""
    return _features(3, x)


def f_4(x):
    """:40-44 columns, rows, 2x2 squares (17 indices)"""
    return _features(4, x)


def f_5(x):
    """:48-54 f_4 + the four centre crosses (21 indices)"""
    return _features(5, x)


def f_6(x):
    """:58-69 f_5 + twelve 3x2 / 2x3 rectangles in base 14 on min(x, 13) (33 indices)"""
    return _features(6, x)


class _WeightsView:
    """Read-only stand-in for the reference's list-of-lists `QAgent.weights` (:136-164) while the tables live in the
    device buffer: weights[i] is table i as a float32 array (a fresh device->host copy of that table, not
    writeable), so `agent.weights[i][f]`, len() and iteration work as before.  Writes go through update() /
    train_run(); assign file-format arrays to `agent.weights` + np_to_list() to replace the tables."""

    def __init__(self, agent):
        self._agent = agent
        self._offsets = cabi.table_offsets(agent.n)

    def __len__(self):
        return len(self._offsets) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        table = self._agent._w[self._offsets[i]:self._offsets[i + 1]].cpu().numpy()
        table.flags.writeable = False
        return table

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class QAgent:
    """:85-406"""

    feature_functions = dict(zip(range(2, 7), (f_2, f_3, f_4, f_5, f_6)))
    parameter_shape = {2: (24, 1 << 8), 3: (52, 1 << 12), 4: (17, 1 << 16), 5: (21, 1 << 20), 6: (33, 0)}

    def __init__(self, name='agent', config_file=None, storage='s3', console='web', log_file=None, n=4, alpha=0.25,
                 decay=0.75, decay_step=10000, low_alpha_limit=0.01, with_weights=True,
                 batch=1024, update_mode='atomic', seed=None):
        self.name, self.file = name, f'{name}.pkl'
        self.game_file = f'best_of_{self.file}'
        self.s3 = storage == 's3'
        self.log_file = log_file
        quiet_web = console != 'local' and log_file is not None
        self.print = Logger(log_file=log_file).add if quiet_web else print

        # hyper-parameters: the stored config (if any) overrides the arguments (:104-113)
        stored = (load_s3(config_file) if config_file else None) or {}
        given = dict(n=n, alpha=alpha, decay=decay, decay_step=decay_step, low_alpha_limit=low_alpha_limit)
        for key, value in given.items():
            setattr(self, key, stored.get(key, value))
        self.num_feat, self.size_feat = QAgent.parameter_shape[self.n]
        self.features = QAgent.feature_functions[self.n]

        # training state (:119-125)
        self.step, self.top_score, self.top_tile = 0, 0, 10
        self.top_game, self.train_history = None, []
        self.next_decay = self.decay_step

        # new knobs (not in the reference): games per lock-step, scatter mode, Philox seed, and the first Philox game
        # id nobody has played yet (with a fixed `seed` a second run must not replay the streams of the first)
        self.batch, self.update_mode, self.seed = batch, update_mode, seed
        self.next_game_id = 0

        self._w = None                      # float32 device buffer (never pickled)
        self._file_weights = None           # file-format arrays, only between load and first use
        self.weight_signature = None
        if with_weights:
            self.init_weights()

    def __str__(self):
        return '\n'.join((f'Agent {self.name}, n={self.n}',
                          f'trained for {self.step} episodes, top score = {self.top_score}'))

    # ------------------------------------------------------------------ weights (:136-164)
    @property
    def weights(self):
        """the reference's attribute: file-format arrays between load and first use (or None), afterwards a read-only
        per-table view of the device buffer (_WeightsView)"""
        return _WeightsView(self) if self._w is not None else self._file_weights

    @weights.setter
    def weights(self, arrays):
        self._file_weights = arrays

    def init_weights(self):
        """np.random.random(shape) / 100 per table group, like :136-149 (same stream under np.random.seed)"""
        self.weight_signature = engine.SIGNATURE[self.n]
        arrays = [(np.random.random((d, s)) / 100).astype(np.float32)
                  for d, s in zip(engine.SIGNATURE[self.n], engine.GROUP_SIZE[self.n])]
        self._upload(arrays)

    def _upload(self, arrays):
        ctx = engine.Context.get()
        flat = engine.flat_from_arrays(arrays)
        if flat.size != cabi.num_weights(self.n):
            raise ValueError(f'weights have {flat.size} entries, n={self.n} needs {cabi.num_weights(self.n)}')
        self._w = ctx.to_device(flat)
        self._file_weights = None

    def _device_weights(self):
        if self._w is None:
            if self._file_weights is None:
                raise ValueError('agent has no weights (with_weights=False and nothing loaded)')
            self.np_to_list()
        return self._w

    def list_to_np(self):
        """:151-158 list of float32 arrays, one per signature group (the weight-file payload)"""
        return engine.arrays_from_flat(self.n, self._device_weights().cpu().numpy())

    def np_to_list(self):
        """:160-164 file-format arrays in self.weights -> working storage (here: the device buffer)"""
        if self._file_weights is not None:
            if self.weight_signature is None:
                self.weight_signature = engine.SIGNATURE[self.n]
            self._upload(self._file_weights)

    def __getstate__(self):
        d = dict(self.__dict__)
        d['weights'] = self.list_to_np() if (self._w is not None or self._file_weights is not None) else None
        d.pop('_w', None)
        d.pop('_file_weights', None)
        return d

    def __setstate__(self, d):
        d = dict(d)
        self._file_weights = d.pop('weights', None)          # the reference's pickles carry 'weights' in __dict__
        self.__dict__.update(d)
        self._w = None
        self.__dict__.setdefault('batch', 1024)
        self.__dict__.setdefault('update_mode', 'atomic')
        self.__dict__.setdefault('seed', None)
        self.__dict__.setdefault('next_game_id', self.__dict__.get('step', 0))
        if isinstance(self._file_weights, list) and self._file_weights and \
                not isinstance(self._file_weights[0], np.ndarray):
            # a reference pickle taken while weights were list-of-lists: regroup by signature
            sig, rows, o = self.weight_signature, self._file_weights, 0
            arrays = []
            for dcount in sig:
                arrays.append(np.array(rows[o:o + dcount], dtype=np.float32))
                o += dcount
            self.weights = arrays

    # ------------------------------------------------------------------ save / load (:166-200)
    def save_agent(self):
        """S3 layout as the reference: the agent without weights under a/, the weight arrays under weights/"""
        if not self.s3:
            with open(self.file, 'wb') as out:
                pickle.dump(self, out, pickle.HIGHEST_PROTOCOL)
            return
        shell = QAgent(name=self.name, with_weights=False)
        shell.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('weights', '_w', '_file_weights')})
        save_s3(shell, f'a/{self.file}')
        save_s3(self.list_to_np(), f'weights/{self.file}')

    def save_game(self, game):
        if not self.s3:
            return game.save_game(self.game_file)
        save_s3(game, f'g/{self.game_file}')

    @staticmethod
    def load_agent_local(file):
        """:188-193 (the reference opens the pickle in text mode, which cannot work on Python 3; fixed)"""
        with open(file, 'rb') as src:
            agent = pickle.load(src)
        agent.np_to_list()
        return agent

    @staticmethod
    def load_agent(file):
        """:195-200 `file` is the S3 key 'a/<name>.pkl'; its weights sit under 'weights/<name>.pkl'"""
        agent = load_s3(file)
        agent.weights = load_s3('weights/' + file[2:])
        agent.np_to_list()
        return agent

    # ------------------------------------------------------------------ evaluate / update (:202-214)
    def evaluate(self, row, score=None):
        ctx = engine.Context.get()
        b = ctx.to_device(np.array([pack_row(row)], dtype=np.uint64))
        return float(ctx.evaluate(self.n, self._device_weights(), b).cpu()[0])

    def evaluate_batch(self, rows):
        """values of many boards at once: rows [m,4,4] -> float32 [m]"""
        ctx = engine.Context.get()
        return ctx.evaluate(self.n, self._device_weights(), ctx.pack(np.asarray(rows, dtype=np.int32))).cpu().numpy()

    def update(self, row, dw):
        ctx = engine.Context.get()
        b = ctx.to_device(np.array([pack_row(row)], dtype=np.uint64))
        ctx.td_update(self.n, self._device_weights(), b, ctx.to_device(np.array([dw], dtype=np.float32)),
                      mode=cabi.UPD_ATOMIC | cabi.UPD_SUM)

    def _mode(self, batch):
        m = cabi.UPD_DETERMINISTIC if self.update_mode == 'deterministic' else cabi.UPD_ATOMIC
        return m | (cabi.UPD_MEAN if batch > 1 else cabi.UPD_SUM)

    def _next_seed(self):
        return random.getrandbits(63) if self.seed is None else int(self.seed)

    # ------------------------------------------------------------------ episode (:224-252)
    def episode(self, trace_len=1 << 15):
        """one TD(0) self-play game, exactly the reference's loop with B = 1 (Philox spawns keyed by a draw from
        `random`); returns the finished Game with moves (incl. the -1 sentinel) and tiles filled in.  The move /
        tile record holds trace_len entries: a longer game raises instead of returning a truncated record."""
        ctx = engine.Context.get()
        games = engine.GameBatch(1, seed=self._next_seed(), id_stride=0, ctx=ctx).init(first_id=self.next_game_id)
        self.next_game_id += 1
        start = unpack_board(games.to_host()['board'][0])
        tr = engine.TDTrainer(ctx, self.n, self._device_weights(), games, self.alpha, self._mode(1))
        import torch
        td = ctx.empty((1, trace_len), torch.int8).fill_(-2)
        ts = ctx.zeros((1, trace_len), torch.int16)
        while True:
            for _ in range(64):
                tr.step(trace=(td, None, None, ts, trace_len))
            if int(games.flags.cpu()[0]) & cabi.F_DONE:
                break
        if int(games.moves.cpu()[0]) >= trace_len:
            raise cabi.B2048Error(f'episode of {int(games.moves.cpu()[0])} moves does not fit trace_len={trace_len}: '
                                  'its moves / tiles record would be truncated; pass a larger trace_len')
        game = Game(row=start)
        game.starting_position = start.copy()
        game.adopt_device_result(games.to_host(), 0, td, ts)
        game.moves.append(-1)
        self.step += 1
        return game

    def _display_lr(self):
        self.print(f'episode = {self.step + 1}, current learning rate = {round(self.alpha, 4)}:')

    def decay_alpha(self):
        """:257-262"""
        decayed = max(self.alpha * self.decay, self.low_alpha_limit)
        self.alpha, self.next_decay = round(decayed, 4), self.step + self.decay_step
        rule = '-' * 6
        self.print(rule)
        self._display_lr()
        self.print(rule)

    # ------------------------------------------------------------------ train_run (:269-346)
    def _account(self, i, game, max_tile, book, saving):
        """the reference's per-episode statistics, reports and save cadence (:298-341) for episode number i"""
        book['ma100'].append(game.score)
        book['last1000'].append(game.score)
        if game.score > book['best'].score:
            book['best'] = game
            if game.score > self.top_score:
                self.top_game, self.top_score = game, game.score
                self.print(f'\nnew best game at episode {i}!\n{game}\n')
                if saving:
                    self.save_game(game)
                    self.print(f'game saved at {self.game_file}')
        if max_tile >= 10:
            book['reached'][min(max_tile, 16) - 10] += 1
        if max_tile > self.top_tile:                       # a new maximum tile also decays the learning rate
            self.top_tile = max_tile
            self.decay_alpha()
        if i % 100 == 0:
            ma = int(np.mean(book['ma100']))
            self.train_history.append(ma)
            self.print(f'episode {i}: score {game.score} reached {1 << max_tile} ma_100 = {ma}')
        if i % 1000:
            return
        now = time.time()
        lines = ['\n------', f'{round((now - book["lap"]) / 60, 2)} min', f'episode = {i}',
                 f'average over last 1000 episodes = {np.mean(book["last1000"])}']
        tail_counts = np.cumsum(book['reached'][::-1])[::-1] / 10          # share (%) reaching >= 2^(10+j)
        lines += [f'{1 << (j + 10)} reached in {float(r)} %' for j, r in enumerate(tail_counts) if r]
        lines += ['best of last 1000:', str(book['best']), 'best of this Agent:', str(self.top_game)]
        for line in lines:
            self.print(line)
        self._display_lr()
        self.print('------\n')
        if saving:
            self.save_agent()
            self.print(f'agent saved in {self.file}')
        book.update(lap=now, last1000=[], reached=[0] * 7, best=Game(row=np.zeros((4, 4), dtype=np.int32)))

    def train_run(self, num_eps=100000, add_weights='already', saving=True, stopper=None, batch=None, chunk=64):
        """Same driver as the reference (episode statistics, learning-rate decay, report and save cadence), with
        `batch` games advancing in lock-step on the GPU; episodes are accounted in completion order."""
        if add_weights == 'add':
            self.init_weights()
        elif add_weights != 'already':
            self.print('loading weights ...')
            self.weights = load_s3(add_weights)
            self.np_to_list()
        owner = _Owner(stopper)
        B = int(batch or self.batch)
        ctx = engine.Context.get()
        games = engine.GameBatch(B, seed=self._next_seed(), ctx=ctx, fin_cap=max(4 * B, 4096)).init(
            first_id=self.next_game_id)
        tr = engine.TDTrainer(ctx, self.n, self._device_weights(), games, self.alpha, self._mode(B))
        began = time.time()
        book = dict(ma100=deque(maxlen=100), last1000=[], reached=[0] * 7, lap=began,
                    best=Game(row=np.zeros((4, 4), dtype=np.int32)))
        self.print(f'Agent {self.name} training session started, current step = {self.step}')
        self.print('Agent will be saved every 1000 episodes and on STOP command')
        i, last = self.step, self.step + num_eps + 1       # the reference runs num_eps + 1 episodes (:284)
        while i < last and owner.go_on():
            if owner.lost():
                return
            tr.alpha = float(self.alpha)
            tr.run(chunk)
            for rec in games.drain_finished()[:last - i]:
                i += 1
                if self.step > self.next_decay and self.alpha > self.low_alpha_limit:
                    self.decay_alpha()
                self.step += 1
                game = Game(score=int(rec[2]), row=unpack_board((int(rec[6]) << 32) | int(rec[5])))
                game.odometer = int(rec[3])
                self._account(i, game, int(rec[4]), book, saving)
        # slot j plays ids first + j, first + j + B, ...: everything up to the largest id now in a slot is used up
        self.next_game_id = int(games.game_id.max().item()) + 1
        spent = int(time.time() - began)
        self.print(f'Total time = {spent // 60} min {spent % 60} sec')
        if saving:
            self.save_agent()
            self.print(f'{self.name} saved at step {self.step} in {self.file}\n------------------------\n')
        return games.read_counters()

    # ------------------------------------------------------------------ trial (:348-406)
    @staticmethod
    def _trial_report(results, elapsed, shuffles):
        """the reference's summary text (:384-399), same fields and the same rounding (2 decimals: a GPU batch shows
        0.0 ms there, so one extra line gives the per-move time in microseconds); `results` sorted by score, best
        first.  `time per shuffle` divides by the process-wide Game.counter like the reference (:398)."""
        reached = np.array([1 << int(np.max(g.row)) for g in results])
        moves = max(sum(g.odometer for g in results), 1)
        parts = ['\nBest games:'] + [f'{g}\n' for g in results[:3]]
        parts.append(f'average score of {len(results)} runs = {np.average([g.score for g in results])}')
        parts += [f'{tile} reached in {np.count_nonzero(reached >= tile) / len(results) * 100}%'
                  for tile in (16384, 8192, 4096, 2048, 1024)]
        parts += [f'total time = {round(elapsed, 2)}',
                  f'average time per move = {round(elapsed / moves * 1000, 2)} ms',
                  f'total number of shuffles = {Game.counter}',
                  f'time per shuffle = {round(elapsed / max(Game.counter, 1) * 1000, 2)} ms',
                  f'(average time per move = {round(elapsed / moves * 1e6, 3)} us, '
                  f'{shuffles} shuffles in this trial)']
        return '\n'.join(parts)

    @staticmethod
    def trial(estimator=None, agent_file=None, limit_tile=0, num=20, game_init=None, depth=0, width=1, since_empty=6,
              storage='s3', console='local', log_file=None, game_file=None, verbose=False, stopper=None, seed=None):
        display = Logger(log_file=log_file).add if console != 'local' else print
        owner = _Owner(stopper)
        if agent_file:
            display(f'Loading Agent from {agent_file} ...')
            agent = QAgent.load_agent(agent_file)
            estimator = agent.evaluate
            display(f'Trial run for {num} games, Agent = {agent.name}\n'
                    f'Looking forward: depth={depth}, width={width}, since_empty={since_empty}')
        else:
            agent = Game._agent_of(None, estimator) if estimator is not None else None
        began, counter0 = time.time(), Game.counter
        batched = agent is not None and not (verbose or stopper) and depth <= 4 and (depth == 0 or 1 <= width <= 4)
        if batched:
            results = QAgent._trial_device(agent, num, limit_tile, game_init, seed, display, depth, width, since_empty)
        else:
            results = []
            look = dict(limit_tile=limit_tile, depth=depth, width=width, since_empty=since_empty, verbose=verbose)
            while len(results) < num and owner.go_on():
                if owner.lost():
                    return
                t0 = time.time()
                game = game_init.copy() if game_init is not None else Game()
                game.trial_run(estimator, **look)
                display(f'game {len(results)}, result {game.score}, moves {game.odometer}, '
                        f'achieved {1 << np.max(game.row)}, time = {(time.time() - t0):.2f}')
                results.append(game)
        if not results:
            return
        results.sort(key=lambda g: -g.score)
        display(QAgent._trial_report(results, time.time() - began, Game.counter - counter0))
        if game_file:
            if storage == 's3':
                save_s3(results[0], game_file)
            else:
                results[0].save_game(file=game_file)
            display(f'Best game saved at {game_file}\n------------------------\n')
        return results

    @staticmethod
    def _trial_device(agent, num, limit_tile, game_init, seed, display, depth=0, width=1, since_empty=6):
        """all `num` games in one batch (b2048_greedy_play, or b2048_expectimax_play when depth > 0); the best game
        is replayed once with tracing so that its moves / tiles can be saved and replayed like a reference game"""
        def play(games, **kw):
            if depth == 0:
                return engine.greedy_play(ctx, agent.n, agent._device_weights(), games, limit_tile=limit_tile, **kw)
            return engine.expectimax_play(ctx, agent.n, agent._device_weights(), games, depth, width, since_empty,
                                          limit_tile=limit_tile, **kw)

        began = time.time()
        ctx = engine.Context.get()
        seed = random.getrandbits(63) if seed is None else int(seed)
        games = engine.GameBatch(num, seed=seed, ctx=ctx).init(first_id=0)
        if game_init is not None:
            games.set_positions(np.full(num, pack_row(game_init.row), dtype=np.uint64), [game_init.score] * num)
            games.game_id.copy_(ctx.to_device(np.arange(num, dtype=np.uint64)))
        starts = games.to_host()['board'].copy()
        play(games)
        h = games.to_host()
        c = games.read_counters()
        Game.counter += 4 * c['moves']                               # pre_move calls the reference would have made
        results = []
        for j in range(num):
            g = Game(score=int(h['score'][j]), row=unpack_board(h['board'][j]))
            g.odometer = int(h['moves'][j])
            g.starting_position = unpack_board(starts[j])
            g._replay_key = (seed, j)
            results.append(g)
        best = int(np.argmax(h['score']))
        one = engine.GameBatch(1, seed=seed, ctx=ctx)
        one.set_positions(starts[best:best + 1], None if game_init is None else [game_init.score])
        one.game_id.fill_(best)
        L = int(h['moves'][best]) + 1
        tdir, _, tsp = play(one, trace_len=L)
        results[best].adopt_device_result(one.to_host(), 0, tdir, tsp)
        per_game = (time.time() - began) / max(num, 1)              # the batch plays all games at once
        for j, g in enumerate(results):
            display(f'game {j}, result {g.score}, moves {g.odometer}, achieved {1 << np.max(g.row)}, '
                    f'time = {per_game:.2f}')
        return results


Q_agent = QAgent          # the reference README's older name (README.md:62)
