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
    """:6-13 heartbeat of the Dash session that owns a worker thread"""
    now = time.time()
    if (now - benchmark) > 2 * dash_intervals['check_run']:
        if RUNNING[parent] == 0:
            return 0
        RUNNING[parent] = 0
        return now
    return benchmark


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


class QAgent:
    """:85-406"""

    feature_functions = {2: f_2, 3: f_3, 4: f_4, 5: f_5, 6: f_6}
    parameter_shape = {2: (24, 16 ** 2), 3: (52, 16 ** 3), 4: (17, 16 ** 4), 5: (21, 16 ** 5), 6: (33, 0)}

    def __init__(self, name='agent', config_file=None, storage='s3', console='web', log_file=None, n=4, alpha=0.25,
                 decay=0.75, decay_step=10000, low_alpha_limit=0.01, with_weights=True,
                 batch=1024, update_mode='atomic', seed=None):
        self.name = name
        self.file = name + '.pkl'
        self.game_file = 'best_of_' + self.file
        self.s3 = (storage == 's3')
        self.log_file = log_file
        self.print = print if (console == 'local' or log_file is None) else Logger(log_file=log_file).add

        config = (load_s3(config_file) or {}) if config_file else {}
        self.n = config.get('n', n)
        self.alpha = config.get('alpha', alpha)
        self.decay = config.get('decay', decay)
        self.decay_step = config.get('decay_step', decay_step)
        self.low_alpha_limit = config.get('low_alpha_limit', low_alpha_limit)

        self.num_feat, self.size_feat = QAgent.parameter_shape[self.n]
        self.features = QAgent.feature_functions[self.n]

        self.step = 0
        self.top_game = None
        self.top_score = 0
        self.train_history = []
        self.next_decay = self.decay_step
        self.top_tile = 10

        # new knobs (not in the reference): games per lock-step, scatter mode, Philox seed
        self.batch = batch
        self.update_mode = update_mode
        self.seed = seed

        self._w = None                      # float32 device buffer (never pickled)
        self.weights = None                 # file-format arrays, only between load and first use
        self.weight_signature = None
        if with_weights:
            self.init_weights()

    def __str__(self):
        return f'Agent {self.name}, n={self.n}\ntrained for {self.step} episodes, top score = {self.top_score}'

    # ------------------------------------------------------------------ weights (:136-164)
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
        self.weights = None

    def _device_weights(self):
        if self._w is None:
            if self.weights is None:
                raise ValueError('agent has no weights (with_weights=False and nothing loaded)')
            self.np_to_list()
        return self._w

    def list_to_np(self):
        """:151-158 list of float32 arrays, one per signature group (the weight-file payload)"""
        return engine.arrays_from_flat(self.n, self._device_weights().cpu().numpy())

    def np_to_list(self):
        """:160-164 file-format arrays in self.weights -> working storage (here: the device buffer)"""
        if self.weights is not None:
            if self.weight_signature is None:
                self.weight_signature = engine.SIGNATURE[self.n]
            self._upload(self.weights)

    def __getstate__(self):
        d = dict(self.__dict__)
        d['weights'] = self.list_to_np() if (self._w is not None or self.weights is not None) else None
        d.pop('_w', None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._w = None
        self.__dict__.setdefault('batch', 1024)
        self.__dict__.setdefault('update_mode', 'atomic')
        self.__dict__.setdefault('seed', None)
        if isinstance(self.weights, list) and self.weights and not isinstance(self.weights[0], np.ndarray):
            # a reference pickle taken while weights were list-of-lists: regroup by signature
            sig, rows, o = self.weight_signature, self.weights, 0
            arrays = []
            for dcount in sig:
                arrays.append(np.array(rows[o:o + dcount], dtype=np.float32))
                o += dcount
            self.weights = arrays

    # ------------------------------------------------------------------ save / load (:166-200)
    def save_agent(self):
        if self.s3:
            nps = self.list_to_np()
            agent_params = QAgent(name=self.name, with_weights=False)
            for key in self.__dict__:
                if key not in ('weights', '_w'):
                    setattr(agent_params, key, getattr(self, key))
            save_s3(agent_params, 'a/' + self.file)
            save_s3(nps, 'weights/' + self.file)
        else:
            with open(self.file, 'wb') as f:
                pickle.dump(self, f, -1)

    def save_game(self, game):
        if self.s3:
            save_s3(game, 'g/' + self.game_file)
        else:
            game.save_game(self.game_file)

    @staticmethod
    def load_agent_local(file):
        """:188-193 (the reference opens the pickle in text mode, which cannot work on Python 3; fixed)"""
        with open(file, 'rb') as f:
            agent = pickle.load(f)
        agent.np_to_list()
        return agent

    @staticmethod
    def load_agent(file):
        agent = load_s3(file)
        agent.weights = load_s3(f'weights/{file[2:]}')
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
        `random`); returns the finished Game with moves (incl. the -1 sentinel) and tiles filled in"""
        ctx = engine.Context.get()
        games = engine.GameBatch(1, seed=self._next_seed(), id_stride=0, ctx=ctx).init(first_id=self.step)
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
        self.alpha = round(max(self.alpha * self.decay, self.low_alpha_limit), 4)
        self.next_decay = self.step + self.decay_step
        self.print('------')
        self._display_lr()
        self.print('------')

    # ------------------------------------------------------------------ train_run (:269-346)
    def train_run(self, num_eps=100000, add_weights='already', saving=True, stopper=None, batch=None, chunk=64):
        """Same driver as the reference (episode statistics, learning-rate decay, report and save cadence), with
        `batch` games advancing in lock-step on the GPU; episodes are accounted in completion order."""
        if add_weights == 'add':
            self.init_weights()
        elif add_weights != 'already':
            self.print('loading weights ...')
            self.weights = load_s3(add_weights)
            self.np_to_list()
        if stopper:
            parent, this_thread = stopper['parent'], stopper['a']
        B = int(batch or self.batch)
        ctx = engine.Context.get()
        games = engine.GameBatch(B, seed=self._next_seed(), ctx=ctx, fin_cap=max(4 * B, 4096)).init(first_id=self.step)
        tr = engine.TDTrainer(ctx, self.n, self._device_weights(), games, self.alpha, self._mode(B))
        av1000, ma100 = [], deque(maxlen=100)
        reached = [0] * 7
        best_of_1000 = Game(row=np.zeros((4, 4), dtype=np.int32))
        global_start = start = benchmark_time = time.time()
        self.print(f'Agent {self.name} training session started, current step = {self.step}')
        self.print('Agent will be saved every 1000 episodes and on STOP command')
        first, last = self.step + 1, self.step + num_eps + 1          # the reference runs num_eps + 1 episodes (:284)
        i = first - 1
        stop = False
        while i < last and not stop:
            if stopper:
                if AGENT_PANE[parent]['id'] != this_thread:
                    break
                benchmark_time = check_thread(parent, benchmark_time)
                if not benchmark_time:
                    return
            tr.alpha = float(self.alpha)
            tr.run(chunk)
            for rec in games.drain_finished():
                i += 1
                if i > last:
                    break
                if self.step > self.next_decay and self.alpha > self.low_alpha_limit:
                    self.decay_alpha()
                self.step += 1
                score, odo, max_tile = int(rec[2]), int(rec[3]), int(rec[4])
                game = Game(score=score, row=unpack_board((int(rec[6]) << 32) | int(rec[5])))
                game.odometer = odo
                ma100.append(score)
                av1000.append(score)
                if score > best_of_1000.score:
                    best_of_1000 = game
                    if score > self.top_score:
                        self.top_game, self.top_score = game, score
                        self.print(f'\nnew best game at episode {i}!\n{game.__str__()}\n')
                        if saving:
                            self.save_game(game)
                            self.print(f'game saved at {self.game_file}')
                if max_tile >= 10:
                    reached[min(max_tile, 16) - 10] += 1
                if max_tile > self.top_tile:
                    self.top_tile = max_tile
                    self.decay_alpha()
                if i % 100 == 0:
                    ma = int(np.mean(ma100))
                    self.train_history.append(ma)
                    self.print(f'episode {i}: score {score} reached {1 << max_tile} ma_100 = {ma}')
                if i % 1000 == 0:
                    average = np.mean(av1000)
                    self.print('\n------')
                    self.print(f'{round((time.time() - start) / 60, 2)} min')
                    start = time.time()
                    self.print(f'episode = {i}')
                    self.print(f'average over last 1000 episodes = {average}')
                    av1000 = []
                    for j in range(7):
                        r = sum(reached[j:]) / 10
                        if r:
                            self.print(f'{1 << (j + 10)} reached in {r} %')
                    reached = [0] * 7
                    self.print('best of last 1000:')
                    self.print(best_of_1000.__str__())
                    self.print('best of this Agent:')
                    self.print(self.top_game.__str__())
                    self._display_lr()
                    self.print('------\n')
                    if saving:
                        self.save_agent()
                        self.print(f'agent saved in {self.file}')
                    best_of_1000 = Game(row=np.zeros((4, 4), dtype=np.int32))
        total_time = int(time.time() - global_start)
        self.print(f'Total time = {total_time // 60} min {total_time % 60} sec')
        if saving:
            self.save_agent()
            self.print(f'{self.name} saved at step {self.step} in {self.file}\n------------------------\n')
        return games.read_counters()

    # ------------------------------------------------------------------ trial (:348-406)
    @staticmethod
    def trial(estimator=None, agent_file=None, limit_tile=0, num=20, game_init=None, depth=0, width=1, since_empty=6,
              storage='s3', console='local', log_file=None, game_file=None, verbose=False, stopper=None, seed=None):
        display = print if console == 'local' else Logger(log_file=log_file).add
        if stopper:
            parent, this_thread = stopper['parent'], stopper['a']
        agent = None
        if agent_file:
            display(f'Loading Agent from {agent_file} ...')
            agent = QAgent.load_agent(agent_file)
            estimator = agent.evaluate
            display(f'Trial run for {num} games, Agent = {agent.name}\n'
                    f'Looking forward: depth={depth}, width={width}, since_empty={since_empty}')
        elif estimator is not None:
            agent = Game._agent_of(None, estimator)
        start = benchmark_time = time.time()
        counter0 = Game.counter
        results = []
        if agent is not None and not verbose and not stopper and depth <= 4 and (depth == 0 or 1 <= width <= 4):
            results = QAgent._trial_device(agent, num, limit_tile, game_init, seed, display, depth, width, since_empty)
        else:
            for i in range(num):
                if stopper:
                    if AGENT_PANE[parent]['id'] != this_thread:
                        break
                    benchmark_time = check_thread(parent, benchmark_time)
                    if not benchmark_time:
                        return
                now = time.time()
                game = Game() if game_init is None else game_init.copy()
                game.trial_run(estimator, limit_tile=limit_tile, depth=depth, width=width, since_empty=since_empty,
                               verbose=verbose)
                display(f'game {i}, result {game.score}, moves {game.odometer}, achieved {1 << np.max(game.row)}, '
                        f'time = {(time.time() - now):.2f}')
                results.append(game)
        if not results:
            return
        average = np.average([v.score for v in results])
        figures = [(1 << np.max(v.row)) for v in results]
        total_odo = sum([v.odometer for v in results])
        results.sort(key=lambda v: v.score, reverse=True)

        def share(limit):
            return len([0 for v in figures if v >= limit]) / len(figures) * 100

        message = '\nBest games:\n'
        for v in results[:3]:
            message += v.__str__() + '\n' + '\n'
        elapsed = time.time() - start
        shuffles = max(Game.counter - counter0, 1)
        message += f'average score of {len(results)} runs = {average}\n' + \
                   f'16384 reached in {share(16384)}%\n' + f'8192 reached in {share(8192)}%\n' + \
                   f'4096 reached in {share(4096)}%\n' + f'2048 reached in {share(2048)}%\n' + \
                   f'1024 reached in {share(1024)}%\n' + f'total time = {round(elapsed, 2)}\n' + \
                   f'average time per move = {round(elapsed / max(total_odo, 1) * 1000, 4)} ms\n' + \
                   f'total number of shuffles = {Game.counter}\n' + \
                   f'time per shuffle = {round(elapsed / shuffles * 1000, 4)} ms'
        display(message)
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
        for j, g in enumerate(results):
            display(f'game {j}, result {g.score}, moves {g.odometer}, achieved {1 << np.max(g.row)}')
        return results


Q_agent = QAgent          # the reference README's older name (README.md:62)
