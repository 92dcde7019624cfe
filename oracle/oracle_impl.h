/*
 * oracle/oracle_impl.h -- TEST INFRASTRUCTURE ONLY (see oracle.c header).
 *
 * Precision-generic half of the oracle.  oracle.c includes this file twice:
 *   REAL=double, SUF=f64  -> the reference's arithmetic (Python floats are float64)
 *   REAL=float,  SUF=f32  -> the same statements in float32, i.e. the arithmetic the
 *                            device path computes in, so that the CUDA deterministic
 *                            mode can be compared BIT-EXACT with this restatement.
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* r_learning.py:202-203  evaluate: sum(weights[i][f_i(row)]) in table order, starting from 0. */
REAL FN(orc_evaluate)(int n, const REAL *w, const int32_t *row)
{
    int32_t f[ORC_MAX_FEAT];
    int F = orc_features(n, row, f);
    REAL v = 0;
    for (int i = 0; i < F; i++) v = v + w[orc_table_offset(n, i) + f[i]];
    return v;
}

/* r_learning.py:207-214  update: the 8 D4 images in the reference's order
 * r, r^T, R r, (R r)^T, ...  with R = rot90(transpose(transpose(.))) = rot90. */
void FN(orc_update)(int n, REAL *w, const int32_t *row_in, REAL dw)
{
    int32_t row[16], t[16], f[ORC_MAX_FEAT];
    memcpy(row, row_in, sizeof row);
    for (int k = 0; k < 4; k++) {
        int F = orc_features(n, row, f);
        for (int i = 0; i < F; i++) w[orc_table_offset(n, i) + f[i]] += dw;
        orc_transpose(row, t);                       /* :211 row = np.transpose(row) */
        F = orc_features(n, t, f);
        for (int i = 0; i < F; i++) w[orc_table_offset(n, i) + f[i]] += dw;
        orc_transpose(t, row);                       /* :214 np.rot90(np.transpose(row)) */
        orc_rot90(row, 1, t);
        memcpy(row, t, sizeof row);
    }
}

/* The (table offset + index) keys one update() touches, in the reference's order (8*F of them). */
int FN(orc_update_keys)(int n, const int32_t *row_in, int64_t *keys)
{
    int32_t row[16], t[16], f[ORC_MAX_FEAT];
    int m = 0;
    memcpy(row, row_in, sizeof row);
    for (int k = 0; k < 4; k++) {
        int F = orc_features(n, row, f);
        for (int i = 0; i < F; i++) keys[m++] = orc_table_offset(n, i) + f[i];
        orc_transpose(row, t);
        F = orc_features(n, t, f);
        for (int i = 0; i < F; i++) keys[m++] = orc_table_offset(n, i) + f[i];
        orc_transpose(t, row);
        orc_rot90(row, 1, t);
        memcpy(row, t, sizeof row);
    }
    return m;
}

typedef struct { REAL *delta; int32_t *gcount; int32_t *last; int64_t *touched; int64_t *qsum; } FN(orc_scratch);
static void FN(scratch_alloc)(FN(orc_scratch) *sc, int n, int64_t m);
static void FN(scratch_free)(FN(orc_scratch) *sc);
static int64_t FN(update_batch_impl)(int n, REAL *w, const uint64_t *boards, const REAL *dw, int64_t m, int rule,
                                     FN(orc_scratch) *sc);

/* update(state, dw) under one of the batch rules (0 = the reference's sequential update) */
static void FN(update_rule)(int n, REAL *w, const int32_t *row, REAL dw, int rule, FN(orc_scratch) *sc)
{
    if (rule == 0) { FN(orc_update)(n, w, row, dw); return; }
    uint64_t b = orc_pack(row);
    FN(update_batch_impl)(n, w, &b, &dw, 1, rule, sc);
}

/* r_learning.py:229-237 (== game_logic.py:150-161 at depth 0): scan d = 0..3, skip unchanged
 * directions, strict '>' so the lowest direction wins ties.  Returns the action (0 if none valid,
 * like the reference's initial 'action = 0'), and the best afterstate/score/value. */
static int FN(best_move)(int n, const REAL *w, const int32_t *row, int64_t score,
                         int32_t *best_row, int64_t *best_score, REAL *best_value, int *n_valid)
{
    int action = 0, any = 0;
    REAL bv = -INFINITY;
    for (int d = 0; d < 4; d++) {
        int32_t nr[16];
        int64_t ns;
        int ch = orc_pre_move(row, score, d, nr, &ns);
        if (ch < 0) return -1;
        if (ch) {
            REAL v = FN(orc_evaluate)(n, w, nr);
            any++;
            if (v > bv) {
                action = d; bv = v;
                memcpy(best_row, nr, sizeof nr);
                *best_score = ns;
            }
        }
    }
    *best_value = bv;
    if (n_valid) *n_valid = any;
    return action;
}

/*
 * r_learning.py:224-252  QAgent.episode, teacher-forced: the spawns come from the recorded
 * `tiles` list (game_logic.py:112-121 recorded at :121) instead of Python's `random`.
 *   start[16]            the reference game's starting_position (game_logic.py:66)
 *   tiles[3*k + 0..2]    (tile, i, j) of the k-th recorded spawn
 * Outputs (each sized for n_tiles + 1 entries): moves (with the -1 sentinel, :247), values
 * (best_value per step), dws (the dw applied at that step; NaN when state is None, :238).
 * Returns the number of moves made (odometer), or -1 on a reference KeyError, -2 when the
 * recorded spawn list is exhausted before the game is over.
 */
static int FN(episode_replay_body)(int n, REAL *w, REAL alpha, const int32_t *start,
                                   const int32_t *tiles, int n_tiles,
                                   int32_t *moves, REAL *values, REAL *dws,
                                   int32_t *final_row, int64_t *final_score, int rule, FN(orc_scratch) *sc);

/* rule: 0 = the reference's update (sequential adds); 1..4 = the batch rules of orc_update_batch with m = 1
 * (what the device does in its deterministic / merged modes), see below */
int FN(orc_episode_replay)(int n, REAL *w, REAL alpha, const int32_t *start,
                           const int32_t *tiles, int n_tiles,
                           int32_t *moves, REAL *values, REAL *dws,
                           int32_t *final_row, int64_t *final_score, int rule)
{
    FN(orc_scratch) sc = {0};
    if (rule) FN(scratch_alloc)(&sc, n, 1);
    int r = FN(episode_replay_body)(n, w, alpha, start, tiles, n_tiles, moves, values, dws, final_row, final_score,
                                    rule, &sc);
    if (rule) FN(scratch_free)(&sc);
    return r;
}

static int FN(episode_replay_body)(int n, REAL *w, REAL alpha, const int32_t *start,
                                   const int32_t *tiles, int n_tiles,
                                   int32_t *moves, REAL *values, REAL *dws,
                                   int32_t *final_row, int64_t *final_score, int rule, FN(orc_scratch) *sc)
{
    int32_t row[16], state[16], best_row[16];
    int64_t score = 0, best_score = 0;
    int have_state = 0, odo = 0;
    REAL old_label = 0;
    int F = orc_num_feat(n);
    memcpy(row, start, sizeof row);
    while (!orc_game_over(row)) {
        REAL best_value;
        int action = FN(best_move)(n, w, row, score, best_row, &best_score, &best_value, NULL);
        if (action < 0) return -1;
        REAL dw = NAN;
        if (have_state) {                                            /* :238-241 */
            dw = ((REAL)(best_score - score) + best_value - old_label) * alpha / (REAL)F;
            FN(update_rule)(n, w, state, dw, rule, sc);
        }
        memcpy(row, best_row, sizeof row);                           /* :242 */
        score = best_score;
        moves[odo] = action; values[odo] = best_value; dws[odo] = dw;
        memcpy(state, row, sizeof row);                              /* :245 */
        old_label = best_value; have_state = 1;
        if (odo >= n_tiles) return -2;
        row[tiles[3 * odo + 1] * 4 + tiles[3 * odo + 2]] = tiles[3 * odo];   /* :246 */
        odo++;
    }
    moves[odo] = -1;                                                 /* :247 */
    {
        REAL dw = -old_label * alpha / (REAL)F;                      /* :248-249 */
        if (have_state) FN(update_rule)(n, w, state, dw, rule, sc);
        values[odo] = 0; dws[odo] = dw;
    }
    memcpy(final_row, row, sizeof row);
    *final_score = score;
    return odo;
}

/*
 * game_logic.py:170-183  Game.trial_run at depth 0 (look_forward :215-216 == estimator call),
 * teacher-forced on the recorded spawns.  Returns odometer; no -1 sentinel (contrast episode).
 */
int FN(orc_trial_replay)(int n, const REAL *w, const int32_t *start,
                         const int32_t *tiles, int n_tiles, int limit_tile, int step_limit,
                         int32_t *moves, REAL *values, int32_t *final_row, int64_t *final_score)
{
    int32_t row[16], best_row[16];
    int64_t score = 0, best_score = 0;
    int odo = 0;
    memcpy(row, start, sizeof row);
    while (odo < step_limit) {
        if (orc_game_over(row)) break;
        if (limit_tile && orc_max_tile(row) >= limit_tile) break;
        REAL best_value;
        int action = FN(best_move)(n, w, row, score, best_row, &best_score, &best_value, NULL);
        if (action < 0) return -1;
        memcpy(row, best_row, sizeof row);
        score = best_score;
        moves[odo] = action; values[odo] = best_value;
        if (odo >= n_tiles) return -2;
        row[tiles[3 * odo + 1] * 4 + tiles[3 * odo + 2]] = tiles[3 * odo];
        odo++;
    }
    memcpy(final_row, row, sizeof row);
    *final_score = score;
    return odo;
}

/*
 * Greedy play of games [first_id, first_id + num) with the counter-based Philox spawn stream
 * (SPEC in oracle.c: orc_spawn_*).  Same loop as orc_trial_replay.  Games are independent, so
 * this is the one place the oracle uses all host threads (OpenMP) -- it is the CPU baseline for
 * the greedy metric.  Per game: final score, moves, max tile, final packed board; total number of
 * evaluate() calls (valid afterstates) is returned through n_eval.
 */
int64_t FN(orc_play_philox)(int n, const REAL *w, uint64_t seed, uint64_t first_id, int64_t num,
                            int limit_tile, int step_limit, int threads,
                            int64_t *scores, int32_t *n_moves, int32_t *max_tile,
                            uint64_t *final_board, int64_t *n_eval)
{
    int64_t total_moves = 0, total_eval = 0;
    int err = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : total_moves, total_eval)
#endif
    for (int64_t g = 0; g < num; g++) {
        int32_t row[16], best_row[16];
        int64_t score = 0, best_score = 0;
        int odo = 0, ovf = 0;
        uint64_t id = first_id + (uint64_t)g;
        orc_spawn_initial(seed, id, row);
        while (odo < step_limit) {
            if (orc_game_over(row)) break;
            if (limit_tile && orc_max_tile(row) >= limit_tile) break;
            REAL best_value;
            int nv = 0;
            int action = FN(best_move)(n, w, row, score, best_row, &best_score, &best_value, &nv);
            if (action < 0) { ovf = 1; break; }
            if (orc_max_tile(best_row) > 15) { ovf = 1; break; }   /* 2^16 escape: flag + stop */
            total_eval += nv;
            memcpy(row, best_row, sizeof row);
            score = best_score;
            odo++;
            orc_spawn_move(seed, id, (uint32_t)odo, row);
        }
        if (ovf) err = 1;
        total_moves += odo;
        if (scores) scores[g] = score;
        if (n_moves) n_moves[g] = odo;
        if (max_tile) max_tile[g] = orc_max_tile(row);
        if (final_board) final_board[g] = orc_pack(row);
    }
    if (n_eval) *n_eval = total_eval;
    (void)err;
    return total_moves;
}

/*
 * game_logic.py:214-243  Game.look_forward: the sampled expectimax above the estimator, with
 * estimator = QAgent.evaluate.  Statement by statement:
 *   :215-216  depth == 0                      -> evaluate(row)
 *   :217-219  empty_count(row) >= since_empty -> evaluate(row)
 *   :220-222  num_tiles = min(width, empty);  tile_positions = random.sample(empty cells, num_tiles)
 *   :226      new_tile = 1 if random.randrange(10) else 2          (drawn per position, in loop order)
 *   :229-231  game over after the spawn       -> best_value = -100
 *   :233-239  else max over the changed directions of look_forward(afterstate, depth - 1)
 *   :241-242  average += max(best_value, 0);  return average / num_tiles
 * Randomness comes from an orc_lf_rng:
 *   replay mode (pinning against the real reference): the logged results of random.sample /
 *     random.randrange, consumed in the reference's depth-first call order;
 *   Philox mode (the device's spec): node-keyed words.  A node's path code starts at 4 + root direction and
 *     becomes path * 16 + 4 * t + d going down through sampled tile t and direction d; positions come from
 *     counter (id_lo, id_hi, move_no, 2 | path << 8): the j-th position is the mulhi(word_j, left)-th of the
 *     `left` = empty - j empty cells not yet taken (row-major order); tiles from counter (.., 3 | path << 8):
 *     "4" iff mulhi(word_j, 10) == 0.  width <= 4.
 */
typedef struct {
    int replay;                  /* 1: consume logs; 0: Philox */
    const int32_t *log_pos;      /* flat positions, in draw order */
    const int32_t *log_tile;     /* tiles (1 | 2), in draw order */
    int64_t n_pos, n_tile, i_pos, i_tile;
    uint64_t seed, id;
    uint32_t move_no;
} FN(orc_lf_rng);

static REAL FN(look_forward_rec)(int n, const REAL *w, const int32_t *row, int64_t score, int depth, int width,
                                 int since_empty, FN(orc_lf_rng) *rng, uint32_t path, int *err)
{
    if (depth == 0) return FN(orc_evaluate)(n, w, row);
    int32_t cells[16];
    int empty = orc_empty(row, cells);
    if (empty >= since_empty) return FN(orc_evaluate)(n, w, row);
    int num = width < empty ? width : empty;
    int32_t pos[4];
    int32_t tile[4];
    if (num > 4) { *err = 1; return 0; }
    if (rng->replay) {
        if (rng->i_pos + num > rng->n_pos) { *err = 2; return 0; }
        for (int j = 0; j < num; j++) pos[j] = rng->log_pos[rng->i_pos++];
    } else {
        uint32_t wp[4], wt[4];
        orc_spawn_words(rng->seed, rng->id, rng->move_no, 2u | (path << 8), wp);
        orc_spawn_words(rng->seed, rng->id, rng->move_no, 3u | (path << 8), wt);
        int left = empty;
        for (int j = 0; j < num; j++, left--) {
            int k = (int)mulhi32(wp[j], (uint32_t)left);
            pos[j] = cells[k];
            for (int q = k; q + 1 < left; q++) cells[q] = cells[q + 1];     /* without replacement */
            tile[j] = mulhi32(wt[j], 10u) == 0 ? 2 : 1;
        }
    }
    REAL average = 0;
    for (int j = 0; j < num; j++) {
        if (rng->replay) {
            if (rng->i_tile >= rng->n_tile) { *err = 2; return 0; }
            tile[j] = rng->log_tile[rng->i_tile++];
        }
        int32_t nr[16];
        memcpy(nr, row, sizeof nr);
        nr[pos[j]] = tile[j];
        REAL best;
        if (orc_game_over(nr)) {
            best = -100;
        } else {
            best = -INFINITY;
            for (int d = 0; d < 4; d++) {
                int32_t tr[16];
                int64_t ts;
                int ch = orc_pre_move(nr, score, d, tr, &ts);
                if (ch < 0) { *err = 3; return 0; }                       /* the reference's KeyError */
                if (ch && orc_max_tile(tr) > 15) ch = 0;   /* 2^16 escape: the reference is undefined there; skipped */
                if (ch) {
                    REAL v = FN(look_forward_rec)(n, w, tr, ts, depth - 1, width, since_empty, rng,
                                                  path * 16u + 4u * (uint32_t)j + (uint32_t)d, err);
                    if (v > best) best = v;
                }
            }
        }
        average = average + (best > 0 ? best : 0);
    }
    return average / (REAL)num;
}

/* game_logic.py:150-161 _find_best_move above look_forward, then trial_run (:170-183) with the Philox spawn
 * stream: games [first_id, first_id + num) played with depth / width / since_empty look-ahead (the device's
 * b2048_expectimax_play).  A direction whose afterstate holds a 2^16 tile is skipped. */
int64_t FN(orc_play_expectimax)(int n, const REAL *w, uint64_t seed, uint64_t first_id, int64_t num, int depth, int width,
                                int since_empty, int limit_tile, int step_limit, int threads, int64_t *scores,
                                int32_t *n_moves, uint64_t *final_board)
{
    int64_t total_moves = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total_moves)
#endif
    for (int64_t g = 0; g < num; g++) {
        int32_t row[16];
        int64_t score = 0;
        int odo = 0;
        uint64_t id = first_id + (uint64_t)g;
        orc_spawn_initial(seed, id, row);
        while (odo < step_limit) {
            if (orc_game_over(row)) break;
            if (limit_tile && orc_max_tile(row) >= limit_tile) break;
            int best_dir = -1, err = 0;
            int32_t best_row[16];
            int64_t best_score = 0;
            REAL best_value = -INFINITY;
            FN(orc_lf_rng) rng = {0, NULL, NULL, 0, 0, 0, 0, seed, id, (uint32_t)odo};
            for (int d = 0; d < 4; d++) {
                int32_t nr[16];
                int64_t ns;
                int ch = orc_pre_move(row, score, d, nr, &ns);
                if (ch <= 0 || orc_max_tile(nr) > 15) continue;
                REAL v = FN(look_forward_rec)(n, w, nr, ns, depth, width, since_empty, &rng, 4u + (uint32_t)d, &err);
                if (v > best_value || best_dir < 0) {               /* :157 strict '>', first valid direction wins ties */
                    if (best_dir < 0 || v > best_value) { best_value = v; best_dir = d; memcpy(best_row, nr, sizeof nr); best_score = ns; }
                }
            }
            if (best_dir < 0 || err) break;
            memcpy(row, best_row, sizeof row);
            score = best_score;
            odo++;
            orc_spawn_move(seed, id, (uint32_t)odo, row);
        }
        total_moves += odo;
        if (scores) scores[g] = score;
        if (n_moves) n_moves[g] = odo;
        if (final_board) final_board[g] = orc_pack(row);
    }
    return total_moves;
}

/* look_forward values of m afterstates (rows [m,16]); Philox mode: root_dir[q] = the direction that produced
 * afterstate q (path code 4 + dir), ids/move_no per afterstate.  Returns 0, or an error code. */
int FN(orc_look_forward)(int n, const REAL *w, const int32_t *rows, const int64_t *scores, int64_t m, int depth,
                         int width, int since_empty, int replay, const int32_t *log_pos, int64_t n_pos,
                         const int32_t *log_tile, int64_t n_tile, uint64_t seed, const uint64_t *ids,
                         const uint32_t *move_no, const int32_t *root_dir, REAL *values)
{
    FN(orc_lf_rng) rng = {replay, log_pos, log_tile, n_pos, n_tile, 0, 0, seed, 0, 0};
    int err = 0;
    for (int64_t q = 0; q < m && !err; q++) {
        if (!replay) { rng.id = ids[q]; rng.move_no = move_no[q]; }
        values[q] = FN(look_forward_rec)(n, w, rows + 16 * q, scores ? scores[q] : 0, depth, width, since_empty,
                                         &rng, 4u + (uint32_t)(replay ? 0 : root_dir[q]), &err);
    }
    return err;
}

/*
 * update() for m (board, dw) entries sharing one table set -- the restatement b2048_td_update is
 * checked against.  Entries with NaN dw are skipped.  rule:
 *   0 : w[k] += dw one contribution at a time, entry order, reference key order (QAgent.update x m)
 *   1 : S[k] = sequential sum (from 0) of the contributions to k in entry order;  w[k] += S[k]
 *   2 : as 1 but w[k] += S[k] / G[k], G[k] = number of distinct entries contributing to k
 *   3 : exact fixed point (the device's DETERMINISTIC mode): Q[k] = sum of llrint(dw * 2^32) over the
 *       contributions (int64, order-independent);  w[k] += (REAL)((double)Q[k] / 2^32)
 *   4 : as 3 but w[k] += (REAL)((double)Q[k] / 2^32 / G[k])
 *       non-finite dw are skipped in rules 3/4
 * scratch: NULL, or {delta[num_weights] zeros, gcount[num_weights] zeros, last[num_weights] = -1,
 * touched[m*8*F]} kept clean across calls (used by orc_td_lockstep to avoid reallocating).
 */
static void FN(scratch_alloc)(FN(orc_scratch) *sc, int n, int64_t m)
{
    size_t nw = (size_t)orc_num_weights(n);
    sc->delta = (REAL *)calloc(nw, sizeof(REAL));
    sc->gcount = (int32_t *)calloc(nw, sizeof(int32_t));
    sc->last = (int32_t *)malloc(nw * sizeof(int32_t));
    memset(sc->last, 0xff, nw * sizeof(int32_t));
    sc->touched = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1) * 8 * ORC_MAX_FEAT);
    sc->qsum = (int64_t *)calloc(nw, sizeof(int64_t));
}

static void FN(scratch_free)(FN(orc_scratch) *sc)
{
    free(sc->delta); free(sc->gcount); free(sc->last); free(sc->touched); free(sc->qsum);
}

static int64_t FN(update_batch_impl)(int n, REAL *w, const uint64_t *boards, const REAL *dw, int64_t m, int rule,
                                     FN(orc_scratch) *sc)
{
    int64_t n_upd = 0, nt = 0;
    for (int64_t j = 0; j < m; j++) {
        if (dw[j] != dw[j]) continue;                 /* NaN = no update */
        if (rule >= 3 && !isfinite((double)dw[j])) continue;
        int32_t row[16];
        orc_unpack(boards[j], row);
        n_upd++;
        if (rule == 0) {
            FN(orc_update)(n, w, row, dw[j]);
        } else {
            int64_t keys[8 * ORC_MAX_FEAT];
            int nk = FN(orc_update_keys)(n, row, keys);
            int64_t qd = rule >= 3 ? (int64_t)llrint((double)dw[j] * 4294967296.0) : 0;
            for (int q = 0; q < nk; q++) {
                sc->delta[keys[q]] += dw[j];
                sc->qsum[keys[q]] += qd;
                if (sc->last[keys[q]] != (int32_t)j) { sc->last[keys[q]] = (int32_t)j; sc->gcount[keys[q]]++; }
                sc->touched[nt++] = keys[q];
            }
        }
    }
    for (int64_t q = 0; q < nt; q++) {
        int64_t k = sc->touched[q];
        if (sc->gcount[k]) {
            if (rule >= 3) {
                double u = (double)sc->qsum[k] / 4294967296.0;
                if (rule == 4) u = u / (double)sc->gcount[k];
                w[k] += (REAL)u;
            } else {
                w[k] += rule == 2 ? sc->delta[k] / (REAL)sc->gcount[k] : sc->delta[k];
            }
            sc->delta[k] = 0; sc->gcount[k] = 0; sc->last[k] = -1; sc->qsum[k] = 0;
        }
    }
    return n_upd;
}

int64_t FN(orc_update_batch)(int n, REAL *w, const uint64_t *boards, const REAL *dw, int64_t m, int rule)
{
    FN(orc_scratch) sc = {0};
    if (rule) FN(scratch_alloc)(&sc, n, m);
    int64_t r = FN(update_batch_impl)(n, w, boards, dw, m, rule, &sc);
    if (rule) FN(scratch_free)(&sc);
    return r;
}

/*
 * Lock-step batched TD(0): B game slots share one weight table (SURVEY 7.2 "B>1 lock-step").
 * One lock-step = for every slot, in slot order, the body of QAgent.episode's while loop
 * (r_learning.py:228-246) or, if the slot's game is over, its terminal update (:247-249) followed by
 * an in-place restart with game id += id_stride.  All slots evaluate with the weights W_t of the
 * start of the lock-step; their updates are applied after all evaluations:
 *   segmented == 0 : w[k] += dw one contribution at a time, slot order, reference key order
 *   segmented == 1 : S[k] = sum of the contributions to key k in slot order (starting from 0),
 *                    w[k] += S[k]                              (W_{t+1} = W_t + sum of deltas)
 *   segmented == 2 : as 1, but w[k] += S[k] / G[k] with G[k] = number of DISTINCT slots that
 *                    contributed to k in this lock-step ("per-key mean over games"; this is what
 *                    keeps B >> 1 stable at the reference's alpha; G = 1 when B = 1)
 * With B == 1 and segmented == 0 this is exactly QAgent.episode repeated.
 * State arrays are in/out so the call can be chained; `init` != 0 starts fresh games with ids
 * first_id + slot.  Returns the number of TD updates (update() call equivalents) performed.
 */
int64_t FN(orc_td_lockstep)(int n, REAL *w, REAL alpha, uint64_t seed, uint64_t first_id,
                            uint64_t id_stride, int B, int steps, int segmented, int init,
                            int threads,
                            uint64_t *board, int64_t *score, int32_t *odo, uint64_t *game_id,
                            uint64_t *state, REAL *old_label, uint8_t *have_state,
                            int64_t *fin_count, int64_t *fin_score_sum, int64_t *fin_moves_sum,
                            int32_t *fin_max_tile_hist /* [17] */, int64_t *n_moves_out)
{
    int F = orc_num_feat(n);
    int64_t n_updates = 0, n_moves = 0;
    uint64_t *upd_board = (uint64_t *)malloc(sizeof(uint64_t) * B);
    REAL *upd_dw = (REAL *)malloc(sizeof(REAL) * B);
    uint8_t *upd_on = (uint8_t *)malloc(B);
    FN(orc_scratch) sc = {0};
    if (segmented) FN(scratch_alloc)(&sc, n, B);
    if (init) {
        for (int j = 0; j < B; j++) {
            int32_t row[16];
            game_id[j] = first_id + (uint64_t)j;
            orc_spawn_initial(seed, game_id[j], row);
            board[j] = orc_pack(row);
            score[j] = 0; odo[j] = 0; state[j] = 0; old_label[j] = 0; have_state[j] = 0;
        }
    }
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    for (int s = 0; s < steps; s++) {
        /* phase A: every slot evaluates against W_t (read-only -> parallel over slots) */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
        for (int j = 0; j < B; j++) {
            int32_t row[16], best_row[16];
            orc_unpack(board[j], row);
            upd_on[j] = 0;
            if (orc_game_over(row)) {
                if (have_state[j]) {
                    upd_on[j] = 1; upd_board[j] = state[j];
                    upd_dw[j] = -old_label[j] * alpha / (REAL)F;
                }
                /* bookkeeping for the finished game is done in the serial phase below */
                upd_on[j] |= 2;
                continue;
            }
            int64_t best_score = 0;
            REAL best_value;
            int action = FN(best_move)(n, w, row, score[j], best_row, &best_score, &best_value, NULL);
            (void)action;
            if (have_state[j]) {
                upd_on[j] = 1; upd_board[j] = state[j];
                upd_dw[j] = ((REAL)(best_score - score[j]) + best_value - old_label[j]) * alpha / (REAL)F;
            }
            score[j] = best_score;
            odo[j] += 1;
            state[j] = orc_pack(best_row);
            old_label[j] = best_value;
            have_state[j] = 1;
            orc_spawn_move(seed, game_id[j], (uint32_t)odo[j], best_row);
            board[j] = orc_pack(best_row);
            upd_on[j] |= 4;
        }
        /* serial bookkeeping (slot order), then phase B = one batch update over the B slots */
        for (int j = 0; j < B; j++) {
            if (upd_on[j] & 4) n_moves++;
            if (upd_on[j] & 2) {
                int32_t row[16];
                orc_unpack(board[j], row);
                if (fin_count) (*fin_count)++;
                if (fin_score_sum) *fin_score_sum += score[j];
                if (fin_moves_sum) *fin_moves_sum += odo[j];
                if (fin_max_tile_hist) fin_max_tile_hist[orc_max_tile(row)]++;
                game_id[j] += id_stride;
                orc_spawn_initial(seed, game_id[j], row);
                board[j] = orc_pack(row);
                score[j] = 0; odo[j] = 0; state[j] = 0; old_label[j] = 0; have_state[j] = 0;
            }
            if (!(upd_on[j] & 1)) upd_dw[j] = NAN;
        }
        n_updates += FN(update_batch_impl)(n, w, upd_board, upd_dw, B, segmented, &sc);
    }
    free(upd_board); free(upd_dw); free(upd_on);
    if (segmented) FN(scratch_free)(&sc);
    if (n_moves_out) *n_moves_out = n_moves;
    return n_updates;
}

#undef FN
#undef CAT
#undef CAT_
