/*
 * b2048.h -- C-ABI of the B200-native 2048 / n-tuple hot path (libb2048.so, sm_100a).
 *
 * The reference (abachurin/2048) is pure Python and has no FFI layer; its boundary for this path
 * is the class surface of game2048/game_logic.py (Game) and game2048/r_learning.py (QAgent).
 * Every entry point below names the reference function (file:line under /root/reference) whose
 * work it performs for a whole batch.  The Python binding a maintainer would add is a ctypes stub
 * (INTEGRATION.md); 2048_b200/game2048/cabi.py is that stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name starts with
 *     h_; the caller allocates and frees everything, the library keeps no mutable global state
 *     (re-entrant, callable from any host thread);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - return value: 0 = ok, < 0 = B2048_E* argument error, > 0 = cudaError_t of the launch;
 *   - packed board: uint64, cell (r,c) = nibble at bit 4*(15-4r-c) holding the exponent
 *     (0 = empty, k = tile 2^k), i.e. the 16 hex digits read in row-major order
 *     (reference board: 4x4 int32 exponents, game_logic.py:41-45,62);
 *   - 2^16 escape: a nibble cannot hold exponent 16.  A move that would create a 2^16 tile sets an
 *     overflow bit and stores that cell saturated at 15; game loops treat the game as finished and
 *     count it in B2048_CTR_OVERFLOW (the reference raises KeyError there, game_logic.py:129);
 *   - directions: 0=left 1=up 2=right 3=down (game_logic.py:50);
 *   - weights: ONE contiguous float32 buffer, table i of agent n at b2048_table_offset(n, i), the
 *     order of the reference's weights[i] list (r_learning.py:136-149).
 */
#ifndef B2048_H
#define B2048_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2048_ABI_VERSION 1

#define B2048_EINVAL (-1)   /* bad argument (NULL pointer, negative size, unknown n) */
#define B2048_ENOTSUP (-2)  /* not supported on this device / configuration */
#define B2048_EWORK (-3)    /* workspace too small */

typedef void *b2048_stream_t;

/* ---- library / layout queries (host only, no device work) -------------------------------- */
int b2048_abi_version(void);
const char *b2048_strerror(int code);
/* QAgent.parameter_shape, r_learning.py:88: 24/52/17/21/33 features for n = 2..6; -1 otherwise */
int b2048_num_feat(int n);
/* first weight of table i (i = num_feat gives the total), r_learning.py:136-149 */
int64_t b2048_table_offset(int n, int i);
int64_t b2048_num_weights(int n);

/* ---- (1) packed boards -------------------------------------------------------------------- */
/* rows: int32 [m,16] row-major exponents  <->  boards: uint64 [m] */
int b2048_pack(const int32_t *rows, uint64_t *boards, int64_t m, b2048_stream_t stream);
int b2048_unpack(const uint64_t *boards, int32_t *rows, int64_t m, b2048_stream_t stream);

/* ---- (2) moves ---------------------------------------------------------------------------- */
/* create_table, game_logic.py:18-39.  lut: uint32 [65536], key = the packed 16-bit line:
 *   bits 0-15 line after "slide+merge left"; bits 16-19 / 20-23 exponents x of the (up to two)
 *   merges (score += 2^(x+1) each, 0 = none); bit 24 changed; bit 25 overflow (a 2^16 was made). */
#define B2048_LUT_ENTRIES 65536
int b2048_lut_build(uint32_t *lut, b2048_stream_t stream);

/* Game.pre_move for all four directions, game_logic.py:123-142, plus game_over (:109-110):
 *   after [m,4] afterstates, gain [m,4] merge score of the move (new_score - score),
 *   flags [m]: bit d = direction d changed the board, bit 4+d = direction d overflowed,
 *   over [m] (may be NULL): 1 iff no empty cell and no equal neighbours. */
int b2048_move4(const uint32_t *lut, const uint64_t *boards, int64_t m, uint64_t *after, uint32_t *gain,
                uint8_t *flags, uint8_t *over, b2048_stream_t stream);

/* Game.empty / empty_count / adjacent_pair_count / game_over, game_logic.py:96-110, and max tile.
 * stats: uint8 [m,4] = (empty_count, adjacent_pair_count, game_over, max exponent);
 * empty_mask: uint16 [m], bit p set iff flat cell p = 4r+c is empty (row-major order). NULL ok. */
int b2048_board_stats(const uint64_t *boards, int64_t m, uint8_t *stats, uint16_t *empty_mask,
                      b2048_stream_t stream);

/* ---- (3) new tiles ------------------------------------------------------------------------- */
/* Game.create_new_tile / new_tile, game_logic.py:112-121, from a counter-based Philox4x32-10 stream
 * instead of Python's `random`: key = seed, counter = (id_lo, id_hi, move_no, purpose 0);
 * tile = 2 ("4") iff mulhi(w0,10) == 0 else 1; cell = the mulhi(w1, n_empty)-th empty cell in
 * row-major order.  boards in/out; spawn (may be NULL) receives (tile << 8) | flat cell, 0xFFFF if
 * the board was full (the reference would raise IndexError). */
int b2048_spawn_philox(uint64_t *boards, int64_t m, uint64_t seed, const uint64_t *game_id,
                       const uint32_t *move_no, uint16_t *spawn, b2048_stream_t stream);
/* Game.__init__, game_logic.py:61-66: empty board + two spawns (words 0,1 then 2,3 of move_no 0);
 * game ids are first_id + i*id_step. */
int b2048_spawn_initial(uint64_t *boards, int64_t m, uint64_t seed, uint64_t first_id, uint64_t id_step,
                        b2048_stream_t stream);
/* replay mode, Game.replay game_logic.py:259-260: boards[i] cell pos[i] := tile[i]; entries with
 * tile == 0 are skipped (ragged batches). */
int b2048_spawn_replay(uint64_t *boards, int64_t m, const uint8_t *tile, const uint8_t *pos,
                       b2048_stream_t stream);

/* BASELINE config 5: move4 + a Philox spawn on every changed, non-overflowing afterstate, one pass.
 * spawn stream: ONE Philox block per board, counter = (index_lo, index_hi, 0, purpose 1), index = first_index + i;
 * direction d takes word d (low 16 bits: tile fraction, high 16 bits: cell fraction).  spawned [m,4] (NULL = no spawn pass).  Row LUTs are staged in shared
 * memory by a persistent grid. */
int b2048_sweep(const uint32_t *lut, const uint64_t *boards, int64_t m, uint64_t seed, uint64_t first_index,
                uint64_t *after, uint32_t *gain, uint8_t *flags, uint64_t *spawned, b2048_stream_t stream);

/* ---- (4) n-tuple agent ---------------------------------------------------------------------- */
/* f_2 .. f_6, r_learning.py:17-69: feat int32 [m, num_feat(n)] table indices in table order */
int b2048_features(int n, const uint64_t *boards, int64_t m, int32_t *feat, b2048_stream_t stream);
/* QAgent.evaluate, r_learning.py:202-203: value[i] = sum over tables, float32, table order */
int b2048_evaluate(int n, const float *weights, const uint64_t *boards, int64_t m, float *value,
                   b2048_stream_t stream);
/* QAgent.update, r_learning.py:207-214, for m (board, dw) entries sharing one table set: for the 8
 * D4 images of each board, weights[table][index] += dw.  Entries whose dw is NaN are skipped.
 * mode = execution | rule:
 *   B2048_UPD_ATOMIC         float32 red.global.add, unordered (run-to-run differences of a few ulp)
 *   B2048_UPD_DETERMINISTIC  exact segmented reduction by key: every contribution is quantised to
 *                            llrint(dw * 2^32) and summed per key in int64, which is order-independent,
 *                            so the result is bit-identical run to run, GPU to GPU and to the CPU
 *                            restatement; applied once per key as float((double)Q / 2^32 [/ G]).
 *                            Non-finite dw are skipped.
 *   | B2048_UPD_SORTED       (with DETERMINISTIC) do the reduction by key as radix sort of (key, entry)
 *                            + chunked segmented sums instead of directly with int64 atomics; same bits
 *   B2048_UPD_SUM            w[k] += S[k], S[k] = sum of the contributions to key k (the reference
 *                            rule; QAgent.update when m == 1)
 *   B2048_UPD_MEAN           w[k] += S[k] / G[k], G[k] = number of DISTINCT entries contributing to
 *                            k: the batched rule that stays stable at the reference's alpha when
 *                            many games hit the same key in one lock-step; == SUM at m == 1
 * One thread handles the 8 images of one (entry, table) pair; the accumulators are replicated per CTA group against
 * same-address atomics on hot keys (empty rows, small tiles).
 * delta (may be NULL) receives the same increments as weights (multi-GPU delta buffer).
 * work: b2048_td_update_workspace(n, m, mode) bytes (b2048_td_update itself ignores it for ATOMIC|SUM; the
 * fused loops below always need it); zero it once before the first call; every call leaves it ready for the
 * next one (accumulators and counters back at zero). */
#define B2048_UPD_ATOMIC 0
#define B2048_UPD_DETERMINISTIC 1
#define B2048_UPD_SUM 0
#define B2048_UPD_MEAN 2
#define B2048_UPD_SORTED 4
size_t b2048_td_update_workspace(int n, int64_t m, int mode);
int b2048_td_update(int n, float *weights, float *delta, const uint64_t *boards, const float *dw, int64_t m, int mode,
                    void *work, size_t work_bytes, b2048_stream_t stream);

/* ---- fused game loops ------------------------------------------------------------------------ */
/* Device-resident state of B game slots (structure of arrays; every pointer is a device pointer). */
typedef struct b2048_games {
    int64_t B;            /* slots */
    uint64_t *board;      /* [B] current board (after the spawn) */
    uint32_t *score;      /* [B] Game.score */
    uint32_t *moves;      /* [B] Game.odometer */
    uint64_t *game_id;    /* [B] Philox stream id of the game in the slot */
    uint64_t *state;      /* [B] TD: previous afterstate (r_learning.py:245 `state`) */
    float *old_label;     /* [B] TD: its value at the time (r_learning.py:245 `old_label`) */
    uint8_t *flags;       /* [B] B2048_F_* */
    uint64_t *counters;   /* [B2048_CTR_COUNT] accumulated by the kernels */
    uint32_t *tile_hist;  /* [17] finished games by max exponent */
    uint64_t seed;        /* Philox key */
    uint64_t id_stride;   /* TD in-place restart: game_id += id_stride (total slots over all ranks);
                             0 = no restart: a finished slot gets B2048_F_DONE (QAgent.episode, B = 1) */
    uint32_t *fin_log;    /* [fin_cap,8] (id lo, id hi, score, moves, max exponent, board lo, board hi, 0)
                             of finished games in completion order; head = counters[B2048_CTR_LOG]; NULL ok */
    int64_t fin_cap;      /* records that fit; later finishes are counted but not stored */
} b2048_games_t;

#define B2048_F_HAVE_STATE 1u /* TD: state/old_label valid */
#define B2048_F_DONE 2u       /* greedy: game finished (game over, limit tile, or overflow) */
#define B2048_F_OVERFLOW 4u   /* a 2^16 tile would have been created */

enum {
    B2048_CTR_MOVES = 0,    /* committed moves */
    B2048_CTR_EVALS = 1,    /* evaluate() calls (valid afterstates scored) */
    B2048_CTR_UPDATES = 2,  /* update() calls (8*num_feat weight RMWs each) */
    B2048_CTR_FINISHED = 3, /* games finished */
    B2048_CTR_SCORE_SUM = 4,
    B2048_CTR_MOVES_SUM = 5,
    B2048_CTR_OVERFLOW = 6,
    B2048_CTR_ACTIVE = 7,   /* greedy_play: slots still playing after the call (overwritten) */
    B2048_CTR_LOG = 8,      /* finished-game records appended to fin_log (host may reset to 0) */
    B2048_CTR_QUEUE = 9,    /* greedy_play: next slot to hand out (work queue of the running launch, overwritten) */
    B2048_CTR_FAULT = 10,   /* != 0: a persistent launch gave up at a grid barrier (a CTA was lost); the results of
                               that call and the workspace are invalid.  Never reset by the kernels. */
    B2048_CTR_COUNT = 16
};

/* Fresh games in every slot: ids first_id + i, Game.__init__ spawns, zero score/moves/flags;
 * zeroes counters and tile_hist when reset_counters != 0. */
int b2048_games_init(const b2048_games_t *g, uint64_t first_id, int reset_counters, b2048_stream_t stream);

/* Game.trial_run at depth 0, game_logic.py:150-183, for every slot that is not DONE: up to
 * max_steps moves per slot in ONE launch (each game advances independently; weights are read-only).
 * limit_tile as in trial_run (exponent; 0 = none); step_limit = total odometer cap (100000 there).
 * replay == NULL: Philox spawns.  Otherwise replay mode: tile/pos are [B, replay_len] recorded
 * spawns (tile 0 = exhausted -> slot stops, stays not DONE), indexed by the slot's odometer.
 * trace_dir / trace_value / trace_spawn (each may be NULL): [B, trace_len] per move: chosen direction,
 * its value, and the spawn that followed ((tile << 8) | flat cell) -- Game.moves / Game.tiles. */
typedef struct b2048_replay {
    const uint8_t *tile;  /* [B, len] */
    const uint8_t *pos;   /* [B, len] flat cell 4r+c */
    int64_t len;
} b2048_replay_t;
int b2048_greedy_play(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, int max_steps,
                      int limit_tile, int step_limit, const b2048_replay_t *replay, int8_t *trace_dir,
                      float *trace_value, uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream);

/* Game.look_forward, game_logic.py:214-243, with estimator = QAgent.evaluate, for m afterstates: the sampled
 * expectimax of the reference (depth 0..4, width 1..4, since_empty as there; a leaf or a node with
 * empty_count >= since_empty is evaluate(); otherwise min(width, empty) empty cells are sampled without
 * replacement with a 2/4 tile each, and the value is the mean of max(0, game over ? -100 : best direction)).
 * The reference draws from Python's `random`; here every node draws node-keyed Philox words: key = seed, counter =
 * (id_lo, id_hi, move_no, purpose | path << 8), purpose 2 = positions, 3 = tiles, path = 4 + root_dir at the root
 * afterstate and path * 16 + 4 * tile_index + direction below it.  root_dir[q] = the direction that produced
 * afterstate q.  One warp per afterstate.  A direction that would create a 2^16 tile is skipped. */
int b2048_look_forward(int n, const float *weights, const uint32_t *lut, const uint64_t *boards, const uint64_t *game_id,
                       const uint32_t *move_no, const uint8_t *root_dir, int64_t m, int depth, int width, int since_empty,
                       uint64_t seed, float *value, b2048_stream_t stream);
/* Game.trial_run with look-ahead, game_logic.py:150-183 (_find_best_move above look_forward, strict '>' over
 * d = 0..3, commit, Philox spawn) for every slot that is not DONE, up to max_steps moves per slot in one launch,
 * one warp per game; move_no of a node = the slot's odometer before the move.  Counters as b2048_greedy_play
 * (B2048_CTR_EVALS counts the evaluate() calls at the leaves and cut-offs of the trees); trace_dir / trace_spawn
 * [B, trace_len] as there (may be NULL). */
int b2048_expectimax_play(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, int max_steps,
                          int limit_tile, int step_limit, int depth, int width, int since_empty, int8_t *trace_dir,
                          uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream);

/* One lock-step of QAgent.episode (r_learning.py:224-252) for all B slots, two launches:
 *   phase A (every slot, weights W_t read-only): game over -> terminal dw = -old_label*alpha/F
 *     (:247-249), statistics, in-place restart; else best afterstate by strict '>' over d=0..3
 *     (:229-237), dw = (gain + best_value - old_label)*alpha/F if a previous afterstate exists
 *     (:238-241), commit, remember (state, old_label) (:242-245), spawn (:246).
 *   phase B: b2048_td_update(previous state, dw) over the B slots (mode, work as there, m = B).
 * upd_board [B], upd_dw [B] are scratch owned by the caller (dw = NaN marks "no update").
 * delta (may be NULL): second accumulation buffer of the same shape as weights that receives the
 * same updates (multi-GPU: allreduced every K steps, then b2048_delta_apply).
 * replay / trace as in b2048_greedy_play (trace_len entries per slot, indexed by odometer; replay
 * mode never restarts a slot: a finished slot gets DONE after its terminal update). */
int b2048_td_step(int n, float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                  int mode, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                  const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value, float *trace_dw,
                  uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream);
/* phase A alone (one launch); b2048_td_step == b2048_td_phase_a + b2048_td_update(upd_board, upd_dw, B).
 * Exposed so that callers can time or overlap the gather and the scatter halves separately. */
int b2048_td_phase_a(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, float alpha,
                     uint64_t *upd_board, float *upd_dw, const b2048_replay_t *replay, int8_t *trace_dir,
                     float *trace_value, float *trace_dw, uint16_t *trace_spawn, int64_t trace_len,
                     b2048_stream_t stream);
/* `steps` lock-steps with no host work in between.  Default: ONE cooperative launch of a persistent kernel
 * (every CTA owns a fixed range of slots; phase A and the accumulation run back to back, a grid barrier, the
 * CTA applies the keys it touched first, a grid barrier; weights, delta and work must be 16-byte aligned) -- same results as
 * `steps` calls of b2048_td_step
 * (bit-identical in the DETERMINISTIC modes).  mode | B2048_RUN_STEPWISE (or B2048_UPD_SORTED, or a device
 * without cooperative launch) enqueues b2048_td_step `steps` times instead (3 launches per lock-step). */
#define B2048_RUN_STEPWISE 8
/* mode | B2048_RUN_GENERIC: the persistent kernel in its generic slot layout (rounds over the staging arrays and per-CTA
 * key lists) even where the one-round register layout would fit; same results, used by the parity tests */
#define B2048_RUN_GENERIC 16
/* the apply phase of the persistent kernel: a dense scan of the accumulators (default when the batch needs the generic slot
 * layout; n <= 5) or the lists of first-touched keys.  mode | B2048_RUN_SCAN /
 * B2048_RUN_LISTS force one of them (same results; the parity tests run both). */
#define B2048_RUN_SCAN 32
#define B2048_RUN_LISTS 64
/* In the generic layout with the scanning apply the persistent kernel splits the slots over its CTAs in proportion to the
 * slots per kilo-cycle each CTA sustained in the PREVIOUS launch on this workspace (from 128 slots per CTA up; SMs differ
 * by up to 1.5x in the atomic rate they sustain; the first launch on a zeroed workspace splits evenly).  Results do not depend on the split
 * (bit-identical in the DETERMINISTIC modes).  mode | B2048_RUN_EVEN keeps the even split. */
#define B2048_RUN_EVEN 128
/* number of kernel launches b2048_td_run(n, B slots, mode, steps) enqueues on this device (1 = persistent) */
int64_t b2048_td_run_launches(int n, int64_t B, int mode, int steps);
int b2048_td_run(int n, float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                 int mode, int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                 b2048_stream_t stream);

/* multi-GPU weight sync.  Every rank accumulates its increments in `delta` (while also applying
 * them locally); every K lock-steps:
 *   b2048_delta_pack:  packed[0..count) = delta, packed[count..2count) = (delta != 0) ? 1 : 0
 *   allreduce(sum) of packed over the ranks (NCCL, in place)
 *   b2048_delta_apply: w_sync[k] += sum_k / max(1, contributors_k);  weights[k] = w_sync[k];  delta[k] = 0
 * i.e. the per-key-mean rule applied across ranks (a key only one rank touched keeps its full update; a key
 * every rank touched gets their mean) -- one fused pass, replicas end bit-identical.  contributors == NULL
 * gives the plain sum  w_sync += delta_sum. */
int b2048_delta_pack(const float *delta, float *packed, int64_t count, b2048_stream_t stream);
/* The same packing without a delta buffer in the training kernels: delta := weights - w_sync (what this rank's
 * weights moved by since the last sync), packed[count..2count) = (weights != w_sync).  With it the hot loop needs
 * no second accumulation (delta = NULL in b2048_td_run) and b2048_delta_apply takes delta = NULL. */
int b2048_delta_pack_diff(const float *weights, const float *w_sync, float *packed, int64_t count,
                          b2048_stream_t stream);
int b2048_delta_apply(float *weights, float *w_sync, float *delta, const float *delta_sum, const float *contributors,
                      int64_t count, b2048_stream_t stream);


/* The same exchange with one BIT per weight instead of a float indicator (4 + 1/8 bytes per weight and rank on the wire
 * instead of 8): delta [count] = weights - w_sync, bits [ceil(count/32)] bit i%32 of word i/32 = (weights[i] != w_sync[i]);
 * allreduce(sum) of delta and allgather of the bit planes over the ranks; then with bits_all = [world, ceil(count/32)]:
 *   w_sync[k] += delta_sum[k] / max(1, number of ranks whose bit k is set);  weights[k] = w_sync[k]. */
int b2048_delta_pack_bits(const float *weights, const float *w_sync, float *delta, uint32_t *bits, int64_t count,
                          b2048_stream_t stream);
int b2048_delta_apply_bits(float *weights, float *w_sync, const float *delta_sum, const uint32_t *bits_all, int world,
                           int64_t count, b2048_stream_t stream);

/* The whole exchange as ONE kernel per rank over NVLink / NVSwitch peer memory (no NCCL): rank r reduces the r-th
 * contiguous slice of the weights -- remote loads of every rank's weights and w_sync, sum of the deltas in rank order,
 * contributor count, w_sync + sum / max(1, contributors) -- and stores the result into weights and w_sync of EVERY rank,
 * so all replicas hold the owner's bits.  w / w_sync / flags: the buffers of rank q as mapped into THIS process (CUDA IPC
 * or symmetric memory; entry `rank` = the local buffers); each flags[q] is B2048_PEER_FLAG_WORDS zero-initialised uint32.
 * Ranks rendezvous inside the kernel through flags: every rank must call with the same count and the same epoch, a
 * value that increases by one per call starting at 1.  max_ctas: grid cap (0 = 4 per SM); all ranks' kernels must be
 * resident at the same time (one per GPU; several emulated ranks on ONE GPU need streams and a cap).
 * flags[rank][B2048_PEER_FAULT] != 0 afterwards: a peer never arrived (2^26 polls), the sync did not complete. */
#define B2048_MAX_PEERS 16
#define B2048_PEER_ARRIVE 0
#define B2048_PEER_DONE 16
#define B2048_PEER_TICKET 32
#define B2048_PEER_FAULT 33
#define B2048_PEER_FLAG_WORDS 64
typedef struct b2048_peers {
    float *w[B2048_MAX_PEERS];
    float *w_sync[B2048_MAX_PEERS];
    uint32_t *flags[B2048_MAX_PEERS];
    int world, rank;
} b2048_peers_t;
int b2048_sync_peers(const b2048_peers_t *peers, int64_t count, uint32_t epoch, int max_ctas, b2048_stream_t stream);

/* b2048_td_run with the exchange INSIDE the persistent launch: `steps` lock-steps in one cooperative kernel, and after
 * every lock-step that completes a period of sync_every (since_sync = lock-steps already done in the current period, so
 * the first sync comes after sync_every - since_sync lock-steps) all CTAs of every rank run the b2048_sync_peers exchange
 * on the spot -- game state stays in registers, no relaunch and no separate sync kernel.  The syncs of the call use the
 * epochs epoch, epoch + 1, ... (same numbering as b2048_sync_peers; the two may be mixed on one flag block).  Every rank
 * must call with the same n, steps, sync_every, since_sync and epoch; weights must be peers->w[peers->rank].
 * Returns B2048_ENOTSUP when the persistent kernel cannot be used (no cooperative launch, or the batch does not fit one
 * wave): the caller then alternates b2048_td_run and b2048_sync_peers.  ATOMIC or DETERMINISTIC, SUM or MEAN. */
int b2048_td_run_peers(int n, float *weights, const uint32_t *lut, const b2048_games_t *g, float alpha, int mode, int steps,
                       uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes, const b2048_peers_t *peers,
                       int sync_every, int since_sync, uint32_t epoch, b2048_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2048_H */
