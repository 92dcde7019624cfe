/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, CPU restatement of the abachurin/2048 hot path (game2048/game_logic.py and
 * game2048/r_learning.py), written from the reference's *behaviour*; every function cites the
 * reference file:line (relative to /root/reference) it follows.  It exists so that the CUDA path
 * can be checked on a GPU box where the Python reference is not available.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs -- as the CHECKER or the timed CPU baseline, never as a fallback for the product.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against tests/golden/ (fixtures produced by
 * tests/golden/gen_golden.py from the real reference imported in the build container: the full
 * 65,536-entry move table, pre_move/game_over/empty/features on seeded random boards, the D4 order
 * of update(), recorded greedy games and recorded TD episodes with per-step values and final
 * weights), and tests/test_oracle_vs_reference.py re-checks it live against /root/reference when
 * that directory exists.  The reference itself has no tests.
 *
 * Board representation here is deliberately the reference's (int32 exponents, 4x4 row-major,
 * game_logic.py:41-45,62), NOT the packed device format, so that pack/unpack is checked too.
 *
 * Not in the reference (new specification, mirrored bit-exactly by the CUDA kernels):
 *   - packed board: cell (r,c) is the nibble at bit 4*(15-4r-c) of a uint64 (reads as 16 hex digits
 *     in row-major order)
 *   - Philox4x32-10 counter-based spawn stream (Salmon et al., SC'11), orc_spawn_*
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_FEAT 52

/* ------------------------------------------------------------------------------------------- */
/* move table: game_logic.py:18-39 (create_table), bound at :51                                  */
/* ------------------------------------------------------------------------------------------- */
static uint8_t T_line[65536][4];
static uint32_t T_score[65536];
static uint8_t T_changed[65536];
static int T_ready = 0;

static void table_init(void)
{
    if (T_ready) return;
    for (int a = 0; a < 16; a++)
        for (int b = 0; b < 16; b++)
            for (int c = 0; c < 16; c++)
                for (int d = 0; d < 16; d++) {
                    int key = (a << 12) | (b << 8) | (c << 4) | d;
                    int line[4] = {a, b, c, d};
                    uint32_t score = 0;
                    /* :26  (len(set(line)) == 4 and min(line)) or (not max(line)) */
                    int distinct = (a != b) && (a != c) && (a != d) && (b != c) && (b != d) && (c != d);
                    int mn = a, mx = a;
                    for (int i = 1; i < 4; i++) { if (line[i] < mn) mn = line[i]; if (line[i] > mx) mx = line[i]; }
                    if ((distinct && mn) || !mx) {
                        for (int i = 0; i < 4; i++) T_line[key][i] = (uint8_t)line[i];
                        T_score[key] = 0; T_changed[key] = 0;
                        continue;
                    }
                    int l1[4], n1 = 0;                         /* :29 drop zeros */
                    for (int i = 0; i < 4; i++) if (line[i]) l1[n1++] = line[i];
                    for (int i = 0; i + 1 < n1; i++) {          /* :30-34 single left-to-right pass */
                        int x = l1[i];
                        if (x == l1[i + 1]) {
                            score += 1u << (x + 1);
                            l1[i] = x + 1; l1[i + 1] = 0;
                        }
                    }
                    int l2[4] = {0, 0, 0, 0}, n2 = 0;          /* :35-36 drop zeros, right-pad */
                    for (int i = 0; i < n1; i++) if (l1[i]) l2[n2++] = l1[i];
                    int changed = 0;
                    for (int i = 0; i < 4; i++) { T_line[key][i] = (uint8_t)l2[i]; if (l2[i] != line[i]) changed = 1; }
                    T_score[key] = score; T_changed[key] = (uint8_t)changed;   /* :37 */
                }
    T_ready = 1;
}

/* Export the whole table (lines may contain 16 = the reference's unrepresentable 2^16 tile). */
void orc_create_table(uint8_t *lines /*65536*4*/, uint32_t *score, uint8_t *changed)
{
    table_init();
    memcpy(lines, T_line, sizeof T_line);
    memcpy(score, T_score, sizeof T_score);
    memcpy(changed, T_changed, sizeof T_changed);
}

/* ------------------------------------------------------------------------------------------- */
/* numpy view ops used by the reference                                                          */
/* ------------------------------------------------------------------------------------------- */
/* np.rot90(m, k): k counter-clockwise quarter turns; k=1: out[i][j] = in[j][3-i] */
void orc_rot90(const int32_t *in, int k, int32_t *out)
{
    int32_t a[16], b[16];
    memcpy(a, in, sizeof a);
    k = ((k % 4) + 4) % 4;
    for (int t = 0; t < k; t++) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) b[i * 4 + j] = a[j * 4 + (3 - i)];
        memcpy(a, b, sizeof a);
    }
    memcpy(out, a, sizeof a);
}

void orc_transpose(const int32_t *in, int32_t *out)
{
    int32_t b[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) b[i * 4 + j] = in[j * 4 + i];
    memcpy(out, b, sizeof b);
}

/* ------------------------------------------------------------------------------------------- */
/* board ops: game_logic.py:96-148                                                               */
/* ------------------------------------------------------------------------------------------- */
/* :96-99  empty: positions of zeros in row-major order (np.where) */
int orc_empty(const int32_t *row, int32_t *pos /*16, flat index 4*i+j*/)
{
    int m = 0;
    for (int p = 0; p < 16; p++) if (row[p] == 0) pos[m++] = p;
    return m;
}

/* :101-103 */
int orc_empty_count(const int32_t *row)
{
    int nz = 0;
    for (int p = 0; p < 16; p++) nz += row[p] != 0;
    return 16 - nz;
}

/* :105-107  24 - nonzero horizontal differences - nonzero vertical differences */
int orc_adjacent_pair_count(const int32_t *row)
{
    int nzh = 0, nzv = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) nzh += (row[i * 4 + j] - row[i * 4 + j + 1]) != 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) nzv += (row[i * 4 + j] - row[(i + 1) * 4 + j]) != 0;
    return 24 - nzh - nzv;
}

/* :109-110 */
int orc_game_over(const int32_t *row)
{
    return !orc_empty_count(row) && !orc_adjacent_pair_count(row);
}

int orc_max_tile(const int32_t *row)
{
    int m = 0;
    for (int p = 0; p < 16; p++) if (row[p] > m) m = row[p];
    return m;
}

/* :123-134  _left.  Returns change (0/1), or -1 where the reference raises KeyError (cell > 15). */
int orc_left(const int32_t *row, int64_t score, int32_t *new_row, int64_t *new_score)
{
    int change = 0;
    table_init();
    memcpy(new_row, row, 16 * sizeof(int32_t));
    *new_score = score;
    for (int i = 0; i < 4; i++) {
        const int32_t *r = row + 4 * i;
        if ((r[0] | r[1] | r[2] | r[3]) & ~15) return -1;
        int key = (r[0] << 12) | (r[1] << 8) | (r[2] << 4) | r[3];
        if (T_changed[key]) {
            change = 1;
            *new_score += T_score[key];
            for (int j = 0; j < 4; j++) new_row[4 * i + j] = T_line[key][j];
        }
    }
    return change;
}

/* :136-142  pre_move: rot90(row, d) -> _left -> rot90(., 4-d).  0=left 1=up 2=right 3=down (:50) */
int orc_pre_move(const int32_t *row, int64_t score, int direction, int32_t *new_row, int64_t *new_score)
{
    int32_t a[16], b[16];
    if (direction) orc_rot90(row, direction, a); else memcpy(a, row, sizeof a);
    int change = orc_left(a, score, b, new_score);
    if (change < 0) return change;
    if (direction) orc_rot90(b, 4 - direction, new_row); else memcpy(new_row, b, sizeof b);
    return change;
}

/* batch helper for the tests: boards[16*m] -> 4 directions each */
int orc_pre_move_batch(const int32_t *rows, const int64_t *scores, int64_t m,
                       int32_t *new_rows /*m*4*16*/, int64_t *new_scores /*m*4*/, int8_t *change /*m*4*/)
{
    for (int64_t q = 0; q < m; q++)
        for (int d = 0; d < 4; d++)
            change[q * 4 + d] = (int8_t)orc_pre_move(rows + 16 * q, scores ? scores[q] : 0, d,
                                                      new_rows + (q * 4 + d) * 16, new_scores + q * 4 + d);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* packed board (new spec): cell (r,c) <-> nibble at bit 4*(15 - 4r - c)                          */
/* ------------------------------------------------------------------------------------------- */
uint64_t orc_pack(const int32_t *row)
{
    uint64_t b = 0;
    for (int p = 0; p < 16; p++) b |= (uint64_t)(row[p] & 15) << (4 * (15 - p));
    return b;
}

void orc_unpack(uint64_t b, int32_t *row)
{
    for (int p = 0; p < 16; p++) row[p] = (int32_t)((b >> (4 * (15 - p))) & 15);
}

/* ------------------------------------------------------------------------------------------- */
/* n-tuple features: r_learning.py:17-69; shapes :88; table layout :136-149                      */
/* ------------------------------------------------------------------------------------------- */
int orc_num_feat(int n)
{
    switch (n) { case 2: return 24; case 3: return 52; case 4: return 17; case 5: return 21; case 6: return 33; }
    return -1;
}

/* size of table i (r_learning.py:88, :136-149; n=6 cutoff 14 hard-coded at :138) */
int64_t orc_table_size(int n, int i)
{
    if (n == 2) return 256;
    if (n == 3) return 4096;
    if (i < 17) return 65536;
    if (i < 21) return 1048576;
    return 7529536; /* 14^6 */
}

int64_t orc_table_offset(int n, int i)
{
    if (n == 2) return 256ll * i;
    if (n == 3) return 4096ll * i;
    if (i <= 17) return 65536ll * i;
    if (i <= 21) return 65536ll * 17 + 1048576ll * (i - 17);
    return 65536ll * 17 + 1048576ll * 4 + 7529536ll * (i - 21);
}

int64_t orc_num_weights(int n) { return orc_table_offset(n, orc_num_feat(n)); }

#define X(r, c) x[(r) * 4 + (c)]
#define Y(r, c) y[(r) * 4 + (c)]

static int feat_4(const int32_t *x, int32_t *f)          /* r_learning.py:40-44 */
{
    int m = 0;
    for (int c = 0; c < 4; c++) f[m++] = (X(0, c) << 12) + (X(1, c) << 8) + (X(2, c) << 4) + X(3, c);
    for (int r = 0; r < 4; r++) f[m++] = (X(r, 0) << 12) + (X(r, 1) << 8) + (X(r, 2) << 4) + X(r, 3);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            f[m++] = (X(r, c) << 12) + (X(r + 1, c) << 8) + (X(r, c + 1) << 4) + X(r + 1, c + 1);
    return m;
}

static int feat_middle(const int32_t *x, int32_t *f)     /* r_learning.py:52-53 */
{
    int m = 0;
    for (int i = 1; i < 3; i++)
        for (int j = 1; j < 3; j++)
            f[m++] = (X(i, j) << 16) + (X(i - 1, j) << 12) + (X(i, j - 1) << 8) + (X(i + 1, j) << 4) + X(i, j + 1);
    return m;
}

int orc_features(int n, const int32_t *x, int32_t *f)
{
    int m = 0;
    if (n == 2) {                                         /* :17-20 */
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) f[m++] = (X(r, c) << 4) + X(r + 1, c);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 3; c++) f[m++] = (X(r, c) << 4) + X(r, c + 1);
        return m;
    }
    if (n == 3) {                                         /* :24-31 */
        for (int r = 0; r < 2; r++) for (int c = 0; c < 4; c++) f[m++] = (X(r, c) << 8) + (X(r + 1, c) << 4) + X(r + 2, c);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 2; c++) f[m++] = (X(r, c) << 8) + (X(r, c + 1) << 4) + X(r, c + 2);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) f[m++] = (X(r + 1, c) << 8) + (X(r + 1, c + 1) << 4) + X(r, c + 1);   /* ex_00 */
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) f[m++] = (X(r, c) << 8) + (X(r + 1, c) << 4) + X(r + 1, c + 1);       /* ex_01 */
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) f[m++] = (X(r, c) << 8) + (X(r, c + 1) << 4) + X(r + 1, c + 1);       /* ex_10 */
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) f[m++] = (X(r, c) << 8) + (X(r + 1, c) << 4) + X(r, c + 1);           /* ex_11 */
        return m;
    }
    m = feat_4(x, f);
    if (n == 4) return m;
    m += feat_middle(x, f + m);                           /* :48-54 */
    if (n == 5) return m;
    if (n == 6) {                                         /* :58-69 */
        int32_t y[16];
        for (int p = 0; p < 16; p++) y[p] = x[p] < 13 ? x[p] : 13;
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++)
                f[m++] = 537824 * Y(r, c) + 38416 * Y(r + 1, c) + 2744 * Y(r + 2, c) + 196 * Y(r, c + 1) +
                         14 * Y(r + 1, c + 1) + Y(r + 2, c + 1);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 2; c++)
                f[m++] = 537824 * Y(r, c) + 38416 * Y(r, c + 1) + 2744 * Y(r, c + 2) + 196 * Y(r + 1, c) +
                         14 * Y(r + 1, c + 1) + Y(r + 1, c + 2);
        return m;
    }
    return -1;
}
#undef X
#undef Y

int orc_features_batch(int n, const int32_t *rows, int64_t m, int32_t *out)
{
    int F = orc_num_feat(n);
    for (int64_t q = 0; q < m; q++) orc_features(n, rows + 16 * q, out + q * F);
    return F;
}

/* ------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",     */
/* SC'11; constants of Random123 philox.h) and the spawn-stream SPEC built on it (new).          */
/* ------------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* words for (game id, move number, purpose): key = seed, counter = (id_lo, id_hi, move_no, purpose) */
void orc_spawn_words(uint64_t seed, uint64_t id, uint32_t move_no, uint32_t purpose, uint32_t out[4])
{
    uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), move_no, purpose};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

/* game_logic.py:112-116 create_new_tile with Philox words in place of random.randrange(10) /
 * random.choice(empties): tile = 2 ("4") iff floor(10*u) == 0 (p = 0.1), position = the
 * floor(n_empty*u')-th empty cell in row-major order.  Returns flat position or -1 if full. */
int orc_spawn_apply(int32_t *row, uint32_t r_tile, uint32_t r_pos)
{
    int32_t pos[16];
    int m = orc_empty(row, pos);
    if (!m) return -1;
    int tile = mulhi32(r_tile, 10) == 0 ? 2 : 1;
    int p = pos[mulhi32(r_pos, (uint32_t)m)];
    row[p] = tile;
    return p | (tile << 8);
}

/* game_logic.py:62-64: empty board + two spawns (words 0,1 then 2,3 of move_no 0) */
void orc_spawn_initial(uint64_t seed, uint64_t id, int32_t *row)
{
    uint32_t w[4];
    memset(row, 0, 16 * sizeof(int32_t));
    orc_spawn_words(seed, id, 0, 0, w);
    orc_spawn_apply(row, w[0], w[1]);
    orc_spawn_apply(row, w[2], w[3]);
}

/* game_logic.py:118-121: the spawn after the move_no-th move (move_no = odometer after increment) */
int orc_spawn_move(uint64_t seed, uint64_t id, uint32_t move_no, int32_t *row)
{
    uint32_t w[4];
    orc_spawn_words(seed, id, move_no, 0, w);
    return orc_spawn_apply(row, w[0], w[1]);
}

/* sweep spawn (BASELINE config 5): ONE Philox block per board index i (purpose 1, counter z = 0); direction d takes
 * word d, low 16 bits -> tile fraction, high 16 bits -> cell fraction (both as the top half of a 32-bit word) */
int orc_spawn_sweep(uint64_t seed, uint64_t index, int d, int32_t *row)
{
    uint32_t w[4];
    orc_spawn_words(seed, index, 0, 1, w);
    return orc_spawn_apply(row, w[d & 3] << 16, w[d & 3] & 0xFFFF0000u);
}

/*
 * Config-5 sweep restated: for every packed board, 4 x (afterstate, score gain), changed mask,
 * overflow mask (an afterstate containing the unrepresentable 2^16 tile; its packed value is
 * stored with that nibble saturated to 15), and a Philox spawn on every changed afterstate
 * (unchanged / overflowing directions: spawned = afterstate).
 */
void orc_sweep(const uint64_t *boards, int64_t m, uint64_t seed, uint64_t first_index, int threads,
               uint64_t *after /*m*4*/, uint32_t *gain /*m*4*/, uint8_t *flags /*m: bit d changed, bit 4+d overflow*/,
               uint64_t *spawned /*m*4*/)
{
    table_init();
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t q = 0; q < m; q++) {
        int32_t row[16], nr[16];
        uint8_t fl = 0;
        orc_unpack(boards[q], row);
        for (int d = 0; d < 4; d++) {
            int64_t ns;
            int ch = orc_pre_move(row, 0, d, nr, &ns);
            int ovf = 0;
            for (int p = 0; p < 16; p++) if (nr[p] > 15) { nr[p] = 15; ovf = 1; }
            after[q * 4 + d] = orc_pack(nr);
            gain[q * 4 + d] = (uint32_t)ns;
            if (ch) fl |= (uint8_t)(1 << d);
            if (ovf) fl |= (uint8_t)(16 << d);
            if (spawned) {
                if (ch && !ovf) orc_spawn_sweep(seed, first_index + (uint64_t)q, d, nr);
                spawned[q * 4 + d] = orc_pack(nr);
            }
        }
        flags[q] = fl;
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------- */
/* precision-generic agent code: float64 = the reference's arithmetic, float32 = the device's     */
/* ------------------------------------------------------------------------------------------- */
#define REAL double
#define SUF f64
#include "oracle_impl.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF f32
#include "oracle_impl.h"
#undef REAL
#undef SUF
