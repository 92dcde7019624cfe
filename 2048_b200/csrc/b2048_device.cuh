// b2048_device.cuh -- device-side building blocks of the 2048 / n-tuple hot path (sm_100a).
//
// Everything here is a restatement, on packed 64-bit boards, of reference behaviour cited per
// function as file:line under /root/reference (game2048/game_logic.py, game2048/r_learning.py).
// Packed board: cell (r,c) = nibble at bit 4*(15-4r-c); row r = bits [63-16r : 48-16r].
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace b2048 {

// portable intrinsics: the same functions compile for the host in tests/host_shim.cu, which lets the
// CPU-only test tier check this header against the oracle without a GPU
__host__ __device__ __forceinline__ int popc64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

__host__ __device__ __forceinline__ uint32_t umulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return uint32_t((uint64_t(a) * b) >> 32);
#endif
}

// ------------------------------------------------------------------------------------------------
// LUT entry (b2048_lut_build): bits 0-15 new line, 16-19 / 20-23 merge exponents, 24 changed, 25 overflow
// ------------------------------------------------------------------------------------------------
constexpr uint32_t LUT_CHANGED = 1u << 24;
constexpr uint32_t LUT_OVERFLOW = 1u << 25;

__host__ __device__ __forceinline__ uint32_t lut_score(uint32_t e)
{
    // score += 1 << (x + 1) per merge of two x tiles (game_logic.py:33)
    // branch-free: 2 << 0 = 2 is the only value with bit 1 set, so masking bit 1 maps "no merge" to 0
    uint32_t a = (e >> 16) & 15u, b = (e >> 20) & 15u;
    return ((2u << a) & ~2u) + ((2u << b) & ~2u);
}

// create_table, game_logic.py:18-39, for one line (a,b,c,d) = nibbles 3..0 of `line`.
__host__ __device__ inline uint32_t lut_entry(uint32_t line)
{
    int v[4] = {int(line >> 12) & 15, int(line >> 8) & 15, int(line >> 4) & 15, int(line) & 15};
    int l1[4] = {0, 0, 0, 0}, n1 = 0;
    for (int i = 0; i < 4; i++)
        if (v[i]) l1[n1++] = v[i];                      // :29 drop zeros
    uint32_t code = 0, nm = 0, ovf = 0;
    for (int i = 0; i + 1 < n1; i++) {                  // :30-34 one left-to-right pass
        int x = l1[i];
        if (x && x == l1[i + 1]) {
            code |= uint32_t(x) << (16 + 4 * nm++);
            l1[i] = x + 1;
            l1[i + 1] = 0;
        }
    }
    uint32_t out = 0;
    int n2 = 0;
    for (int i = 0; i < n1; i++)                        // :35-36 drop zeros again, right-pad
        if (l1[i]) {
            int x = l1[i];
            if (x > 15) { x = 15; ovf = 1; }            // 2^16 escape: saturate + flag
            out |= uint32_t(x) << (12 - 4 * n2++);
        }
    uint32_t e = out | code;
    // :37 changed = line != line_2 (on the unsaturated line: an overflowing line always changed)
    if (out != line || ovf) e |= LUT_CHANGED;
    if (ovf) e |= LUT_OVERFLOW;
    return e;
}

// ------------------------------------------------------------------------------------------------
// board symmetries (np.transpose / np.rot90 views of r_learning.py:207-214, game_logic.py:138-141)
// ------------------------------------------------------------------------------------------------
// On the device the symmetries work on the two 32-bit halves (hi = rows 0-1, lo = rows 2-3) with byte permutes (PRMT)
// for everything that moves whole bytes: transpose 10 instructions instead of ~24, flip_h 8 instead of ~16, flip_v 2.
__host__ __device__ __forceinline__ uint64_t transpose(uint64_t x)
{
#ifdef __CUDA_ARCH__
    uint32_t lo = uint32_t(x), hi = uint32_t(x >> 32);
    // stage 1: transpose the nibbles inside every 2x2 block (same masks for both halves)
    lo = (lo & 0xF0F00F0Fu) | ((lo & 0x0000F0F0u) << 12) | ((lo & 0x0F0F0000u) >> 12);
    hi = (hi & 0xF0F00F0Fu) | ((hi & 0x0000F0F0u) << 12) | ((hi & 0x0F0F0000u) >> 12);
    // stage 2: swap the off-diagonal 2x2 blocks = bytes B1<->B4, B3<->B6 of [B0 B1 B2 B3 | B4 B5 B6 B7]
    const uint32_t nh = __byte_perm(hi, lo, 0x3715), nl = __byte_perm(hi, lo, 0x2604);
    return (uint64_t(nh) << 32) | nl;
#else
    uint64_t a1 = x & 0xF0F00F0FF0F00F0FULL;
    uint64_t a2 = x & 0x0000F0F00000F0F0ULL;
    uint64_t a3 = x & 0x0F0F00000F0F0000ULL;
    uint64_t a = a1 | (a2 << 12) | (a3 >> 12);
    uint64_t b1 = a & 0xFF00FF0000FF00FFULL;
    uint64_t b2 = a & 0x00FF00FF00000000ULL;
    uint64_t b3 = a & 0x00000000FF00FF00ULL;
    return b1 | (b2 >> 24) | (b3 << 24);
#endif
}

// mirror the columns (reverse every row)
__host__ __device__ __forceinline__ uint64_t flip_h(uint64_t x)
{
#ifdef __CUDA_ARCH__
    uint32_t lo = uint32_t(x), hi = uint32_t(x >> 32);
    lo = ((lo << 4) & 0xF0F0F0F0u) | ((lo >> 4) & 0x0F0F0F0Fu);       // swap the nibbles of every byte
    hi = ((hi << 4) & 0xF0F0F0F0u) | ((hi >> 4) & 0x0F0F0F0Fu);
    return (uint64_t(__byte_perm(hi, 0, 0x2301)) << 32) | __byte_perm(lo, 0, 0x2301);   // and the bytes of every row
#else
    return ((x & 0xF000F000F000F000ULL) >> 12) | ((x & 0x0F000F000F000F00ULL) >> 4) |
           ((x & 0x00F000F000F000F0ULL) << 4) | ((x & 0x000F000F000F000FULL) << 12);
#endif
}

// mirror the rows (reverse the row order)
__host__ __device__ __forceinline__ uint64_t flip_v(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return (uint64_t(__byte_perm(uint32_t(x), 0, 0x1032)) << 32) | __byte_perm(uint32_t(x >> 32), 0, 0x1032);
#else
    return (x >> 48) | ((x >> 16) & 0x00000000FFFF0000ULL) | ((x << 16) & 0x0000FFFF00000000ULL) | (x << 48);
#endif
}

__host__ __device__ __forceinline__ uint32_t reverse_line(uint32_t r)
{
    return ((r & 0xF000u) >> 12) | ((r & 0x0F00u) >> 4) | ((r & 0x00F0u) << 4) | ((r & 0x000Fu) << 12);
}

// The 8 D4 images QAgent.update visits (r_learning.py:207-214: r, r^T, Rr, (Rr)^T, ... R = rot90).
// As a set this is {id, flip_h, flip_v, flip_h.flip_v} x {id, transpose}; the order is irrelevant
// because all 8 receive the same dw.  s in 0..7.
__host__ __device__ __forceinline__ uint64_t d4_image(uint64_t b, int s)
{
    if (s & 1) b = flip_h(b);
    if (s & 2) b = flip_v(b);
    if (s & 4) b = transpose(b);
    return b;
}

// ------------------------------------------------------------------------------------------------
// board predicates (game_logic.py:96-110)
// ------------------------------------------------------------------------------------------------
// one bit per nibble (bit 4k set iff nibble k is zero)
__host__ __device__ __forceinline__ uint64_t zero_nibbles(uint64_t x)
{
    uint64_t t = x | (x >> 1);
    t |= t >> 2;
    return ~t & 0x1111111111111111ULL;
}

__host__ __device__ __forceinline__ int empty_count(uint64_t b) { return popc64(zero_nibbles(b)); }   // :101-103

// number of equal adjacent pairs (zeros included, like the reference's difference count), :105-107
__host__ __device__ __forceinline__ int adjacent_pair_count(uint64_t b)
{
    uint64_t h = b ^ (b << 4);                       // nibble k vs nibble k-1 (same row if k%4 != 0)
    uint64_t hz = zero_nibbles(h) & 0x1110111011101110ULL;
    uint64_t v = b ^ (b << 16);                      // row r vs row r+1
    uint64_t vz = zero_nibbles(v) & 0x1111111111110000ULL;
    return popc64(hz) + popc64(vz);
}

__host__ __device__ __forceinline__ bool game_over(uint64_t b)                                           // :109-110
{
    return zero_nibbles(b) == 0 && adjacent_pair_count(b) == 0;
}

__host__ __device__ __forceinline__ int max_tile(uint64_t b)
{
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int v = int((b >> (4 * k)) & 15);
        best = v > best ? v : best;
    }
    return best;
}

// ------------------------------------------------------------------------------------------------
// moves: Game._left / pre_move (game_logic.py:123-142).  LutT provides operator()(line) -> entry
// ------------------------------------------------------------------------------------------------
struct LutGlobal {
    const uint32_t *__restrict__ p;
    __device__ __forceinline__ uint32_t operator()(uint32_t line) const { return __ldg(p + line); }
};

// slide+merge every row to the left.  ok bits: 1 changed, 2 overflow
template <class LutT>
__host__ __device__ __forceinline__ uint64_t move_left(const LutT &lut, uint64_t b, uint32_t &gain, uint32_t &flags)
{
    uint64_t out = 0;
    uint32_t fl = 0, g = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint32_t e = lut(uint32_t(b >> (48 - 16 * r)) & 0xFFFFu);
        out |= uint64_t(e & 0xFFFFu) << (48 - 16 * r);
        g += lut_score(e);
        fl |= e >> 24;
    }
    gain = g;
    flags = fl & 3u;
    return out;
}

template <class LutT>
__host__ __device__ __forceinline__ uint64_t move_right(const LutT &lut, uint64_t b, uint32_t &gain, uint32_t &flags)
{
    uint64_t out = 0;
    uint32_t fl = 0, g = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint32_t e = lut(reverse_line(uint32_t(b >> (48 - 16 * r)) & 0xFFFFu));
        out |= uint64_t(reverse_line(e & 0xFFFFu)) << (48 - 16 * r);
        g += lut_score(e);
        fl |= e >> 24;
    }
    gain = g;
    flags = fl & 3u;
    return out;
}

// pre_move(row, score, d): d = 0 left, 1 up, 2 right, 3 down (game_logic.py:50,136-142):
// rot90(row, d) -> left -> rot90 back  ==  up: transpose/left/transpose, right: mirror/left/mirror,
// down: transpose/right/transpose.
template <class LutT>
__host__ __device__ __forceinline__ uint64_t move_dir(const LutT &lut, uint64_t b, int d, uint32_t &gain, uint32_t &flags)
{
    // One straight-line path for all four directions (the lanes of a warp hold all of them, so branches would diverge and
    // every divergent path is issued anyway): masks select the transposed / mirrored board, ONE slide-left, masks undo.
    const uint64_t mt = 0 - uint64_t(d & 1), mf = 0 - uint64_t((d >> 1) & 1);
    uint64_t x = b ^ ((b ^ transpose(b)) & mt);
    x ^= (x ^ flip_h(x)) & mf;
    uint64_t y = move_left(lut, x, gain, flags);
    y ^= (y ^ flip_h(y)) & mf;
    return y ^ ((y ^ transpose(y)) & mt);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) and the spawn rule (game_logic.py:112-121 with Philox words)
// ------------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = uint64_t(0xD2511F53u) * c0;
        uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        c1 = uint32_t(p1);
        c3 = uint32_t(p0);
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ Philox4 spawn_words(uint64_t seed, uint64_t id, uint32_t move_no, uint32_t purpose)
{
    return philox4x32_10(uint32_t(id), uint32_t(id >> 32), move_no, purpose, uint32_t(seed), uint32_t(seed >> 32));
}

__host__ __device__ __forceinline__ int popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// position (bit index, multiple of 4) of the k-th set bit of a one-bit-per-nibble mask, counting
// nibbles in row-major order, i.e. from the MOST significant nibble (0 <= k < popcount).
// Binary search on popcounts (4 levels) instead of a 16-step scan: the spawn is on the critical path of
// every game step and costs a quarter of the instructions of the board sweep.
__host__ __device__ __forceinline__ int kth_empty_shift(uint64_t zmask, int k)
{
    const uint32_t lo = uint32_t(zmask), hi = uint32_t(zmask >> 32);
    int c = popc32(hi);
    const bool in_hi = k < c;                        // row-major order = descending bit position
    uint32_t x = in_hi ? hi : lo;
    int base = in_hi ? 32 : 0;
    k = in_hi ? k : k - c;
    c = popc32(x >> 16);
    bool up = k < c;
    base += up ? 16 : 0;
    k = up ? k : k - c;
    x = up ? x >> 16 : x & 0xFFFFu;
    c = popc32(x >> 8);
    up = k < c;
    base += up ? 8 : 0;
    k = up ? k : k - c;
    x = up ? x >> 8 : x & 0xFFu;
    up = k < int(x >> 4);                            // one bit per nibble: the upper nibble holds 0 or 1
    return base + (up ? 4 : 0);
}

// create_new_tile: tile = 2 iff floor(10 u0) == 0, cell = floor(n_empty u1)-th empty cell, row-major.
// Returns (tile << 8) | flat cell, or 0xFFFF if the board is full.
__host__ __device__ __forceinline__ uint32_t spawn_apply(uint64_t &b, uint32_t r_tile, uint32_t r_pos)
{
    uint64_t z = zero_nibbles(b);
    int m = popc64(z);
    if (m == 0) return 0xFFFFu;
    uint32_t tile = umulhi32(r_tile, 10u) == 0 ? 2u : 1u;
    int k = int(umulhi32(r_pos, uint32_t(m)));
    int sh = kth_empty_shift(z, k);
    b |= uint64_t(tile) << sh;
    return (tile << 8) | uint32_t(15 - sh / 4);
}

// The same spawn for a board that holds at least one tile (every afterstate of a legal move does), without the binary
// search and without a variable 64-bit shift: z * 0x1111...1 puts in nibble j the number of empty cells among nibbles
// 0..j (at most 15, so no nibble overflows); the k-th empty cell in row-major order is the (m - k)-th from the least
// significant end, i.e. the one empty nibble whose prefix count equals m - k.  Returns nothing (callers that record
// the spawn use spawn_apply).
// zero-nibble mask at bit 3 of every nibble (0x8 per empty cell), without shifts: ((x & 7..7) + 7..7) | x has bit 3 set
// iff the nibble is non-zero, and the add never carries across nibbles.  (The integer-logic pipe, which executes shifts,
// is the busy one in the sweep; the add runs on the other pipe.)
__host__ __device__ __forceinline__ uint64_t zero_nibbles_hi(uint64_t x)
{
    // on the halves: no carry can cross the 32-bit boundary, so no add-with-carry chain is needed
    const uint32_t lo = uint32_t(x), hi = uint32_t(x >> 32);
    const uint32_t zl = ~(((lo & 0x77777777u) + 0x77777777u) | lo) & 0x88888888u;
    const uint32_t zh = ~(((hi & 0x77777777u) + 0x77777777u) | hi) & 0x88888888u;
    return (uint64_t(zh) << 32) | zl;
}

__host__ __device__ __forceinline__ void spawn_apply_nonempty(uint64_t &b, uint32_t r_tile, uint32_t r_pos)
{
    const uint64_t z8 = zero_nibbles_hi(b);
    const int m = popc64(z8);                                         // (a full board falls through: hit = 0)
    const uint32_t k = umulhi32(r_pos, uint32_t(m));
    const uint64_t prefix = (z8 >> 3) * 0x1111111111111111ULL;
    const uint64_t want = uint64_t(uint32_t(m) - k) * 0x1111111111111111ULL;
    const uint64_t hit8 = zero_nibbles_hi(prefix ^ want) & z8;        // one bit: bit 4j + 3 of the chosen nibble j
    b |= (hit8 >> 3) * uint64_t(umulhi32(r_tile, 10u) == 0 ? 2u : 1u);   // (a multiply: the shifter pipe is the busy one)
}

// BASELINE config 5 (b2048_sweep) draws ONE Philox block per board, counter (index_lo, index_hi, 0, purpose 1), and
// gives direction d the word d: its low half decides the tile, its high half the cell (16-bit fractions in the top
// bits of the two arguments of spawn_apply: P("4") = 6554 / 65536).
__host__ __device__ __forceinline__ uint32_t sweep_tile_word(uint32_t w) { return w << 16; }
__host__ __device__ __forceinline__ uint32_t sweep_pos_word(uint32_t w) { return w & 0xFFFF0000u; }

// Game.__init__ (game_logic.py:61-66): two spawns on the empty board
__host__ __device__ __forceinline__ uint64_t spawn_initial(uint64_t seed, uint64_t id)
{
    Philox4 w = spawn_words(seed, id, 0u, 0u);
    uint64_t b = 0;
    spawn_apply(b, w.x, w.y);
    spawn_apply(b, w.z, w.w);
    return b;
}

// ------------------------------------------------------------------------------------------------
// n-tuple features (r_learning.py:17-69), compile-time cell lists
// ------------------------------------------------------------------------------------------------
struct FeatSpec {
    int ncell;      // cells per tuple
    int cell[6];    // flat positions 4r+c, most significant first
    int base;       // 16 (shift/or) or 14 (n=6 extra tuples on y = min(x, 13))
};

__host__ __device__ constexpr int num_feat(int n) { return n == 2 ? 24 : n == 3 ? 52 : n == 4 ? 17 : n == 5 ? 21 : n == 6 ? 33 : -1; }

__host__ __device__ constexpr int P(int r, int c) { return 4 * r + c; }

__host__ __device__ constexpr FeatSpec feat_spec(int n, int i)
{
    if (n == 2) {                                    // f_2 :17-20
        if (i < 12) return FeatSpec{2, {P(i / 4, i % 4), P(i / 4 + 1, i % 4), 0, 0, 0, 0}, 16};
        int k = i - 12;
        return FeatSpec{2, {P(k / 3, k % 3), P(k / 3, k % 3 + 1), 0, 0, 0, 0}, 16};
    }
    if (n == 3) {                                    // f_3 :24-31
        if (i < 8) return FeatSpec{3, {P(i / 4, i % 4), P(i / 4 + 1, i % 4), P(i / 4 + 2, i % 4), 0, 0, 0}, 16};
        if (i < 16) {
            int k = i - 8;
            return FeatSpec{3, {P(k / 2, k % 2), P(k / 2, k % 2 + 1), P(k / 2, k % 2 + 2), 0, 0, 0}, 16};
        }
        int k = i - 16, g = k / 9, r = (k % 9) / 3, c = k % 3;
        if (g == 0) return FeatSpec{3, {P(r + 1, c), P(r + 1, c + 1), P(r, c + 1), 0, 0, 0}, 16};   // ex_00
        if (g == 1) return FeatSpec{3, {P(r, c), P(r + 1, c), P(r + 1, c + 1), 0, 0, 0}, 16};       // ex_01
        if (g == 2) return FeatSpec{3, {P(r, c), P(r, c + 1), P(r + 1, c + 1), 0, 0, 0}, 16};       // ex_10
        return FeatSpec{3, {P(r, c), P(r + 1, c), P(r, c + 1), 0, 0, 0}, 16};                       // ex_11
    }
    // f_4 :40-44 is the prefix of f_5 :48-54 and f_6 :58-69
    if (i < 4) return FeatSpec{4, {P(0, i), P(1, i), P(2, i), P(3, i), 0, 0}, 16};                  // columns
    if (i < 8) return FeatSpec{4, {P(i - 4, 0), P(i - 4, 1), P(i - 4, 2), P(i - 4, 3), 0, 0}, 16};  // rows
    if (i < 17) {
        int r = (i - 8) / 3, c = (i - 8) % 3;                                                       // squares
        return FeatSpec{4, {P(r, c), P(r + 1, c), P(r, c + 1), P(r + 1, c + 1), 0, 0}, 16};
    }
    if (i < 21) {
        int a = 1 + (i - 17) / 2, b = 1 + (i - 17) % 2;                                             // crosses
        return FeatSpec{5, {P(a, b), P(a - 1, b), P(a, b - 1), P(a + 1, b), P(a, b + 1), 0}, 16};
    }
    if (i < 27) {
        int r = (i - 21) / 3, c = (i - 21) % 3;                                                     // 3x2 rects
        return FeatSpec{6, {P(r, c), P(r + 1, c), P(r + 2, c), P(r, c + 1), P(r + 1, c + 1), P(r + 2, c + 1)}, 14};
    }
    int r = (i - 27) / 2, c = (i - 27) % 2;                                                         // 2x3 rects
    return FeatSpec{6, {P(r, c), P(r, c + 1), P(r, c + 2), P(r + 1, c), P(r + 1, c + 1), P(r + 1, c + 2)}, 14};
}

__host__ __device__ constexpr int64_t table_size(int n, int i)
{
    return n == 2 ? 256 : n == 3 ? 4096 : i < 17 ? 65536 : i < 21 ? 1048576 : 7529536;   // r_learning.py:88,136-149
}

__host__ __device__ constexpr int64_t table_offset(int n, int i)
{
    int64_t o = 0;
    for (int k = 0; k < i; k++) o += table_size(n, k);
    return o;
}

// y = min(x, 13) on all 16 nibbles at once (r_learning.py:64)
__host__ __device__ __forceinline__ uint64_t clamp13(uint64_t x)
{
    uint64_t t = (x >> 3) & (x >> 2) & (x >> 1) & 0x1111111111111111ULL;   // nibble >= 14
    return (x & ~(t << 1)) | t;                                               // 111x -> 1101
}

template <int N, int I>
__host__ __device__ __forceinline__ uint32_t feat_index(uint64_t b, uint64_t y)
{
    constexpr FeatSpec s = feat_spec(N, I);
    uint32_t idx = 0;
    if constexpr (s.base == 16) {
#pragma unroll
        for (int k = 0; k < s.ncell; k++)
            idx |= (uint32_t(b >> (4 * (15 - s.cell[k]))) & 15u) << (4 * (s.ncell - 1 - k));
    } else {
#pragma unroll
        for (int k = 0; k < s.ncell; k++) idx = idx * 14u + (uint32_t(y >> (4 * (15 - s.cell[k]))) & 15u);
    }
    return idx;
}

// calls f(std::integral_constant<int, I>) for I = 0..F-1 (compile-time unrolled)
template <int N, int I = 0, class Fn>
__host__ __device__ __forceinline__ void for_each_feature(Fn &&f)
{
    if constexpr (I < num_feat(N)) {
        f(std::integral_constant<int, I>{});
        for_each_feature<N, I + 1>(f);
    }
}

// ------------------------------------------------------------------------------------------------
// All F table indices of one board at once for n = 4, 5, 6 (f_4 / f_5 / f_6, r_learning.py:40-69), sharing work
// between the tuples instead of extracting every cell of every tuple (feat_index): the same indices with a third
// of the instructions, which is half of what greedy play executes.
//   rows / columns   16-bit fields of the board and of its transpose
//   2x2 squares      (r,c),(r+1,c) is a byte of column c in the transpose: index = pair(c,r) << 8 | pair(c+1,r)
//   crosses          3 cells of column j and 3 cells of row i around the centre
//   3x2 / 2x3 rects  base-14 value of 3 consecutive cells of a column / row (16 of them), index = a * 14^3 + b
// ------------------------------------------------------------------------------------------------
template <int N>
__host__ __device__ __forceinline__ void feature_indices_fast(uint64_t b, uint32_t (&idx)[num_feat(N)])
{
    static_assert(N >= 4 && N <= 6, "n = 4, 5, 6");
    const uint64_t bt = transpose(b);
    const uint32_t row[4] = {uint32_t(b >> 48), uint32_t(b >> 32) & 0xFFFFu, uint32_t(b >> 16) & 0xFFFFu,
                             uint32_t(b) & 0xFFFFu};
    const uint32_t col[4] = {uint32_t(bt >> 48), uint32_t(bt >> 32) & 0xFFFFu, uint32_t(bt >> 16) & 0xFFFFu,
                             uint32_t(bt) & 0xFFFFu};
#pragma unroll
    for (int c = 0; c < 4; c++) idx[c] = col[c];                       // x[0,c] x[1,c] x[2,c] x[3,c]
#pragma unroll
    for (int r = 0; r < 4; r++) idx[4 + r] = row[r];                   // x[r,0] x[r,1] x[r,2] x[r,3]
    uint32_t pair[4][3];                                               // x[r,c] << 4 | x[r+1,c]
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) pair[c][r] = (col[c] >> (8 - 4 * r)) & 0xFFu;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) idx[8 + 3 * r + c] = (pair[c][r] << 8) | pair[c + 1][r];
    if constexpr (N >= 5) {
#pragma unroll
        for (int q = 0; q < 4; q++) {                                  // centre (i,j), i,j in {1,2}, row-major
            const int i = 1 + q / 2, j = 1 + q % 2;
            const uint32_t v3 = (col[j] >> (4 * (2 - i))) & 0xFFFu;    // x[i-1,j] x[i,j] x[i+1,j]
            const uint32_t h3 = (row[i] >> (4 * (2 - j))) & 0xFFFu;    // x[i,j-1] x[i,j] x[i,j+1]
            idx[17 + q] = ((v3 & 0x0F0u) << 12) | ((v3 & 0xF00u) << 4) | (h3 & 0xF00u) | ((v3 & 0x00Fu) << 4) | (h3 & 0x00Fu);
        }
    }
    if constexpr (N == 6) {
        const uint64_t y = clamp13(b), yt = clamp13(bt);               // min(x, 13) commutes with the transpose
        uint32_t v3[4][2], h3[4][2];                                   // base-14 value of 3 consecutive cells
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t yc = uint32_t(yt >> (48 - 16 * k)) & 0xFFFFu, yr = uint32_t(y >> (48 - 16 * k)) & 0xFFFFu;
#pragma unroll
            for (int o = 0; o < 2; o++) {
                const uint32_t fc = (yc >> (4 - 4 * o)) & 0xFFFu, fr = (yr >> (4 - 4 * o)) & 0xFFFu;
                v3[k][o] = (fc >> 8) * 196u + ((fc >> 4) & 15u) * 14u + (fc & 15u);
                h3[k][o] = (fr >> 8) * 196u + ((fr >> 4) & 15u) * 14u + (fr & 15u);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)                                    // 3x2: column c rows r..r+2, then column c+1
#pragma unroll
            for (int c = 0; c < 3; c++) idx[21 + 3 * r + c] = v3[c][r] * 2744u + v3[c + 1][r];
#pragma unroll
        for (int r = 0; r < 3; r++)                                    // 2x3: row r columns c..c+2, then row r+1
#pragma unroll
            for (int c = 0; c < 2; c++) idx[27 + 2 * r + c] = h3[r][c] * 2744u + h3[r + 1][c];
    }
}

// all F table indices of a board (any n), table order
template <int N>
__host__ __device__ __forceinline__ void feature_indices(uint64_t b, uint32_t (&idx)[num_feat(N)])
{
    if constexpr (N >= 4) {
        feature_indices_fast<N>(b, idx);
    } else {
        for_each_feature<N>([&](auto I) {
            constexpr int i = decltype(I)::value;
            idx[i] = feat_index<N, i>(b, 0);
        });
    }
}

// sum of the F weights at precomputed indices: every gather in flight first, then the sequential float32 sum in
// table order, from 0 (QAgent.evaluate, r_learning.py:202-203)
template <int N, bool COHERENT = false>
__device__ __forceinline__ float gather_sum(const float *__restrict__ w, const uint32_t (&idx)[num_feat(N)])
{
    constexpr int F = num_feat(N);
    float v[F];
    for_each_feature<N>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const float *p = w + table_offset(N, i) + idx[i];
        v[i] = COHERENT ? __ldcg(p) : __ldg(p);
    });
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < F; i++) acc = __fadd_rn(acc, v[i]);
    return acc;
}

// QAgent.evaluate (r_learning.py:202-203): sequential float32 sum in table order, from 0.
// COHERENT: read through L2 (ld.global.cg) instead of the non-coherent L1 path -- required inside the
// persistent training kernel, where other SMs update the tables between lock-steps of the same launch.
template <int N, bool COHERENT = false>
__device__ __forceinline__ float evaluate(const float *__restrict__ w, uint64_t b)
{
    uint32_t idx[num_feat(N)];
    feature_indices<N>(b, idx);
    return gather_sum<N, COHERENT>(w, idx);
}

}   // namespace b2048
