// tests/host_shim.cu -- TEST ONLY.  Compiles the __host__ __device__ building blocks of
// 2048_b200/csrc/b2048_device.cuh for the HOST, so that the CPU-only test tier can check the packed
// board logic (LUT entries, 4-direction moves, predicates, D4 images, features, Philox spawns) against
// the oracle without a GPU.  It is not linked into libb2048.so and is never used by the product.
#include <cstdint>
#include "../2048_b200/csrc/b2048_device.cuh"

using namespace b2048;

struct LutHost {
    const uint32_t *p;
    __host__ __device__ uint32_t operator()(uint32_t line) const { return p[line]; }
};

// the shared-memory decode used by the sweep kernel, on host arrays
struct LutSplitHost {
    const uint16_t *row;
    const uint8_t *code;
    __host__ __device__ uint32_t operator()(uint32_t line) const
    {
        uint32_t r = row[line], c = code[line];
        uint32_t ovf = ((c & 15u) == 15u) | ((c >> 4) == 15u);
        uint32_t ch = (r != line) | ovf;
        return r | (c << 16) | (ch << 24) | (ovf << 25);
    }
};

template <int N>
static void features_n(const uint64_t *boards, int64_t m, int32_t *feat)
{
    constexpr int F = num_feat(N);
    for (int64_t q = 0; q < m; q++) {
        uint64_t b = boards[q], y = clamp13(b);
        int32_t *o = feat + q * F;
        for_each_feature<N>([&](auto I) {
            constexpr int k = decltype(I)::value;
            o[k] = int32_t(feat_index<N, k>(b, y));
        });
    }
}

// the shared-work index routine the evaluate kernels use for n >= 4 (must equal feat_index table by table)
template <int N>
static void features_fast_n(const uint64_t *boards, int64_t m, int32_t *feat)
{
    constexpr int F = num_feat(N);
    for (int64_t q = 0; q < m; q++) {
        uint32_t idx[F];
        feature_indices_fast<N>(boards[q], idx);
        for (int i = 0; i < F; i++) feat[q * F + i] = int32_t(idx[i]);
    }
}

extern "C" {

void hs_lut(uint32_t *lut)
{
    for (uint32_t l = 0; l < 65536; l++) lut[l] = lut_entry(l);
}

void hs_move4(const uint32_t *lut, int split, const uint64_t *boards, int64_t m, uint64_t *after, uint32_t *gain,
              uint8_t *flags, uint8_t *over)
{
    static uint16_t row[65536];
    static uint8_t code[65536];
    for (int l = 0; l < 65536; l++) { row[l] = uint16_t(lut[l]); code[l] = uint8_t(lut[l] >> 16); }
    LutHost L{lut};
    LutSplitHost L2{row, code};
    for (int64_t q = 0; q < m; q++) {
        uint32_t fl = 0;
        for (int d = 0; d < 4; d++) {
            uint32_t f, g;
            after[4 * q + d] = split ? move_dir(L2, boards[q], d, g, f) : move_dir(L, boards[q], d, g, f);
            gain[4 * q + d] = g;
            fl |= (f & 1u) << d;
            fl |= ((f >> 1) & 1u) << (4 + d);
        }
        flags[q] = uint8_t(fl);
        over[q] = game_over(boards[q]);
    }
}

void hs_stats(const uint64_t *boards, int64_t m, uint8_t *stats)
{
    for (int64_t q = 0; q < m; q++) {
        stats[4 * q + 0] = uint8_t(empty_count(boards[q]));
        stats[4 * q + 1] = uint8_t(adjacent_pair_count(boards[q]));
        stats[4 * q + 2] = uint8_t(game_over(boards[q]));
        stats[4 * q + 3] = uint8_t(max_tile(boards[q]));
    }
}

int hs_features(int n, const uint64_t *boards, int64_t m, int32_t *feat)
{
    switch (n) {
    case 2: features_n<2>(boards, m, feat); break;
    case 3: features_n<3>(boards, m, feat); break;
    case 4: features_n<4>(boards, m, feat); break;
    case 5: features_n<5>(boards, m, feat); break;
    case 6: features_n<6>(boards, m, feat); break;
    default: return -1;
    }
    return num_feat(n);
}

int hs_features_fast(int n, const uint64_t *boards, int64_t m, int32_t *feat)
{
    switch (n) {
    case 4: features_fast_n<4>(boards, m, feat); break;
    case 5: features_fast_n<5>(boards, m, feat); break;
    case 6: features_fast_n<6>(boards, m, feat); break;
    default: return -1;
    }
    return num_feat(n);
}

int64_t hs_table_offset(int n, int i) { return table_offset(n, i); }

void hs_d4(uint64_t b, uint64_t *out)
{
    for (int s = 0; s < 8; s++) out[s] = d4_image(b, s);
}

void hs_philox(const uint32_t *c, const uint32_t *k, uint32_t *out)
{
    Philox4 r = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

uint64_t hs_spawn_initial(uint64_t seed, uint64_t id) { return spawn_initial(seed, id); }

uint32_t hs_spawn_move(uint64_t seed, uint64_t id, uint32_t move_no, uint64_t *b)
{
    Philox4 r = spawn_words(seed, id, move_no, 0u);
    return spawn_apply(*b, r.x, r.y);
}

uint32_t hs_spawn_sweep(uint64_t seed, uint64_t index, int d, uint64_t *b)
{
    Philox4 r = spawn_words(seed, index, 0u, 1u);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    return spawn_apply(*b, sweep_tile_word(w[d & 3]), sweep_pos_word(w[d & 3]));
}

// the two spawn routines must agree on every board that holds a tile
int hs_spawn_nonempty_agrees(uint64_t b, uint32_t r_tile, uint32_t r_pos)
{
    uint64_t b1 = b, b2 = b;
    spawn_apply(b1, r_tile, r_pos);
    spawn_apply_nonempty(b2, r_tile, r_pos);
    return b1 == b2;
}
}
