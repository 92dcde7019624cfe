// b2048_kernels.cu -- CUDA kernels (sm_100a) + the C-ABI of include/b2048.h.
//
// Kernel groups (BASELINE.json north_star): (1) packed boards, (2) LUT moves, (3) Philox / replay
// spawns, (4) n-tuple gather (evaluate) and TD scatter (atomic | deterministic, sum | per-key mean),
// plus the fused game loops built from the same device functions (b2048_device.cuh).
// Nothing here is a dense contraction: no tensor cores by design (HBM/L2-latency- and issue-bound).
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <type_traits>

#include <cuda_runtime.h>

#include "b2048_device.cuh"
#include "../../include/b2048.h"

using namespace b2048;

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}

inline cudaStream_t S(b2048_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per device id; immutable)
int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------------
// (1) pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const int32_t *__restrict__ rows, uint64_t *__restrict__ boards, int64_t m)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    const int4 *r = reinterpret_cast<const int4 *>(rows + 16 * i);
    uint64_t b = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int4 v = __ldg(r + q);
        uint64_t line = (uint64_t(v.x & 15) << 12) | (uint64_t(v.y & 15) << 8) | (uint64_t(v.z & 15) << 4) | uint64_t(v.w & 15);
        b |= line << (48 - 16 * q);
    }
    boards[i] = b;
}

__global__ void unpack_kernel(const uint64_t *__restrict__ boards, int32_t *__restrict__ rows, int64_t m)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    uint64_t b = __ldg(boards + i);
    int4 *r = reinterpret_cast<int4 *>(rows + 16 * i);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t line = uint32_t(b >> (48 - 16 * q)) & 0xFFFFu;
        r[q] = make_int4(int(line >> 12) & 15, int(line >> 8) & 15, int(line >> 4) & 15, int(line) & 15);
    }
}

// ------------------------------------------------------------------------------------------------
// (2) moves
// ------------------------------------------------------------------------------------------------
__global__ void lut_build_kernel(uint32_t *__restrict__ lut)
{
    uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line < B2048_LUT_ENTRIES) lut[line] = lut_entry(line);
}

__global__ void move4_kernel(const uint32_t *__restrict__ lut, const uint64_t *__restrict__ boards, int64_t m,
                             uint64_t *__restrict__ after, uint32_t *__restrict__ gain, uint8_t *__restrict__ flags,
                             uint8_t *__restrict__ over)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    LutGlobal L{lut};
    uint64_t b = __ldg(boards + i);
    uint64_t a[4];
    uint32_t g[4], fl = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        uint32_t f;
        a[d] = move_dir(L, b, d, g[d], f);
        fl |= (f & 1u) << d;
        fl |= ((f >> 1) & 1u) << (4 + d);
    }
    ulonglong2 *ap = reinterpret_cast<ulonglong2 *>(after + 4 * i);
    ap[0] = make_ulonglong2(a[0], a[1]);
    ap[1] = make_ulonglong2(a[2], a[3]);
    *reinterpret_cast<uint4 *>(gain + 4 * i) = make_uint4(g[0], g[1], g[2], g[3]);
    flags[i] = uint8_t(fl);
    if (over) over[i] = game_over(b) ? 1 : 0;
}

__global__ void board_stats_kernel(const uint64_t *__restrict__ boards, int64_t m, uint8_t *__restrict__ stats,
                                   uint16_t *__restrict__ empty_mask)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    uint64_t b = __ldg(boards + i);
    uint64_t z = zero_nibbles(b);
    int ne = __popcll(z), np = adjacent_pair_count(b);
    if (stats) {
        uchar4 s = make_uchar4((unsigned char)ne, (unsigned char)np, (unsigned char)(ne == 0 && np == 0),
                               (unsigned char)max_tile(b));
        reinterpret_cast<uchar4 *>(stats)[i] = s;
    }
    if (empty_mask) {
        uint32_t mask = 0;
#pragma unroll
        for (int p = 0; p < 16; p++) mask |= uint32_t((z >> (4 * (15 - p))) & 1u) << p;
        empty_mask[i] = uint16_t(mask);
    }
}

// ------------------------------------------------------------------------------------------------
// (3) spawns
// ------------------------------------------------------------------------------------------------
__global__ void spawn_philox_kernel(uint64_t *__restrict__ boards, int64_t m, uint64_t seed,
                                    const uint64_t *__restrict__ game_id, const uint32_t *__restrict__ move_no,
                                    uint16_t *__restrict__ spawn)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    uint64_t b = boards[i];
    Philox4 w = spawn_words(seed, __ldg(game_id + i), __ldg(move_no + i), 0u);
    uint32_t res = spawn_apply(b, w.x, w.y);
    boards[i] = b;
    if (spawn) spawn[i] = uint16_t(res);
}

__global__ void spawn_initial_kernel(uint64_t *__restrict__ boards, int64_t m, uint64_t seed, uint64_t first_id,
                                     uint64_t id_step)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    boards[i] = spawn_initial(seed, first_id + uint64_t(i) * id_step);
}

__global__ void spawn_replay_kernel(uint64_t *__restrict__ boards, int64_t m, const uint8_t *__restrict__ tile,
                                    const uint8_t *__restrict__ pos)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    uint32_t t = tile[i];
    if (!t) return;
    int sh = 4 * (15 - int(pos[i] & 15));
    uint64_t b = boards[i];
    boards[i] = (b & ~(0xFULL << sh)) | (uint64_t(t & 15u) << sh);     // row[pos] = tile (game_logic.py:260)
}

// config-5 sweep: persistent CTAs, row LUT (u16) + merge-code LUT (u8) staged in 192 KB of shared memory
struct LutShared {
    const uint16_t *row;
    const uint8_t *code;
    __device__ __forceinline__ uint32_t operator()(uint32_t line) const
    {
        uint32_t r = row[line], c = code[line];
        uint32_t ovf = ((c & 15u) == 15u) | ((c >> 4) == 15u);
        uint32_t ch = (r != line) | ovf;
        return r | (c << 16) | (ch << 24) | (ovf << 25);
    }
};

constexpr int SWEEP_THREADS = 1024;
constexpr size_t SWEEP_SMEM = 65536 * 2 + 65536;

__global__ void __launch_bounds__(SWEEP_THREADS, 1)
sweep_kernel(const uint32_t *__restrict__ lut, const uint64_t *__restrict__ boards, int64_t m, uint64_t seed,
             uint64_t first_index, uint64_t *__restrict__ after, uint32_t *__restrict__ gain,
             uint8_t *__restrict__ flags, uint64_t *__restrict__ spawned)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t *srow = reinterpret_cast<uint16_t *>(smem);
    uint8_t *scode = smem + 65536 * 2;
    // stage: 4 entries per thread per iteration (16 B global load -> 8 B + 4 B shared stores)
    for (int q = threadIdx.x; q < 65536 / 4; q += SWEEP_THREADS) {
        uint4 e = __ldg(reinterpret_cast<const uint4 *>(lut) + q);
        uint2 r = make_uint2((e.x & 0xFFFFu) | (e.y << 16), (e.z & 0xFFFFu) | (e.w << 16));
        reinterpret_cast<uint2 *>(srow)[q] = r;
        uint32_t c = ((e.x >> 16) & 0xFFu) | (((e.y >> 16) & 0xFFu) << 8) | (((e.z >> 16) & 0xFFu) << 16) |
                     (((e.w >> 16) & 0xFFu) << 24);
        reinterpret_cast<uint32_t *>(scode)[q] = c;
    }
    __syncthreads();
    LutShared L{srow, scode};
    const int64_t stride = int64_t(gridDim.x) * SWEEP_THREADS;
    for (int64_t i = blockIdx.x * int64_t(SWEEP_THREADS) + threadIdx.x; i < m; i += stride) {
        uint64_t b = __ldg(boards + i);
        uint64_t a[4];
        uint32_t g[4], fl = 0, ok = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            uint32_t f;
            a[d] = move_dir(L, b, d, g[d], f);
            fl |= (f & 1u) << d;
            fl |= ((f >> 1) & 1u) << (4 + d);
            ok |= uint32_t((f & 3u) == 1u) << d;
        }
        ulonglong2 *ap = reinterpret_cast<ulonglong2 *>(after + 4 * i);
        ap[0] = make_ulonglong2(a[0], a[1]);
        ap[1] = make_ulonglong2(a[2], a[3]);
        *reinterpret_cast<uint4 *>(gain + 4 * i) = make_uint4(g[0], g[1], g[2], g[3]);
        flags[i] = uint8_t(fl);
        if (spawned) {
            uint64_t idx = first_index + uint64_t(i);
            if (ok & 3u) {
                Philox4 w = spawn_words(seed, idx, 0u, 1u);
                if (ok & 1u) spawn_apply(a[0], w.x, w.y);
                if (ok & 2u) spawn_apply(a[1], w.z, w.w);
            }
            if (ok & 12u) {
                Philox4 w = spawn_words(seed, idx, 1u, 1u);
                if (ok & 4u) spawn_apply(a[2], w.x, w.y);
                if (ok & 8u) spawn_apply(a[3], w.z, w.w);
            }
            ulonglong2 *sp = reinterpret_cast<ulonglong2 *>(spawned + 4 * i);
            sp[0] = make_ulonglong2(a[0], a[1]);
            sp[1] = make_ulonglong2(a[2], a[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// (4) features / evaluate
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void features_kernel(const uint64_t *__restrict__ boards, int64_t m, int32_t *__restrict__ feat)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    constexpr int F = num_feat(N);
    uint64_t b = __ldg(boards + i);
    uint64_t y = (N == 6) ? clamp13(b) : 0;
    int32_t *o = feat + i * F;
    for_each_feature<N>([&](auto I) {
        constexpr int k = decltype(I)::value;
        o[k] = int32_t(feat_index<N, k>(b, y));
    });
}

template <int N>
__global__ void evaluate_kernel(const float *__restrict__ w, const uint64_t *__restrict__ boards, int64_t m,
                                float *__restrict__ value)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    value[i] = evaluate<N>(w, __ldg(boards + i));
}

// ------------------------------------------------------------------------------------------------
// TD update: 8 D4 images x F tables per (board, dw)
// ------------------------------------------------------------------------------------------------
// red.global.add.f32 (no return value -> fire and forget)
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

// ---- accumulate pass -------------------------------------------------------------------------
// Thread per (entry j, image s); the 8 images of an entry are 8 adjacent lanes, a warp holds 4 entries.
// For every table i the warp first merges lanes that hit the same key (__match_any_sync): hot keys
// (empty rows/squares early in a game) would otherwise serialise thousands of same-address atomics in L2.
//   DIRECT          atomic + sum rule: the merged contribution goes straight into w (and delta)
//   otherwise       acc[k] += contribution (float RED, or exact int64 fixed point when EXACT),
//                   cnt[k] += number of distinct entries (MEAN) or 1; the lane that sees cnt go 0 -> >0
//                   appends k to the touched list, so the apply pass needs no atomics at all.
constexpr double FIX_SCALE = 4294967296.0;      // 2^32: exact-mode contributions are llrint(dw * 2^32)

struct UpdCtrl {
    uint32_t count;     // touched keys of the running update
    uint32_t ticket;    // apply pass: blocks done (the last one resets both)
    uint32_t pad[2];
};

// The accumulator is replicated ACC_REPLICAS(nw) times (replica = CTA index mod R): contributions are shared
// so broadly (only 16-30 % of the keys of a lock-step are distinct) that same-address atomics, which L2
// serialises at ~1.4 ns each, would otherwise dominate.  The apply pass sums the replicas of a touched key.
__host__ __device__ constexpr int acc_replicas(int64_t nw) { return nw <= (int64_t(1) << 23) ? 8 : 2; }

__device__ __forceinline__ long long quantize(float d) { return __double2ll_rn(double(d) * FIX_SCALE); }

__device__ __forceinline__ void accumulate_key(bool exact, void *__restrict__ acc, uint32_t *__restrict__ cnt,
                                               uint32_t *__restrict__ touched, UpdCtrl *__restrict__ ctrl, int64_t k,
                                               float fsum, long long qsum, uint32_t nfirst, int64_t replica_off)
{
    if (exact)
        atomicAdd(reinterpret_cast<unsigned long long *>(acc) + replica_off + k, (unsigned long long)qsum);
    else
        atomicAdd(reinterpret_cast<float *>(acc) + replica_off + k, fsum);
    uint32_t old = atomicAdd(cnt + k, nfirst);
    if (old == 0) touched[atomicAdd(&ctrl->count, 1u)] = uint32_t(k);
}

// The contributions of one (entry, image) lane for the tables i with i % CH == chunk (CH = 1: all tables).
// Must be called by all 32 lanes of a warp with a warp-uniform `chunk`; lanes 8k..8k+7 hold the 8 images of
// one entry.
template <int N, bool EXACT, bool MEAN, bool DIRECT, int CH>
__device__ __forceinline__ void accum_features(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc,
                                               uint32_t *__restrict__ cnt, uint32_t *__restrict__ touched,
                                               uint32_t *__restrict__ count, uint64_t b, float d, bool live, int s,
                                               int lane, int chunk, int64_t replica_off)
{
    const uint64_t y = (N == 6) ? clamp13(b) : 0;
    const long long q = (EXACT && live) ? quantize(d) : 0;
    for_each_feature<N>([&](auto I) {
        constexpr int i = decltype(I)::value;
        if (CH > 1 && (i % CH) != chunk) return;
        const uint32_t f = feat_index<N, i>(b, y);
        const uint32_t k = uint32_t(table_offset(N, i)) + f;
        uint32_t first = 1;
        if (MEAN) {                                   // is a lower image of the same entry on the same key?
#pragma unroll
            for (int o = 1; o < 8; o++) {
                uint32_t fo = __shfl_xor_sync(FULL, f, o);
                if (fo == f && (s ^ o) < s) first = 0;
            }
        }
        const uint32_t peers = __match_any_sync(FULL, live ? k : 0xFFFFFFFFu - uint32_t(lane));
        float fsum = d;
        long long qsum = q;
        uint32_t nf = first;
        if (peers & (peers - 1)) {                    // more than one lane on this key: merge (group-uniform branch)
            fsum = 0.0f; qsum = 0; nf = 0;
            for (uint32_t mm = peers; mm; mm &= mm - 1) {
                const int src = __ffs(mm) - 1;
                if (EXACT) {
                    qsum += __shfl_sync(peers, q, src);
                } else {
                    fsum += __shfl_sync(peers, d, src);
                }
                if (MEAN) nf += __shfl_sync(peers, first, src);
            }
            if (!MEAN) nf = 1;
        }
        if (live && lane == __ffs(peers) - 1) {
            if (DIRECT) {
                atomicAdd(w + k, fsum);
                if (delta) atomicAdd(delta + k, fsum);
            } else {
                if (EXACT)
                    atomicAdd(reinterpret_cast<unsigned long long *>(acc) + replica_off + k, (unsigned long long)qsum);
                else
                    atomicAdd(reinterpret_cast<float *>(acc) + replica_off + k, fsum);
                if (atomicAdd(cnt + k, nf) == 0) touched[atomicAdd(count, 1u)] = k;
            }
        }
    });
}

template <int N, bool EXACT, bool MEAN, bool DIRECT>
__global__ void __launch_bounds__(128)
td_accum_kernel(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc, uint32_t *__restrict__ cnt,
                uint32_t *__restrict__ touched, UpdCtrl *__restrict__ ctrl, const uint64_t *__restrict__ boards,
                const float *__restrict__ dw, int64_t m)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t j = t >> 3;
    const int s = int(t & 7), lane = threadIdx.x & 31;
    const bool on = j < m;
    const float d = on ? __ldg(dw + j) : NAN;
    const bool live = on && (EXACT ? isfinite(d) : !isnan(d));
    if (!__any_sync(FULL, live)) return;
    const uint64_t b = d4_image(on ? __ldg(boards + j) : 0, s);
    constexpr int64_t NW = table_offset(N, num_feat(N));
    const int64_t replica_off = int64_t(blockIdx.x % acc_replicas(NW)) * NW;
    accum_features<N, EXACT, MEAN, DIRECT, 1>(w, delta, acc, cnt, touched, ctrl ? &ctrl->count : nullptr, b, d, live, s,
                                              lane, 0, replica_off);
}

// ---- apply pass: one thread per touched key, plain loads/stores --------------------------------
template <bool EXACT, bool MEAN, bool COHERENT>
__device__ __forceinline__ void apply_key(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc,
                                          uint32_t *__restrict__ cnt, uint32_t k, int64_t nw, int R)
{
    const uint32_t c = COHERENT ? __ldcg(cnt + k) : cnt[k];
    float u;
    if (EXACT) {
        long long *a = reinterpret_cast<long long *>(acc) + k;
        long long qs = 0;
        for (int r = 0; r < R; r++) {
            long long v = COHERENT ? __ldcg(a + r * nw) : a[r * nw];
            if (v) { qs += v; a[r * nw] = 0; }
        }
        double x = double(qs) / FIX_SCALE;
        if (MEAN) x = x / double(c);
        u = __double2float_rn(x);
    } else {
        float *a = reinterpret_cast<float *>(acc) + k;
        float fs = 0.0f;
        for (int r = 0; r < R; r++) {
            float v = COHERENT ? __ldcg(a + r * nw) : a[r * nw];
            if (v != 0.0f) { fs += v; a[r * nw] = 0.0f; }
        }
        u = MEAN ? __fdiv_rn(fs, float(c)) : fs;
    }
    cnt[k] = 0;
    w[k] = __fadd_rn(COHERENT ? __ldcg(w + k) : w[k], u);
    if (delta) delta[k] = __fadd_rn(COHERENT ? __ldcg(delta + k) : delta[k], u);
}

template <bool EXACT, bool MEAN>
__global__ void __launch_bounds__(256)
td_apply_kernel(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc, uint32_t *__restrict__ cnt,
                const uint32_t *__restrict__ touched, UpdCtrl *__restrict__ ctrl, int64_t nw)
{
    const int R = acc_replicas(nw);
    const uint32_t count = *reinterpret_cast<volatile uint32_t *>(&ctrl->count);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x)
        apply_key<EXACT, MEAN, false>(w, delta, acc, cnt, touched[t], nw, R);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctrl->ticket, 1u) == gridDim.x - 1) {       // last block: ready for the next update
            ctrl->count = 0;
            ctrl->ticket = 0;
        }
    }
}

// deterministic, SORTED variant: key generation -> stable LSD radix sort of (key, entry | first << 31) ->
// chunked per-key partial sums (exact int64) -> the same acc/cnt/touched/apply tail as the direct variant
template <int N>
__global__ void td_keys_kernel(const uint64_t *__restrict__ boards, const float *__restrict__ dw, int64_t m,
                               uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    int64_t j = t >> 3;
    constexpr int F = num_feat(N);
    const int s = int(t & 7);
    const bool on = j < m;
    float d = on ? __ldg(dw + j) : NAN;
    const bool live = on && isfinite(d);
    uint64_t b = d4_image(on ? __ldg(boards + j) : 0, s);
    uint64_t y = (N == 6) ? clamp13(b) : 0;
    uint32_t *ko = keys + (j * 8 + s) * F;
    uint32_t *vo = vals + (j * 8 + s) * F;
    for_each_feature<N>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const uint32_t f = feat_index<N, i>(b, y);
        uint32_t first = 1;
#pragma unroll
        for (int o = 1; o < 8; o++) {
            uint32_t fo = __shfl_xor_sync(FULL, f, o);
            if (fo == f && (s ^ o) < s) first = 0;
        }
        if (on) {
            ko[i] = live ? uint32_t(table_offset(N, i)) + f : 0xFFFFFFFFu;
            vo[i] = uint32_t(j) | (first << 31);
        }
    });
}

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_THREADS / 32;

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t M, int shift, uint32_t *__restrict__ hist, int nblocks)
{
    constexpr int RADIX = 1 << BITS;
    __shared__ uint32_t h[RADIX];
    for (int q = threadIdx.x; q < RADIX; q += SORT_THREADS) h[q] = 0;
    __syncthreads();
    int64_t base = int64_t(blockIdx.x) * SORT_TILE;
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        int64_t idx = base + it * SORT_THREADS + threadIdx.x;
        if (idx < M) atomicAdd(&h[(__ldg(keys + idx) >> shift) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < RADIX; q += SORT_THREADS) hist[int64_t(q) * nblocks + blockIdx.x] = h[q];
}

// exclusive scan of `count` uint32 in place, one block
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t *__restrict__ data, int64_t count)
{
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    int64_t chunk = (count + 1023) / 1024;
    int64_t lo = t * chunk, hi = lo + chunk < count ? lo + chunk : count;
    uint32_t s = 0;
    for (int64_t q = lo; q < hi; q++) s += data[q];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {       // Hillis-Steele inclusive scan
        uint32_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (int64_t q = lo; q < hi; q++) {
        uint32_t v = data[q];
        data[q] = run;
        run += v;
    }
}

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
                     uint32_t *__restrict__ vout, const uint32_t *__restrict__ offs, int64_t M, int shift, int nblocks)
{
    constexpr int RADIX = 1 << BITS;
    __shared__ uint32_t wcount[SORT_WARPS][RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = threadIdx.x; q < SORT_WARPS * RADIX; q += SORT_THREADS) (&wcount[0][0])[q] = 0;
    __syncthreads();
    // warp-striped tile: warp w owns 256 consecutive elements; item it covers 32 consecutive ones,
    // so (it, lane) order == input order (needed for stability)
    const int64_t wbase = int64_t(blockIdx.x) * SORT_TILE + warp * (32 * SORT_ITEMS);
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        int64_t idx = wbase + it * 32 + lane;
        bool valid = idx < M;
        key[it] = valid ? __ldg(kin + idx) : 0u;
        val[it] = valid ? __ldg(vin + idx) : 0u;
        uint32_t digit = valid ? ((key[it] >> shift) & (RADIX - 1)) : RADIX;      // RADIX = "no element"
        uint32_t peers = __match_any_sync(FULL, digit);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = wcount[warp][digit];
            wcount[warp][digit] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        rank[it] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    for (int q = threadIdx.x; q < RADIX; q += SORT_THREADS) {
        uint32_t run = __ldg(offs + int64_t(q) * nblocks + blockIdx.x);
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            uint32_t c = wcount[w][q];
            wcount[w][q] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        int64_t idx = wbase + it * 32 + lane;
        if (idx < M) {
            uint32_t pos = wcount[warp][(key[it] >> shift) & (RADIX - 1)] + rank[it];
            kout[pos] = key[it];
            vout[pos] = val[it];
        }
    }
}

// sorted (key, value): every run of equal keys is cut into chunks of <= SEG_CHUNK positions; the thread at the
// head of a chunk sums it (exact int64, so the association order is irrelevant) and merges it into acc/cnt
constexpr int SEG_CHUNK = 32;

template <bool MEAN>
__global__ void td_sorted_accum_kernel(void *__restrict__ acc, uint32_t *__restrict__ cnt, uint32_t *__restrict__ touched,
                                       UpdCtrl *__restrict__ ctrl, const uint32_t *__restrict__ keys,
                                       const uint32_t *__restrict__ vals, const float *__restrict__ dw, int64_t M)
{
    int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (p >= M) return;
    const uint32_t k = __ldg(keys + p);
    if (k == 0xFFFFFFFFu) return;
    if ((p % SEG_CHUNK) != 0 && __ldg(keys + p - 1) == k) return;
    const int64_t end = (p / SEG_CHUNK + 1) * SEG_CHUNK < M ? (p / SEG_CHUNK + 1) * SEG_CHUNK : M;
    long long qs = 0;
    uint32_t nf = 0;
    for (int64_t q = p; q < end && __ldg(keys + q) == k; q++) {
        const uint32_t v = __ldg(vals + q);
        qs += quantize(__ldg(dw + (v & 0x7FFFFFFFu)));
        nf += v >> 31;
    }
    // a chunk that continues a run started in an earlier chunk may hold no 'first' contribution: it must not
    // be mistaken for an untouched key, so the touched marker is "count of contributions" when !MEAN
    if (MEAN) {
        atomicAdd(reinterpret_cast<unsigned long long *>(acc) + k, (unsigned long long)qs);
        const bool run_head = (p == 0) || __ldg(keys + p - 1) != k;
        uint32_t old = atomicAdd(cnt + k, nf);
        (void)old;
        if (run_head) touched[atomicAdd(&ctrl->count, 1u)] = k;       // exactly one head per key run
    } else {
        accumulate_key(true, acc, cnt, touched, ctrl, k, 0.0f, qs, 1u, 0);
    }
}

// ------------------------------------------------------------------------------------------------
// fused game loops
// ------------------------------------------------------------------------------------------------
__global__ void games_init_kernel(b2048_games_t g, uint64_t first_id, int reset_counters)
{
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (reset_counters && i < B2048_CTR_COUNT) g.counters[i] = 0;
    if (reset_counters && i < 17) g.tile_hist[i] = 0;
    if (i >= g.B) return;
    uint64_t id = first_id + uint64_t(i);
    g.game_id[i] = id;
    g.board[i] = spawn_initial(g.seed, id);
    g.score[i] = 0;
    g.moves[i] = 0;
    g.state[i] = 0;
    g.old_label[i] = 0.0f;
    g.flags[i] = 0;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src, int width)
{
    uint32_t lo = __shfl_sync(FULL, uint32_t(v), src, width);
    uint32_t hi = __shfl_sync(FULL, uint32_t(v >> 32), src, width);
    return (uint64_t(hi) << 32) | lo;
}

__device__ __forceinline__ void warp_add_counter(uint64_t *ctr, uint32_t v)
{
    v = __reduce_add_sync(FULL, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(reinterpret_cast<unsigned long long *>(ctr), (unsigned long long)v);
}

__device__ __forceinline__ void log_finished(const b2048_games_t &g, uint64_t id, uint32_t score, uint32_t moves,
                                             uint32_t max_exp, uint64_t board)
{
    if (!g.fin_log) return;
    unsigned long long idx = atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_LOG), 1ULL);
    if (int64_t(idx) < g.fin_cap) {
        uint4 *rec = reinterpret_cast<uint4 *>(g.fin_log) + 2 * idx;
        rec[0] = make_uint4(uint32_t(id), uint32_t(id >> 32), score, moves);
        rec[1] = make_uint4(max_exp, uint32_t(board), uint32_t(board >> 32), 0u);
    }
}

// One afterstate per lane (lane d of a 4-lane group = direction d), value by n-tuple gather, then a
// width-4 shuffle argmax with the reference's tie rule (strict '>' scanning d = 0..3: lowest d wins).
template <int N, bool COHERENT = false>
__device__ __forceinline__ void best_move(const float *__restrict__ w, const LutGlobal &L, uint64_t board, int d,
                                          bool run, uint64_t &best_after, uint32_t &best_gain, float &best_value,
                                          int &best_dir, uint32_t &best_flags, uint32_t &n_valid)
{
    uint32_t gain = 0, fl = 0;
    uint64_t after = move_dir(L, board, d, gain, fl);
    const bool valid = run && (fl & 1u);
    float v = valid ? evaluate<N, COHERENT>(w, after) : -INFINITY;
    // a direction that would create 2^16 is kept valid here; the caller stops the game if it wins
    float bv = v;
    int bd = valid ? d : 4;                           // invalid lanes never win ties
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
        float ov = __shfl_xor_sync(FULL, bv, off, 4);
        int od = __shfl_xor_sync(FULL, bd, off, 4);
        if (od < 4 && (bd == 4 || ov > bv || (ov == bv && od < bd))) { bv = ov; bd = od; }
    }
    n_valid = __popc(__ballot_sync(FULL, valid) >> ((threadIdx.x & 31) & ~3) & 0xFu);
    const int src = bd & 3;
    best_after = shfl64(after, src, 4);
    best_gain = __shfl_sync(FULL, gain, src, 4);
    best_flags = __shfl_sync(FULL, fl, src, 4);
    best_value = bv;
    best_dir = bd;
}

template <int N>
__global__ void __launch_bounds__(128)
greedy_play_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, b2048_games_t g, int max_steps,
                   int limit_tile, int step_limit, b2048_replay_t rp, int has_replay, int8_t *__restrict__ trace_dir,
                   float *__restrict__ trace_value, uint16_t *__restrict__ trace_spawn, int64_t trace_len)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t slot = t >> 2;
    const int d = int(t & 3);
    const bool in = slot < g.B;
    LutGlobal L{lut};
    uint64_t board = in ? g.board[slot] : 0;
    uint32_t score = in ? g.score[slot] : 0;
    uint32_t odo = in ? g.moves[slot] : 0;
    uint32_t flags = in ? g.flags[slot] : B2048_F_DONE;
    const uint64_t id = in ? g.game_id[slot] : 0;
    bool run = in && !(flags & B2048_F_DONE);
    uint32_t c_moves = 0, c_evals = 0, c_fin = 0, c_score = 0, c_msum = 0, c_ovf = 0;
    for (int step = 0; step < max_steps; step++) {
        if (!__any_sync(FULL, run)) break;
        if (run) {
            bool stop = game_over(board) || (limit_tile && max_tile(board) >= limit_tile) || int(odo) >= step_limit;
            if (stop) {
                flags |= B2048_F_DONE;
                run = false;
                if (d == 0) {
                    c_fin++; c_score += score; c_msum += odo;
                    atomicAdd(g.tile_hist + max_tile(board), 1u);
                    log_finished(g, id, score, odo, max_tile(board), board);
                }
            }
        }
        if (run && has_replay) {                                    // recorded spawns exhausted -> pause
            if (int64_t(odo) >= rp.len || __ldg(rp.tile + slot * rp.len + odo) == 0) run = false;
        }
        uint64_t ba;
        uint32_t bg, bf, nv;
        float bv;
        int bd;
        best_move<N>(w, L, board, d, run, ba, bg, bv, bd, bf, nv);
        if (run) {
            if (bf & 2u) {                                          // 2^16 escape: flag + stop
                flags |= B2048_F_DONE | B2048_F_OVERFLOW;
                run = false;
                if (d == 0) {
                    c_fin++; c_score += score; c_msum += odo; c_ovf++;
                    atomicAdd(g.tile_hist + 16, 1u);
                    log_finished(g, id, score, odo, 16u, board);
                }
            } else {
                if (d == 0) {
                    c_moves++; c_evals += nv;
                    if (trace_dir && int64_t(odo) < trace_len) trace_dir[slot * trace_len + odo] = int8_t(bd);
                    if (trace_value && int64_t(odo) < trace_len) trace_value[slot * trace_len + odo] = bv;
                }
                board = ba;
                score += bg;
                uint32_t sp;
                if (has_replay) {
                    uint32_t tl = __ldg(rp.tile + slot * rp.len + odo);
                    uint32_t ps = __ldg(rp.pos + slot * rp.len + odo) & 15u;
                    int sh = 4 * (15 - int(ps));
                    board = (board & ~(0xFULL << sh)) | (uint64_t(tl & 15u) << sh);
                    sp = (tl << 8) | ps;
                    odo++;
                } else {
                    odo++;
                    Philox4 r = spawn_words(g.seed, id, odo, 0u);
                    sp = spawn_apply(board, r.x, r.y);
                }
                if (d == 0 && trace_spawn && int64_t(odo) <= trace_len) trace_spawn[slot * trace_len + odo - 1] = uint16_t(sp);
            }
        }
    }
    if (in && d == 0) {
        g.board[slot] = board;
        g.score[slot] = score;
        g.moves[slot] = odo;
        g.flags[slot] = uint8_t(flags);
    }
    warp_add_counter(g.counters + B2048_CTR_MOVES, c_moves);
    warp_add_counter(g.counters + B2048_CTR_EVALS, c_evals);
    warp_add_counter(g.counters + B2048_CTR_FINISHED, c_fin);
    warp_add_counter(g.counters + B2048_CTR_SCORE_SUM, c_score);
    warp_add_counter(g.counters + B2048_CTR_MOVES_SUM, c_msum);
    warp_add_counter(g.counters + B2048_CTR_OVERFLOW, c_ovf);
    warp_add_counter(g.counters + B2048_CTR_ACTIVE, (in && d == 0 && !(flags & B2048_F_DONE)) ? 1u : 0u);
}

// TD lock-step, phase A (see b2048.h).  4 lanes per slot (lane d = direction d); all 32 lanes of a warp must
// call this together (width-4 shuffles inside).
struct StepCounters {
    uint32_t moves = 0, evals = 0, upd = 0, fin = 0, score = 0, msum = 0, ovf = 0;
};

__device__ __forceinline__ void flush_counters(uint64_t *counters, StepCounters &c)
{
    warp_add_counter(counters + B2048_CTR_MOVES, c.moves);
    warp_add_counter(counters + B2048_CTR_EVALS, c.evals);
    warp_add_counter(counters + B2048_CTR_UPDATES, c.upd);
    warp_add_counter(counters + B2048_CTR_FINISHED, c.fin);
    warp_add_counter(counters + B2048_CTR_SCORE_SUM, c.score);
    warp_add_counter(counters + B2048_CTR_MOVES_SUM, c.msum);
    warp_add_counter(counters + B2048_CTR_OVERFLOW, c.ovf);
    c = StepCounters{};
}

template <int N, bool COHERENT>
__device__ __forceinline__ void phase_a_slot(const float *__restrict__ w, const LutGlobal &L, const b2048_games_t &g,
                                             float alpha, int64_t slot, int d, bool in, uint64_t *__restrict__ upd_board,
                                             float *__restrict__ upd_dw, const b2048_replay_t &rp, int has_replay,
                                             int8_t *__restrict__ trace_dir, float *__restrict__ trace_value,
                                             float *__restrict__ trace_dw, uint16_t *__restrict__ trace_spawn,
                                             int64_t trace_len, StepCounters &c)
{
    constexpr int F = num_feat(N);
    uint64_t board = in ? g.board[slot] : 0;
    uint32_t score = in ? g.score[slot] : 0;
    uint32_t odo = in ? g.moves[slot] : 0;
    uint32_t flags = in ? g.flags[slot] : B2048_F_DONE;
    uint64_t id = in ? g.game_id[slot] : 0;
    uint64_t state = in ? g.state[slot] : 0;
    float old_label = in ? g.old_label[slot] : 0.0f;
    bool run = in && !(flags & B2048_F_DONE);
    if (run && has_replay && !game_over(board)) {
        if (int64_t(odo) >= rp.len || __ldg(rp.tile + slot * rp.len + odo) == 0) run = false;   // spawns exhausted
    }
    const bool over = run && game_over(board);
    uint64_t ba;
    uint32_t bg, bf, nv;
    float bv;
    int bd;
    best_move<N, COHERENT>(w, L, board, d, run && !over, ba, bg, bv, bd, bf, nv);
    float dw = NAN;
    uint64_t ub = 0;
    if (run) {
        const bool finished = over || (bf & 2u);
        if (finished) {
            if (flags & B2048_F_HAVE_STATE) {                        // r_learning.py:248-249
                dw = __fdiv_rn(__fmul_rn(-old_label, alpha), float(F));
                ub = state;
            }
            if (d == 0) {
                c.fin++; c.score += score; c.msum += odo;
                if (!over) c.ovf++;
                atomicAdd(g.tile_hist + (over ? max_tile(board) : 16), 1u);
                log_finished(g, id, score, odo, over ? uint32_t(max_tile(board)) : 16u, board);
                if (trace_dir && int64_t(odo) < trace_len) {
                    trace_dir[slot * trace_len + odo] = -1;          // :247 sentinel
                    if (trace_value) trace_value[slot * trace_len + odo] = 0.0f;
                    if (trace_dw) trace_dw[slot * trace_len + odo] = dw;
                }
            }
            if (has_replay || g.id_stride == 0) {                    // single-episode mode: stop, no restart
                flags = (flags | B2048_F_DONE) & ~B2048_F_HAVE_STATE;
                if (!over) flags |= B2048_F_OVERFLOW;
            } else {                                                 // in-place restart
                id += g.id_stride;
                board = spawn_initial(g.seed, id);
                score = 0; odo = 0; state = 0; old_label = 0.0f; flags = 0;
            }
        } else {
            if (flags & B2048_F_HAVE_STATE) {                        // :238-241
                float x = __fadd_rn(float(bg), bv);                  // (best_score - score) + best_value
                x = __fsub_rn(x, old_label);
                dw = __fdiv_rn(__fmul_rn(x, alpha), float(F));
                ub = state;
            }
            if (d == 0) {
                c.moves++; c.evals += nv;
                if (trace_dir && int64_t(odo) < trace_len) {
                    trace_dir[slot * trace_len + odo] = int8_t(bd);
                    if (trace_value) trace_value[slot * trace_len + odo] = bv;
                    if (trace_dw) trace_dw[slot * trace_len + odo] = dw;
                }
            }
            board = ba;                                              // :242-245
            score += bg;
            state = ba;
            old_label = bv;
            flags |= B2048_F_HAVE_STATE;
            uint32_t sp;
            if (has_replay) {                                        // :246 new_tile
                uint32_t tl = __ldg(rp.tile + slot * rp.len + odo);
                uint32_t ps = __ldg(rp.pos + slot * rp.len + odo) & 15u;
                int sh = 4 * (15 - int(ps));
                board = (board & ~(0xFULL << sh)) | (uint64_t(tl & 15u) << sh);
                sp = (tl << 8) | ps;
                odo++;
            } else {
                odo++;
                Philox4 r = spawn_words(g.seed, id, odo, 0u);
                sp = spawn_apply(board, r.x, r.y);
            }
            if (d == 0 && trace_spawn && int64_t(odo) <= trace_len) trace_spawn[slot * trace_len + odo - 1] = uint16_t(sp);
        }
        if (d == 0 && !isnan(dw)) c.upd++;
    }
    if (in && d == 0) {
        upd_board[slot] = ub;
        upd_dw[slot] = dw;
        if (run) {
            g.board[slot] = board;
            g.score[slot] = score;
            g.moves[slot] = odo;
            g.game_id[slot] = id;
            g.state[slot] = state;
            g.old_label[slot] = old_label;
            g.flags[slot] = uint8_t(flags);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(128)
td_phase_a_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, b2048_games_t g, float alpha,
                  uint64_t *__restrict__ upd_board, float *__restrict__ upd_dw, b2048_replay_t rp, int has_replay,
                  int8_t *__restrict__ trace_dir, float *__restrict__ trace_value, float *__restrict__ trace_dw,
                  uint16_t *__restrict__ trace_spawn, int64_t trace_len)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t slot = t >> 2;
    LutGlobal L{lut};
    StepCounters c;
    phase_a_slot<N, false>(w, L, g, alpha, slot, int(t & 3), slot < g.B, upd_board, upd_dw, rp, has_replay, trace_dir,
                           trace_value, trace_dw, trace_spawn, trace_len, c);
    flush_counters(g.counters, c);
}

__global__ void delta_pack_kernel(const float *__restrict__ delta, float *__restrict__ packed, int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
        float d = delta[i];
        packed[i] = d;
        packed[count + i] = d != 0.0f ? 1.0f : 0.0f;
    }
}

__global__ void delta_apply_kernel(float *__restrict__ w, float *__restrict__ w_sync, float *__restrict__ delta,
                                   const float *__restrict__ delta_sum, const float *__restrict__ contributors,
                                   int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
        float s = delta_sum[i];
        if (contributors) {
            float c = contributors[i];
            if (c > 1.0f) s = __fdiv_rn(s, c);
        }
        float v = __fadd_rn(w_sync[i], s);
        w_sync[i] = v;
        w[i] = v;
        delta[i] = 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch helpers
// ------------------------------------------------------------------------------------------------
#define DISPATCH_N(n, ...)                               \
    switch (n) {                                         \
    case 2: { constexpr int N = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int N = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int N = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int N = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int N = 6; __VA_ARGS__; } break; \
    default: return B2048_EINVAL;                        \
    }

inline int key_bits(int n)
{
    int64_t nw = table_offset(n, num_feat(n));
    int b = 1;
    while ((int64_t(1) << b) < nw + 1) b++;          // +1: the 0xFFFFFFFF sentinel must sort last
    return b;
}

struct WorkLayout {
    int64_t M, nw;      // contributions, weights
    int nblocks;        // sort tiles
    size_t acc, cnt, touched, ctrl, keys_a, keys_b, vals_a, vals_b, hist, total;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

inline WorkLayout work_layout(int n, int64_t m, int mode)
{
    WorkLayout L{};
    L.M = m * 8 * num_feat(n);
    L.nw = table_offset(n, num_feat(n));
    L.nblocks = int(cdiv(L.M, SORT_TILE));
    size_t o = 0;
    if (mode == (B2048_UPD_ATOMIC | B2048_UPD_SUM)) { L.total = 0; return L; }
    L.ctrl = o; o += 256;
    L.acc = o; o += align256(size_t(L.nw) * 8 * acc_replicas(L.nw));
    L.cnt = o; o += align256(size_t(L.nw) * 4);
    int64_t cap = L.M < L.nw ? L.M : L.nw;
    L.touched = o; o += align256(size_t(cap > 0 ? cap : 1) * 4);
    if (mode & B2048_UPD_SORTED) {
        size_t kb = align256(size_t(L.M > 0 ? L.M : 1) * 4);
        L.keys_a = o; o += kb;
        L.keys_b = o; o += kb;
        L.vals_a = o; o += kb;
        L.vals_b = o; o += kb;
        L.hist = o; o += align256(size_t(256) * size_t(L.nblocks > 0 ? L.nblocks : 1) * 4);
    }
    L.total = o;
    return L;
}

template <int BITS>
int radix_pass(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, uint32_t *hist, int64_t M,
               int shift, int nblocks, cudaStream_t st)
{
    radix_hist_kernel<BITS><<<nblocks, SORT_THREADS, 0, st>>>(kin, M, shift, hist, nblocks);
    scan_kernel<<<1, 1024, 0, st>>>(hist, int64_t(1 << BITS) * nblocks);
    radix_scatter_kernel<BITS><<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, hist, M, shift, nblocks);
    return launch_status();
}

template <int N, bool EXACT, bool MEAN, bool DIRECT>
void launch_accum(unsigned grid, cudaStream_t st, float *w, float *delta, void *acc, uint32_t *cnt, uint32_t *touched,
                  UpdCtrl *ctrl, const uint64_t *boards, const float *dw, int64_t m)
{
    td_accum_kernel<N, EXACT, MEAN, DIRECT><<<grid, 128, 0, st>>>(w, delta, acc, cnt, touched, ctrl, boards, dw, m);
}

int td_update_impl(int n, float *weights, float *delta, const uint64_t *boards, const float *dw, int64_t m, int mode,
                   void *work, size_t work_bytes, cudaStream_t st)
{
    if (m == 0) return 0;
    const bool det = mode & B2048_UPD_DETERMINISTIC, mean = mode & B2048_UPD_MEAN, sorted = mode & B2048_UPD_SORTED;
    if (sorted && !det) return B2048_EINVAL;
    const unsigned grid = unsigned(cdiv(m * 8, 128));
    if (!det && !mean) {
        DISPATCH_N(n, launch_accum<N, false, false, true>(grid, st, weights, delta, nullptr, nullptr, nullptr, nullptr,
                                                          boards, dw, m));
        return launch_status();
    }
    WorkLayout L = work_layout(n, m, mode);
    if (!work || work_bytes < L.total) return B2048_EWORK;
    unsigned char *base = reinterpret_cast<unsigned char *>(work);
    void *acc = base + L.acc;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(base + L.cnt), *touched = reinterpret_cast<uint32_t *>(base + L.touched);
    UpdCtrl *ctrl = reinterpret_cast<UpdCtrl *>(base + L.ctrl);
    if (!sorted) {
        if (det && mean) { DISPATCH_N(n, launch_accum<N, true, true, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m)); }
        else if (det)    { DISPATCH_N(n, launch_accum<N, true, false, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m)); }
        else             { DISPATCH_N(n, launch_accum<N, false, true, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m)); }
    } else {
        uint32_t *ka = reinterpret_cast<uint32_t *>(base + L.keys_a), *kb = reinterpret_cast<uint32_t *>(base + L.keys_b);
        uint32_t *va = reinterpret_cast<uint32_t *>(base + L.vals_a), *vb = reinterpret_cast<uint32_t *>(base + L.vals_b);
        uint32_t *hist = reinterpret_cast<uint32_t *>(base + L.hist);
        DISPATCH_N(n, td_keys_kernel<N><<<grid, 128, 0, st>>>(boards, dw, m, ka, va));
        const int bits = key_bits(n);
        const int passes = (bits + 7) / 8;
        const int per = (bits + passes - 1) / passes;       // digit width, equal for all passes
        int shift = 0, rc = 0;
        for (int p = 0; p < passes && !rc; p++, shift += per) {
            switch (per) {
            case 8: rc = radix_pass<8>(ka, va, kb, vb, hist, L.M, shift, L.nblocks, st); break;
            case 7: rc = radix_pass<7>(ka, va, kb, vb, hist, L.M, shift, L.nblocks, st); break;
            case 6: rc = radix_pass<6>(ka, va, kb, vb, hist, L.M, shift, L.nblocks, st); break;
            default: rc = radix_pass<5>(ka, va, kb, vb, hist, L.M, shift, L.nblocks, st); break;
            }
            uint32_t *tk = ka; ka = kb; kb = tk;
            uint32_t *tv = va; va = vb; vb = tv;
        }
        if (rc) return rc;
        if (mean) td_sorted_accum_kernel<true><<<unsigned(cdiv(L.M, 256)), 256, 0, st>>>(acc, cnt, touched, ctrl, ka, va, dw, L.M);
        else      td_sorted_accum_kernel<false><<<unsigned(cdiv(L.M, 256)), 256, 0, st>>>(acc, cnt, touched, ctrl, ka, va, dw, L.M);
    }
    int rc = launch_status();
    if (rc) return rc;
    const unsigned agrid = unsigned(2 * sm_count());
    if (det && mean)  td_apply_kernel<true, true><<<agrid, 256, 0, st>>>(weights, delta, acc, cnt, touched, ctrl, L.nw);
    else if (det)     td_apply_kernel<true, false><<<agrid, 256, 0, st>>>(weights, delta, acc, cnt, touched, ctrl, L.nw);
    else              td_apply_kernel<false, true><<<agrid, 256, 0, st>>>(weights, delta, acc, cnt, touched, ctrl, L.nw);
    return launch_status();
}

bool games_ok(const b2048_games_t *g)
{
    return g && g->B >= 0 && g->board && g->score && g->moves && g->game_id && g->state && g->old_label && g->flags &&
           g->counters && g->tile_hist;
}

}   // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int b2048_abi_version(void) { return B2048_ABI_VERSION; }

const char *b2048_strerror(int code)
{
    if (code == 0) return "ok";
    if (code == B2048_EINVAL) return "b2048: invalid argument";
    if (code == B2048_ENOTSUP) return "b2048: not supported";
    if (code == B2048_EWORK) return "b2048: workspace too small";
    if (code > 0) return cudaGetErrorString(cudaError_t(code));
    return "b2048: unknown error";
}

int b2048_num_feat(int n) { return num_feat(n); }

int64_t b2048_table_offset(int n, int i)
{
    int F = num_feat(n);
    if (F < 0 || i < 0 || i > F) return -1;
    return table_offset(n, i);
}

int64_t b2048_num_weights(int n)
{
    int F = num_feat(n);
    return F < 0 ? -1 : table_offset(n, F);
}

int b2048_pack(const int32_t *rows, uint64_t *boards, int64_t m, b2048_stream_t stream)
{
    if (m < 0 || (m && (!rows || !boards))) return B2048_EINVAL;
    if (!m) return 0;
    pack_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(rows, boards, m);
    return launch_status();
}

int b2048_unpack(const uint64_t *boards, int32_t *rows, int64_t m, b2048_stream_t stream)
{
    if (m < 0 || (m && (!rows || !boards))) return B2048_EINVAL;
    if (!m) return 0;
    unpack_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(boards, rows, m);
    return launch_status();
}

int b2048_lut_build(uint32_t *lut, b2048_stream_t stream)
{
    if (!lut) return B2048_EINVAL;
    lut_build_kernel<<<B2048_LUT_ENTRIES / 256, 256, 0, S(stream)>>>(lut);
    return launch_status();
}

int b2048_move4(const uint32_t *lut, const uint64_t *boards, int64_t m, uint64_t *after, uint32_t *gain,
                uint8_t *flags, uint8_t *over, b2048_stream_t stream)
{
    if (m < 0 || !lut || (m && (!boards || !after || !gain || !flags))) return B2048_EINVAL;
    if (!m) return 0;
    move4_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(lut, boards, m, after, gain, flags, over);
    return launch_status();
}

int b2048_board_stats(const uint64_t *boards, int64_t m, uint8_t *stats, uint16_t *empty_mask, b2048_stream_t stream)
{
    if (m < 0 || (m && !boards)) return B2048_EINVAL;
    if (!m) return 0;
    board_stats_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(boards, m, stats, empty_mask);
    return launch_status();
}

int b2048_spawn_philox(uint64_t *boards, int64_t m, uint64_t seed, const uint64_t *game_id, const uint32_t *move_no,
                       uint16_t *spawn, b2048_stream_t stream)
{
    if (m < 0 || (m && (!boards || !game_id || !move_no))) return B2048_EINVAL;
    if (!m) return 0;
    spawn_philox_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(boards, m, seed, game_id, move_no, spawn);
    return launch_status();
}

int b2048_spawn_initial(uint64_t *boards, int64_t m, uint64_t seed, uint64_t first_id, uint64_t id_step,
                        b2048_stream_t stream)
{
    if (m < 0 || (m && !boards)) return B2048_EINVAL;
    if (!m) return 0;
    spawn_initial_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(boards, m, seed, first_id, id_step);
    return launch_status();
}

int b2048_spawn_replay(uint64_t *boards, int64_t m, const uint8_t *tile, const uint8_t *pos, b2048_stream_t stream)
{
    if (m < 0 || (m && (!boards || !tile || !pos))) return B2048_EINVAL;
    if (!m) return 0;
    spawn_replay_kernel<<<unsigned(cdiv(m, 256)), 256, 0, S(stream)>>>(boards, m, tile, pos);
    return launch_status();
}

int b2048_sweep(const uint32_t *lut, const uint64_t *boards, int64_t m, uint64_t seed, uint64_t first_index,
                uint64_t *after, uint32_t *gain, uint8_t *flags, uint64_t *spawned, b2048_stream_t stream)
{
    if (m < 0 || !lut || (m && (!boards || !after || !gain || !flags))) return B2048_EINVAL;
    if (!m) return 0;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SWEEP_SMEM));
        if (e != cudaSuccess) return int(e);
        attr_set[dev] = true;
    }
    int64_t want = cdiv(m, SWEEP_THREADS);
    unsigned grid = unsigned(want < sm_count() ? want : sm_count());
    sweep_kernel<<<grid, SWEEP_THREADS, SWEEP_SMEM, S(stream)>>>(lut, boards, m, seed, first_index, after, gain, flags,
                                                                 spawned);
    return launch_status();
}

int b2048_features(int n, const uint64_t *boards, int64_t m, int32_t *feat, b2048_stream_t stream)
{
    if (m < 0 || num_feat(n) < 0 || (m && (!boards || !feat))) return B2048_EINVAL;
    if (!m) return 0;
    DISPATCH_N(n, features_kernel<N><<<unsigned(cdiv(m, 128)), 128, 0, S(stream)>>>(boards, m, feat));
    return launch_status();
}

int b2048_evaluate(int n, const float *weights, const uint64_t *boards, int64_t m, float *value, b2048_stream_t stream)
{
    if (m < 0 || num_feat(n) < 0 || !weights || (m && (!boards || !value))) return B2048_EINVAL;
    if (!m) return 0;
    DISPATCH_N(n, evaluate_kernel<N><<<unsigned(cdiv(m, 128)), 128, 0, S(stream)>>>(weights, boards, m, value));
    return launch_status();
}

size_t b2048_td_update_workspace(int n, int64_t m, int mode)
{
    if (num_feat(n) < 0 || m < 0) return 0;
    return work_layout(n, m, mode).total;
}

int b2048_td_update(int n, float *weights, float *delta, const uint64_t *boards, const float *dw, int64_t m, int mode,
                    void *work, size_t work_bytes, b2048_stream_t stream)
{
    if (m < 0 || num_feat(n) < 0 || !weights || (m && (!boards || !dw)) || (mode & ~7)) return B2048_EINVAL;
    return td_update_impl(n, weights, delta, boards, dw, m, mode, work, work_bytes, S(stream));
}

int b2048_games_init(const b2048_games_t *g, uint64_t first_id, int reset_counters, b2048_stream_t stream)
{
    if (!games_ok(g)) return B2048_EINVAL;
    int64_t threads = g->B > 32 ? g->B : 32;
    games_init_kernel<<<unsigned(cdiv(threads, 256)), 256, 0, S(stream)>>>(*g, first_id, reset_counters);
    return launch_status();
}

int b2048_greedy_play(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, int max_steps,
                      int limit_tile, int step_limit, const b2048_replay_t *replay, int8_t *trace_dir,
                      float *trace_value, uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream)
{
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || max_steps < 0) return B2048_EINVAL;
    if (replay && (!replay->tile || !replay->pos || replay->len < 0)) return B2048_EINVAL;
    if (g->B == 0) return 0;
    cudaError_t e = cudaMemsetAsync(g->counters + B2048_CTR_ACTIVE, 0, sizeof(uint64_t), S(stream));
    if (e != cudaSuccess) return int(e);
    b2048_replay_t rp = replay ? *replay : b2048_replay_t{nullptr, nullptr, 0};
    unsigned grid = unsigned(cdiv(g->B * 4, 128));
    DISPATCH_N(n, greedy_play_kernel<N><<<grid, 128, 0, S(stream)>>>(weights, lut, *g, max_steps, limit_tile, step_limit,
                                                                     rp, replay ? 1 : 0, trace_dir, trace_value,
                                                                     trace_spawn, trace_len));
    return launch_status();
}

int b2048_td_phase_a(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, float alpha,
                     uint64_t *upd_board, float *upd_dw, const b2048_replay_t *replay, int8_t *trace_dir,
                     float *trace_value, float *trace_dw, uint16_t *trace_spawn, int64_t trace_len,
                     b2048_stream_t stream)
{
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || !upd_board || !upd_dw) return B2048_EINVAL;
    if (replay && (!replay->tile || !replay->pos || replay->len < 0)) return B2048_EINVAL;
    if (g->B == 0) return 0;
    b2048_replay_t rp = replay ? *replay : b2048_replay_t{nullptr, nullptr, 0};
    unsigned grid = unsigned(cdiv(g->B * 4, 128));
    DISPATCH_N(n, td_phase_a_kernel<N><<<grid, 128, 0, S(stream)>>>(weights, lut, *g, alpha, upd_board, upd_dw, rp,
                                                                   replay ? 1 : 0, trace_dir, trace_value, trace_dw,
                                                                   trace_spawn, trace_len));
    return launch_status();
}

int b2048_td_step(int n, float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                  int mode, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                  const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value, float *trace_dw,
                  uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream)
{
    if (mode & ~7) return B2048_EINVAL;
    int rc = b2048_td_phase_a(n, weights, lut, g, alpha, upd_board, upd_dw, replay, trace_dir, trace_value, trace_dw,
                              trace_spawn, trace_len, stream);
    if (rc || g->B == 0) return rc;
    return td_update_impl(n, weights, delta, upd_board, upd_dw, g->B, mode, work, work_bytes, S(stream));
}

int b2048_td_run(int n, float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                 int mode, int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                 b2048_stream_t stream)
{
    if (steps < 0) return B2048_EINVAL;
    for (int s = 0; s < steps; s++) {
        int rc = b2048_td_step(n, weights, delta, lut, g, alpha, mode, upd_board, upd_dw, work, work_bytes, nullptr,
                               nullptr, nullptr, nullptr, nullptr, 0, stream);
        if (rc) return rc;
    }
    return 0;
}

int b2048_delta_pack(const float *delta, float *packed, int64_t count, b2048_stream_t stream)
{
    if (count < 0 || (count && (!delta || !packed))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_pack_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(delta, packed, count);
    return launch_status();
}

int b2048_delta_apply(float *weights, float *w_sync, float *delta, const float *delta_sum, const float *contributors,
                      int64_t count, b2048_stream_t stream)
{
    if (count < 0 || (count && (!weights || !w_sync || !delta || !delta_sum))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_apply_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(weights, w_sync, delta, delta_sum,
                                                                              contributors, count);
    return launch_status();
}

}   // extern "C"
