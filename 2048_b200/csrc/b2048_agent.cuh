// b2048_agent.cuh -- the n-tuple agent kernels (templates on the tuple size N) and their launchers:
// features / evaluate, TD update (atomic | deterministic, sum | per-key mean, direct | sorted), the fused
// game loops (greedy play, lock-step phase A) and the persistent lock-step trainer.  Instantiated once per N
// by b2048_agent_inst.cu.
#pragma once
#include <cmath>
#include <type_traits>
#include <vector>

#include "b2048_host.cuh"

namespace {

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

// ---- accumulate pass -------------------------------------------------------------------------
// Thread per (entry j, image s); the 8 images of an entry are 8 adjacent lanes, a warp holds 4 entries.
// For every table i the warp first merges lanes that hit the same key (__match_any_sync): hot keys
// (empty rows/squares early in a game) would otherwise serialise thousands of same-address atomics in L2.
//   DIRECT          atomic + sum rule: the merged contribution goes straight into w (and delta)
//   otherwise       acc[k] += contribution (float RED, or exact int64 fixed point when EXACT),
//                   cnt[k] += number of distinct entries (MEAN) or 1; the lane that sees cnt go 0 -> >0
//                   appends k to the touched list, so the apply pass needs no atomics at all.
constexpr double FIX_SCALE = 4294967296.0;      // 2^32: exact-mode contributions are llrint(dw * 2^32)


// The accumulator is replicated ACC_REPLICAS(nw) times (replica = CTA index mod R): contributions are shared
// so broadly (only 16-30 % of the keys of a lock-step are distinct) that same-address atomics, which L2
// serialises at ~1.4 ns each, would otherwise dominate.  The apply pass sums the replicas of a touched key.

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

// ------------------------------------------------------------------------------------------------
// persistent lock-step trainer: `steps` lock-steps of QAgent.episode for all slots in ONE cooperative
// launch (b2048_td_run).  The stepwise path costs 3 launches (~4.5 us each) plus three pipeline
// fill/drain phases per lock-step; here every CTA owns a fixed range of game slots and runs
//     phase A (own slots, weights read through L2)                 -> upd_board / upd_dw
//     phase B (own entries: 8 images x F tables, warp-merged)      -> acc / cnt (+ the CTA's own key list)
//     ---- grid barrier ----
//     apply   (the keys this CTA touched FIRST: w += S [/ G], accumulators back to zero)
//     ---- grid barrier ----
// Phase B writes only the accumulators, never w, so no barrier is needed between A and B of different
// CTAs; the DIRECT variant (atomic + sum rule, increments straight into w) needs one there instead.
// One accumulator copy (no replicas): the first-touch test must see every contribution to a key.
//   float modes   acc = float2[nw] {sum, count}: ONE returning ATOMG.ADD.F32x2 per merged contribution
//   exact modes   acc = int64[nw] (RED.64) + cnt = u32[nw] (returning ATOMG), as in the stepwise path
// A contribution that finds count == 0 appends its key to the CTA-private list (shared-memory cursor), so the
// apply phase has no global counter, no global list and no atomics.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of a cooperative launch; `target` is the CTA's running arrival total (thread 0 only)
__device__ __forceinline__ void grid_barrier(uint32_t *bar, uint32_t &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();                              // release: the CTA's writes and atomics before the arrival
        atomicAdd(bar, 1u);
        // a wait of 2^25 polls (seconds) can only be a lost CTA: trap instead of hanging the device
        uint32_t polls = 0;
        while (ld_acquire_gpu(bar) < target)
            if (++polls > (1u << 25)) __trap();
    }
    __syncthreads();
}

template <int N, int I, int CH, class Fn>
__device__ __forceinline__ void for_each_feature_strided(Fn &&f)
{
    if constexpr (I < num_feat(N)) {
        f(std::integral_constant<int, I>{});
        for_each_feature_strided<N, I + CH, CH>(f);
    }
}

struct AccumTarget {
    float *w, *delta;
    void *acc;
    uint32_t *cnt;
    uint32_t *list;         // this CTA's key list (global memory segment)
    uint32_t *cursor;       // its length (shared memory)
};

// Tables i = C, C+CH, ... of the 4 entries x 8 images held by one warp (lanes 8e..8e+7 = entry e).
// One __match_any_sync per table yields everything the merge needs: the lanes sharing the key (any entry,
// any image), hence per entry e the number of its images on the key (c_e) -- sum = sum_e c_e * dw_e,
// distinct entries G = #{e : c_e > 0} -- and the leader lane that issues the atomic.
// Pass 1 issues every atomic of the chunk back to back; pass 2 consumes the returned counts (a single loop
// would expose one L2 round trip per table).
template <int N, bool EXACT, bool MEAN, bool DIRECT, int C, int CH>
__device__ __forceinline__ void accum_chunk(const AccumTarget &t, uint64_t b, const float (&de)[4],
                                            const long long (&qe)[4], bool live, int lane)
{
    constexpr int F = num_feat(N);
    constexpr int P = (F - C + CH - 1) / CH;
    const uint64_t y = (N == 6) ? clamp13(b) : 0;
    uint32_t key[P];
    float oldf[P];
    uint32_t oldu[P];
    for_each_feature_strided<N, C, CH>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int p = (i - C) / CH;
        const uint32_t k = uint32_t(table_offset(N, i)) + feat_index<N, i>(b, y);
        const uint32_t peers = __match_any_sync(FULL, live ? k : ~uint32_t(lane));
        const int c0 = __popc(peers & 0x000000FFu), c1 = __popc(peers & 0x0000FF00u);
        const int c2 = __popc(peers & 0x00FF0000u), c3 = __popc(peers & 0xFF000000u);
        key[p] = k;
        oldf[p] = 1.0f;
        oldu[p] = 1u;
        if (live && lane == __ffs(peers) - 1) {
            const uint32_t nf = MEAN ? uint32_t((c0 != 0) + (c1 != 0) + (c2 != 0) + (c3 != 0)) : 1u;
            if (EXACT) {
                const long long qs = c0 * qe[0] + c1 * qe[1] + c2 * qe[2] + c3 * qe[3];
                atomicAdd(reinterpret_cast<unsigned long long *>(t.acc) + k, (unsigned long long)qs);
                oldu[p] = atomicAdd(t.cnt + k, nf);
            } else {
                float fs = c0 ? float(c0) * de[0] : 0.0f;
                if (c1) fs += float(c1) * de[1];
                if (c2) fs += float(c2) * de[2];
                if (c3) fs += float(c3) * de[3];
                if (DIRECT) {
                    atomicAdd(t.w + k, fs);
                    if (t.delta) atomicAdd(t.delta + k, fs);
                } else {
                    oldf[p] = atomicAdd(reinterpret_cast<float2 *>(t.acc) + k, make_float2(fs, float(nf))).y;
                }
            }
        }
    });
    if (DIRECT) return;
    uint32_t first = 0, total = 0;                    // bit p: this lane saw count 0 -> >0 on key[p]
    uint32_t before[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        const bool f = EXACT ? (oldu[p] == 0u) : (oldf[p] == 0.0f);
        const uint32_t m = __ballot_sync(FULL, f);
        first |= uint32_t(f) << p;
        before[p] = total + __popc(m & ((1u << lane) - 1u));
        total += __popc(m);
    }
    if (total) {                                      // warp-uniform
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(t.cursor, total);
        base = __shfl_sync(FULL, base, 0);
#pragma unroll
        for (int p = 0; p < P; p++)
            if ((first >> p) & 1u) t.list[base + before[p]] = key[p];
    }
}

template <bool EXACT, bool MEAN>
__device__ __forceinline__ void apply_key_single(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc,
                                                 uint32_t *__restrict__ cnt, uint32_t k)
{
    float u;
    if (EXACT) {
        long long *a = reinterpret_cast<long long *>(acc) + k;
        const long long qs = __ldcg(a);
        const uint32_t c = __ldcg(cnt + k);
        double x = double(qs) / FIX_SCALE;
        if (MEAN) x = x / double(c);
        u = __double2float_rn(x);
        __stcg(a, 0LL);
        __stcg(cnt + k, 0u);
    } else {
        float2 *a = reinterpret_cast<float2 *>(acc) + k;
        const float2 v = __ldcg(a);
        u = MEAN ? __fdiv_rn(v.x, v.y) : v.x;
        __stcg(a, make_float2(0.0f, 0.0f));
    }
    __stcg(w + k, __fadd_rn(__ldcg(w + k), u));
    if (delta) __stcg(delta + k, __fadd_rn(__ldcg(delta + k), u));
}


template <int N, bool EXACT, bool MEAN, bool DIRECT, int CH>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
td_persist_kernel(float *w, float *delta, void *acc, uint32_t *cnt, uint32_t *lists, PersistCtrl *ctrl,
                  const uint32_t *__restrict__ lut, b2048_games_t g, float alpha, int steps, uint64_t *upd_board,
                  float *upd_dw, int spc, int64_t list_cap, long long *tlog)
{
    // tlog (debug, B2048_PERSIST_TLOG): clock64 at the 6 phase boundaries of the last 16 steps, per CTA
    __shared__ uint32_t s_cursor;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t slot0 = int64_t(blockIdx.x) * spc;
    const int nslots = int(g.B - slot0 < spc ? (g.B - slot0 > 0 ? g.B - slot0 : 0) : spc);
    const AccumTarget tgt{w, delta, acc, cnt, lists + int64_t(blockIdx.x) * list_cap, &s_cursor};
    const b2048_replay_t no_replay{nullptr, nullptr, 0};
    LutGlobal L{lut};
    StepCounters c;
    uint32_t bar_target = 0;
    for (int step = 0; step < steps; step++) {
        if (threadIdx.x == 0) s_cursor = 0;
        long long *tl = (tlog && threadIdx.x == 0 && step >= steps - 16) ?
                        tlog + (int64_t(blockIdx.x) * 16 + (step - (steps - 16))) * 6 : nullptr;
        if (tl) tl[0] = clock64();
        // ---- phase A: 4 lanes per slot, 8 slots per warp
        for (int base = warp * 8; base < nslots; base += nwarps * 8) {
            const int ls = base + (lane >> 2);
            phase_a_slot<N, true>(w, L, g, alpha, slot0 + ls, lane & 3, ls < nslots, upd_board, upd_dw, no_replay, 0,
                                  nullptr, nullptr, nullptr, nullptr, 0, c);
        }
        if (DIRECT) grid_barrier(&ctrl->bar, bar_target);     // every slot has read W_t before anyone adds to it
        else __syncthreads();
        if (tl) tl[1] = clock64();
        // ---- phase B: 8 lanes per entry, 4 entries per warp, CH table chunks per entry group
        const int ngroups = (nslots + 3) >> 2;
        for (int it = warp; it < ngroups * CH; it += nwarps) {
            const int eg = it % ngroups, chunk = it / ngroups;
            const int le = eg * 4 + (lane >> 3);
            const bool on = le < nslots;
            const float d = on ? __ldcg(upd_dw + slot0 + le) : NAN;
            const bool live = on && (EXACT ? isfinite(d) : !isnan(d));
            if (!__any_sync(FULL, live)) continue;
            const uint64_t b = d4_image(on ? __ldcg(upd_board + slot0 + le) : 0, lane & 7);
            float de[4];
            long long qe[4];
            const long long q = (EXACT && live) ? quantize(d) : 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                de[e] = __shfl_sync(FULL, d, 8 * e);
                qe[e] = EXACT ? (long long)shfl64(uint64_t(q), 8 * e, 32) : 0;
            }
            if constexpr (CH == 1) {
                accum_chunk<N, EXACT, MEAN, DIRECT, 0, 1>(tgt, b, de, qe, live, lane);
            } else if constexpr (CH == 2) {
                if (chunk == 0) accum_chunk<N, EXACT, MEAN, DIRECT, 0, 2>(tgt, b, de, qe, live, lane);
                else            accum_chunk<N, EXACT, MEAN, DIRECT, 1, 2>(tgt, b, de, qe, live, lane);
            } else {
                static_assert(CH == 4, "table chunks: 1, 2 or 4");
                if (chunk == 0)      accum_chunk<N, EXACT, MEAN, DIRECT, 0, 4>(tgt, b, de, qe, live, lane);
                else if (chunk == 1) accum_chunk<N, EXACT, MEAN, DIRECT, 1, 4>(tgt, b, de, qe, live, lane);
                else if (chunk == 2) accum_chunk<N, EXACT, MEAN, DIRECT, 2, 4>(tgt, b, de, qe, live, lane);
                else                 accum_chunk<N, EXACT, MEAN, DIRECT, 3, 4>(tgt, b, de, qe, live, lane);
            }
        }
        __syncthreads();
        if (tl) tl[2] = clock64();
        grid_barrier(&ctrl->bar, bar_target);                 // every contribution of the step has landed
        if (tl) tl[3] = clock64();
        if (!DIRECT) {
            const uint32_t mine = s_cursor;
            for (uint32_t q = threadIdx.x; q < mine; q += blockDim.x)
                apply_key_single<EXACT, MEAN>(w, delta, acc, cnt, __ldcg(tgt.list + q));
            __syncthreads();
            if (tl) tl[4] = clock64();
            grid_barrier(&ctrl->bar, bar_target);             // W_{t+1} complete
        }
        if (tl) tl[5] = clock64();
    }
    flush_counters(g.counters, c);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctrl->exit_ticket, 1u) == gridDim.x - 1) {   // every CTA is past its last barrier
            ctrl->bar = 0;
            ctrl->exit_ticket = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
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

template <int N>
int td_update_impl(float *weights, float *delta, const uint64_t *boards, const float *dw, int64_t m, int mode,
                   void *work, size_t work_bytes, cudaStream_t st)
{
    if (m == 0) return 0;
    const bool det = mode & B2048_UPD_DETERMINISTIC, mean = mode & B2048_UPD_MEAN, sorted = mode & B2048_UPD_SORTED;
    if (sorted && !det) return B2048_EINVAL;
    const unsigned grid = unsigned(cdiv(m * 8, 128));
    if (!det && !mean) {
        launch_accum<N, false, false, true>(grid, st, weights, delta, nullptr, nullptr, nullptr, nullptr,
                                                          boards, dw, m);
        return launch_status();
    }
    WorkLayout L = work_layout(N, m, mode);
    if (!work || work_bytes < L.total) return B2048_EWORK;
    unsigned char *base = reinterpret_cast<unsigned char *>(work);
    void *acc = base + L.acc;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(base + L.cnt), *touched = reinterpret_cast<uint32_t *>(base + L.touched);
    UpdCtrl *ctrl = reinterpret_cast<UpdCtrl *>(base + L.ctrl);
    if (!sorted) {
        if (det && mean) { launch_accum<N, true, true, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m); }
        else if (det)    { launch_accum<N, true, false, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m); }
        else             { launch_accum<N, false, true, false>(grid, st, weights, delta, acc, cnt, touched, ctrl, boards, dw, m); }
    } else {
        uint32_t *ka = reinterpret_cast<uint32_t *>(base + L.keys_a), *kb = reinterpret_cast<uint32_t *>(base + L.keys_b);
        uint32_t *va = reinterpret_cast<uint32_t *>(base + L.vals_a), *vb = reinterpret_cast<uint32_t *>(base + L.vals_b);
        uint32_t *hist = reinterpret_cast<uint32_t *>(base + L.hist);
        td_keys_kernel<N><<<grid, 128, 0, st>>>(boards, dw, m, ka, va);
        const int bits = key_bits(N);
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

// ---- persistent trainer launch ---------------------------------------------------------------------
template <int N, bool EXACT, bool MEAN, bool DIRECT, int CH>
int launch_persist(int grid, cudaStream_t st, void **args)
{
    auto kern = td_persist_kernel<N, EXACT, MEAN, DIRECT, CH>;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PERSIST_THREADS, 0);
    if (e != cudaSuccess) return int(e);
    if (occ < 1 || grid > occ * sm_count()) return B2048_ENOTSUP;       // all CTAs must be co-resident
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern), dim3(unsigned(grid)), dim3(PERSIST_THREADS), args, 0, st);
    return e == cudaSuccess ? 0 : int(e);
}

template <int N, int CH>
int launch_persist_mode(bool det, bool mean, int grid, cudaStream_t st, void **args)
{
    if (!det && !mean) return launch_persist<N, false, false, true, CH>(grid, st, args);
    if (!det) return launch_persist<N, false, true, false, CH>(grid, st, args);
    if (mean) return launch_persist<N, true, true, false, CH>(grid, st, args);
    return launch_persist<N, true, false, false, CH>(grid, st, args);
}

template <int N>
int td_run_persistent(float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                      int mode, int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                      cudaStream_t st)
{
    const bool det = mode & B2048_UPD_DETERMINISTIC, mean = mode & B2048_UPD_MEAN;
    const int64_t B = g->B;
    WorkLayout L = work_layout(N, B, mode | B2048_UPD_MEAN);       // DIRECT uses only the control block
    if (!work || work_bytes < work_layout(N, B, mode).total) return B2048_EWORK;
    // one CTA per SM (tunable): slots per CTA = a multiple of 4 (one warp = 4 entries in phase B)
    int max_grid = sm_count() * env_int("B2048_PERSIST_CTAS_PER_SM", 1);
    if (max_grid > PERSIST_MAX_GRID) max_grid = PERSIST_MAX_GRID;
    int64_t spc = cdiv(cdiv(B, max_grid), 4) * 4;
    const int grid = int(cdiv(B, spc));
    const int groups = int(spc / 4);
    // table chunks: spread the entry groups of a CTA over its 16 warps
    int ch = env_int("B2048_PERSIST_CHUNKS", 0);
    if (ch != 1 && ch != 2 && ch != 4) ch = groups * 4 <= PERSIST_THREADS / 32 ? 4 : groups * 2 <= PERSIST_THREADS / 32 ? 2 : 1;
    if (N == 6 && ch == 1) ch = 2;                                 // bounds the registers of a chunk
    unsigned char *base = reinterpret_cast<unsigned char *>(work);
    void *acc = base + L.acc;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(base + L.cnt), *lists = reinterpret_cast<uint32_t *>(base + L.lists);
    PersistCtrl *ctrl = reinterpret_cast<PersistCtrl *>(base + L.ctrl);
    b2048_games_t games = *g;
    int spc_i = int(spc);
    int64_t list_cap = spc * 8 * num_feat(N);
    long long *tlog = nullptr;
    const char *tlog_path = getenv("B2048_PERSIST_TLOG");          // debug: per-phase clocks -> text file
    if (tlog_path && *tlog_path && steps >= 16) {
        if (cudaMalloc(&tlog, size_t(grid) * 16 * 6 * sizeof(long long)) != cudaSuccess) tlog = nullptr;
        else cudaMemsetAsync(tlog, 0, size_t(grid) * 16 * 6 * sizeof(long long), st);
    }
    void *args[] = {&weights, &delta, &acc, &cnt, &lists, &ctrl, &lut, &games, &alpha, &steps, &upd_board, &upd_dw,
                    &spc_i, &list_cap, &tlog};
    int rc = B2048_EINVAL;
    switch (ch) {
    case 1: rc = launch_persist_mode<N, 1>(det, mean, grid, st, args); break;
    case 2: rc = launch_persist_mode<N, 2>(det, mean, grid, st, args); break;
    default: rc = launch_persist_mode<N, 4>(det, mean, grid, st, args); break;
    }
    if (tlog) {
        std::vector<long long> h(size_t(grid) * 16 * 6);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), tlog, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(tlog);
        if (FILE *f = fopen(tlog_path, "a")) {
            fprintf(f, "# grid %d spc %d ch %d mode %d n %d steps %d: per CTA, per step: A B bar1 apply bar2 (cycles)\n", grid,
                    spc_i, ch, mode, N, steps);
            for (int c = 0; c < grid; c++)
                for (int q = 0; q < 16; q++) {
                    const long long *t = &h[(size_t(c) * 16 + q) * 6];
                    fprintf(f, "%d %d %lld %lld %lld %lld %lld\n", c, q, t[1] - t[0], t[2] - t[1], t[3] - t[2],
                            t[4] ? t[4] - t[3] : 0, t[4] ? t[5] - t[4] : 0);
                }
            fclose(f);
        }
    }
    return rc;
}


template <int N>
int features_impl(const uint64_t *boards, int64_t m, int32_t *feat, cudaStream_t st)
{
    features_kernel<N><<<unsigned(cdiv(m, 128)), 128, 0, st>>>(boards, m, feat);
    return launch_status();
}

template <int N>
int evaluate_impl(const float *w, const uint64_t *boards, int64_t m, float *value, cudaStream_t st)
{
    evaluate_kernel<N><<<unsigned(cdiv(m, 128)), 128, 0, st>>>(w, boards, m, value);
    return launch_status();
}

template <int N>
int greedy_play_impl(const float *w, const uint32_t *lut, const b2048_games_t *g, int max_steps, int limit_tile,
                     int step_limit, const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value,
                     uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st)
{
    b2048_replay_t rp = replay ? *replay : b2048_replay_t{nullptr, nullptr, 0};
    unsigned grid = unsigned(cdiv(g->B * 4, 128));
    greedy_play_kernel<N><<<grid, 128, 0, st>>>(w, lut, *g, max_steps, limit_tile, step_limit, rp, replay ? 1 : 0, trace_dir,
                                               trace_value, trace_spawn, trace_len);
    return launch_status();
}

template <int N>
int td_phase_a_impl(const float *w, const uint32_t *lut, const b2048_games_t *g, float alpha, uint64_t *upd_board,
                    float *upd_dw, const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value, float *trace_dw,
                    uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st)
{
    b2048_replay_t rp = replay ? *replay : b2048_replay_t{nullptr, nullptr, 0};
    unsigned grid = unsigned(cdiv(g->B * 4, 128));
    td_phase_a_kernel<N><<<grid, 128, 0, st>>>(w, lut, *g, alpha, upd_board, upd_dw, rp, replay ? 1 : 0, trace_dir,
                                              trace_value, trace_dw, trace_spawn, trace_len);
    return launch_status();
}

}   // namespace
