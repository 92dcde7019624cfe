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

// ---- accumulate pass (stepwise path: b2048_td_update / b2048_td_step) ---------------------------
// Thread per (entry j, table i): the keys of the 8 D4 images, duplicates among them found in registers.
//   DIRECT          atomic + sum rule: every contribution goes straight into w (and delta)
//   otherwise       acc[k] += contribution (float RED, or exact int64 fixed point when EXACT),
//                   cnt[k] += 1 per entry and key (G of the MEAN rule; a touched marker otherwise); the thread that
//                   sees cnt go 0 -> >0 appends k to the touched list, so the apply pass needs no atomics at all.
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

// One thread per (entry, table): the 8 D4 image keys of the table, a duplicate test among them in registers (an
// entry counts once per key in G), one atomic per key into the CTA's accumulator replica, and the keys whose count
// went 0 -> >0 appended to the touched list (one warp scan + one global atomic per warp).  No warp-level key
// matching: __match_any_sync costs ~12 cycles per distinct value (profiles/r01b_microbench_warpops.txt).
template <int N>
__device__ __forceinline__ uint32_t table_offset_rt0(int i)
{
    if (N <= 4) return uint32_t(i) * uint32_t(table_size(N, 0));
    return i <= 17 ? uint32_t(i) * 65536u : i <= 21 ? 17u * 65536u + uint32_t(i - 17) * 1048576u
                                                    : 17u * 65536u + 4u * 1048576u + uint32_t(i - 21) * 7529536u;
}

template <int N, bool EXACT, bool MEAN, bool DIRECT>
__global__ void __launch_bounds__(128)
td_accum_kernel(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc, uint32_t *__restrict__ cnt,
                uint32_t *__restrict__ touched, UpdCtrl *__restrict__ ctrl, const uint64_t *__restrict__ boards,
                const float *__restrict__ dw, int64_t m)
{
    constexpr int F = num_feat(N);
    constexpr int MAXC = N;
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t j = t / F;
    const int tab = int(t % F), lane = threadIdx.x & 31;
    const bool on = j < m;
    const float d = on ? __ldg(dw + j) : NAN;
    const bool live = on && (EXACT ? isfinite(d) : !isnan(d));
    if (!__any_sync(FULL, live)) return;
    const FeatSpec spec = feat_spec(N, tab);
    const bool base14 = spec.base == 14;
    const uint32_t key_off = table_offset_rt0<N>(tab);
    uint64_t img[8];
    const uint64_t b0 = on ? __ldg(boards + j) : 0;
    img[0] = base14 ? clamp13(b0) : b0;
    img[1] = flip_h(img[0]);
    img[2] = flip_v(img[0]);
    img[3] = flip_v(img[1]);
#pragma unroll
    for (int s = 0; s < 4; s++) img[4 + s] = transpose(img[s]);
    constexpr int64_t NW = table_offset(N, num_feat(N));
    const int64_t replica_off = int64_t(blockIdx.x % acc_replicas(NW)) * NW;
    const long long qd = (EXACT && live) ? quantize(d) : 0;
    uint32_t idx[8], old[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
            if (k < spec.ncell) {
                const uint32_t cell = uint32_t(img[s] >> (4 * (15 - spec.cell[k]))) & 15u;
                v = base14 ? v * 14u + cell : (v << 4) | cell;
            }
        idx[s] = v;
        bool ld = true;                                   // no lower image of this entry has the same key
#pragma unroll
        for (int o = 0; o < s; o++) ld &= idx[o] != v;
        old[s] = 1u;
        if (!live) continue;
        const uint32_t k = key_off + v;
        if (DIRECT) {
            atomicAdd(w + k, d);
            if (delta) atomicAdd(delta + k, d);
        } else {
            if (EXACT) atomicAdd(reinterpret_cast<unsigned long long *>(acc) + replica_off + k, (unsigned long long)qd);
            else atomicAdd(reinterpret_cast<float *>(acc) + replica_off + k, d);
            if (ld) old[s] = atomicAdd(cnt + k, 1u);
        }
    }
    if (DIRECT) return;
    uint32_t first = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) first |= uint32_t(old[s] == 0u) << s;
    const uint32_t mine = __popc(first);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += up;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (total) {                                          // warp-uniform
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctrl->count, total);
        uint32_t pos = __shfl_sync(FULL, base, 0) + incl - mine;
#pragma unroll
        for (int s = 0; s < 8; s++)
            if ((first >> s) & 1u) touched[pos++] = key_off + idx[s];
    }
}

// ---- apply pass: one thread per touched key, plain loads/stores --------------------------------
// Every load of a key (count, the R accumulator replicas, the old weight) is issued before the first store: stores to
// addresses the compiler cannot tell apart from the next loads would otherwise put one L2 round trip per replica on
// the critical path (8 of them at n <= 5).
template <bool EXACT, bool MEAN, int R>
__device__ __forceinline__ void apply_key(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc,
                                          uint32_t *__restrict__ cnt, uint32_t k, int64_t nw)
{
    const uint32_t c = cnt[k];
    const float wk = w[k];
    const float dk = delta ? delta[k] : 0.0f;
    float u;
    if (EXACT) {
        long long *a = reinterpret_cast<long long *>(acc) + k;
        long long v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = a[r * nw];
        long long qs = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            qs += v[r];
            if (v[r]) a[r * nw] = 0;
        }
        double x = double(qs) / FIX_SCALE;
        if (MEAN) x = x / double(c);
        u = __double2float_rn(x);
    } else {
        float *a = reinterpret_cast<float *>(acc) + k;
        float v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = a[r * nw];
        float fs = 0.0f;
#pragma unroll
        for (int r = 0; r < R; r++)
            if (v[r] != 0.0f) { fs += v[r]; a[r * nw] = 0.0f; }
        u = MEAN ? __fdiv_rn(fs, float(c)) : fs;
    }
    cnt[k] = 0;
    w[k] = __fadd_rn(wk, u);
    if (delta) delta[k] = __fadd_rn(dk, u);
}

template <bool EXACT, bool MEAN>
__global__ void __launch_bounds__(256)
td_apply_kernel(float *__restrict__ w, float *__restrict__ delta, void *__restrict__ acc, uint32_t *__restrict__ cnt,
                const uint32_t *__restrict__ touched, UpdCtrl *__restrict__ ctrl, int64_t nw)
{
    const uint32_t count = *reinterpret_cast<volatile uint32_t *>(&ctrl->count);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        if (acc_replicas(nw) == 8) apply_key<EXACT, MEAN, 8>(w, delta, acc, cnt, touched[t], nw);
        else apply_key<EXACT, MEAN, 2>(w, delta, acc, cnt, touched[t], nw);
    }
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

struct LutShared {
    const uint16_t *row;
    const uint8_t *code;
    // the entry in b2048_lut_build's format, rebuilt from the two shared-memory tables
    __device__ __forceinline__ uint32_t operator()(uint32_t line) const
    {
        const uint32_t r = row[line], c = code[line];
        uint32_t t = c & (c >> 1);
        t &= t >> 2;                                       // bit 0 / 4: that merge exponent is 15 (a 2^16 would appear)
        const uint32_t ovf = (t & 0x11u) ? 1u : 0u;
        return r | (c << 16) | (uint32_t((r != line) | ovf) << 24) | (ovf << 25);
    }
};

constexpr int LUT_SMEM_BYTES = 65536 * 2 + 65536 + 256 * 4;        // lines u16, merge codes u8, score of a code byte u32

__device__ __forceinline__ void stage_lut_shared(const uint32_t *__restrict__ lut, unsigned char *smem)
{
    uint16_t *srow = reinterpret_cast<uint16_t *>(smem);
    uint8_t *scode = smem + 65536 * 2;
    for (int q = threadIdx.x; q < 65536 / 4; q += blockDim.x) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(lut) + q);
        reinterpret_cast<uint2 *>(srow)[q] = make_uint2((e.x & 0xFFFFu) | (e.y << 16), (e.z & 0xFFFFu) | (e.w << 16));
        reinterpret_cast<uint32_t *>(scode)[q] = ((e.x >> 16) & 0xFFu) | (((e.y >> 16) & 0xFFu) << 8) |
                                                 (((e.z >> 16) & 0xFFu) << 16) | (((e.w >> 16) & 0xFFu) << 24);
    }
    // merge score of a code byte (exponents a | b << 4): 2^(a+1) + 2^(b+1), 0 for "no merge"
    uint32_t *sscore = reinterpret_cast<uint32_t *>(smem + 65536 * 3);
    for (int q = threadIdx.x; q < 256; q += blockDim.x) sscore[q] = ((2u << (q & 15)) & ~2u) + ((2u << (q >> 4)) & ~2u);
    __syncthreads();
}

// One afterstate per lane (lane d of a 4-lane group = direction d), value by n-tuple gather, then a
// width-4 shuffle argmax with the reference's tie rule (strict '>' scanning d = 0..3: lowest d wins).
// Split in two so that the persistent trainer can do the weight-independent half (LUT move, table indices)
// while it waits at the grid barrier that precedes the next lock-step.
template <int N>
struct MovePrep {
    uint64_t after;
    uint32_t gain, fl;
    uint32_t idx[num_feat(N)];
};

template <int N, class LutT>
__device__ __forceinline__ void move_prepare(const LutT &L, uint64_t board, int d, MovePrep<N> &p)
{
    p.gain = 0;
    p.fl = 0;
    p.after = move_dir(L, board, d, p.gain, p.fl);
    feature_indices<N>(p.after, p.idx);
}

template <int N, bool COHERENT = false>
__device__ __forceinline__ void best_move_finish(const float *__restrict__ w, const MovePrep<N> &p, int d, bool run,
                                                 uint64_t &best_after, uint32_t &best_gain, float &best_value,
                                                 int &best_dir, uint32_t &best_flags, uint32_t &n_valid)
{
    const bool valid = run && (p.fl & 1u);
    float v = valid ? gather_sum<N, COHERENT>(w, p.idx) : -INFINITY;
    // a direction that would create 2^16 is kept valid here; the caller stops the game if it wins
    float bv = v;
    int bd = valid ? d : 4;                           // invalid lanes never win ties
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, off, 4);
        const int od = __shfl_xor_sync(FULL, bd, off, 4);
        const bool take = (od < 4) & ((bd == 4) | (ov > bv) | ((ov == bv) & (od < bd)));   // selects, no divergent branch
        bv = take ? ov : bv;
        bd = take ? od : bd;
    }
    n_valid = __popc(__ballot_sync(FULL, valid) >> ((threadIdx.x & 31) & ~3) & 0xFu);
    const int src = bd & 3;
    best_after = shfl64(p.after, src, 4);
    best_gain = __shfl_sync(FULL, p.gain, src, 4);
    best_flags = __shfl_sync(FULL, p.fl, src, 4);
    best_value = bv;
    best_dir = bd;
}

template <int N, bool COHERENT = false, class LutT = LutGlobal>
__device__ __forceinline__ void best_move(const float *__restrict__ w, const LutT &L, uint64_t board, int d,
                                          bool run, uint64_t &best_after, uint32_t &best_gain, float &best_value,
                                          int &best_dir, uint32_t &best_flags, uint32_t &n_valid)
{
    MovePrep<N> p;
    move_prepare<N>(L, board, d, p);
    best_move_finish<N, COHERENT>(w, p, d, run, best_after, best_gain, best_value, best_dir, best_flags, n_valid);
}

#ifndef B2048_GREEDY_MINBLOCKS
#define B2048_GREEDY_MINBLOCKS 6
#endif
// Game.trial_run at depth 0 (game_logic.py:170-183).  4 lanes per game (lane d = direction d).  The grid is
// persistent and the 4-lane groups take game slots from a queue (g.counters[B2048_CTR_QUEUE]): games end after
// very different numbers of moves, and with a fixed slot per group a warp would idle until the longest of its 8
// games is over.  A slot is played until it is DONE or has made max_steps moves in this launch, then written back.
// SMEM_LUT: one CTA of GREEDY_WIDE_THREADS per SM with the row LUT staged in 192 KB of shared memory, instead of several
// 128-thread CTAs reading it through L1 (lab variant: profiles/r02_greedy_lut_placement.txt has the comparison)
constexpr int GREEDY_WIDE_THREADS = 768;
template <int N, bool SMEM_LUT>
__global__ void __launch_bounds__(SMEM_LUT ? GREEDY_WIDE_THREADS : 128, SMEM_LUT ? 1 : B2048_GREEDY_MINBLOCKS)
greedy_play_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, b2048_games_t g, int max_steps,
                   int limit_tile, int step_limit, b2048_replay_t rp, int has_replay, int8_t *__restrict__ trace_dir,
                   float *__restrict__ trace_value, uint16_t *__restrict__ trace_spawn, int64_t trace_len)
{
    extern __shared__ __align__(16) unsigned char greedy_smem[];
    if (SMEM_LUT) stage_lut_shared(lut, greedy_smem);
    const int lane = threadIdx.x & 31, d = lane & 3;
    const unsigned gmask = 0xFu << (lane & ~3);
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_QUEUE);
    using LutT = std::conditional_t<SMEM_LUT, LutShared, LutGlobal>;
    LutT L;
    if constexpr (SMEM_LUT) L = LutShared{reinterpret_cast<const uint16_t *>(greedy_smem), greedy_smem + 65536 * 2};
    else L = LutGlobal{lut};
    int64_t slot = -1;
    uint64_t board = 0, id = 0;
    uint32_t score = 0, odo = 0, flags = B2048_F_DONE;
    bool in = false, run = false, more = true;
    int steps_done = 0;
    uint32_t c_moves = 0, c_evals = 0, c_fin = 0, c_score = 0, c_msum = 0, c_ovf = 0, c_active = 0;
    while (true) {
        // a group whose game is over (or paused, or out of moves for this launch) writes it back and takes the next slot
        while (more && !(run && steps_done < max_steps)) {
            if (in && d == 0) {
                g.board[slot] = board;
                g.score[slot] = score;
                g.moves[slot] = odo;
                g.flags[slot] = uint8_t(flags);
                if (!(flags & B2048_F_DONE)) c_active++;
            }
            unsigned long long nx = 0;
            if (d == 0) nx = atomicAdd(queue, 1ULL);
            nx = (unsigned long long)__shfl_sync(gmask, (long long)nx, lane & ~3);
            slot = int64_t(nx);
            in = slot < g.B;
            more = in;
            board = in ? g.board[slot] : 0;
            score = in ? g.score[slot] : 0;
            odo = in ? g.moves[slot] : 0;
            flags = in ? g.flags[slot] : B2048_F_DONE;
            id = in ? g.game_id[slot] : 0;
            run = in && !(flags & B2048_F_DONE);
            steps_done = 0;
            if (in && !run) in = false;                             // already finished: nothing to write back
        }
        const bool go = run && steps_done < max_steps;
        if (!__any_sync(FULL, go)) break;                           // every group has drained the queue
        run = go;
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
        if (run && has_replay) {                                    // recorded spawns exhausted -> pause (stays not DONE)
            if (int64_t(odo) >= rp.len || __ldg(rp.tile + slot * rp.len + odo) == 0) run = false;
        }
        uint64_t ba;
        uint32_t bg, bf, nv;
        float bv;
        int bd;
        best_move<N, false, LutT>(w, L, board, d, run, ba, bg, bv, bd, bf, nv);
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
                    sp = 0;
                    if (trace_spawn) sp = spawn_apply(board, r.x, r.y);
                    else spawn_apply_nonempty(board, r.x, r.y);       // the search-free form (no record wanted)
                }
                if (d == 0 && trace_spawn && int64_t(odo) <= trace_len) trace_spawn[slot * trace_len + odo - 1] = uint16_t(sp);
                steps_done++;
            }
        }
    }
    if (in && d == 0) {                                             // the game the group still holds
        g.board[slot] = board;
        g.score[slot] = score;
        g.moves[slot] = odo;
        g.flags[slot] = uint8_t(flags);
        if (!(flags & B2048_F_DONE)) c_active++;
    }
    warp_add_counter(g.counters + B2048_CTR_MOVES, c_moves);
    warp_add_counter(g.counters + B2048_CTR_EVALS, c_evals);
    warp_add_counter(g.counters + B2048_CTR_FINISHED, c_fin);
    warp_add_counter(g.counters + B2048_CTR_SCORE_SUM, c_score);
    warp_add_counter(g.counters + B2048_CTR_MOVES_SUM, c_msum);
    warp_add_counter(g.counters + B2048_CTR_OVERFLOW, c_ovf);
    warp_add_counter(g.counters + B2048_CTR_ACTIVE, c_active);
}

// ------------------------------------------------------------------------------------------------
// Small batches (BASELINE configs[0]: 1,000 games): greedy play is then bound by the LATENCY of one move -- LUT move,
// table indices, one L2 round trip for the gather, argmax, spawn -- times the length of the longest game, with most
// of the chip idle.  This kernel spends the idle lanes on looking one move ahead:
//   * 16 lanes per game, lane = (d, e) = (direction of this move, direction of the next one);
//   * the spawn after a move depends only on the afterstate and on Philox words keyed by (game id, move number), so
//     for each of the four candidate afterstates a_d the board b_d = spawn(a_d) of the NEXT move is known before any
//     value is: lane (d, e) computes a_d, b_d, and the second-level afterstate a_de = move(b_d, e);
//   * the gathers of all 4 + 16 afterstates are issued together: ONE L2 round trip decides TWO moves
//     (argmax over d among the lanes (d, 0), then over e inside the winning group);
//   * the Philox words of the following two moves are computed in the shadow of the gathers;
//   * the row LUT sits in shared memory (128 KB of lines + 64 KB of merge codes, staged once per CTA).
// The decisions, sums (table order) and tie rules are those of greedy_play_kernel move for move, so the games are
// bit-identical; 5x the gather traffic is what the otherwise idle L2 pays for halving the latency per move.
// No replay / trace support: the launcher takes greedy_play_kernel for those.
// ------------------------------------------------------------------------------------------------
// slide+merge in direction d from the two shared-memory tables: the afterstate, whether it differs from the board, and
// the four merge-code bytes of its lines (spec_gain / spec_overflow decode them later, off the critical path)
__device__ __forceinline__ uint64_t spec_move(const uint16_t *__restrict__ srow, const uint8_t *__restrict__ scode,
                                              uint64_t b, int d, uint32_t &codes)
{
    // branch-free: the lanes of a warp hold all four directions, so a branch around a transform would diverge anyway
    // and, with one warp per scheduler, cost ~10 cycles of the chain each; masks select instead
    const uint64_t mt = 0 - uint64_t(d & 1), mf = 0 - uint64_t((d >> 1) & 1);
    uint64_t x = b ^ ((b ^ transpose(b)) & mt);
    x ^= (x ^ flip_h(x)) & mf;
    uint64_t out = 0;
    codes = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t line = uint32_t(x >> (48 - 16 * r)) & 0xFFFFu;
        out |= uint64_t(srow[line]) << (48 - 16 * r);
        codes |= uint32_t(scode[line]) << (8 * r);
    }
    out ^= (out ^ flip_h(out)) & mf;
    return out ^ ((out ^ transpose(out)) & mt);
}

__device__ __forceinline__ bool spec_overflow(uint32_t codes)        // some merge exponent is 15: a 2^16 would appear
{
    uint32_t t = codes & (codes >> 1);
    t &= t >> 2;
    return (t & 0x11111111u) != 0;
}

// score += 2^(x+1) per merge (game_logic.py:33): four lookups in the 256-entry score table of the code bytes
__device__ __forceinline__ uint32_t spec_gain(const uint32_t *__restrict__ sscore, uint32_t codes)
{
    return sscore[codes & 0xFFu] + sscore[(codes >> 8) & 0xFFu] + sscore[(codes >> 16) & 0xFFu] + sscore[codes >> 24];
}

constexpr int SPEC_THREADS = 128;

template <int N>
__global__ void __launch_bounds__(SPEC_THREADS, 1)
greedy_spec_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, b2048_games_t g, int max_steps,
                   int limit_tile, int step_limit)
{
    constexpr int F = num_feat(N);
    extern __shared__ __align__(16) unsigned char lut_smem[];
    stage_lut_shared(lut, lut_smem);
    const LutShared L{reinterpret_cast<const uint16_t *>(lut_smem), lut_smem + 65536 * 2};
    const int lane = threadIdx.x & 31, hl = lane & 15, d = hl >> 2, e = hl & 3;
    const int gbase = lane & 16;                                    // first lane of this game's half-warp
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_QUEUE);
    int64_t slot = -1;
    uint64_t board = 0, id = 0;
    uint32_t score = 0, odo = 0, flags = B2048_F_DONE;
    bool in = false, run = false, more = true;
    int steps_done = 0;
    Philox4 p1{}, p2{};                                             // spawn words of moves odo + 1, odo + 2
    uint32_t c_moves = 0, c_evals = 0, c_fin = 0, c_score = 0, c_msum = 0, c_ovf = 0, c_active = 0;
    while (true) {
        while (more && !(run && steps_done < max_steps)) {
            if (in && hl == 0) {
                g.board[slot] = board;
                g.score[slot] = score;
                g.moves[slot] = odo;
                g.flags[slot] = uint8_t(flags);
                if (!(flags & B2048_F_DONE)) c_active++;
            }
            unsigned long long nx = 0;
            if (hl == 0) nx = atomicAdd(queue, 1ULL);
            nx = (unsigned long long)__shfl_sync(0xFFFFu << gbase, (long long)nx, gbase);
            slot = int64_t(nx);
            in = slot < g.B;
            more = in;
            board = in ? g.board[slot] : 0;
            score = in ? g.score[slot] : 0;
            odo = in ? g.moves[slot] : 0;
            flags = in ? g.flags[slot] : B2048_F_DONE;
            id = in ? g.game_id[slot] : 0;
            run = in && !(flags & B2048_F_DONE);
            steps_done = 0;
            if (in && !run) in = false;
            if (run) {
                p1 = spawn_words(g.seed, id, odo + 1u, 0u);
                p2 = spawn_words(g.seed, id, odo + 2u, 0u);
            }
        }
        const bool go = run && steps_done < max_steps;
        if (!__any_sync(FULL, go)) break;
        run = go;
        // "game over" is not tested up front: a board is over iff no direction changes it, i.e. iff the argmax below
        // finds no candidate -- the same for the board after the first move and its second-level moves
        const bool limit_hit = run && ((limit_tile && max_tile(board) >= limit_tile) || int(odo) >= step_limit);
        // ---- first level (everything below up to the commits is executed by all 32 lanes: shuffles inside).  The order
        //      of the code is the order of issue: the first-level gathers leave before the second level is even computed,
        //      and the first-level argmax runs while the second-level gathers are in flight.
        uint32_t codes1 = 0, codes2 = 0;
        const uint64_t a1 = spec_move(L.row, L.code, board, d, codes1);
        const bool valid1 = run && !limit_hit && (a1 != board || spec_overflow(codes1));
        const bool need1 = valid1 && e == 0;                           // lane (d, 0) scores a_d
        // (the gathers are unconditional: a lane without a legal afterstate reads entry 0 of every table and its sum is
        //  ignored -- a predicated load would put a divergent branch around each group of loads, and with one warp per
        //  scheduler every branch costs ~10 cycles of the chain)
        uint32_t idx1[F];
        feature_indices<N>(need1 ? a1 : 0ULL, idx1);
        float v1[F];
        for_each_feature<N>([&](auto I) {
            constexpr int i = decltype(I)::value;
            v1[i] = __ldg(w + table_offset(N, i) + idx1[i]);
        });
        // ---- second level: the board after the spawn that WOULD follow a_d, and its move in direction e
        const bool ovf1 = spec_overflow(codes1);
        uint64_t b1 = a1;
        spawn_apply_nonempty(b1, p1.x, p1.y);                           // (only used where the first move is legal)
        // the second move exists only if the first one is legal, the game may go on after it and this launch may play it
        const bool stop2 = !valid1 || ovf1 || (limit_tile && max_tile(b1) >= limit_tile) ||
                           int(odo + 1u) >= step_limit || steps_done + 1 >= max_steps;
        const uint64_t a2 = spec_move(L.row, L.code, b1, e, codes2);
        const bool valid2 = !stop2 && (a2 != b1 || spec_overflow(codes2));
        uint32_t idx2[F];
        feature_indices<N>(valid2 ? a2 : 0ULL, idx2);
        float v2[F];
        for_each_feature<N>([&](auto I) {
            constexpr int i = decltype(I)::value;
            v2[i] = __ldg(w + table_offset(N, i) + idx2[i]);
        });
        // ---- in the shadow of the gathers: the spawn words of the two moves after these, the merge scores
        const Philox4 p3 = spawn_words(g.seed, id, odo + 3u, 0u), p4 = spawn_words(g.seed, id, odo + 4u, 0u);
        const uint32_t *sscore = reinterpret_cast<const uint32_t *>(lut_smem + 65536 * 3);
        const uint32_t g1 = spec_gain(sscore, codes1), g2 = spec_gain(sscore, codes2);
        const uint32_t f1 = ovf1 ? 2u : 0u, f2 = spec_overflow(codes2) ? 2u : 0u;
        // ---- first move: argmax over d among the lanes (d, 0) (xor 4, xor 8), then everybody reads lane (0, 0)
        float s1 = 0.0f;
#pragma unroll
        for (int i = 0; i < F; i++) s1 = __fadd_rn(s1, v1[i]);
        float bv = need1 ? s1 : -INFINITY;
        int bd = need1 ? d : 4;
#pragma unroll
        for (int off = 4; off < 16; off <<= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, off, 16);
            const int od = __shfl_xor_sync(FULL, bd, off, 16);
            const bool take = (od < 4) & ((bd == 4) | (ov > bv) | ((ov == bv) & (od < bd)));     // no short circuit
            bv = take ? ov : bv;
            bd = take ? od : bd;
        }
        bd = __shfl_sync(FULL, bd, 0, 16);
        const uint32_t nv1 = __popc((__ballot_sync(FULL, need1) >> gbase) & 0xFFFFu);
        const int src1 = (bd & 3) << 2;                                // lane (bd, 0) of the half-warp
        const uint64_t w_b1 = shfl64(b1, src1, 16);
        const uint32_t w_g1 = __shfl_sync(FULL, g1, src1, 16), w_f1 = __shfl_sync(FULL, f1, src1, 16);
        const bool w_stop2 = __shfl_sync(FULL, int(stop2), src1, 16) != 0;
        float s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < F; i++) s2 = __fadd_rn(s2, v2[i]);
        // ---- second move: argmax over e inside the winning group (every lane reads the lane (bd, its own e))
        float cv = __shfl_sync(FULL, valid2 ? s2 : -INFINITY, src1 + e, 16);
        int ce = __shfl_sync(FULL, valid2 ? e : 4, src1 + e, 16);
#pragma unroll
        for (int off = 1; off < 4; off <<= 1) {
            const float ov = __shfl_xor_sync(FULL, cv, off, 16);
            const int oe = __shfl_xor_sync(FULL, ce, off, 16);
            const bool take = (oe < 4) & ((ce == 4) | (ov > cv) | ((ov == cv) & (oe < ce)));
            cv = take ? ov : cv;
            ce = take ? oe : ce;
        }
        const uint32_t nv2 = __popc((__ballot_sync(FULL, valid2) >> (gbase + src1)) & 0xFu);
        const int src2 = src1 + (ce & 3);
        uint64_t w_a2 = shfl64(a2, src2, 16);
        const uint32_t w_g2 = __shfl_sync(FULL, g2, src2, 16), w_f2 = __shfl_sync(FULL, f2, src2, 16);
        int advanced = 0;
        if (run && bd == 4) {                                       // no legal move (game over), limit tile or step limit
            flags |= B2048_F_DONE;
            run = false;
            if (hl == 0) {
                c_fin++; c_score += score; c_msum += odo;
                atomicAdd(g.tile_hist + max_tile(board), 1u);
                log_finished(g, id, score, odo, max_tile(board), board);
            }
        }
        if (run) {
            if (w_f1 & 2u) {                                        // the first move would create 2^16: flag + stop
                flags |= B2048_F_DONE | B2048_F_OVERFLOW;
                run = false;
                if (hl == 0) {
                    c_fin++; c_score += score; c_msum += odo; c_ovf++;
                    atomicAdd(g.tile_hist + 16, 1u);
                    log_finished(g, id, score, odo, 16u, board);
                }
            } else {
                if (hl == 0) { c_moves++; c_evals += nv1; }
                board = w_b1;
                score += w_g1;
                odo++;
                steps_done++;
                advanced = 1;
                if (!w_stop2 && ce != 4) {                          // (ce == 4: the board after the first move is over)
                    if (w_f2 & 2u) {
                        flags |= B2048_F_DONE | B2048_F_OVERFLOW;
                        run = false;
                        if (hl == 0) {
                            c_fin++; c_score += score; c_msum += odo; c_ovf++;
                            atomicAdd(g.tile_hist + 16, 1u);
                            log_finished(g, id, score, odo, 16u, board);
                        }
                    } else {
                        if (hl == 0) { c_moves++; c_evals += nv2; }
                        spawn_apply_nonempty(w_a2, p2.x, p2.y);
                        board = w_a2;
                        score += w_g2;
                        odo++;
                        steps_done++;
                        advanced = 2;
                    }
                }
            }
        }
        if (advanced == 2) { p1 = p3; p2 = p4; }
        else if (advanced == 1) { p1 = p2; p2 = p3; }
    }
    if (in && hl == 0) {
        g.board[slot] = board;
        g.score[slot] = score;
        g.moves[slot] = odo;
        g.flags[slot] = uint8_t(flags);
        if (!(flags & B2048_F_DONE)) c_active++;
    }
    warp_add_counter(g.counters + B2048_CTR_MOVES, c_moves);
    warp_add_counter(g.counters + B2048_CTR_EVALS, c_evals);
    warp_add_counter(g.counters + B2048_CTR_FINISHED, c_fin);
    warp_add_counter(g.counters + B2048_CTR_SCORE_SUM, c_score);
    warp_add_counter(g.counters + B2048_CTR_MOVES_SUM, c_msum);
    warp_add_counter(g.counters + B2048_CTR_OVERFLOW, c_ovf);
    warp_add_counter(g.counters + B2048_CTR_ACTIVE, c_active);
}

// ------------------------------------------------------------------------------------------------
// sampled expectimax above the estimator: Game.look_forward (game_logic.py:214-243), estimator = evaluate.
//   depth == 0 or empty_count >= since_empty  -> evaluate(afterstate)                           (:215-219)
//   else sample min(width, empty) empty cells without replacement and a 2/4 tile for each      (:220-226),
//        value = mean over them of max(0, game over ? -100 : max over changed directions of
//        look_forward(afterstate', depth - 1))                                                 (:229-242)
// The reference draws from Python's `random`; the device uses node-keyed Philox words (oracle_impl.h states the
// spec: path code, purposes 2 / 3, the j-th position = the mulhi(word_j, left)-th remaining empty cell in row-major
// order).  A direction that would create a 2^16 tile is skipped (reference undefined there).
// Parallel shape: 16 lanes per root afterstate = (sampled tile j, direction d) of the first level; every lane walks
// its subtree depth-first (template recursion on the remaining depth, one non-inlined body per level).
// ------------------------------------------------------------------------------------------------
struct LfParams {
    uint64_t seed, id;
    uint32_t move_no;
    int width, since_empty;
    uint32_t *evals;             // per-thread count of evaluate() calls (may be NULL)
};

template <int N>
__device__ __forceinline__ float lf_leaf(const float *__restrict__ w, uint64_t b, const LfParams &P)
{
    if (P.evals) ++*P.evals;
    return evaluate<N>(w, b);
}

template <int N, int D>
__device__ __noinline__ float lf_node(const float *__restrict__ w, const uint32_t *__restrict__ lut, uint64_t b,
                                      uint32_t path, LfParams P)
{
    if constexpr (D == 0) {
        return lf_leaf<N>(w, b, P);
    } else {
        const uint64_t z = zero_nibbles(b);
        const int empty = popc64(z);
        if (empty >= P.since_empty) return lf_leaf<N>(w, b, P);
        const int num = P.width < empty ? P.width : empty;
        const Philox4 wp = spawn_words(P.seed, P.id, P.move_no, 2u | (path << 8));
        const Philox4 wt = spawn_words(P.seed, P.id, P.move_no, 3u | (path << 8));
        const uint32_t rp[4] = {wp.x, wp.y, wp.z, wp.w}, rt[4] = {wt.x, wt.y, wt.z, wt.w};
        LutGlobal L{lut};
        uint64_t left_mask = z;
        float average = 0.0f;
#pragma unroll 1
        for (int j = 0; j < num; j++) {
            const int sh = kth_empty_shift(left_mask, int(umulhi32(rp[j], uint32_t(empty - j))));
            left_mask &= ~(1ULL << sh);                                     // without replacement
            const uint64_t nb = b | (uint64_t(umulhi32(rt[j], 10u) == 0 ? 2u : 1u) << sh);
            float best;
            if (game_over(nb)) {
                best = -100.0f;
            } else {
                best = -INFINITY;
#pragma unroll 1
                for (int d = 0; d < 4; d++) {
                    uint32_t gain, fl;
                    const uint64_t a = move_dir(L, nb, d, gain, fl);
                    if ((fl & 3u) == 1u) {
                        const float v = lf_node<N, D - 1>(w, lut, a, path * 16u + 4u * uint32_t(j) + uint32_t(d), P);
                        if (v > best) best = v;
                    }
                }
            }
            average = __fadd_rn(average, best > 0.0f ? best : 0.0f);
        }
        return __fdiv_rn(average, float(num));
    }
}

template <int N>
__device__ __forceinline__ float lf_dispatch(const float *__restrict__ w, const uint32_t *__restrict__ lut, uint64_t b,
                                             int depth, uint32_t path, const LfParams &P)
{
    switch (depth) {
    case 0: return lf_node<N, 0>(w, lut, b, path, P);
    case 1: return lf_node<N, 1>(w, lut, b, path, P);
    case 2: return lf_node<N, 2>(w, lut, b, path, P);
    default: return lf_node<N, 3>(w, lut, b, path, P);
    }
}

constexpr int LF_MAX_DEPTH = 4;

// look_forward value of one afterstate, computed by the 16 lanes `hmask` of a half-warp together (all of them
// call with the same arguments; hl = lane index inside the half).  1 <= depth <= LF_MAX_DEPTH or 0.
template <int N>
__device__ __forceinline__ float lf_top(const float *__restrict__ w, const uint32_t *__restrict__ lut, uint64_t b, int depth,
                                        uint32_t path, const LfParams &P, int hl, unsigned hmask)
{
    const uint64_t z = zero_nibbles(b);
    const int empty = popc64(z);
    if (depth == 0 || empty >= P.since_empty) {
        if (P.evals && hl == 0) ++*P.evals;               // the 16 lanes compute the same value: counted once
        return evaluate<N>(w, b);
    }
    const int num = P.width < empty ? P.width : empty;
    const Philox4 wp = spawn_words(P.seed, P.id, P.move_no, 2u | (path << 8));
    const Philox4 wt = spawn_words(P.seed, P.id, P.move_no, 3u | (path << 8));
    const uint32_t rp[4] = {wp.x, wp.y, wp.z, wp.w}, rt[4] = {wt.x, wt.y, wt.z, wt.w};
    const int j = hl >> 2, d = hl & 3;
    uint64_t left_mask = z, nb = b;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (q < num) {
            const int sh = kth_empty_shift(left_mask, int(umulhi32(rp[q], uint32_t(empty - q))));
            left_mask &= ~(1ULL << sh);
            if (q == j) nb = b | (uint64_t(umulhi32(rt[q], 10u) == 0 ? 2u : 1u) << sh);
        }
    }
    const bool active = j < num;
    const bool over = active && game_over(nb);
    LutGlobal L{lut};
    uint32_t gain, fl;
    const uint64_t a = move_dir(L, nb, d, gain, fl);
    float best = -INFINITY;
    if (active && !over && (fl & 3u) == 1u) best = lf_dispatch<N>(w, lut, a, depth - 1, path * 16u + 4u * uint32_t(j) + uint32_t(d), P);
    __syncwarp(hmask);
    best = fmaxf(best, __shfl_xor_sync(hmask, best, 1, 16));
    best = fmaxf(best, __shfl_xor_sync(hmask, best, 2, 16));
    if (over) best = -100.0f;
    const float c = best > 0.0f ? best : 0.0f;
    float average = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float cq = __shfl_sync(hmask, c, 4 * q, 16);
        if (q < num) average = __fadd_rn(average, cq);
    }
    return __fdiv_rn(average, float(num));
}

// The same value for depth >= 2 by a whole warp: the two upper levels of the tree are laid out explicitly -- 16
// level-1 nodes (tile j, direction d) and up to 256 level-2 afterstates (j, d, j2, d2).  The valid ones are compacted
// with a ballot per 32 items and shared evenly by the 32 lanes, each walking the subtree below its items
// depth-first.  With 16 items of very different size only (lf_top) a third of the lanes were active per instruction;
// without the compaction the lanes holding invalid items idled.  Arena: per-warp shared memory.
struct LfArena {
    uint64_t board[16];      // level-1 afterstates
    uint32_t pos2[16];       // their 4 sampled cells (6-bit shifts)
    uint32_t meta[16];       // bit 0-1 state (0 invalid, 1 evaluate directly, 2 expand), bits 4-6 num2, 8-11 tile2 ("4" bits),
                             // 12-15 game over after spawn j2
    float value1[16];        // direct values
    float value2[256];       // level-2 values
    // the valid level-2 afterstates, compacted (item = (j, d, j2, d2) index), so that the lanes share them evenly
    uint64_t a2[256];
    uint8_t item[256];
    // depth 3: the leaves below 8 level-2 afterstates at a time, compacted again before the gathers
    uint64_t leaf[128];
    float lval[128];         // value per (item in chunk, j3, d3)
    uint32_t pos3[32], meta3[32];   // sampled for 32 afterstates at a time (all lanes busy)
    uint8_t lslot[128];
};

// BATCH: at depth 3 lay the third level out as well and run the gathers on compacted leaves (pays when the SMs are
// full of games: +18 % at 16,384 games; with one warp per scheduler the extra round trips per chunk cost 4-8 %)
template <int N, bool BATCH>
__device__ __forceinline__ float lf_top2(const float *__restrict__ w, const uint32_t *__restrict__ lut, uint64_t b, int depth,
                                         uint32_t path, const LfParams &P, int lane, LfArena &A)
{
    const uint64_t z = zero_nibbles(b);
    const int empty = popc64(z);
    if (empty >= P.since_empty) {
        if (P.evals && lane == 0) ++*P.evals;
        return evaluate<N>(w, b);
    }
    const int num = P.width < empty ? P.width : empty;
    LutGlobal L{lut};
    // ---- level 1: lanes 0..15 = (j, d)
    const int j = (lane >> 2) & 3, d = lane & 3;
    bool over1 = false;
    if (lane < 16) {
        const Philox4 wp = spawn_words(P.seed, P.id, P.move_no, 2u | (path << 8));
        const Philox4 wt = spawn_words(P.seed, P.id, P.move_no, 3u | (path << 8));
        const uint32_t rp[4] = {wp.x, wp.y, wp.z, wp.w}, rt[4] = {wt.x, wt.y, wt.z, wt.w};
        uint64_t left_mask = z, nb = b;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (q < num) {
                const int sh = kth_empty_shift(left_mask, int(umulhi32(rp[q], uint32_t(empty - q))));
                left_mask &= ~(1ULL << sh);
                if (q == j) nb = b | (uint64_t(umulhi32(rt[q], 10u) == 0 ? 2u : 1u) << sh);
            }
        const bool active = j < num;
        over1 = active && game_over(nb);
        uint32_t gain, fl;
        const uint64_t a1 = move_dir(L, nb, d, gain, fl);
        uint32_t meta = 0, pos2 = 0;
        if (active && !over1 && (fl & 3u) == 1u) {
            const uint64_t z1 = zero_nibbles(a1);
            const int empty1 = popc64(z1);
            if (empty1 >= P.since_empty) {                      // depth - 1 >= 1 here
                meta = 1u;
                A.value1[lane] = lf_leaf<N>(w, a1, P);
            } else {
                const uint32_t path1 = path * 16u + 4u * uint32_t(j) + uint32_t(d);
                const int num2 = P.width < empty1 ? P.width : empty1;
                const Philox4 vp = spawn_words(P.seed, P.id, P.move_no, 2u | (path1 << 8));
                const Philox4 vt = spawn_words(P.seed, P.id, P.move_no, 3u | (path1 << 8));
                const uint32_t sp[4] = {vp.x, vp.y, vp.z, vp.w}, st[4] = {vt.x, vt.y, vt.z, vt.w};
                uint64_t lm = z1;
                meta = 2u | (uint32_t(num2) << 4);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q < num2) {
                        const int sh = kth_empty_shift(lm, int(umulhi32(sp[q], uint32_t(empty1 - q))));
                        lm &= ~(1ULL << sh);
                        const uint32_t four = umulhi32(st[q], 10u) == 0 ? 1u : 0u;
                        pos2 |= uint32_t(sh) << (6 * q);
                        meta |= four << (8 + q);
                        if (game_over(a1 | (uint64_t(1u + four) << sh))) meta |= 1u << (12 + q);
                    }
            }
        }
        A.board[lane] = a1;
        A.pos2[lane] = pos2;
        A.meta[lane] = meta;
    }
    __syncwarp();
    // ---- level 2: the valid ones of the 256 items (j, d, j2, d2), compacted with a ballot per 32 items
    const unsigned lt_mask = (1u << lane) - 1u;
    int nvalid = 0;
#pragma unroll 1
    for (int it = lane; it < 256; it += 32) {
        const int l = it >> 4, j2 = (it >> 2) & 3, d2 = it & 3;
        const uint32_t meta = A.meta[l];
        bool ok = false;
        uint64_t a2 = 0;
        if ((meta & 3u) == 2u && j2 < int((meta >> 4) & 7u) && !((meta >> (12 + j2)) & 1u)) {
            const int sh = int((A.pos2[l] >> (6 * j2)) & 63u);
            const uint64_t nb2 = A.board[l] | (uint64_t(1u + ((meta >> (8 + j2)) & 1u)) << sh);
            uint32_t gain, fl;
            a2 = move_dir(L, nb2, d2, gain, fl);
            ok = (fl & 3u) == 1u;
        }
        A.value2[it] = -INFINITY;
        const unsigned m = __ballot_sync(FULL, ok);
        if (ok) {
            const int pos = nvalid + __popc(m & lt_mask);
            A.a2[pos] = a2;
            A.item[pos] = uint8_t(it);
        }
        nvalid += __popc(m);
    }
    __syncwarp();
    if (!BATCH || depth != 3) {
        // each lane walks the subtrees below its share of the valid afterstates depth-first (depth 2: they are leaves)
#pragma unroll 1
        for (int q = lane; q < nvalid; q += 32) {
            const uint32_t it = A.item[q];
            const uint32_t path2 = (path * 16u + (it >> 4)) * 16u + (it & 15u);
            A.value2[it] = lf_dispatch<N>(w, lut, A.a2[q], depth - 2, path2, P);
        }
    } else {
        // depth 3: one more level laid out: (a) every lane samples the tiles of one afterstate (32 at a time: the two
        // Philox calls per node are a third of the instructions of this branch); then 8 afterstates (<= 128 leaves) at
        // a time: (b) all lanes make the candidate leaves (item, j3, d3) and compact the valid ones, (c) the gathers run
        // on the compacted leaves, (d) lanes 0..7 back the values up (lf_node<N, 1>'s arithmetic).
#pragma unroll 1
        for (int qs = 0; qs < nvalid; qs += 32) {
            {
                uint32_t meta = 0, pos3 = 0;
                if (qs + lane < nvalid) {
                    const uint64_t a2 = A.a2[qs + lane];
                    const uint64_t z2 = zero_nibbles(a2);
                    const int empty2 = popc64(z2);
                    if (empty2 >= P.since_empty) {
                        meta = 1u;                                         // the afterstate itself is the leaf
                    } else {
                        const uint32_t it = A.item[qs + lane];
                        const uint32_t path2 = (path * 16u + (it >> 4)) * 16u + (it & 15u);
                        const int num3 = P.width < empty2 ? P.width : empty2;
                        const Philox4 vp = spawn_words(P.seed, P.id, P.move_no, 2u | (path2 << 8));
                        const Philox4 vt = spawn_words(P.seed, P.id, P.move_no, 3u | (path2 << 8));
                        const uint32_t sp[4] = {vp.x, vp.y, vp.z, vp.w}, st[4] = {vt.x, vt.y, vt.z, vt.w};
                        uint64_t lm = z2;
                        meta = 2u | (uint32_t(num3) << 4);
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (q < num3) {
                                const int sh = kth_empty_shift(lm, int(umulhi32(sp[q], uint32_t(empty2 - q))));
                                lm &= ~(1ULL << sh);
                                const uint32_t four = umulhi32(st[q], 10u) == 0 ? 1u : 0u;
                                pos3 |= uint32_t(sh) << (6 * q);
                                meta |= four << (8 + q);
                                if (game_over(a2 | (uint64_t(1u + four) << sh))) meta |= 1u << (12 + q);
                            }
                    }
                }
                A.pos3[lane] = pos3;
                A.meta3[lane] = meta;
            }
            __syncwarp();
#pragma unroll 1
            for (int q0 = qs; q0 < nvalid && q0 < qs + 32; q0 += 8) {
            const int m0 = q0 - qs;                                        // first pos3 / meta3 entry of this group of 8
            int nleaf = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int slot = lane + 32 * k, ql = slot >> 4, j3 = (slot >> 2) & 3, d3 = slot & 3;
                const uint32_t meta = A.meta3[m0 + ql];
                bool ok = false;
                uint64_t a3 = 0;
                if ((meta & 3u) == 1u) {
                    ok = (slot & 15) == 0;
                    a3 = A.a2[q0 + ql];
                } else if ((meta & 3u) == 2u && j3 < int((meta >> 4) & 7u) && !((meta >> (12 + j3)) & 1u)) {
                    const int sh = int((A.pos3[m0 + ql] >> (6 * j3)) & 63u);
                    const uint64_t nb3 = A.a2[q0 + ql] | (uint64_t(1u + ((meta >> (8 + j3)) & 1u)) << sh);
                    uint32_t gain, fl;
                    a3 = move_dir(L, nb3, d3, gain, fl);
                    ok = (fl & 3u) == 1u;
                }
                A.lval[slot] = -INFINITY;
                const unsigned m = __ballot_sync(FULL, ok);
                if (ok) {
                    const int pos = nleaf + __popc(m & lt_mask);
                    A.leaf[pos] = a3;
                    A.lslot[pos] = uint8_t(slot);
                }
                nleaf += __popc(m);
            }
            __syncwarp();
#pragma unroll 1
            for (int p = lane; p < nleaf; p += 32) A.lval[A.lslot[p]] = lf_leaf<N>(w, A.leaf[p], P);
            __syncwarp();
            if (lane < 8 && q0 + lane < nvalid) {
                const uint32_t meta = A.meta3[m0 + lane];
                float v;
                if ((meta & 3u) == 1u) {
                    v = A.lval[lane * 16];
                } else {
                    const int num3 = int((meta >> 4) & 7u);
                    float avg = 0.0f;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q < num3) {
                            float bq = fmaxf(fmaxf(A.lval[lane * 16 + 4 * q], A.lval[lane * 16 + 4 * q + 1]),
                                             fmaxf(A.lval[lane * 16 + 4 * q + 2], A.lval[lane * 16 + 4 * q + 3]));
                            if ((meta >> (12 + q)) & 1u) bq = -100.0f;
                            avg = __fadd_rn(avg, bq > 0.0f ? bq : 0.0f);
                        }
                    v = __fdiv_rn(avg, float(num3));
                }
                A.value2[A.item[q0 + lane]] = v;
            }
            __syncwarp();
            }
        }
    }
    __syncwarp();
    // ---- back up: level 2 -> level 1 (lanes 0..15), level 1 -> root
    float best = -INFINITY;
    if (lane < 16) {
        const uint32_t meta = A.meta[lane];
        if ((meta & 3u) == 1u) {
            best = A.value1[lane];
        } else if ((meta & 3u) == 2u) {
            const int num2 = int((meta >> 4) & 7u);
            float avg = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < num2) {
                    float bq = fmaxf(fmaxf(A.value2[lane * 16 + 4 * q], A.value2[lane * 16 + 4 * q + 1]),
                                     fmaxf(A.value2[lane * 16 + 4 * q + 2], A.value2[lane * 16 + 4 * q + 3]));
                    if ((meta >> (12 + q)) & 1u) bq = -100.0f;
                    avg = __fadd_rn(avg, bq > 0.0f ? bq : 0.0f);
                }
            best = __fdiv_rn(avg, float(num2));
        }
    }
    best = fmaxf(best, __shfl_xor_sync(FULL, best, 1));
    best = fmaxf(best, __shfl_xor_sync(FULL, best, 2));
    if (over1) best = -100.0f;
    const float c = best > 0.0f ? best : 0.0f;
    float average = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float cq = __shfl_sync(FULL, c, 4 * q);
        if (q < num) average = __fadd_rn(average, cq);
    }
    __syncwarp();                                                  // the arena is reused by the next call
    return __fdiv_rn(average, float(num));
}

// m afterstates, one warp each (depth <= 1: the two half-warps compute the same value with lf_top)
template <int N>
__global__ void __launch_bounds__(128)
look_forward_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, const uint64_t *__restrict__ boards,
                    const uint64_t *__restrict__ game_id, const uint32_t *__restrict__ move_no,
                    const uint8_t *__restrict__ root_dir, int64_t m, int depth, int width, int since_empty, uint64_t seed,
                    float *__restrict__ value)
{
    __shared__ LfArena arena[4];
    const int64_t q = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= m) return;                                                   // warp-uniform
    const LfParams P{seed, __ldg(game_id + q), __ldg(move_no + q), width, since_empty, nullptr};
    const uint32_t path = 4u + uint32_t(__ldg(root_dir + q) & 3);
    const float v = depth >= 2 ? lf_top2<N, false>(w, lut, __ldg(boards + q), depth, path, P, lane, arena[threadIdx.x >> 5])
                               : lf_top<N>(w, lut, __ldg(boards + q), depth, path, P, lane & 15, 0xFFFFu << (lane & 16));
    if (lane == 0) value[q] = v;
}

// Game.trial_run with look-ahead (game_logic.py:150-183): one warp per game slot, the two half-warps score root
// directions (0, 1) then (2, 3); strict '>' scanning d = 0..3, commit, Philox spawn.  Whole games per launch.
// MINB: resident CTAs per SM the register allocation must allow.  Few games (a warp each) run fastest with all the
// registers the tree walk wants (166, MINB = 1: 2.0 M moves/s at 1,024 games vs 1.4 M capped); many games want the
// occupancy (MINB = 6, <= 80 registers: 3.7 M moves/s at 8,192 games vs 2.5 M uncapped).
template <int N, int MINB>
__global__ void __launch_bounds__(128, MINB)
expectimax_play_kernel(const float *__restrict__ w, const uint32_t *__restrict__ lut, b2048_games_t g, int max_steps,
                       int limit_tile, int step_limit, int depth, int width, int since_empty,
                       int8_t *__restrict__ trace_dir, uint16_t *__restrict__ trace_spawn, int64_t trace_len)
{
    __shared__ LfArena arena[4];
    const int64_t slot = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const unsigned hmask = 0xFFFFu << (lane & 16);
    const bool in = slot < g.B;
    LutGlobal L{lut};
    uint64_t board = in ? g.board[slot] : 0;
    uint32_t score = in ? g.score[slot] : 0;
    uint32_t odo = in ? g.moves[slot] : 0;
    uint32_t flags = in ? g.flags[slot] : B2048_F_DONE;
    const uint64_t id = in ? g.game_id[slot] : 0;
    bool run = in && !(flags & B2048_F_DONE);
    uint32_t c_moves = 0, c_evals = 0;
    for (int step = 0; step < max_steps && run; step++) {                 // warp-uniform: one game per warp
        if (game_over(board) || (limit_tile && max_tile(board) >= limit_tile) || int(odo) >= step_limit) {
            flags |= B2048_F_DONE;
            break;
        }
        const LfParams P{g.seed, id, odo, width, since_empty, &c_evals};
        float v4[4];
        if (depth >= 2) {                                                 // the whole warp on one root direction at a time
#pragma unroll 1
            for (int rd = 0; rd < 4; rd++) {
                uint32_t gain, fl;
                const uint64_t a = move_dir(L, board, rd, gain, fl);
                v4[rd] = (fl & 3u) == 1u ? lf_top2<N, MINB != 1>(w, lut, a, depth, 4u + uint32_t(rd), P, lane, arena[threadIdx.x >> 5])
                                         : -INFINITY;
            }
        } else {                                                          // half-warps: root directions (0, 1) then (2, 3)
            float val[2];
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const int rd = 2 * pass + half;
                uint32_t gain, fl;
                const uint64_t a = move_dir(L, board, rd, gain, fl);
                val[pass] = (fl & 3u) == 1u ? lf_top<N>(w, lut, a, depth, 4u + uint32_t(rd), P, hl, hmask) : -INFINITY;
                __syncwarp();
            }
            v4[0] = __shfl_sync(FULL, val[0], 0);
            v4[1] = __shfl_sync(FULL, val[0], 16);
            v4[2] = __shfl_sync(FULL, val[1], 0);
            v4[3] = __shfl_sync(FULL, val[1], 16);
        }
        int bd = -1;
        float bv = -INFINITY;
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (v4[d] != -INFINITY || bd < 0) {
                uint32_t gain, fl;
                move_dir(L, board, d, gain, fl);
                if ((fl & 3u) == 1u && (bd < 0 || v4[d] > bv)) { bd = d; bv = v4[d]; }
            }
        if (bd < 0) {                                                     // only 2^16-creating moves left: flag + stop
            flags |= B2048_F_DONE | B2048_F_OVERFLOW;
            break;
        }
        uint32_t gain, fl;
        board = move_dir(L, board, bd, gain, fl);
        score += gain;
        if (lane == 0 && trace_dir && int64_t(odo) < trace_len) trace_dir[slot * trace_len + odo] = int8_t(bd);
        odo++;
        c_moves++;
        const Philox4 r = spawn_words(g.seed, id, odo, 0u);
        const uint32_t sp = spawn_apply(board, r.x, r.y);
        if (lane == 0 && trace_spawn && int64_t(odo) <= trace_len) trace_spawn[slot * trace_len + odo - 1] = uint16_t(sp);
    }
    warp_add_counter(g.counters + B2048_CTR_EVALS, c_evals);
    if (in && lane == 0) {
        g.board[slot] = board;
        g.score[slot] = score;
        g.moves[slot] = odo;
        g.flags[slot] = uint8_t(flags);
        if (c_moves) atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_MOVES), (unsigned long long)c_moves);
        if (run && (flags & B2048_F_DONE)) {
            atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_FINISHED), 1ULL);
            atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_SCORE_SUM), (unsigned long long)score);
            atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_MOVES_SUM), (unsigned long long)odo);
            if (flags & B2048_F_OVERFLOW) atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_OVERFLOW), 1ULL);
            atomicAdd(g.tile_hist + ((flags & B2048_F_OVERFLOW) ? 16 : max_tile(board)), 1u);
            log_finished(g, id, score, odo, (flags & B2048_F_OVERFLOW) ? 16u : uint32_t(max_tile(board)), board);
        }
        if (!(flags & B2048_F_DONE)) atomicAdd(reinterpret_cast<unsigned long long *>(g.counters + B2048_CTR_ACTIVE), 1ULL);
    }
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

// one game slot in registers (b2048_games_t, structure of arrays in memory)
struct SlotState {
    uint64_t board, id, state;
    uint32_t score, odo, flags;
    float old_label;
};

__device__ __forceinline__ SlotState slot_load(const b2048_games_t &g, int64_t slot, bool in)
{
    SlotState s;
    s.board = in ? g.board[slot] : 0;
    s.score = in ? g.score[slot] : 0;
    s.odo = in ? g.moves[slot] : 0;
    s.flags = in ? g.flags[slot] : B2048_F_DONE;
    s.id = in ? g.game_id[slot] : 0;
    s.state = in ? g.state[slot] : 0;
    s.old_label = in ? g.old_label[slot] : 0.0f;
    return s;
}

__device__ __forceinline__ void slot_store(const b2048_games_t &g, int64_t slot, const SlotState &s)
{
    g.board[slot] = s.board;
    g.score[slot] = s.score;
    g.moves[slot] = s.odo;
    g.game_id[slot] = s.id;
    g.state[slot] = s.state;
    g.old_label[slot] = s.old_label;
    g.flags[slot] = uint8_t(s.flags);
}

// One lock-step of one slot (QAgent.episode body, r_learning.py:228-249) on the register copy `s`: returns true if
// the slot was live (its state changed); (ub, dw) = the TD update to apply (dw = NaN: none).  4 lanes per slot
// (lane d = direction d); all 32 lanes of a warp must call this together (width-4 shuffles inside).
struct NoPublish {
    __device__ __forceinline__ void operator()(uint64_t, float) const {}
};

// `publish(ub, dw)` is called by every lane as soon as the update of the step is known (afterstate to update and
// dw; the bookkeeping, the restart and the spawn follow): the persistent kernel hands them to phase B there.
template <int N, bool COHERENT, class Publish = NoPublish>
__device__ __forceinline__ bool phase_a_compute(const float *__restrict__ w, const LutGlobal &L, const b2048_games_t &g,
                                                float alpha, int64_t slot, int d, bool in, SlotState &s, uint64_t &ub,
                                                float &dw, const b2048_replay_t &rp, int has_replay,
                                                int8_t *__restrict__ trace_dir, float *__restrict__ trace_value,
                                                float *__restrict__ trace_dw, uint16_t *__restrict__ trace_spawn,
                                                int64_t trace_len, StepCounters &c, const MovePrep<N> *prep = nullptr,
                                                Publish publish = Publish())
{
    constexpr int F = num_feat(N);
    uint64_t board = s.board;
    uint32_t score = s.score;
    uint32_t odo = s.odo;
    uint32_t flags = s.flags;
    uint64_t id = s.id;
    uint64_t state = s.state;
    float old_label = s.old_label;
    bool run = in && !(flags & B2048_F_DONE);
    if (run && has_replay && !game_over(board)) {
        if (int64_t(odo) >= rp.len || __ldg(rp.tile + slot * rp.len + odo) == 0) run = false;   // spawns exhausted
    }
    const bool over = run && game_over(board);
    uint64_t ba;
    uint32_t bg, bf, nv;
    float bv;
    int bd;
    if (prep) best_move_finish<N, COHERENT>(w, *prep, d, run && !over, ba, bg, bv, bd, bf, nv);   // moves of s.board
    else best_move<N, COHERENT>(w, L, board, d, run && !over, ba, bg, bv, bd, bf, nv);
    dw = NAN;
    ub = 0;
    const bool finished = run && (over || (bf & 2u));
    if (run && (flags & B2048_F_HAVE_STATE)) {
        float x = -old_label;                                        // terminal update, r_learning.py:248-249
        if (!finished) {                                             // :238-241
            x = __fadd_rn(float(bg), bv);                            // (best_score - score) + best_value
            x = __fsub_rn(x, old_label);
        }
        dw = __fdiv_rn(__fmul_rn(x, alpha), float(F));
        ub = state;
    }
    publish(ub, dw);
    if (run) {
        if (finished) {
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
                sp = 0;
                if (trace_spawn) sp = spawn_apply(board, r.x, r.y);
                else spawn_apply_nonempty(board, r.x, r.y);
            }
            if (d == 0 && trace_spawn && int64_t(odo) <= trace_len) trace_spawn[slot * trace_len + odo - 1] = uint16_t(sp);
        }
        if (d == 0 && !isnan(dw)) c.upd++;
    }
    if (run) {
        s.board = board;
        s.score = score;
        s.odo = odo;
        s.id = id;
        s.state = state;
        s.old_label = old_label;
        s.flags = flags;
    }
    return run;
}

template <int N, bool COHERENT>
__device__ __forceinline__ void phase_a_slot(const float *__restrict__ w, const LutGlobal &L, const b2048_games_t &g,
                                             float alpha, int64_t slot, int d, bool in, uint64_t *__restrict__ upd_board,
                                             float *__restrict__ upd_dw, const b2048_replay_t &rp, int has_replay,
                                             int8_t *__restrict__ trace_dir, float *__restrict__ trace_value,
                                             float *__restrict__ trace_dw, uint16_t *__restrict__ trace_spawn,
                                             int64_t trace_len, StepCounters &c)
{
    SlotState s = slot_load(g, slot, in);
    uint64_t ub;
    float dw;
    const bool run = phase_a_compute<N, COHERENT>(w, L, g, alpha, slot, d, in, s, ub, dw, rp, has_replay, trace_dir,
                                                  trace_value, trace_dw, trace_spawn, trace_len, c);
    if (in && d == 0) {
        upd_board[slot] = ub;
        upd_dw[slot] = dw;
        if (run) slot_store(g, slot, s);
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
//     phase A  own slots, weights read through L2 (ld.cg)             -> upd_board / upd_dw
//     phase B  own entries, one thread per (entry, table): 8 D4 images -> accumulators
//     flush    the CTA's shared-memory table of small-exponent keys   -> dense global hot table (RED)
//     ---- grid barrier ----
//     apply    the keys this CTA touched FIRST + its slice of the hot table: w += S [/ G], accumulators := 0
//     ---- grid barrier ----
// Phase B writes only accumulators, never w, so no barrier is needed between A and B of different CTAs; the
// DIRECT variant (atomic + sum rule, increments straight into w) needs one there instead.
//
// Accumulators (one copy, no replicas: the first-touch test must see every contribution to a key):
//   float modes   acc = float2[nw] {sum, count}: ONE returning ATOMG.ADD.F32x2 per contribution
//   exact modes   acc = int64[nw] (RED.64) + cnt = u32[nw] (returning ATOMG), as in the stepwise path
// The contribution that finds count == 0 appends its key to the CTA-private list (shared-memory cursor), so
// the apply phase needs no global counter, no global list and no atomics.
//
// Hot keys.  Keys whose cells are all <= 3 (empty, 2, 4, 8) take ~30 % of all contributions and include every
// really hot address (thousands of games per lock-step on "empty row").  They are dense: 4^cells per table.
// Each CTA accumulates them in shared memory, then adds its non-zero entries to a dense global hot table with
// fire-and-forget REDs (<= one per CTA and key), which the apply phase scans in slices.  No first-touch
// bookkeeping, no same-address pile-up in L2.
//
// No warp-level key matching here: __match_any_sync costs ~12 cycles per distinct value (380 cycles for 32
// distinct keys, measured, profiles/microbench/warpops.cu) and serialises per SM sub-partition.  A thread owns
// one table of one entry and compares its own 8 image keys in registers (an entry counts once per key in G).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of a cooperative launch; `target` is the CTA's running arrival total (thread 0 only).
// Arrival = red.release.gpu (MEMBAR.ALL.GPU + RED: the CTA's earlier writes and atomics are performed first);
// the wait polls with relaxed loads and does NOT invalidate L1 (no CCTL.IVALL): everything another CTA may have
// written is read with ld.cg (L2) afterwards, while the row LUT and the CTA's own game slots stay L1-resident.
__device__ __forceinline__ void grid_arrive(uint32_t *bar, uint32_t &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    }
}

// Besides the formal acquire below, every datum another CTA may have written before the barrier -- weights, accumulators,
// counts, hot table, key lists, staged updates -- is read after it with ld.cg / an atomic, i.e. at L2; ld.ca /
// ld.global.nc are used only for the row LUT, which nobody writes, and for the CTA's own game slots.
// Returns false when the launch is dead: a wait of 2^25 polls (seconds) can only be a lost CTA.  The CTA that times out
// raises g.counters[B2048_CTR_FAULT]; every other CTA sees the flag within 2^12 polls and leaves too, so the launch
// ends instead of hanging the device, and the host finds the fault in the counters it reads anyway (engine raises).
__device__ __forceinline__ bool grid_wait(const uint32_t *bar, uint32_t target, uint64_t *fault)
{
    __shared__ int s_dead;
    if (threadIdx.x == 0) {
        uint32_t polls = 0;
        int dead = 0;
        while (ld_relaxed_gpu(bar) < target) {
            if ((++polls & 0xFFFu) == 0 &&
                (polls > (1u << 25) || *reinterpret_cast<volatile unsigned long long *>(fault) != 0ULL)) {
                dead = 1;
                break;
            }
        }
        if (dead) atomicExch(reinterpret_cast<unsigned long long *>(fault), 1ULL);
        // the acquire of the PTX memory model: one more read of the counter after the relaxed polling, synchronising with
        // the arrivals' red.release.gpu (thread 0 acquires, the bar.sync below extends it to the CTA).  It compiles to
        // LDG.STRONG.GPU + CCTL.IVALL, i.e. it drops the SM's L1 once per barrier; measured cost 1 % at the headline shape
        // (12.02 -> 12.14 us per lock-step), 0.1 % elsewhere (profiles/r02_barrier_acquire.txt).
        asm volatile("{ .reg .u32 seen; ld.acquire.gpu.global.u32 seen, [%0]; }" ::"l"(bar) : "memory");
        s_dead = dead;
    }
    __syncthreads();
    return s_dead == 0;
}

__device__ __forceinline__ bool grid_barrier(uint32_t *bar, uint32_t &target, uint64_t *fault)
{
    grid_arrive(bar, target);
    return grid_wait(bar, target, fault);
}

// Weight exchange with the other GPUs from INSIDE the persistent launch (b2048_td_run_peers), called by every thread of
// every CTA right after the grid barrier that completes W_{t+1}: announce (this rank's lock-steps are done, its weight
// stores are ordered before the signal), wait for every peer's announcement, reduce this rank's slice over peer memory
// with all CTAs (peer_reduce_slice), grid barrier, tell the peers that the stores are out, wait for theirs.  The game
// state stays in registers and the launch goes on with the next lock-step: no relaunch, no separate sync kernel.
__device__ __forceinline__ bool peer_sync_in_kernel(const PeerSync &ps, uint32_t epoch, uint32_t *bar, uint32_t &bar_target,
                                                    uint64_t *fault)
{
    __shared__ int s_ok;
    const b2048_peers_t &P = ps.peers;
    const int W = P.world, rank = P.rank;
    uint32_t *mine = P.flags[rank];
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (blockIdx.x == 0 && int(threadIdx.x) < W) {
        __threadfence_system();
        st_release_sys(P.flags[threadIdx.x] + B2048_PEER_ARRIVE + rank, epoch);
    }
    if (int(threadIdx.x) < W && !wait_epoch(mine + B2048_PEER_ARRIVE + threadIdx.x, epoch)) s_ok = 0;
    __syncthreads();
    if (s_ok) peer_reduce_slice_scalar(P, ps.count, int64_t(blockIdx.x) * blockDim.x + threadIdx.x, int64_t(gridDim.x) * blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();             // the CTA's remote stores are performed (cumulative)
    if (!grid_barrier(bar, bar_target, fault)) return false;  // ... and every other CTA's of this rank
    if (blockIdx.x == 0 && int(threadIdx.x) < W && s_ok) st_release_sys(P.flags[threadIdx.x] + B2048_PEER_DONE + rank, epoch);
    if (int(threadIdx.x) < W && s_ok && !wait_epoch(mine + B2048_PEER_DONE + threadIdx.x, epoch)) s_ok = 0;
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) {
            atomicExch(reinterpret_cast<unsigned long long *>(fault), 1ULL);
            mine[B2048_PEER_FAULT] = epoch;
        }
        return false;
    }
    return true;
}

// ---- dense small-exponent key space --------------------------------------------------------------
// Keys whose cells are all below HL = hot_levels(n) ("small" keys) are accumulated per CTA in shared memory and merged
// through a dense global table.  HL = 5 (cells 0..4, radix-5 index) where the table fits next to everything else
// (n <= 4: 10,625 keys at n = 4), HL = 4 (cells 0..3, two bits per cell) otherwise.  Measured (r02): HL = 5 makes the
// generic layout 5-6 % faster at n <= 4 but the one-round layout of the headline shape 13 % slower (more shared-memory
// CAS atomics and flush REDs than L2 chains saved), so HL = 4 is the default everywhere; B2048_HOT_LEVELS_SMALL_N=5 selects
// the n <= 4 choice (lab builds).
#ifndef B2048_HOT_LEVELS_SMALL_N
#define B2048_HOT_LEVELS_SMALL_N 4
#endif
__host__ __device__ constexpr int hot_levels(int n) { return n <= 4 ? B2048_HOT_LEVELS_SMALL_N : 4; }
__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }
__host__ __device__ constexpr int small_tables(int n) { return n == 6 ? 21 : num_feat(n); }   // base-16 tables only
__host__ __device__ constexpr int tuple_cells(int n, int i) { return n <= 3 ? n : i < 17 ? 4 : i < 21 ? 5 : 6; }
__host__ __device__ constexpr int small_offset(int n, int i)            // first dense index of table i
{
    int o = 0;
    for (int k = 0; k < i && k < small_tables(n); k++) o += ipow(hot_levels(n), tuple_cells(n, k));
    return o;
}
__host__ __device__ constexpr int small_count(int n) { return small_offset(n, small_tables(n)); }

template <int N>
__device__ __forceinline__ uint32_t table_offset_rt(int i)
{
    if (N <= 4) return uint32_t(i) * uint32_t(table_size(N, 0));
    return i <= 17 ? uint32_t(i) * 65536u : i <= 21 ? 17u * 65536u + uint32_t(i - 17) * 1048576u
                                                    : 17u * 65536u + 4u * 1048576u + uint32_t(i - 21) * 7529536u;
}

template <int N>
__device__ __forceinline__ int small_offset_rt(int i)
{
    if (N <= 4) return i * ipow(hot_levels(N), N);
    return i <= 17 ? i << 8 : (17 << 8) + ((i - 17) << 10);
}

// is every cell of the base-16 key v (ncell nibbles) small?
template <int N>
__device__ __forceinline__ bool small_key(uint32_t v, uint32_t big_mask)
{
    if (hot_levels(N) == 4) return (v & big_mask) == 0;
    // HL = 5: no nibble >= 8, and no nibble in 5..7 (bit 2 set together with bit 1 or bit 0)
    return ((v & 0x88888888u) | ((v >> 2) & ((v >> 1) | v) & 0x11111111u)) == 0;
}

// dense index of a small key inside its table
template <int N, int MAXC>
__device__ __forceinline__ uint32_t small_compact(uint32_t v)
{
    uint32_t cmp = 0;
    if (hot_levels(N) == 4) {
#pragma unroll
        for (int k = 0; k < MAXC; k++) cmp |= ((v >> (4 * k)) & 3u) << (2 * k);
    } else {
#pragma unroll
        for (int k = MAXC - 1; k >= 0; k--) cmp = cmp * 5u + ((v >> (4 * k)) & 15u);
    }
    return cmp;
}

// dense index -> weight index (inverse of the compaction in phase B)
template <int N>
__device__ __forceinline__ uint32_t small_to_key(int dense)
{
    if (hot_levels(N) != 4) {                          // n <= 4: every tuple has N cells, radix hot_levels(N)
        constexpr int P = ipow(hot_levels(N), N);
        const int tab = dense / P;
        int compact = dense - tab * P;
        uint32_t idx = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
            idx |= uint32_t(compact % hot_levels(N)) << (4 * k);
            compact /= hot_levels(N);
        }
        return table_offset_rt<N>(tab) + idx;
    }
    int tab, compact, cells;
    if (N <= 4) { tab = dense >> (2 * N); compact = dense & ((1 << (2 * N)) - 1); cells = N; }
    else if (dense < (17 << 8)) { tab = dense >> 8; compact = dense & 255; cells = 4; }
    else { tab = 17 + ((dense - (17 << 8)) >> 10); compact = (dense - (17 << 8)) & 1023; cells = 5; }
    uint32_t idx = 0;
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (k < cells) idx |= uint32_t((compact >> (2 * k)) & 3) << (4 * k);
    return table_offset_rt<N>(tab) + idx;
}

struct PersistBuffers {
    float *w, *delta;
    void *acc;              // float2[nw] | int64[nw]
    uint32_t *cnt;          // exact modes: u32[nw]
    uint32_t *lists;        // per-CTA key lists, list_cap keys each
    void *hot;              // float2[NS] | int64[NS] followed by u32[NS]
    int64_t list_cap;
    float *tune;            // [grid] slots per kilo-cycle each CTA sustained in the previous launch (0 = not measured yet)
};

template <bool EXACT, bool MEAN>
__device__ __forceinline__ float update_value(long long q, float fs, float c)
{
    if (EXACT) {
        double x = double(q) / FIX_SCALE;
        if (MEAN) x = x / double(c);
        return __double2float_rn(x);
    }
    return MEAN ? __fdiv_rn(fs, c) : fs;
}

__device__ __forceinline__ void add_weight(float *__restrict__ w, float *__restrict__ delta, uint32_t k, float u)
{
    __stcg(w + k, __fadd_rn(__ldcg(w + k), u));
    if (delta) __stcg(delta + k, __fadd_rn(__ldcg(delta + k), u));
}

constexpr int PERSIST_TILE = 32;           // FAST path: at most this many slots per CTA

// FAST: the CTA's slots fit one phase-B round (slots <= 512 / F) -- the headline shape, 4,096 games on 148 SMs.
// Game state then lives in registers for the whole launch, phase A hands (dw, 8 D4 images) to phase B through
// shared memory, and a thread keeps the keys it touched first (and the small-key entries it made non-zero) in
// registers across the barrier and applies / flushes them itself: no lists, no cursors, no global round trip
// except the weight gathers and the atomics.  Otherwise slots are processed in rounds through the b2048_td_step
// staging arrays and per-CTA key lists in global memory.
// SCAN (n <= 5; the launcher's choice in the generic layout):
// no first-touch bookkeeping at all.  Phase B fires non-returning REDs (1.5 instead of 2.3 LSU cycles per lane), and the
// apply phase is a dense, coalesced scan of the accumulators by all CTAs (8.9 MB at n = 4: ~60 KB per SM and lock-step),
// which finds the touched keys by their non-zero contributor count.
template <int N, bool EXACT, bool MEAN, bool DIRECT, bool FAST, bool PEERS, bool SCAN>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
td_persist_kernel(PersistBuffers pb, PersistCtrl *ctrl, const uint32_t *__restrict__ lut, b2048_games_t g, float alpha,
                  int steps, uint64_t *upd_board, float *upd_dw, int spc, long long *tlog, const __grid_constant__ PeerSync ps)
{
    // tlog (debug, B2048_PERSIST_TLOG): clock64 at the phase boundaries of the last 16 steps, per CTA
    constexpr int F = num_feat(N);
    constexpr int NS = small_count(N);
    constexpr int MAXC = N;                                // cells of the widest tuple
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // small-key accumulators: sums (float | int64), counts (u32), dirty list (u16 dense indices)
    float *s_sum = reinterpret_cast<float *>(smem_raw);
    unsigned long long *s_q = reinterpret_cast<unsigned long long *>(smem_raw);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem_raw + size_t(NS) * (EXACT ? 8 : 4));
    uint16_t *s_dirty = reinterpret_cast<uint16_t *>(smem_raw + size_t(NS) * (EXACT ? 12 : 8));
    __shared__ uint32_t s_cursor, s_ndirty;
    __shared__ uint64_t s_img[FAST ? PERSIST_TILE * 8 : 1];
    __shared__ float s_dw[FAST ? PERSIST_TILE : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int64_t slot0 = int64_t(blockIdx.x) * spc;
    int nslots = int(g.B - slot0 < spc ? (g.B - slot0 > 0 ? g.B - slot0 : 0) : spc);
    // Self-tuning partition (generic layout with the scanning apply: slots live in global memory and nothing per CTA is
    // sized by its slot count).  SMs do not sustain the same atomic / gather rate -- with 65,536 games at n = 5 the slowest
    // CTA needs 1.5x the cycles of the fastest, always the same ones (far from the L2 slices that hold the accumulators) --
    // and the grid barrier waits for the slowest.  Every launch leaves each CTA's measured slots per kilo-cycle in the
    // workspace; the next launch splits the slots in proportion (all CTAs read the same table before anyone rewrites it).
    constexpr bool TUNE = !FAST && SCAN;
    __shared__ long long s_part[2];
    if (TUNE && pb.tune) {
        if (threadIdx.x == 0) {
            double total = 0.0, before = 0.0, mine = 0.0, mean = 0.0;
            bool known = true;
            for (unsigned t = 0; t < gridDim.x; t++) {
                const float r = __ldcg(pb.tune + t);
                known = known && r > 0.0f;
                mean += r;
            }
            mean /= double(gridDim.x);
            for (unsigned t = 0; known && t < gridDim.x; t++) {      // clamped to 0.6 .. 1.6 of the mean rate
                double r = __ldcg(pb.tune + t);
                r = r < 0.6 * mean ? 0.6 * mean : r > 1.6 * mean ? 1.6 * mean : r;
                total += r;
                if (t < blockIdx.x) before += r;
                if (t == blockIdx.x) mine = r;
            }
            long long lo = slot0, hi = slot0 + nslots;
            if (known) {                                   // identical arithmetic in every CTA: the ranges tile [0, B)
                lo = (long long)(double(g.B) * (before / total));
                hi = blockIdx.x + 1 == gridDim.x ? (long long)g.B : (long long)(double(g.B) * ((before + mine) / total));
            }
            s_part[0] = lo;
            s_part[1] = hi;
        }
        __syncthreads();
        slot0 = s_part[0];
        nslots = int(s_part[1] - s_part[0]);
    }
    long long tune_cycles = 0;
    uint32_t *list = pb.lists + int64_t(blockIdx.x) * pb.list_cap;     // generic path only
    float2 *acc2 = reinterpret_cast<float2 *>(pb.acc);
    unsigned long long *accq = reinterpret_cast<unsigned long long *>(pb.acc);
    float2 *hot2 = reinterpret_cast<float2 *>(pb.hot);
    unsigned long long *hotq = reinterpret_cast<unsigned long long *>(pb.hot);
    uint32_t *hotc = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(pb.hot) + size_t(NS) * 8);
    const b2048_replay_t no_replay{nullptr, nullptr, 0};
    LutGlobal L{lut};
    StepCounters c;
    uint32_t bar_target = 0;
    uint64_t *fault = g.counters + B2048_CTR_FAULT;

    for (int q = threadIdx.x; q < NS; q += blockDim.x) {
        if (EXACT) s_q[q] = 0; else s_sum[q] = 0.0f;
        s_cnt[q] = 0;
    }
    // phase A role (FAST): lane d of slot threadIdx.x / 4, state in registers
    const int a_slot = int(threadIdx.x) >> 2, a_dir = int(threadIdx.x) & 3;
    const bool a_warp = FAST && warp * 8 < nslots, a_in = FAST && a_slot < nslots;
    SlotState st{};
    bool st_dirty = false;
    if (FAST) st = slot_load(g, slot0 + a_slot, a_in);
    // phase B role: table `tab` of entries erow, erow + EPR, ...
    const int EPR = blockDim.x / F;
    const bool worker = int(threadIdx.x) < EPR * F;
    const int tab = int(threadIdx.x) % F, erow = int(threadIdx.x) / F;
    const FeatSpec spec = feat_spec(N, tab);
    int sh[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; k++) sh[k] = 4 * (15 - spec.cell[k < spec.ncell ? k : 0]);
    const int ncell = spec.ncell;
    const bool base14 = spec.base == 14;
    const uint32_t key_off = table_offset_rt<N>(tab);
    const int dense_off = small_offset_rt<N>(tab);
    const uint32_t big_mask = base14 ? 0u : (0xCCCCCCu & ((1u << (4 * ncell)) - 1u));   // any cell > 3
    const int rounds = (nslots + EPR - 1) / EPR;
    // FAST: the weight-independent half of phase A (LUT moves of the 4 directions, table indices of the afterstates)
    // is done one step ahead, between the arrival at and the wait for the grid barrier that precedes phase A
    constexpr bool PREP = FAST && N <= 5 && !EXACT;        // n = 6 / exact modes: the extra registers would spill
    MovePrep<N> prep;
    if (PREP && a_warp) move_prepare<N>(L, st.board, a_dir, prep);
    __syncthreads();

    for (int step = 0; step < steps; step++) {
        if (threadIdx.x == 0) { s_cursor = 0; s_ndirty = 0; }
        long long *tl = (tlog && threadIdx.x == 0 && step >= steps - 16) ?
                        tlog + (int64_t(blockIdx.x) * 16 + (step - (steps - 16))) * 8 : nullptr;
        if (tl) tl[0] = clock64();
        const long long tune_t0 = (TUNE && threadIdx.x == 0) ? clock64() : 0;
        // ---- phase A: 4 lanes per slot, 8 slots per warp
        if (FAST) {
            // lane d stages images d and 4 + d (d4_image order).  Except in the DIRECT mode the hand-over to phase B
            // (the CTA barrier) happens inside phase A, before the spawn and the bookkeeping of the move: the other
            // 12 warps start on the keys while these 4 finish the step.
            auto stage = [&](uint64_t ub, float dw) {
                if (a_in) {
                    uint64_t im = (a_dir & 1) ? flip_h(ub) : ub;
                    if (a_dir & 2) im = flip_v(im);
                    s_img[a_slot * 8 + a_dir] = im;
                    s_img[a_slot * 8 + 4 + a_dir] = transpose(im);
                    if (a_dir == 0) s_dw[a_slot] = dw;
                }
                if (!DIRECT) __syncthreads();
            };
            if (a_warp) {
                uint64_t ub;
                float dw;
                st_dirty |= phase_a_compute<N, true>(pb.w, L, g, alpha, slot0 + a_slot, a_dir, a_in, st, ub, dw, no_replay, 0,
                                                     nullptr, nullptr, nullptr, nullptr, 0, c, PREP ? &prep : nullptr, stage);
            } else if (!DIRECT) {
                __syncthreads();
            }
        } else {
            for (int base = warp * 8; base < nslots; base += nwarps * 8) {
                const int ls = base + (lane >> 2);
                phase_a_slot<N, true>(pb.w, L, g, alpha, slot0 + ls, lane & 3, ls < nslots, upd_board, upd_dw, no_replay,
                                      0, nullptr, nullptr, nullptr, nullptr, 0, c);
            }
        }
        if (DIRECT && !grid_barrier(&ctrl->bar, bar_target, fault)) return;   // every slot has read W_t before anyone adds to it
        else if (!FAST) __syncthreads();
        if (tl) tl[1] = clock64();
        // ---- phase B: one thread per (entry, table); image s: key, duplicate test against the lower images,
        //      atomic -- issued image by image so that the key arithmetic overlaps the atomics in flight
        // FAST: these survive the barrier (the thread applies / flushes its own keys).  The values returned by the
        // global atomics are NOT consumed before the grid barrier: the wait for them overlaps the flush and the
        // barrier itself (the barrier's release fence orders every atomic of the CTA before the arrival anyway).
        uint32_t idx[8], first = 0, dirty = 0;
        float oldf[8];
        uint32_t oldu[8], oldc[8];
        for (int r = 0; r < rounds; r++) {
            const int e = r * EPR + erow;
            const bool on = worker && e < nslots;
            float d = NAN;
            uint64_t img[8];
            if (FAST) {
                if (on) d = s_dw[e];
#pragma unroll
                for (int s = 0; s < 8; s++) img[s] = on ? s_img[e * 8 + s] : 0;
                if (base14) {
#pragma unroll
                    for (int s = 0; s < 8; s++) img[s] = clamp13(img[s]);
                }
            } else {
                if (on) d = __ldcg(upd_dw + slot0 + e);
                const uint64_t b0 = on ? __ldcg(reinterpret_cast<const unsigned long long *>(upd_board) + slot0 + e) : 0;
                img[0] = base14 ? clamp13(b0) : b0;
                img[1] = flip_h(img[0]);
                img[2] = flip_v(img[0]);
                img[3] = flip_v(img[1]);
#pragma unroll
                for (int s = 0; s < 4; s++) img[4 + s] = transpose(img[s]);
            }
            const bool live = on && (EXACT ? isfinite(d) : !isnan(d));
            const long long qd = (EXACT && live) ? quantize(d) : 0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < MAXC; k++)
                    if (k < ncell) {
                        const uint32_t cell = uint32_t(img[s] >> sh[k]) & 15u;
                        v = base14 ? v * 14u + cell : (v << 4) | cell;
                    }
                idx[s] = v;
                bool ld = true;                               // no lower image of this entry has the same key
#pragma unroll
                for (int o = 0; o < s; o++) ld &= idx[o] != v;
                oldf[s] = 1.0f;
                oldu[s] = 1u;
                oldc[s] = 1u;
                if (!live) continue;
                if (!base14 && small_key<N>(v, big_mask)) {   // small-exponent key: shared memory
                    const int di = dense_off + int(small_compact<N, MAXC>(v));
                    if (EXACT) atomicAdd(s_q + di, (unsigned long long)qd); else atomicAdd(s_sum + di, d);
                    if (ld) oldc[s] = atomicAdd(s_cnt + di, 1u);
                } else {
                    const uint32_t k = key_off + v;
                    if (DIRECT) {
                        atomicAdd(pb.w + k, d);
                        if (pb.delta) atomicAdd(pb.delta + k, d);
                    } else if (EXACT) {
                        atomicAdd(accq + k, (unsigned long long)qd);
                        if (ld) {
                            if (SCAN) atomicAdd(pb.cnt + k, 1u);
                            else oldu[s] = atomicAdd(pb.cnt + k, 1u);
                        }
                    } else if (SCAN) {
                        atomicAdd(acc2 + k, make_float2(d, ld ? 1.0f : 0.0f));     // result unused: RED
                    } else if (ld) {
                        oldf[s] = atomicAdd(acc2 + k, make_float2(d, 1.0f)).y;
                    } else {
                        atomicAdd(acc2 + k, make_float2(d, 0.0f));
                    }
                }
            }
            dirty = 0;
#pragma unroll
            for (int s = 0; s < 8; s++) dirty |= uint32_t(oldc[s] == 0u) << s;
            if (!FAST) {
                first = 0;
                if (!DIRECT && !SCAN) {
#pragma unroll
                    for (int s = 0; s < 8; s++) first |= uint32_t(EXACT ? (oldu[s] == 0u) : (oldf[s] == 0.0f)) << s;
                }
                // first touches -> the CTA's key list; first small-key touches -> its dirty list (one warp scan)
                const uint32_t mine = __popc(first) | (__popc(dirty) << 16);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += up;
                }
                const uint32_t total = __shfl_sync(FULL, incl, 31);
                if (total) {                                  // warp-uniform
                    uint32_t base_f = 0, base_d = 0;
                    if (lane == 0) {
                        if (total & 0xFFFFu) base_f = atomicAdd(&s_cursor, total & 0xFFFFu);
                        if (total >> 16) base_d = atomicAdd(&s_ndirty, total >> 16);
                    }
                    uint32_t pf = __shfl_sync(FULL, base_f, 0) + ((incl - mine) & 0xFFFFu);
                    uint32_t pd = __shfl_sync(FULL, base_d, 0) + ((incl - mine) >> 16);
#pragma unroll
                    for (int s = 0; s < 8; s++) {
                        if ((first >> s) & 1u) list[pf++] = key_off + idx[s];
                        if ((dirty >> s) & 1u) s_dirty[pd++] = uint16_t(dense_off + int(small_compact<N, MAXC>(idx[s])));
                    }
                }
            }
        }
        __syncthreads();
        if (tl) tl[2] = clock64();
        // ---- flush the dirty entries of the small-key table (and zero them for the next step): FAST, the thread
        //      that made an entry non-zero flushes it; otherwise from the CTA's dirty list
        {
            auto flush_entry = [&](int q) {
                const uint32_t cv = s_cnt[q];
                s_cnt[q] = 0;
                if (DIRECT) {
                    const uint32_t k = small_to_key<N>(q);
                    atomicAdd(pb.w + k, s_sum[q]);
                    if (pb.delta) atomicAdd(pb.delta + k, s_sum[q]);
                    s_sum[q] = 0.0f;
                } else if (EXACT) {
                    atomicAdd(hotq + q, s_q[q]);
                    atomicAdd(hotc + q, cv);
                    s_q[q] = 0;
                } else {
                    atomicAdd(hot2 + q, make_float2(s_sum[q], float(cv)));
                    s_sum[q] = 0.0f;
                }
            };
            if (FAST) {
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if ((dirty >> s) & 1u) flush_entry(dense_off + int(small_compact<N, MAXC>(idx[s])));
            } else {
                const uint32_t nd = s_ndirty;
                for (uint32_t t = threadIdx.x; t < nd; t += blockDim.x) flush_entry(s_dirty[t]);
            }
        }
        if (tl) tl[3] = clock64();
        if (TUNE && threadIdx.x == 0) tune_cycles += clock64() - tune_t0;
        grid_arrive(&ctrl->bar, bar_target);                  // every contribution of the step has landed
        if (DIRECT && PREP && a_warp) move_prepare<N>(L, st.board, a_dir, prep);   // W_{t+1} is complete after this one
        // FAST apply, the half that does not depend on other CTAs, done while waiting: which keys this thread
        // touched first (the values returned by its atomics), and the old weights of those keys and of its hot entry
        // (a key is applied by exactly one thread, so nobody else writes them before this thread does)
        const int hq = blockIdx.x * blockDim.x + threadIdx.x;
        const bool hot_on = FAST && !DIRECT && hq < NS;
        const uint32_t hk = hot_on ? small_to_key<N>(hq) : 0u;
        float hw = 0.0f, hd = 0.0f, wv[8], dv[8];
        if (FAST && !DIRECT) {
            first = 0;
            if (!SCAN) {
#pragma unroll
                for (int s = 0; s < 8; s++) first |= uint32_t(EXACT ? (oldu[s] == 0u) : (oldf[s] == 0.0f)) << s;
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if ((first >> s) & 1u) {
                        wv[s] = __ldcg(pb.w + key_off + idx[s]);
                        if (pb.delta) dv[s] = __ldcg(pb.delta + key_off + idx[s]);
                    }
            }
            if (hot_on) {
                hw = __ldcg(pb.w + hk);
                if (pb.delta) hd = __ldcg(pb.delta + hk);
            }
        }
        if (!grid_wait(&ctrl->bar, bar_target, fault)) return;
        if (tl) tl[4] = clock64();
        if (!DIRECT) {
            float2 hv = make_float2(0.0f, 0.0f);
            unsigned long long hqv = 0;
            uint32_t hcv = 0;
            if (hot_on) {
                if (EXACT) { hqv = __ldcg(hotq + hq); hcv = __ldcg(hotc + hq); }
                else hv = __ldcg(hot2 + hq);
            }
            if (SCAN) {
                // dense scan of the accumulators, SB x 16 bytes per thread in flight; a key is touched iff its contributor
                // count is non-zero (float modes: the .y of its {sum, count} pair; exact modes: cnt[k])
                constexpr int64_t NWT = table_offset(N, F);
#ifndef B2048_SCAN_BATCH
#define B2048_SCAN_BATCH 8
#endif
                constexpr int SB = B2048_SCAN_BATCH;           // 16-byte loads in flight per thread (the scan is latency-bound)
                const int64_t gsz = int64_t(gridDim.x) * blockDim.x, g0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
                // Two rounds of loads per batch, each with everything in flight: the accumulators, then -- for the units that
                // hold a touched key -- the old weights (and the exact sums).  Stores are per key: a neighbour in the same
                // 16-byte unit may be a small-exponent key that the hot-table pass updates at the same time.
                if (EXACT) {
                    static_assert(NWT % 4 == 0, "cnt is scanned four keys at a time");
                    constexpr int SE = SB < 4 ? SB : 4;
                    const uint4 *c4 = reinterpret_cast<const uint4 *>(pb.cnt);
                    for (int64_t q0 = g0; q0 < NWT / 4; q0 += SE * gsz) {
                        uint4 cv[SE];
#pragma unroll
                        for (int a = 0; a < SE; a++) {
                            const int64_t q = q0 + a * gsz;
                            cv[a] = q < NWT / 4 ? __ldcg(c4 + q) : make_uint4(0u, 0u, 0u, 0u);
                        }
                        ulonglong2 s01[SE], s23[SE];
                        float4 wv[SE];
#pragma unroll
                        for (int a = 0; a < SE; a++)
                            if (cv[a].x | cv[a].y | cv[a].z | cv[a].w) {
                                const int64_t k = 4 * (q0 + a * gsz);
                                s01[a] = __ldcg(reinterpret_cast<const ulonglong2 *>(accq + k));
                                s23[a] = __ldcg(reinterpret_cast<const ulonglong2 *>(accq + k + 2));
                                wv[a] = __ldcg(reinterpret_cast<const float4 *>(pb.w + k));
                            }
#pragma unroll
                        for (int a = 0; a < SE; a++)
                            if (cv[a].x | cv[a].y | cv[a].z | cv[a].w) {
                                const uint32_t k0 = uint32_t(4 * (q0 + a * gsz));
                                const uint32_t cc[4] = {cv[a].x, cv[a].y, cv[a].z, cv[a].w};
                                const unsigned long long qq[4] = {s01[a].x, s01[a].y, s23[a].x, s23[a].y};
                                const float ww[4] = {wv[a].x, wv[a].y, wv[a].z, wv[a].w};
#pragma unroll
                                for (int e = 0; e < 4; e++)
                                    if (cc[e]) {
                                        const float u = update_value<true, MEAN>((long long)qq[e], 0.0f, float(cc[e]));
                                        __stcg(accq + k0 + e, 0ULL);
                                        __stcg(pb.cnt + k0 + e, 0u);
                                        __stcg(pb.w + k0 + e, __fadd_rn(ww[e], u));
                                        if (pb.delta) __stcg(pb.delta + k0 + e, __fadd_rn(__ldcg(pb.delta + k0 + e), u));
                                    }
                            }
                    }
                } else {
                    static_assert(NWT % 2 == 0, "the {sum, count} pairs are scanned two keys at a time");
                    const float4 *a4 = reinterpret_cast<const float4 *>(pb.acc);
                    for (int64_t q0 = g0; q0 < NWT / 2; q0 += SB * gsz) {
                        float4 v[SB];
#pragma unroll
                        for (int a = 0; a < SB; a++) {
                            const int64_t q = q0 + a * gsz;
                            v[a] = q < NWT / 2 ? __ldcg(a4 + q) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        }
                        float2 wv[SB];
#pragma unroll
                        for (int a = 0; a < SB; a++)
                            if (v[a].y != 0.0f || v[a].w != 0.0f)
                                wv[a] = __ldcg(reinterpret_cast<const float2 *>(pb.w + 2 * (q0 + a * gsz)));
#pragma unroll
                        for (int a = 0; a < SB; a++) {
                            const uint32_t k = uint32_t(2 * (q0 + a * gsz));
                            if (v[a].y != 0.0f) {
                                const float u = update_value<false, MEAN>(0, v[a].x, v[a].y);
                                __stcg(acc2 + k, make_float2(0.0f, 0.0f));
                                __stcg(pb.w + k, __fadd_rn(wv[a].x, u));
                                if (pb.delta) __stcg(pb.delta + k, __fadd_rn(__ldcg(pb.delta + k), u));
                            }
                            if (v[a].w != 0.0f) {
                                const float u = update_value<false, MEAN>(0, v[a].z, v[a].w);
                                __stcg(acc2 + k + 1, make_float2(0.0f, 0.0f));
                                __stcg(pb.w + k + 1, __fadd_rn(wv[a].y, u));
                                if (pb.delta) __stcg(pb.delta + k + 1, __fadd_rn(__ldcg(pb.delta + k + 1), u));
                            }
                        }
                    }
                }
            } else if (FAST) {
                // the keys this thread touched first (registers): every accumulator load first, then the arithmetic
                // and the stores -- one L2 round trip for up to 8 keys instead of one per key
                float2 av[8];
                long long aq[8];
                uint32_t ac[8];
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const uint32_t k = key_off + idx[s];
                    if ((first >> s) & 1u) {
                        if (EXACT) { aq[s] = (long long)__ldcg(accq + k); ac[s] = __ldcg(pb.cnt + k); }
                        else av[s] = __ldcg(acc2 + k);
                    }
                }
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    const uint32_t k = key_off + idx[s];
                    if ((first >> s) & 1u) {
                        float u;
                        if (EXACT) {
                            u = update_value<true, MEAN>(aq[s], 0.0f, float(ac[s]));
                            __stcg(accq + k, 0ULL);
                            __stcg(pb.cnt + k, 0u);
                        } else {
                            u = update_value<false, MEAN>(0, av[s].x, av[s].y);
                            __stcg(acc2 + k, make_float2(0.0f, 0.0f));
                        }
                        __stcg(pb.w + k, __fadd_rn(wv[s], u));
                        if (pb.delta) __stcg(pb.delta + k, __fadd_rn(dv[s], u));
                    }
                }
            } else {
                // generic layout: the CTA's key list, 4 keys per thread and pass with every load of a pass in flight
                // together (keys, then accumulators + weights, then the arithmetic and the stores)
                const uint32_t mine = s_cursor;
                constexpr int AB = 4;
                for (uint32_t q0 = threadIdx.x; q0 < mine; q0 += AB * blockDim.x) {
                    uint32_t kk[AB];
                    bool on[AB];
#pragma unroll
                    for (int a = 0; a < AB; a++) {
                        const uint32_t q = q0 + a * blockDim.x;
                        on[a] = q < mine;
                        kk[a] = on[a] ? __ldcg(list + q) : 0u;
                    }
                    float2 av[AB];
                    long long aq[AB];
                    uint32_t ac[AB];
                    float wv2[AB], dv2[AB];
#pragma unroll
                    for (int a = 0; a < AB; a++)
                        if (on[a]) {
                            if (EXACT) { aq[a] = (long long)__ldcg(accq + kk[a]); ac[a] = __ldcg(pb.cnt + kk[a]); }
                            else av[a] = __ldcg(acc2 + kk[a]);
                            wv2[a] = __ldcg(pb.w + kk[a]);
                            if (pb.delta) dv2[a] = __ldcg(pb.delta + kk[a]);
                        }
#pragma unroll
                    for (int a = 0; a < AB; a++)
                        if (on[a]) {
                            float u;
                            if (EXACT) {
                                u = update_value<true, MEAN>(aq[a], 0.0f, float(ac[a]));
                                __stcg(accq + kk[a], 0ULL);
                                __stcg(pb.cnt + kk[a], 0u);
                            } else {
                                u = update_value<false, MEAN>(0, av[a].x, av[a].y);
                                __stcg(acc2 + kk[a], make_float2(0.0f, 0.0f));
                            }
                            __stcg(pb.w + kk[a], __fadd_rn(wv2[a], u));
                            if (pb.delta) __stcg(pb.delta + kk[a], __fadd_rn(dv2[a], u));
                        }
                }
            }
            if (hot_on && (EXACT ? hcv != 0u : hv.y != 0.0f)) {
                const float u = EXACT ? update_value<true, MEAN>((long long)hqv, 0.0f, float(hcv))
                                      : update_value<false, MEAN>(0, hv.x, hv.y);
                if (EXACT) { __stcg(hotq + hq, 0ULL); __stcg(hotc + hq, 0u); }
                else __stcg(hot2 + hq, make_float2(0.0f, 0.0f));
                __stcg(pb.w + hk, __fadd_rn(hw, u));
                if (pb.delta) __stcg(pb.delta + hk, __fadd_rn(hd, u));
            }
            for (int q = blockIdx.x * blockDim.x + threadIdx.x + (FAST ? gridDim.x * blockDim.x : 0); q < NS;
                 q += gridDim.x * blockDim.x) {
                float u;
                if (EXACT) {
                    const uint32_t cv = __ldcg(hotc + q);
                    if (!cv) continue;
                    u = update_value<true, MEAN>((long long)__ldcg(hotq + q), 0.0f, float(cv));
                    __stcg(hotq + q, 0ULL);
                    __stcg(hotc + q, 0u);
                } else {
                    const float2 v = __ldcg(hot2 + q);
                    if (v.y == 0.0f) continue;
                    u = update_value<false, MEAN>(0, v.x, v.y);
                    __stcg(hot2 + q, make_float2(0.0f, 0.0f));
                }
                add_weight(pb.w, pb.delta, small_to_key<N>(q), u);
            }
            if (tl) tl[5] = clock64();
            grid_arrive(&ctrl->bar, bar_target);              // W_{t+1} complete once everybody has arrived
            if (PREP && a_warp) move_prepare<N>(L, st.board, a_dir, prep);
            if (!grid_wait(&ctrl->bar, bar_target, fault)) return;
        }
        if (tl) tl[6] = clock64();
        // ---- multi-GPU: every sync_every lock-steps the replicas exchange what their weights moved by, in this launch
        if (PEERS && ps.sync_every > 0) {                     // (W_{t+1} is complete here in every mode)
            const int done = ps.since_sync + step + 1;        // lock-steps since the sync before this launch
            if (done % ps.sync_every == 0 &&
                !peer_sync_in_kernel(ps, ps.epoch + uint32_t(done / ps.sync_every) - 1u, &ctrl->bar, bar_target, fault))
                return;
        }
    }
    if (FAST && a_in && a_dir == 0 && st_dirty) slot_store(g, slot0 + a_slot, st);
    if (TUNE && pb.tune && threadIdx.x == 0 && tune_cycles > 0 && nslots > 0) {
        // slots per kilo-cycle of this launch, averaged with the previous figure (all CTAs are past the last barrier of
        // the launch when the first of them gets here, so nobody still reads the old table)
        float rate = float(double(nslots) * double(steps) * 1000.0 / double(tune_cycles));
        const float old = __ldcg(pb.tune + blockIdx.x);
        if (old > 0.0f) rate = 0.5f * (rate + old);
        __stcg(pb.tune + blockIdx.x, rate);
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
    const unsigned grid = unsigned(cdiv(m * num_feat(N), 128));       // accumulate: one thread per (entry, table)
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
        td_keys_kernel<N><<<unsigned(cdiv(m * 8, 128)), 128, 0, st>>>(boards, dw, m, ka, va);   // thread per (entry, image)
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
template <int N, bool EXACT, bool MEAN, bool DIRECT, bool FAST, bool PEERS, bool SCAN>
int launch_persist(int grid, cudaStream_t st, void **args)
{
    auto kern = td_persist_kernel<N, EXACT, MEAN, DIRECT, FAST, PEERS, SCAN>;
    const int smem = small_count(N) * (EXACT ? 14 : 10);           // sums, counts, dirty list
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per call: no cache
    if (e != cudaSuccess) return int(e);
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PERSIST_THREADS, smem);
    if (e != cudaSuccess) return int(e);
    if (occ < 1 || grid > occ * sm_count()) return B2048_ENOTSUP;       // all CTAs must be co-resident
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern), dim3(unsigned(grid)), dim3(PERSIST_THREADS), args,
                                    size_t(smem), st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        cudaGetLastError();                                             // not sticky: the caller takes the stepwise path
        return B2048_ENOTSUP;
    }
    return e == cudaSuccess ? 0 : int(e);
}

// instantiated combinations: the in-launch exchange (PEERS) exists for the per-key-mean rule only, the scanning apply
// (SCAN) for n <= 5 and never for the direct mode (which has no apply phase)
template <int N, bool FAST, bool PEERS, bool SCAN>
int launch_persist_mode(bool det, bool mean, int grid, cudaStream_t st, void **args)
{
    if constexpr (PEERS) {
        if (!mean) return B2048_ENOTSUP;
        if (!det) return launch_persist<N, false, true, false, FAST, true, SCAN>(grid, st, args);
        return launch_persist<N, true, true, false, FAST, true, SCAN>(grid, st, args);
    } else {
        if (!det && !mean) return launch_persist<N, false, false, true, FAST, false, false>(grid, st, args);
        if (!det) return launch_persist<N, false, true, false, FAST, false, SCAN>(grid, st, args);
        if (mean) return launch_persist<N, true, true, false, FAST, false, SCAN>(grid, st, args);
        return launch_persist<N, true, false, false, FAST, false, SCAN>(grid, st, args);
    }
}

template <int N, bool FAST, bool PEERS>
int launch_persist_scan(bool scan, bool det, bool mean, int grid, cudaStream_t st, void **args)
{
    if constexpr (N <= 5) {
        if (scan) return launch_persist_mode<N, FAST, PEERS, true>(det, mean, grid, st, args);
    }
    return launch_persist_mode<N, FAST, PEERS, false>(det, mean, grid, st, args);
}

template <int N>
int td_run_persistent(float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                      int mode, int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                      const PeerSync *peer_sync, cudaStream_t st)
{
    const bool det = mode & B2048_UPD_DETERMINISTIC, mean = mode & B2048_UPD_MEAN;
    const int64_t B = g->B;
    const int umode = mode & 7;                                    // mode may carry B2048_RUN_GENERIC
    WorkLayout L = work_layout(N, B, umode | B2048_UPD_MEAN);      // DIRECT uses only the control block
    if (!work || work_bytes < work_layout(N, B, umode).total) return B2048_EWORK;
    // one CTA per SM: slots per CTA rounded up to a multiple of 4
    int max_grid = sm_count();
    if (max_grid > PERSIST_MAX_GRID) max_grid = PERSIST_MAX_GRID;
    const int64_t spc = cdiv(cdiv(B, max_grid), 4) * 4;
    const int grid = int(cdiv(B, spc));
    unsigned char *base = reinterpret_cast<unsigned char *>(work);
    PersistBuffers pb{weights, delta, base + L.acc, reinterpret_cast<uint32_t *>(base + L.cnt),
                      reinterpret_cast<uint32_t *>(base + L.lists), base + L.hot, spc * 8 * num_feat(N),
                      // (measured: +3 % from 32,768 games up, nothing to gain below ~128 slots per CTA)
                      ((mode & B2048_RUN_EVEN) || spc < 128) ? nullptr : reinterpret_cast<float *>(base + L.tune)};
    PersistCtrl *ctrl = reinterpret_cast<PersistCtrl *>(base + L.ctrl);
    b2048_games_t games = *g;
    int spc_i = int(spc);
    long long *tlog = nullptr;
#ifdef B2048_PERSIST_TLOG                                          // profiling builds only (profiles/tlog_summary.py)
    const char *tlog_path = getenv("B2048_PERSIST_TLOG");          // per-phase clocks of the last 16 steps -> text file
    if (tlog_path && *tlog_path && steps >= 16) {
        if (cudaMalloc(&tlog, size_t(grid) * 16 * 8 * sizeof(long long)) != cudaSuccess) tlog = nullptr;
        else cudaMemsetAsync(tlog, 0, size_t(grid) * 16 * 8 * sizeof(long long), st);
    }
#endif
    PeerSync ps{};                                                 // sync_every = 0: single GPU
    if (peer_sync) ps = *peer_sync;
    void *args[] = {&pb, &ctrl, &lut, &games, &alpha, &steps, &upd_board, &upd_dw, &spc_i, &tlog, &ps};
    const bool with_peers = ps.sync_every > 0;
    // FAST: one phase-B round per lock-step, state in registers, staging and key list in shared memory
    const bool fast = spc <= PERSIST_THREADS / num_feat(N) && spc <= PERSIST_TILE && !(mode & B2048_RUN_GENERIC);
    // scanning apply: every mode in the generic layout (measured, profiles/r02_td_scan_vs_lists.txt: float modes 15-50 %
    // faster there, exact modes 1-17 %; the one-round register layout keeps its lists, 5 % faster at 4,096 games), or
    // when forced either way (the parity tests run both)
    bool scan = N <= 5 && !fast;
    if (N <= 3 && !det) scan = true;       // tiny tables (<= 1.7 MB of accumulators): 11.3 vs 16.9 us at n = 3, 1,000 games
    if (mode & B2048_RUN_SCAN) scan = N <= 5;
    if (mode & B2048_RUN_LISTS) scan = false;
    const int rc = fast ? (with_peers ? launch_persist_scan<N, true, true>(scan, det, mean, grid, st, args)
                                      : launch_persist_scan<N, true, false>(scan, det, mean, grid, st, args))
                        : (with_peers ? launch_persist_scan<N, false, true>(scan, det, mean, grid, st, args)
                                      : launch_persist_scan<N, false, false>(scan, det, mean, grid, st, args));
#ifdef B2048_PERSIST_TLOG
    if (tlog) {
        std::vector<long long> h(size_t(grid) * 16 * 8);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), tlog, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(tlog);
        if (FILE *f = fopen(tlog_path, "a")) {
            fprintf(f, "# grid %d spc %d mode %d n %d steps %d: per CTA, per step: A B flush bar1 apply bar2 (cycles)\n", grid,
                    spc_i, mode, N, steps);
            for (int c = 0; c < grid; c++)
                for (int q = 0; q < 16; q++) {
                    const long long *t = &h[(size_t(c) * 16 + q) * 8];
                    fprintf(f, "%d %d %lld %lld %lld %lld %lld %lld\n", c, q, t[1] - t[0], t[2] - t[1], t[3] - t[2],
                            t[4] - t[3], t[5] ? t[5] - t[4] : 0, t[5] ? t[6] - t[5] : 0);
                }
            fclose(f);
        }
    }
#endif
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
    // small batches without replay / traces: the latency-oriented kernel (two moves per L2 round trip, 16 lanes per game)
    if (!replay && !trace_dir && !trace_value && !trace_spawn && g->B <= int64_t(2) * sm_count() * (SPEC_THREADS / 16) &&
        max_steps >= 2) {
        cudaError_t e = cudaFuncSetAttribute(greedy_spec_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_SMEM_BYTES);
        if (e != cudaSuccess) return int(e);
        const int64_t want = cdiv(g->B * 16, SPEC_THREADS);
        const unsigned grid = unsigned(want < sm_count() ? want : sm_count());
        greedy_spec_kernel<N><<<grid, SPEC_THREADS, LUT_SMEM_BYTES, st>>>(w, lut, *g, max_steps, limit_tile, step_limit);
        return launch_status();
    }
#ifdef B2048_GREEDY_SMEM_LUT                                          // lab build: row LUT in shared memory, 1 wide CTA per SM
    {
        cudaError_t e = cudaFuncSetAttribute(greedy_play_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_SMEM_BYTES);
        if (e != cudaSuccess) return int(e);
        const int64_t want = cdiv(g->B * 4, GREEDY_WIDE_THREADS);
        const unsigned grid = unsigned(want < sm_count() ? want : sm_count());
        greedy_play_kernel<N, true><<<grid, GREEDY_WIDE_THREADS, LUT_SMEM_BYTES, st>>>(w, lut, *g, max_steps, limit_tile, step_limit, rp,
                                                                                   replay ? 1 : 0, trace_dir, trace_value, trace_spawn, trace_len);
        return launch_status();
    }
#endif
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, greedy_play_kernel<N, false>, 128, 0);
    if (e != cudaSuccess) return int(e);
    const int64_t want = cdiv(g->B * 4, 128), cap = int64_t(sm_count()) * (occ > 0 ? occ : 1);
    const unsigned grid = unsigned(want < cap ? want : cap);           // persistent: groups pull slots from the queue
    greedy_play_kernel<N, false><<<grid, 128, 0, st>>>(w, lut, *g, max_steps, limit_tile, step_limit, rp, replay ? 1 : 0, trace_dir,
                                                      trace_value, trace_spawn, trace_len);
    return launch_status();
}

template <int N>
int look_forward_impl(const float *w, const uint32_t *lut, const uint64_t *boards, const uint64_t *game_id,
                      const uint32_t *move_no, const uint8_t *root_dir, int64_t m, int depth, int width, int since_empty,
                      uint64_t seed, float *value, cudaStream_t st)
{
    look_forward_kernel<N><<<unsigned(cdiv(m * 32, 128)), 128, 0, st>>>(w, lut, boards, game_id, move_no, root_dir, m, depth,
                                                                       width, since_empty, seed, value);
    return launch_status();
}

template <int N>
int expectimax_play_impl(const float *w, const uint32_t *lut, const b2048_games_t *g, int max_steps, int limit_tile,
                         int step_limit, int depth, int width, int since_empty, int8_t *trace_dir, uint16_t *trace_spawn,
                         int64_t trace_len, cudaStream_t st)
{
    const unsigned grid = unsigned(cdiv(g->B * 32, 128));
    if (g->B <= 2048)
        expectimax_play_kernel<N, 1><<<grid, 128, 0, st>>>(w, lut, *g, max_steps, limit_tile, step_limit, depth, width,
                                                          since_empty, trace_dir, trace_spawn, trace_len);
    else
        expectimax_play_kernel<N, 6><<<grid, 128, 0, st>>>(w, lut, *g, max_steps, limit_tile, step_limit, depth, width,
                                                          since_empty, trace_dir, trace_spawn, trace_len);
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
