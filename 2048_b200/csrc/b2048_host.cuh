// b2048_host.cuh -- host-side helpers shared by the translation units of libb2048.so (internal).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#include "b2048_device.cuh"
#include "../../include/b2048.h"

namespace {

using namespace b2048;

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int B2048_RUN_LAYOUT = B2048_RUN_GENERIC | B2048_RUN_SCAN | B2048_RUN_LISTS | B2048_RUN_EVEN;   // layout hints of b2048_td_run

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}

inline cudaStream_t S(b2048_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (a driver attribute query, ~100 ns: nothing is cached, the library keeps no
// mutable state of its own)
int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}

bool cooperative_ok()
{
    int dev = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    return cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop;
}

// ---- TD update workspace (b2048_td_update_workspace) ------------------------------------------------
struct UpdCtrl {
    uint32_t count;     // touched keys of the running update
    uint32_t ticket;    // apply pass: blocks done (the last one resets both)
    uint32_t pad[2];
};

struct PersistCtrl {
    UpdCtrl upd;            // stepwise path (offset 0)
    uint32_t bar;           // grid barrier arrivals of the running launch
    uint32_t exit_ticket;   // CTAs that left; the last one resets bar and exit_ticket
};

// stepwise path: the accumulator is replicated (replica = CTA index mod R) against same-address atomics
__host__ __device__ constexpr int acc_replicas(int64_t nw) { return nw <= (int64_t(1) << 23) ? 8 : 2; }

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int PERSIST_THREADS = 512;

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
    size_t acc, cnt, touched, ctrl, keys_a, keys_b, vals_a, vals_b, hist, lists, hot, tune, total;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

constexpr int PERSIST_MAX_GRID = 1024;

inline WorkLayout work_layout(int n, int64_t m, int mode)
{
    WorkLayout L{};
    L.M = m * 8 * num_feat(n);
    L.nw = table_offset(n, num_feat(n));
    L.nblocks = int(cdiv(L.M, SORT_TILE));
    size_t o = 0;
    L.ctrl = o; o += 256;
    if (mode == (B2048_UPD_ATOMIC | B2048_UPD_SUM)) { L.total = o; return L; }    // control block only
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
    // persistent trainer: per-CTA key lists, entries_per_CTA * 8F keys each (PERSIST_MAX_GRID CTAs at most)
    L.lists = o; o += align256(size_t(m + 4 * PERSIST_MAX_GRID) * 8 * num_feat(n) * 4);
    L.hot = o; o += align256(size_t(10752) * 12);                    // dense table of the small-exponent keys (<= 10,625)
    L.tune = o; o += align256(size_t(PERSIST_MAX_GRID) * 4);         // per-CTA slots per kilo-cycle of the previous launch
    L.total = o;
    return L;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float4 ld_sys_f4(const float *p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float ld_sys_f1(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// spin until *flag has reached `epoch` (wrap-safe); false after ~2^26 polls: a peer is gone
__device__ __forceinline__ bool wait_epoch(const uint32_t *flag, uint32_t epoch)
{
    for (uint32_t polls = 0; int32_t(ld_acquire_sys(flag) - epoch) < 0;)
        if (++polls > (1u << 26)) return false;
    return true;
}


// One pass of rank P.rank over ITS contiguous slice of the weights (see peer_sync_kernel): remote loads of every rank's
// w and w_sync, sum of the deltas in rank order, contributor count, result stored into w and w_sync of every rank.
// Run-time world size (the persistent trainer calls it from inside its lock-step loop); `nthreads` threads with ids
// 0 <= tid < nthreads share the slice.
__device__ __forceinline__ void peer_reduce_slice(const b2048_peers_t &P, int64_t count, int64_t tid, int64_t nthreads)
{
    const int W = P.world, rank = P.rank;
    const int64_t n4 = count >> 2;
    const int64_t per = (n4 + W - 1) / W;
    const int64_t lo = rank * per, hi = (lo + per < n4) ? lo + per : n4;
    for (int64_t v = lo + tid; v < hi; v += nthreads) {
        float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, base[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll 2
        for (int q = 0; q < W; q++) {                                      // rank order: a fixed association
            const float4 a = ld_sys_f4(P.w[q] + 4 * v), b = ld_sys_f4(P.w_sync[q] + 4 * v);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float d = __fsub_rn(av[k], bv[k]);
                sum[k] = q == 0 ? d : __fadd_rn(sum[k], d);
                c[k] += av[k] != bv[k];
                if (q == rank) base[k] = bv[k];
            }
        }
        if (__all_sync(__activemask(), (c[0] | c[1] | c[2] | c[3]) == 0)) continue;   // nobody moved the warp's 128 weights
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = __fadd_rn(base[k], c[k] > 1u ? __fdiv_rn(sum[k], float(c[k])) : sum[k]);
        const float4 out = make_float4(r[0], r[1], r[2], r[3]);
        for (int q = 0; q < W; q++) {
            *reinterpret_cast<float4 *>(P.w[q] + 4 * v) = out;
            *reinterpret_cast<float4 *>(P.w_sync[q] + 4 * v) = out;
        }
    }
    if (rank == W - 1)                                                     // count % 4 trailing weights
        for (int64_t i = (n4 << 2) + tid; i < count; i += nthreads) {
            float sum = 0.0f, base = 0.0f;
            uint32_t c = 0;
            for (int q = 0; q < W; q++) {
                const float av = ld_sys_f1(P.w[q] + i), bv = ld_sys_f1(P.w_sync[q] + i);
                const float d = __fsub_rn(av, bv);
                sum = q == 0 ? d : __fadd_rn(sum, d);
                c += av != bv;
                if (q == rank) base = bv;
            }
            if (c == 0) continue;
            const float out = __fadd_rn(base, c > 1u ? __fdiv_rn(sum, float(c)) : sum);
            for (int q = 0; q < W; q++) { P.w[q][i] = out; P.w_sync[q][i] = out; }
        }
}

// The same pass one weight per thread and iteration (4-byte accesses, coalesced over the warp): a tenth of the registers of
// the 16-byte version, for the exchange inside the persistent trainer, whose hot loop owns the register file.
__device__ __forceinline__ void peer_reduce_slice_scalar(const b2048_peers_t &P, int64_t count, int64_t tid, int64_t nthreads)
{
    const int W = P.world, rank = P.rank;
    const int64_t per = (count + W - 1) / W;
    const int64_t lo = rank * per, hi = (lo + per < count) ? lo + per : count;
    for (int64_t i = lo + tid; i < hi; i += nthreads) {
        float sum = 0.0f, base = 0.0f;
        uint32_t c = 0;
        for (int q = 0; q < W; q++) {                                      // rank order: a fixed association
            const float av = ld_sys_f1(P.w[q] + i), bv = ld_sys_f1(P.w_sync[q] + i);
            const float d = __fsub_rn(av, bv);
            sum = q == 0 ? d : __fadd_rn(sum, d);
            c += av != bv;
            if (q == rank) base = bv;
        }
        if (__all_sync(__activemask(), c == 0)) continue;                  // nobody moved the warp's 32 weights
        const float out = __fadd_rn(base, c > 1u ? __fdiv_rn(sum, float(c)) : sum);
        for (int q = 0; q < W; q++) { P.w[q][i] = out; P.w_sync[q][i] = out; }
    }
}

// in-kernel weight exchange of the persistent trainer (b2048_td_run_peers): which lock-steps end with a sync
struct PeerSync {
    b2048_peers_t peers;
    int64_t count;          // weights
    int sync_every;         // 0 = never (single GPU)
    int since_sync;         // lock-steps already done since the last sync when the launch starts
    uint32_t epoch;         // epoch of the first sync of this launch
};

bool games_ok(const b2048_games_t *g)
{
    if (!g || g->B < 0 || !g->counters || !g->tile_hist) return false;
    if (g->B == 0) return true;                        // an empty batch may carry null slot arrays
    return g->board && g->score && g->moves && g->game_id && g->state && g->old_label && g->flags;
}


}   // namespace

// per-n entry points: every agent size is compiled in its own translation unit (b2048_agent_inst.cu with
// -DB2048_N=n, built in parallel) and reached through this table
struct b2048_agent_ops {
    int (*features)(const uint64_t *boards, int64_t m, int32_t *feat, cudaStream_t st);
    int (*evaluate)(const float *w, const uint64_t *boards, int64_t m, float *value, cudaStream_t st);
    int (*td_update)(float *w, float *delta, const uint64_t *boards, const float *dw, int64_t m, int mode, void *work,
                     size_t work_bytes, cudaStream_t st);
    int (*greedy_play)(const float *w, const uint32_t *lut, const b2048_games_t *g, int max_steps, int limit_tile,
                       int step_limit, const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value,
                       uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st);
    int (*td_phase_a)(const float *w, const uint32_t *lut, const b2048_games_t *g, float alpha, uint64_t *upd_board,
                      float *upd_dw, const b2048_replay_t *replay, int8_t *trace_dir, float *trace_value,
                      float *trace_dw, uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st);
    int (*td_run_persistent)(float *w, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha, int mode,
                             int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                             const PeerSync *peer_sync, cudaStream_t st);
    int (*look_forward)(const float *w, const uint32_t *lut, const uint64_t *boards, const uint64_t *game_id,
                        const uint32_t *move_no, const uint8_t *root_dir, int64_t m, int depth, int width,
                        int since_empty, uint64_t seed, float *value, cudaStream_t st);
    int (*expectimax_play)(const float *w, const uint32_t *lut, const b2048_games_t *g, int max_steps, int limit_tile,
                           int step_limit, int depth, int width, int since_empty, int8_t *trace_dir,
                           uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st);
};
extern const b2048_agent_ops b2048_agent_ops_2, b2048_agent_ops_3, b2048_agent_ops_4, b2048_agent_ops_5,
    b2048_agent_ops_6;
