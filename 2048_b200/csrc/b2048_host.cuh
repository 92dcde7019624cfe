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
    size_t acc, cnt, touched, ctrl, keys_a, keys_b, vals_a, vals_b, hist, lists, hot, total;
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
    L.hot = o; o += align256(size_t(17 * 256 + 4 * 1024) * 12);      // dense table of the small-exponent keys
    L.total = o;
    return L;
}

bool games_ok(const b2048_games_t *g)
{
    return g && g->B >= 0 && g->board && g->score && g->moves && g->game_id && g->state && g->old_label && g->flags &&
           g->counters && g->tile_hist;
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
                             cudaStream_t st);
    int (*look_forward)(const float *w, const uint32_t *lut, const uint64_t *boards, const uint64_t *game_id,
                        const uint32_t *move_no, const uint8_t *root_dir, int64_t m, int depth, int width,
                        int since_empty, uint64_t seed, float *value, cudaStream_t st);
    int (*expectimax_play)(const float *w, const uint32_t *lut, const b2048_games_t *g, int max_steps, int limit_tile,
                           int step_limit, int depth, int width, int since_empty, int8_t *trace_dir,
                           uint16_t *trace_spawn, int64_t trace_len, cudaStream_t st);
};
extern const b2048_agent_ops b2048_agent_ops_2, b2048_agent_ops_3, b2048_agent_ops_4, b2048_agent_ops_5,
    b2048_agent_ops_6;
