// b2048_kernels.cu -- the C-ABI of include/b2048.h and the kernels that do not depend on the tuple size:
// (1) packed boards, (2) LUT moves, (3) Philox / replay spawns, the config-5 sweep, games init, multi-GPU delta
// pack/apply.  The n-tuple agent kernels (4) live in b2048_agent.cuh, one translation unit per tuple size.
// Nothing here is a dense contraction: no tensor cores by design (HBM/L2-latency- and issue-bound).
#include "b2048_host.cuh"

namespace {

const b2048_agent_ops *agent_ops(int n)
{
    switch (n) {
    case 2: return &b2048_agent_ops_2;
    case 3: return &b2048_agent_ops_3;
    case 4: return &b2048_agent_ops_4;
    case 5: return &b2048_agent_ops_5;
    case 6: return &b2048_agent_ops_6;
    default: return nullptr;
    }
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
constexpr int SWEEP_THREADS = 1024;
constexpr size_t SWEEP_SMEM = 65536 * 2 + 65536 + 256 * 4;   // lines u16, merge codes u8, score of a code u32
#ifndef B2048_SWEEP_SWZ
#define B2048_SWEEP_SWZ 1
#endif
// shared-memory slot of a line.  Lines of real boards are far from uniform (small tiles and empty cells dominate the low
// nibbles, which select the bank), so with the identity mapping a warp's 32 lookups pile up on a few banks (6 wavefronts
// per request on iid boards, more on boards from games).  XOR-ing the bank bits with the upper byte of the line spreads
// them; the map is a bijection (each bit is XOR-ed only with higher bits).  The kernel applies it to the four lines of a
// board at once (sweep_slots: 4 instructions per direction).
__device__ __forceinline__ uint32_t sweep_slot(uint32_t line)
{
    return B2048_SWEEP_SWZ ? line ^ ((line >> 7) & 0x1FEu) : line;
}
__device__ __forceinline__ uint64_t sweep_slots(uint64_t x)
{
    return B2048_SWEEP_SWZ ? x ^ ((x >> 7) & 0x01FE01FE01FE01FEULL) : x;
}

__global__ void __launch_bounds__(SWEEP_THREADS, 1)
sweep_kernel(const uint32_t *__restrict__ lut, const uint64_t *__restrict__ boards, int64_t m, uint64_t seed,
             uint64_t first_index, uint64_t *__restrict__ after, uint32_t *__restrict__ gain,
             uint8_t *__restrict__ flags, uint64_t *__restrict__ spawned)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t *srow = reinterpret_cast<uint16_t *>(smem);
    uint8_t *scode = smem + 65536 * 2;
    // score of a merge-code byte (two exponents): bits 0-23 the merge score 2^(a+1) + 2^(b+1), bits 24+ the number of
    // exponents equal to 15 (a 2^16 would appear); the four lines of a direction just add their entries.  The sweep is
    // bound by the integer-logic pipe (ncu: ALU pipe 81 % busy), so a second small lookup beats decoding 8 nibbles.
    uint32_t *sscore = reinterpret_cast<uint32_t *>(smem + 65536 * 3);
    if (threadIdx.x < 256) {
        const uint32_t a = threadIdx.x & 15u, b = threadIdx.x >> 4;
        sscore[threadIdx.x] = (((2u << a) & ~2u) + ((2u << b) & ~2u)) | ((uint32_t(a == 15u) + uint32_t(b == 15u)) << 24);
    }
    // stage: 4 entries per thread per iteration (16 B global load -> 8 B + 4 B shared stores)
    for (int q = threadIdx.x; q < 65536 / 4; q += SWEEP_THREADS) {
        uint4 e = __ldg(reinterpret_cast<const uint4 *>(lut) + q);
        if (B2048_SWEEP_SWZ) {
            const uint32_t ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t slot = sweep_slot(4u * q + j);
                srow[slot] = uint16_t(ev[j]);
                scode[slot] = uint8_t(ev[j] >> 16);
            }
        } else {
            uint2 r = make_uint2((e.x & 0xFFFFu) | (e.y << 16), (e.z & 0xFFFFu) | (e.w << 16));
            reinterpret_cast<uint2 *>(srow)[q] = r;
            uint32_t c = ((e.x >> 16) & 0xFFu) | (((e.y >> 16) & 0xFFu) << 8) | (((e.z >> 16) & 0xFFu) << 16) |
                         (((e.w >> 16) & 0xFFu) << 24);
            reinterpret_cast<uint32_t *>(scode)[q] = c;
        }
    }
    __syncthreads();
    const int64_t stride = int64_t(gridDim.x) * SWEEP_THREADS;
    for (int64_t i = blockIdx.x * int64_t(SWEEP_THREADS) + threadIdx.x; i < m; i += stride) {
        const uint64_t b = __ldg(boards + i);
        // every direction is "slide left" on a transformed board (pre_move = rot90 / left / rot90 back,
        // game_logic.py:136-142): up = transpose, right = mirror, down = transpose then mirror
        const uint64_t bt = transpose(b);
        const uint64_t x[4] = {b, bt, flip_h(b), flip_h(bt)};
        uint64_t a[4];
        uint32_t g[4], fl = 0, ok = 0;
        bool ovf2[2];
#pragma unroll
        for (int d = 0; d < 4; d++) {
            uint64_t out = 0;
            uint32_t tot = 0;                                    // sum of the four lines' score entries
            const uint64_t xs = sweep_slots(x[d]);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t line = uint32_t(xs >> (48 - 16 * r)) & 0xFFFFu;
                out |= uint64_t(srow[line]) << (48 - 16 * r);
                if (d < 2) tot += sscore[scode[line]];
            }
            // Merges pair up equal neighbours inside maximal runs of equal tiles, floor(L / 2) per run from either
            // end, so the merge score and the 2^16 escape of right / down equal those of left / up: only the
            // afterstate and the changed flag depend on the side.  (Checked exhaustively against the oracle.)
            if (d < 2) {
                ovf2[d] = (tot >> 24) != 0;                      // a 15+15 merge
                g[d] = tot & 0xFFFFFFu;                          // score += 2^(e+1) per merge (game_logic.py:33)
            } else {
                g[d] = g[d - 2];
            }
            const bool ovf = ovf2[d & 1];
            const bool ch = out != x[d] || ovf;
            if (d & 2) out = flip_h(out);
            a[d] = (d & 1) ? transpose(out) : out;
            fl |= uint32_t(ch) << d;
            fl |= uint32_t(ovf) << (4 + d);
            ok |= uint32_t(ch && !ovf) << d;
        }
        ulonglong2 *ap = reinterpret_cast<ulonglong2 *>(after + 4 * i);
        ap[0] = make_ulonglong2(a[0], a[1]);
        ap[1] = make_ulonglong2(a[2], a[3]);
        *reinterpret_cast<uint4 *>(gain + 4 * i) = make_uint4(g[0], g[1], g[2], g[3]);
        flags[i] = uint8_t(fl);
        if (spawned) {
            uint64_t idx = first_index + uint64_t(i);
            if (ok) {                                            // one Philox call per board, one word per direction
                Philox4 w = spawn_words(seed, idx, 0u, 1u);
                const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int d = 0; d < 4; d++)
                    if ((ok >> d) & 1u) spawn_apply_nonempty(a[d], sweep_tile_word(wd[d]), sweep_pos_word(wd[d]));
            }
            ulonglong2 *sp = reinterpret_cast<ulonglong2 *>(spawned + 4 * i);
            sp[0] = make_ulonglong2(a[0], a[1]);
            sp[1] = make_ulonglong2(a[2], a[3]);
        }
    }
}

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

__global__ void delta_pack_kernel(const float *__restrict__ delta, float *__restrict__ packed, int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
        float d = delta[i];
        packed[i] = d;
        packed[count + i] = d != 0.0f ? 1.0f : 0.0f;
    }
}

__global__ void delta_pack_diff_kernel(const float *__restrict__ w, const float *__restrict__ w_sync,
                                       float *__restrict__ packed, int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
        const float a = w[i], b = w_sync[i];
        packed[i] = __fsub_rn(a, b);
        packed[count + i] = a != b ? 1.0f : 0.0f;
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
        if (delta) delta[i] = 0.0f;
    }
}


// ---- multi-GPU sync, compact form: delta = w - w_sync as float32 plus ONE BIT per weight (w != w_sync) instead of a
// float indicator: the message is 4 + 1/8 bytes per weight and rank (allreduce of the deltas, allgather of the bit
// planes) instead of 8.  One weight per thread, the 32 lanes of a warp own one bitmask word.
__global__ void delta_pack_bits_kernel(const float *__restrict__ w, const float *__restrict__ w_sync,
                                       float *__restrict__ delta, uint32_t *__restrict__ bits, int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;                  // a multiple of 32
    const int64_t padded = (count + 31) & ~int64_t(31);
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < padded; i += stride) {
        const bool in = i < count;
        const float a = in ? w[i] : 0.0f, b = in ? w_sync[i] : 0.0f;
        if (in) delta[i] = __fsub_rn(a, b);
        const uint32_t word = __ballot_sync(FULL, a != b);
        if ((threadIdx.x & 31) == 0) bits[i >> 5] = word;
    }
}

// w_sync += delta_sum / max(1, contributors), w = w_sync, contributors = number of ranks whose bit is set
__global__ void delta_apply_bits_kernel(float *__restrict__ w, float *__restrict__ w_sync,
                                        const float *__restrict__ delta_sum, const uint32_t *__restrict__ bits_all,
                                        int world, int64_t words, int64_t count)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < count; i += stride) {
        uint32_t c = 0;
        for (int r = 0; r < world; r++) c += (__ldg(bits_all + r * words + (i >> 5)) >> (i & 31)) & 1u;
        float s = delta_sum[i];
        if (c > 1u) s = __fdiv_rn(s, float(c));
        const float v = __fadd_rn(w_sync[i], s);
        w_sync[i] = v;
        w[i] = v;
    }
}

// ---- multi-GPU sync, fused: ONE kernel per rank over NVLink peer memory (b2048_sync_peers) ------------------------
// Every rank's w, w_sync and flag block are mapped into every process (CUDA IPC / symmetric memory: the host passes the
// peer pointers).  Rank r owns the r-th contiguous slice of the weights: it reads w_q and w_sync_q of that slice from
// every rank q (remote loads), forms delta_q = w_q - w_sync_q and the contributor count, and writes
// w_sync + sum / max(1, contributors) back into w_q and w_sync_q of every rank (remote stores) -- pack, reduce-scatter,
// apply and all-gather of the NCCL path in one pass, 2 x 4 bytes per weight and rank each way, and every replica
// receives the owner's bits.  Ranks rendezvous through epoch flags in peer memory: arrive[q] before the first remote
// load (rank q has finished the lock-steps before its sync kernel and will not touch w until the sync is over), done[q]
// after the last remote store has been fenced (the last CTA to finish signals and waits, the others just exit).
constexpr int PEER_THREADS = 256;

template <int W>   // W = world size (compile-time: the per-rank loads are all issued before the first use)
__global__ void __launch_bounds__(PEER_THREADS)
peer_sync_kernel(b2048_peers_t P, int64_t count, uint32_t epoch)
{
    __shared__ int s_ok;
    const int rank = P.rank;
    uint32_t *mine = P.flags[rank];
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(P.flags[threadIdx.x] + B2048_PEER_ARRIVE + rank, epoch);
    if (threadIdx.x < W && !wait_epoch(mine + B2048_PEER_ARRIVE + threadIdx.x, epoch)) s_ok = 0;
    __syncthreads();
    if (s_ok) {
        const int64_t n4 = count >> 2;                                    // float4 units; the tail goes to the last rank
        const int64_t per = (n4 + W - 1) / W;
        const int64_t lo = rank * per, hi = (lo + per < n4) ? lo + per : n4;
        const int64_t stride = int64_t(gridDim.x) * PEER_THREADS;
        for (int64_t v = lo + blockIdx.x * int64_t(PEER_THREADS) + threadIdx.x; v < hi; v += stride) {
            float4 a[W], b[W];
#pragma unroll
            for (int q = 0; q < W; q++) {
                a[q] = ld_sys_f4(P.w[q] + 4 * v);
                b[q] = ld_sys_f4(P.w_sync[q] + 4 * v);
            }
            float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, base[4];
            uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < W; q++) {                                  // rank order: a fixed association
                const float av[4] = {a[q].x, a[q].y, a[q].z, a[q].w}, bv[4] = {b[q].x, b[q].y, b[q].z, b[q].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float d = __fsub_rn(av[k], bv[k]);
                    sum[k] = q == 0 ? d : __fadd_rn(sum[k], d);
                    c[k] += av[k] != bv[k];
                    if (q == P.rank) base[k] = bv[k];
                }
            }
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; k++) r[k] = __fadd_rn(base[k], c[k] > 1u ? __fdiv_rn(sum[k], float(c[k])) : sum[k]);
            // skip the stores where nobody moved any of the 128 weights the warp covers (every replica already holds
            // w = w_sync = the result there: whole key ranges of large exponents are never touched); decided per warp so
            // that the remote stores stay full 512-byte runs
            if (__all_sync(__activemask(), (c[0] | c[1] | c[2] | c[3]) == 0)) continue;
            const float4 out = make_float4(r[0], r[1], r[2], r[3]);
#pragma unroll
            for (int q = 0; q < W; q++) {
                *reinterpret_cast<float4 *>(P.w[q] + 4 * v) = out;
                *reinterpret_cast<float4 *>(P.w_sync[q] + 4 * v) = out;
            }
        }
        if (rank == W - 1 && blockIdx.x == 0) {                            // count % 4 trailing weights
            for (int64_t i = (n4 << 2) + threadIdx.x; i < count; i += PEER_THREADS) {
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
    }
    // completion: the last CTA of this rank tells every peer that its stores are out, then waits for theirs
    __syncthreads();
    __shared__ uint32_t s_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(mine + B2048_PEER_TICKET, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();                            // acquire side of the ticket chain: every CTA's fenced stores are ordered before the signal below
    if (threadIdx.x == 0) mine[B2048_PEER_TICKET] = 0;
    if (threadIdx.x < W) {
        if (s_ok) {
            st_release_sys(P.flags[threadIdx.x] + B2048_PEER_DONE + rank, epoch);
            if (!wait_epoch(mine + B2048_PEER_DONE + threadIdx.x, epoch)) s_ok = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !s_ok) mine[B2048_PEER_FAULT] = epoch;         // host-visible: this sync did not complete
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
    // per call (a few hundred ns): the attribute belongs to the current context, and the library caches nothing
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SWEEP_SMEM));
    if (e != cudaSuccess) return int(e);
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
    return agent_ops(n)->features(boards, m, feat, S(stream));
}

int b2048_evaluate(int n, const float *weights, const uint64_t *boards, int64_t m, float *value, b2048_stream_t stream)
{
    if (m < 0 || num_feat(n) < 0 || !weights || (m && (!boards || !value))) return B2048_EINVAL;
    if (!m) return 0;
    return agent_ops(n)->evaluate(weights, boards, m, value, S(stream));
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
    return agent_ops(n)->td_update(weights, delta, boards, dw, m, mode, work, work_bytes, S(stream));
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
    if (e == cudaSuccess) e = cudaMemsetAsync(g->counters + B2048_CTR_QUEUE, 0, sizeof(uint64_t), S(stream));
    if (e != cudaSuccess) return int(e);
    if (max_steps == 0) {                                                    // nothing to play: every open slot stays active
        return agent_ops(n)->greedy_play(weights, lut, g, 0, limit_tile, step_limit, replay, trace_dir, trace_value,
                                         trace_spawn, trace_len, S(stream));
    }
    return agent_ops(n)->greedy_play(weights, lut, g, max_steps, limit_tile, step_limit, replay, trace_dir, trace_value,
                                     trace_spawn, trace_len, S(stream));
}

int b2048_look_forward(int n, const float *weights, const uint32_t *lut, const uint64_t *boards, const uint64_t *game_id,
                       const uint32_t *move_no, const uint8_t *root_dir, int64_t m, int depth, int width, int since_empty,
                       uint64_t seed, float *value, b2048_stream_t stream)
{
    if (m < 0 || num_feat(n) < 0 || !weights || !lut || depth < 0 || depth > 4 || width < 1 || width > 4 || since_empty < 0 ||
        (m && (!boards || !game_id || !move_no || !root_dir || !value)))
        return B2048_EINVAL;
    if (!m) return 0;
    return agent_ops(n)->look_forward(weights, lut, boards, game_id, move_no, root_dir, m, depth, width, since_empty, seed,
                                      value, S(stream));
}

int b2048_expectimax_play(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, int max_steps,
                          int limit_tile, int step_limit, int depth, int width, int since_empty, int8_t *trace_dir,
                          uint16_t *trace_spawn, int64_t trace_len, b2048_stream_t stream)
{
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || max_steps < 0 || depth < 0 || depth > 4 || width < 1 ||
        width > 4 || since_empty < 0)
        return B2048_EINVAL;
    if (g->B == 0) return 0;
    cudaError_t e = cudaMemsetAsync(g->counters + B2048_CTR_ACTIVE, 0, sizeof(uint64_t), S(stream));
    if (e != cudaSuccess) return int(e);
    return agent_ops(n)->expectimax_play(weights, lut, g, max_steps, limit_tile, step_limit, depth, width, since_empty,
                                         trace_dir, trace_spawn, trace_len, S(stream));
}

int b2048_td_phase_a(int n, const float *weights, const uint32_t *lut, const b2048_games_t *g, float alpha,
                     uint64_t *upd_board, float *upd_dw, const b2048_replay_t *replay, int8_t *trace_dir,
                     float *trace_value, float *trace_dw, uint16_t *trace_spawn, int64_t trace_len,
                     b2048_stream_t stream)
{
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || !upd_board || !upd_dw) return B2048_EINVAL;
    if (replay && (!replay->tile || !replay->pos || replay->len < 0)) return B2048_EINVAL;
    if (g->B == 0) return 0;
    return agent_ops(n)->td_phase_a(weights, lut, g, alpha, upd_board, upd_dw, replay, trace_dir, trace_value, trace_dw,
                                    trace_spawn, trace_len, S(stream));
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
    return agent_ops(n)->td_update(weights, delta, upd_board, upd_dw, g->B, mode, work, work_bytes, S(stream));
}

int b2048_td_run(int n, float *weights, float *delta, const uint32_t *lut, const b2048_games_t *g, float alpha,
                 int mode, int steps, uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes,
                 b2048_stream_t stream)
{
    if (steps < 0 || (mode & ~(7 | B2048_RUN_STEPWISE | B2048_RUN_LAYOUT))) return B2048_EINVAL;
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || !upd_board || !upd_dw) return B2048_EINVAL;
    if ((reinterpret_cast<uintptr_t>(weights) & 15) || (delta && (reinterpret_cast<uintptr_t>(delta) & 15)) ||
        (reinterpret_cast<uintptr_t>(work) & 15))
        return B2048_EINVAL;                                       // the apply phase uses 16-byte vector loads
    if (steps == 0 || g->B == 0) return 0;
    const bool stepwise = mode & (B2048_RUN_STEPWISE | B2048_UPD_SORTED);
    const int layout = mode & B2048_RUN_LAYOUT;
    mode &= 7;
    if (!stepwise && cooperative_ok()) {
        int rc = agent_ops(n)->td_run_persistent(weights, delta, lut, g, alpha, mode | layout, steps, upd_board, upd_dw, work,
                                                 work_bytes, nullptr, S(stream));
        if (rc != B2048_ENOTSUP) return rc;
    }
    for (int s = 0; s < steps; s++) {
        int rc = b2048_td_step(n, weights, delta, lut, g, alpha, mode, upd_board, upd_dw, work, work_bytes, nullptr,
                               nullptr, nullptr, nullptr, nullptr, 0, stream);
        if (rc) return rc;
    }
    return 0;
}

int b2048_td_run_peers(int n, float *weights, const uint32_t *lut, const b2048_games_t *g, float alpha, int mode, int steps,
                       uint64_t *upd_board, float *upd_dw, void *work, size_t work_bytes, const b2048_peers_t *peers,
                       int sync_every, int since_sync, uint32_t epoch, b2048_stream_t stream)
{
    if (steps < 0 || (mode & ~(7 | B2048_RUN_LAYOUT)) || (mode & B2048_UPD_SORTED)) return B2048_EINVAL;
    if (!games_ok(g) || num_feat(n) < 0 || !weights || !lut || !upd_board || !upd_dw) return B2048_EINVAL;
    if (!peers || peers->world < 1 || peers->world > B2048_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world ||
        sync_every < 1 || since_sync < 0 || since_sync >= sync_every || epoch == 0)
        return B2048_EINVAL;
    for (int q = 0; q < peers->world; q++)
        if (!peers->w[q] || !peers->w_sync[q] || !peers->flags[q]) return B2048_EINVAL;
    if (peers->w[peers->rank] != weights) return B2048_EINVAL;      // the trained buffer must be the mapped one
    if ((reinterpret_cast<uintptr_t>(weights) & 15) || (reinterpret_cast<uintptr_t>(work) & 15)) return B2048_EINVAL;
    if (steps == 0) return 0;
    if (g->B == 0 || !cooperative_ok()) return B2048_ENOTSUP;       // every rank needs the persistent launch
    PeerSync ps{};
    ps.peers = *peers;
    ps.count = table_offset(n, num_feat(n));
    ps.sync_every = sync_every;
    ps.since_sync = since_sync;
    ps.epoch = epoch;
    return agent_ops(n)->td_run_persistent(weights, nullptr, lut, g, alpha, mode, steps, upd_board, upd_dw, work, work_bytes,
                                           &ps, S(stream));
}

int64_t b2048_td_run_launches(int n, int64_t B, int mode, int steps)
{
    if (num_feat(n) < 0 || B < 0 || steps < 0 || (mode & ~(7 | B2048_RUN_STEPWISE | B2048_RUN_LAYOUT))) return -1;
    if (B == 0 || steps == 0) return 0;
    const bool stepwise = mode & (B2048_RUN_STEPWISE | B2048_UPD_SORTED);
    if (!stepwise && cooperative_ok()) return 1;                   // the persistent kernel
    mode &= 7;
    int64_t per_step = 2;                                          // phase A + direct accumulate
    if (mode != (B2048_UPD_ATOMIC | B2048_UPD_SUM)) per_step = 3;  // + apply
    if (mode & B2048_UPD_SORTED) per_step = 1 + 1 + 3 * ((key_bits(n) + 7) / 8) + 1 + 1;
    return per_step * steps;
}

int b2048_delta_pack(const float *delta, float *packed, int64_t count, b2048_stream_t stream)
{
    if (count < 0 || (count && (!delta || !packed))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_pack_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(delta, packed, count);
    return launch_status();
}

int b2048_delta_pack_diff(const float *weights, const float *w_sync, float *packed, int64_t count, b2048_stream_t stream)
{
    if (count < 0 || (count && (!weights || !w_sync || !packed))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_pack_diff_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(weights, w_sync, packed, count);
    return launch_status();
}

int b2048_delta_apply(float *weights, float *w_sync, float *delta, const float *delta_sum, const float *contributors,
                      int64_t count, b2048_stream_t stream)
{
    if (count < 0 || (count && (!weights || !w_sync || !delta_sum))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_apply_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(weights, w_sync, delta, delta_sum,
                                                                              contributors, count);
    return launch_status();
}


int b2048_delta_pack_bits(const float *weights, const float *w_sync, float *delta, uint32_t *bits, int64_t count,
                          b2048_stream_t stream)
{
    if (count < 0 || (count && (!weights || !w_sync || !delta || !bits))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_pack_bits_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(weights, w_sync, delta, bits, count);
    return launch_status();
}

int b2048_delta_apply_bits(float *weights, float *w_sync, const float *delta_sum, const uint32_t *bits_all, int world,
                           int64_t count, b2048_stream_t stream)
{
    if (count < 0 || world < 1 || (count && (!weights || !w_sync || !delta_sum || !bits_all))) return B2048_EINVAL;
    if (!count) return 0;
    int64_t want = cdiv(count, 256), cap = int64_t(sm_count()) * 8;
    delta_apply_bits_kernel<<<unsigned(want < cap ? want : cap), 256, 0, S(stream)>>>(weights, w_sync, delta_sum, bits_all,
                                                                                   world, cdiv(count, 32), count);
    return launch_status();
}

int b2048_sync_peers(const b2048_peers_t *peers, int64_t count, uint32_t epoch, int max_ctas, b2048_stream_t stream)
{
    if (!peers || count < 0 || peers->world < 1 || peers->world > B2048_MAX_PEERS || peers->rank < 0 ||
        peers->rank >= peers->world || epoch == 0)
        return B2048_EINVAL;
    for (int q = 0; q < peers->world; q++)
        if (!peers->w[q] || !peers->w_sync[q] || !peers->flags[q]) return B2048_EINVAL;
    if ((count >> 2) && ((reinterpret_cast<uintptr_t>(peers->w[peers->rank]) | reinterpret_cast<uintptr_t>(peers->w_sync[peers->rank])) & 15))
        return B2048_EINVAL;                                               // float4 access
    const int64_t slice4 = cdiv(count >> 2, peers->world);
    int64_t grid = cdiv(slice4 > 0 ? slice4 : 1, PEER_THREADS);
    int64_t cap = max_ctas > 0 ? max_ctas : int64_t(sm_count()) * 4;        // every CTA of every rank must be resident
    if (grid > cap) grid = cap;
    switch (peers->world) {
#define B2048_PEER_CASE(Wn) case Wn: peer_sync_kernel<Wn><<<unsigned(grid), PEER_THREADS, 0, S(stream)>>>(*peers, count, epoch); break;
    B2048_PEER_CASE(1) B2048_PEER_CASE(2) B2048_PEER_CASE(3) B2048_PEER_CASE(4) B2048_PEER_CASE(5) B2048_PEER_CASE(6)
    B2048_PEER_CASE(7) B2048_PEER_CASE(8)
#undef B2048_PEER_CASE
    default: return B2048_ENOTSUP;
    }
    return launch_status();
}

}   // extern "C"

