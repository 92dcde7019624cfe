// Microbenchmark (not part of the product): L2 float/int atomic throughput on B200 for the access patterns of
// the TD scatter: 557,056 contributions into a 4.46 MB table, with a fraction of them on a few hot addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu && ./atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// mode 0: RED float, 1: ATOM float (returning), 2: RED u64, 3: RED float + ATOM u32 (accum pattern)
template <int MODE>
__global__ void scatter(float *w, unsigned long long *w64, uint32_t *cnt, int n_per_thread, uint32_t nw, uint32_t hot_pct,
                        uint32_t n_hot, float *sink)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0;
    for (int i = 0; i < n_per_thread; i++) {
        uint32_t h = hash32(t * 131u + i);
        uint32_t k = (h % 100u) < hot_pct ? (hash32(h) % n_hot) * 65536u % nw : hash32(h ^ 0x9e3779b9u) % nw;
        if (MODE == 0) atomicAdd(w + k, 1.0f);
        if (MODE == 1) acc += atomicAdd(w + k, 1.0f);
        if (MODE == 2) atomicAdd(w64 + k, 1ull);
        if (MODE == 3) { atomicAdd(w + k, 1.0f); acc += atomicAdd(cnt + k, 1u); }
    }
    if (acc == 12345.f) *sink = acc;
}

__global__ void smem_atomics(float *out, int iters, int same)
{
    __shared__ float s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = 0;
    __syncthreads();
    for (int i = 0; i < iters; i++) atomicAdd(&s[same ? 0 : (threadIdx.x * 33 + i) & 1023], 1.0f);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

template <int MODE>
void run(const char *name, float *w, unsigned long long *w64, uint32_t *cnt, float *sink, uint32_t nw, uint32_t hot_pct,
         uint32_t n_hot)
{
    const int threads = 32768, per = 17;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++) scatter<MODE><<<threads / 128, 128>>>(w, w64, cnt, per, nw, hot_pct, n_hot, sink);
    cudaEventRecord(a);
    const int reps = 20;
    for (int rep = 0; rep < reps; rep++) scatter<MODE><<<threads / 128, 128>>>(w, w64, cnt, per, nw, hot_pct, n_hot, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double us = ms * 1e3 / reps;
    printf("%-34s hot=%2u%% on %3u addrs: %8.2f us per 557k ops  (%6.1f G ops/s)\n", name, hot_pct, n_hot, us,
           threads * per / us * 1e-3);
}

int main()
{
    const uint32_t nw = 1114112;
    float *w, *sink; unsigned long long *w64; uint32_t *cnt;
    cudaMalloc(&w, nw * 4); cudaMalloc(&w64, nw * 8); cudaMalloc(&cnt, nw * 4); cudaMalloc(&sink, 4096);
    cudaMemset(w, 0, nw * 4); cudaMemset(w64, 0, nw * 8); cudaMemset(cnt, 0, nw * 4);
    for (uint32_t hot : {0u, 10u, 30u, 60u}) {
        for (uint32_t nh : {17u, 136u}) {
            if (hot == 0 && nh != 17u) continue;
            run<0>("RED.f32", w, w64, cnt, sink, nw, hot, nh);
            run<1>("ATOM.f32 (returning)", w, w64, cnt, sink, nw, hot, nh);
            run<2>("RED.u64", w, w64, cnt, sink, nw, hot, nh);
            run<3>("RED.f32 + ATOM.u32 (accum)", w, w64, cnt, sink, nw, hot, nh);
        }
    }
    // launch overhead reference: empty-ish kernel
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 100; i++) scatter<0><<<1, 32>>>(w, w64, cnt, 0, nw, 0, 1, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("back-to-back tiny launches: %.2f us each\n", ms * 10);
    for (int same = 0; same < 2; same++) {
        cudaEventRecord(a);
        smem_atomics<<<148, 256>>>(sink, 1000, same);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("smem atomicAdd.f32 %s: 256 thr x 1000 iters per CTA: %.1f us -> %.2f cycles/warp-op @1.9GHz\n",
               same ? "same address" : "spread", ms * 1e3, ms * 1e-3 * 1.9e9 / (8 * 1000));
    }
    return 0;
}
