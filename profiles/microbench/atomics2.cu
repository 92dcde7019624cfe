// Microbenchmark (not part of the product): STEADY-STATE L2 atomic / gather throughput and grid-barrier cost on
// B200 for the persistent lock-step trainer: a resident grid (148 CTAs) loops over rounds, every thread issues
// BATCH independent operations per round on pseudo-random keys of a 1,114,112-entry table (n=4), optionally
// with a share of the keys drawn from a small hot set.  Unlike atomics.cu (one op per thread, dominated by
// launch ramp), this measures what a persistent kernel can sustain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics2 atomics2.cu && ./atomics2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

constexpr int BATCH = 17;

// MODE 0 RED.f32 | 1 ATOM.f32 ret | 2 RED.f32x2 | 3 ATOM.f32x2 ret | 4 RED.u64 | 5 RED.u64 + ATOM.u32 ret
//      6 LDG.cg gather (f32) | 7 LDG.cg gather (float2)
template <int MODE>
__global__ void __launch_bounds__(1024, 1)
steady(float *w, float2 *w2, unsigned long long *w64, uint32_t *cnt, int rounds, uint32_t nw, uint32_t hot_pct,
       uint32_t n_hot, float *sink)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0;
    for (int r = 0; r < rounds; r++) {
        uint32_t k[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; i++) {
            uint32_t h = hash32((t * 131u + i) * 2654435761u + r);
            k[i] = (h % 100u) < hot_pct ? (hash32(h) % n_hot) : hash32(h ^ 0x9e3779b9u) % nw;
        }
        float ret[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; i++) {
            ret[i] = 0;
            if (MODE == 0) atomicAdd(w + k[i], 1.0f);
            if (MODE == 1) ret[i] = atomicAdd(w + k[i], 1.0f);
            if (MODE == 2) atomicAdd(w2 + k[i], make_float2(1.0f, 1.0f));
            if (MODE == 3) ret[i] = atomicAdd(w2 + k[i], make_float2(1.0f, 1.0f)).y;
            if (MODE == 4) atomicAdd(w64 + k[i], 1ull);
            if (MODE == 5) { atomicAdd(w64 + k[i], 1ull); ret[i] = float(atomicAdd(cnt + k[i], 1u)); }
            if (MODE == 6) ret[i] = __ldcg(w + k[i]);
            if (MODE == 7) ret[i] = __ldcg(w2 + k[i]).y;
        }
#pragma unroll
        for (int i = 0; i < BATCH; i++) acc += ret[i];
    }
    if (acc == 12345.678f) *sink = acc;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024, 1) barriers(uint32_t *bar, int rounds)
{
    uint32_t target = 0;
    for (int r = 0; r < rounds; r++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(bar, 1u);
            while (ld_acquire_gpu(bar) < target) { }
        }
        __syncthreads();
    }
}

template <int MODE>
void run(const char *name, int threads, uint32_t hot_pct, uint32_t n_hot, float *w, float2 *w2, unsigned long long *w64,
         uint32_t *cnt, float *sink, uint32_t nw)
{
    const int grid = 148, rounds = 200;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    steady<MODE><<<grid, threads>>>(w, w2, w64, cnt, 20, nw, hot_pct, n_hot, sink);
    cudaEventRecord(a);
    steady<MODE><<<grid, threads>>>(w, w2, w64, cnt, rounds, nw, hot_pct, n_hot, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    double ops = double(grid) * threads * rounds * BATCH;
    printf("%-28s thr=%4d hot=%2u%% on %5u: %7.1f G ops/s  (557k ops = %6.2f us)\n", name, threads, hot_pct, n_hot,
           ops / (ms * 1e-3) / 1e9, 557056.0 / (ops / (ms * 1e-3)) * 1e6);
}

int main()
{
    const uint32_t nw = 1114112;
    float *w, *sink; float2 *w2; unsigned long long *w64; uint32_t *cnt, *bar;
    cudaMalloc(&w, nw * 4); cudaMalloc(&w2, nw * 8); cudaMalloc(&w64, nw * 8); cudaMalloc(&cnt, nw * 4);
    cudaMalloc(&sink, 4); cudaMalloc(&bar, 4);
    cudaMemset(w, 0, nw * 4); cudaMemset(w2, 0, nw * 8); cudaMemset(w64, 0, nw * 8); cudaMemset(cnt, 0, nw * 4);
    for (int threads : {256, 512, 1024}) {
        for (uint32_t hot : {0u, 30u}) {
            run<0>("RED.f32", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<1>("ATOM.f32 ret", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<2>("RED.f32x2", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<3>("ATOM.f32x2 ret", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<4>("RED.u64", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<5>("RED.u64 + ATOM.u32 ret", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<6>("LDG.cg f32 gather", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
            run<7>("LDG.cg f32x2 gather", threads, hot, 4352, w, w2, w64, cnt, sink, nw);
        }
    }
    run<3>("ATOM.f32x2 ret (64 hot)", 512, 30, 64, w, w2, w64, cnt, sink, nw);
    run<0>("RED.f32 (64 hot)", 512, 30, 64, w, w2, w64, cnt, sink, nw);
    for (int threads : {128, 512, 1024}) {
        cudaMemset(bar, 0, 4);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        void *args[] = {&bar, nullptr};
        int rounds = 2000;
        args[1] = &rounds;
        cudaLaunchCooperativeKernel((void *)barriers, dim3(148), dim3(threads), args, 0, 0);
        cudaDeviceSynchronize();
        cudaMemset(bar, 0, 4);
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void *)barriers, dim3(148), dim3(threads), args, 0, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        printf("grid barrier, 148 CTAs x %4d threads: %.2f us each (%s)\n", threads, ms * 1e3 / rounds,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
