// Microbenchmark (not part of the product): cost of one grid-wide barrier of a resident cooperative grid
// (148 CTAs x 512 threads) on B200, three protocols:
//   0  one counter: red.release.gpu arrive by thread 0, thread 0 polls ld.relaxed.gpu   (td_persist_kernel today)
//   1  flag array: thread 0 stores the epoch to flag[cta] (st.release.gpu); threads 0..G-1 each poll one flag
//   3  flag array packed (4 bytes per CTA)
//   2  two-level counters: 8 group counters (arrive), the last arriver of a group bumps the root; poll the root
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gridbar gridbar.cu && ./gridbar
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) barriers(uint32_t *bar, uint32_t *flags, int rounds, float *sink, float *data)
{
    const uint32_t G = gridDim.x;
    uint32_t target = 0, epoch = 0;
    float acc = 0;
    for (int r = 0; r < rounds; r++) {
        // a little memory traffic before the arrival, like the real kernel (one RED per thread)
        atomicAdd(data + ((blockIdx.x * 512 + threadIdx.x) * 37 + r * 101) % (1 << 20), 1.0f);
        __syncthreads();
        if (MODE == 0) {
            if (threadIdx.x == 0) {
                target += G;
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                while (ld_relaxed(bar) < target) { }
            }
        } else if (MODE == 1 || MODE == 3) {
            constexpr int STRIDE = MODE == 1 ? 32 : 1;            // one 128-byte line per flag | packed
            epoch++;
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * STRIDE), "r"(epoch) : "memory");
            if (threadIdx.x < G) while (ld_relaxed(flags + threadIdx.x * STRIDE) < epoch) { }
        } else {
            if (threadIdx.x == 0) {
                target += 1;
                const uint32_t grp = blockIdx.x >> 5, gsize = min(32u, G - grp * 32u);
                uint32_t old;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar + 32 + grp * 32) : "memory");
                if ((old + 1) % gsize == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
                const uint32_t groups = (G + 31) / 32;
                while (ld_relaxed(bar) < target * groups) { }
            }
        }
        __syncthreads();
        acc += __ldcg(data + (threadIdx.x + r) % 1024);
    }
    if (acc == 12345.678f) *sink = acc;
}

template <int MODE>
void run(const char *name, int flagstride_note)
{
    uint32_t *bar, *flags;
    float *sink, *data;
    cudaMalloc(&bar, 4096 * 4); cudaMalloc(&flags, 148 * 32 * 4); cudaMalloc(&sink, 4); cudaMalloc(&data, (1 << 20) * 4);
    cudaMemset(data, 0, (1 << 20) * 4);
    int rounds = 4000;
    void *args[] = {&bar, &flags, &rounds, &sink, &data};
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(bar, 0, 4096 * 4); cudaMemset(flags, 0, 148 * 32 * 4);
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void *)barriers<MODE>, dim3(148), dim3(512), args, 0, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %.3f us per round (%s)\n", name, ms * 1e3 / rounds, cudaGetErrorString(cudaGetLastError()));
    cudaFree(bar); cudaFree(flags); cudaFree(sink); cudaFree(data);
}

__global__ void __launch_bounds__(512, 1) nobarrier(int rounds, float *sink, float *data)
{
    float acc = 0;
    for (int r = 0; r < rounds; r++) {
        atomicAdd(data + ((blockIdx.x * 512 + threadIdx.x) * 37 + r * 101) % (1 << 20), 1.0f);
        __syncthreads();
        __syncthreads();
        acc += __ldcg(data + (threadIdx.x + r) % 1024);
    }
    if (acc == 12345.678f) *sink = acc;
}

int main()
{
    run<0>("counter: red.release + thread-0 poll", 0);
    run<1>("flags: st.release + one poller per CTA flag", 0);
    run<3>("flags packed in 5 lines", 0);
    run<2>("two-level counters (groups of 32)", 0);
    float *sink, *data;
    cudaMalloc(&sink, 4); cudaMalloc(&data, (1 << 20) * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    nobarrier<<<148, 512>>>(4000, sink, data);
    cudaEventRecord(a);
    nobarrier<<<148, 512>>>(4000, sink, data);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %.3f us per round\n", "same loop without the grid barrier", ms * 1e3 / 4000);
    return 0;
}
