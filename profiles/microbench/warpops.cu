// Microbenchmark (not part of the product): latency of __match_any_sync / __shfl_sync / __ballot_sync on B200,
// single warp dependent chain and 16 warps per SM, with 32 distinct / 8 distinct / 1 distinct value per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warpops warpops.cu && ./warpops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(uint32_t *out, long long *cycles, int iters, int distinct)
{
    uint32_t lane = threadIdx.x & 31;
    uint32_t v = (lane % distinct) * 2654435761u + blockIdx.x;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t r;
        if (OP == 0) r = __match_any_sync(0xFFFFFFFFu, v + (acc & 1u) * 0u + acc * 0u + (acc >> 31));
        if (OP == 1) r = __shfl_xor_sync(0xFFFFFFFFu, v + (acc >> 31), 5);
        if (OP == 2) r = __ballot_sync(0xFFFFFFFFu, (v + (acc >> 31)) & 1u);
        acc += r >> 1;                                  // dependent chain
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, int threads, int distinct)
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    chain<OP><<<148, threads>>>(out, cyc, iters, distinct);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; i++) s += h[i];
    printf("%-10s %4d threads/SM, %2d distinct values: %.1f cycles per op per warp\n", name, threads, distinct, s / 148 / iters);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int threads : {32, 128, 512}) {
        for (int d : {32, 8, 1}) run<0>("match_any", threads, d);
        run<1>("shfl_xor", threads, 32);
        run<2>("ballot", threads, 32);
    }
    return 0;
}
