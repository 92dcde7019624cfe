// Random 4-byte gather rate of a B200 over footprints like the n-tuple weight tables (4.46 MB n=4, 21.2 MB n=5, 382.7 MB
// n=6): the physical roof of evaluate() -- every gather moves one 32-byte sector from L2 (or HBM) to the SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
// Each thread keeps U independent loads in flight per iteration (evaluate keeps F = 17..33); indices come from a
// counter hash, so nothing is coalesced.  Reports G gathers/s and the sector bandwidth they imply.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int U, bool L1>
__global__ void __launch_bounds__(128) gather_kernel(const float *__restrict__ w, uint32_t n, int iters, float *out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    for (int it = 0; it < iters; it++) {
        float v[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const uint32_t idx = uint32_t((uint64_t(mix(t * 977u + it * 131071u + j * 7919u)) * n) >> 32);
            v[j] = L1 ? __ldg(w + idx) : __ldcg(w + idx);
        }
#pragma unroll
        for (int j = 0; j < U; j++) acc += v[j];
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int U, bool L1>
double run(const float *w, uint32_t n, float *out, int ctas_per_sm)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm, iters = 256;
    gather_kernel<U, L1><<<grid, 128>>>(w, n, 16, out);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    gather_kernel<U, L1><<<grid, 128>>>(w, n, iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return double(grid) * 128 * iters * U / (ms * 1e-3);
}

int main()
{
    const size_t sizes[] = {1114112, 5308416, 25000000, 95662848};
    float *w, *out;
    cudaMalloc(&w, sizes[3] * 4);
    cudaMalloc(&out, 4);
    cudaMemset(w, 0, sizes[3] * 4);
    printf("# random 4-byte gathers, 128-thread CTAs; G gathers/s (x 32 B sectors = GB/s)\n");
    printf("# footprint_MB  path  loads_in_flight  CTAs/SM  G_gathers/s  sector_GB/s\n");
    for (size_t n : sizes)
        for (int occ : {6, 12}) {
            double r;
            r = run<17, true>(w, uint32_t(n), out, occ);  printf("%8.1f  L1+L2  17  %2d  %8.1f  %8.0f\n", n * 4e-6, occ, r / 1e9, r * 32 / 1e9);
            r = run<33, true>(w, uint32_t(n), out, occ);  printf("%8.1f  L1+L2  33  %2d  %8.1f  %8.0f\n", n * 4e-6, occ, r / 1e9, r * 32 / 1e9);
            r = run<17, false>(w, uint32_t(n), out, occ); printf("%8.1f  L2     17  %2d  %8.1f  %8.0f\n", n * 4e-6, occ, r / 1e9, r * 32 / 1e9);
        }
    return 0;
}
