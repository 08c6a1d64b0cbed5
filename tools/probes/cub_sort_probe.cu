// Comparison probe (NOT part of the product, which uses no CUB): CUB DeviceRadixSort on the same
// GPU for the bench workload, 2^28 (u64 key, u64 value) pairs = 16-byte records, device resident.
// "Best existing kernel on this box" per SURVEY.md 8(d).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cub_sort_probe cub_sort_probe.cu
#include <cstdio>
#include <cstdint>
#include <cub/cub.cuh>

__global__ void fill(unsigned long long * k, unsigned long long * v, size_t n)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = i + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    k[i] = z; v[i] = i;
}

int main(int argc, char ** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const size_t n = (size_t) 1 << lg;
    unsigned long long *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&v0, n * 8); cudaMalloc(&v1, n * 8);
    void * tmp = nullptr; size_t tmpb = 0;
    cub::DeviceRadixSort::SortPairs(tmp, tmpb, k0, k1, v0, v1, n);
    cudaMalloc(&tmp, tmpb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; it++) {
        fill<<<(unsigned) ((n + 255) / 256), 256>>>(k0, v0, n);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortPairs(tmp, tmpb, k0, k1, v0, v1, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    printf("CUB DeviceRadixSort::SortPairs u64/u64 n=2^%d: best %.3f ms = %.2f Grec/s (temp %.1f MB) err=%s\n",
           lg, best, n / best / 1e6, tmpb / 1e6, cudaGetErrorString(cudaDeviceSynchronize()));
    // keys only
    best = 1e9f;
    cub::DeviceRadixSort::SortKeys(nullptr, tmpb, k0, k1, n);
    for (int it = 0; it < 4; it++) {
        fill<<<(unsigned) ((n + 255) / 256), 256>>>(k0, v0, n);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortKeys(tmp, tmpb, k0, k1, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    printf("CUB DeviceRadixSort::SortKeys  u64     n=2^%d: best %.3f ms = %.2f Gkeys/s\n", lg, best, n / best / 1e6);
    return 0;
}
