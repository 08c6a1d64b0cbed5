// Probe: how much DRAM traffic does a random 16-byte gather cost on B200, and does
// any load flavour (nc / cg / cs / lu / L2 prefetch hints / evict-first) change it?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ uint4 ld_variant(const uint4 * p, int v)
{
    uint4 r;
    switch (v) {
    case 0: r = *p; break;
    case 1: asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 2: asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 3: asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 4: asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 5: asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 6: asm volatile("ld.global.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 7: asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    case 8: asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); break;
    default: {
        u64 pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    } break;
    }
    return r;
}

template <int V, int UNROLL>
__global__ void __launch_bounds__(256) gather16(const uint4 * __restrict__ base, const u32 * __restrict__ idx, uint4 * __restrict__ out, size_t n)
{
    const size_t t0 = ((size_t) blockIdx.x * blockDim.x) * UNROLL + threadIdx.x;
    u32 ix[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) { const size_t t = t0 + (size_t) k * blockDim.x; ix[k] = t < n ? idx[t] : 0; }
    uint4 v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) v[k] = ld_variant(base + ix[k], V);
#pragma unroll
    for (int k = 0; k < UNROLL; k++) { const size_t t = t0 + (size_t) k * blockDim.x; if (t < n) out[t] = v[k]; }
}

// scatter variant: sequential read, random 16-byte write
template <int UNROLL>
__global__ void __launch_bounds__(256) scatter16(const uint4 * __restrict__ base, const u32 * __restrict__ idx, uint4 * __restrict__ out, size_t n)
{
    const size_t t0 = ((size_t) blockIdx.x * blockDim.x) * UNROLL + threadIdx.x;
#pragma unroll
    for (int k = 0; k < UNROLL; k++) { const size_t t = t0 + (size_t) k * blockDim.x; if (t < n) out[idx[t]] = base[t]; }
}

__global__ void make_idx(u32 * idx, size_t n, u32 a, u32 b, int window_log2)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (window_log2 == 0) { idx[i] = (u32) (((u64) a * i + b) & (n - 1)); return; }
    // permutation that is random only inside windows of 2^window_log2 records
    const size_t w = (size_t) 1 << window_log2;
    idx[i] = (u32) ((i & ~(w - 1)) | (((u64) a * i + b) & (w - 1)));
}

template <int V> float run(const uint4 * base, const u32 * idx, uint4 * out, size_t n)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int U = 4; const size_t blocks = (n + 256 * U - 1) / (256 * U);
    gather16<V, U><<<(unsigned) blocks, 256>>>(base, idx, out, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) gather16<V, U><<<(unsigned) blocks, 256>>>(base, idx, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}

int main(int argc, char ** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const size_t n = (size_t) 1 << lg;
    uint4 * base, * out; u32 * idx;
    cudaMalloc(&base, n * 16); cudaMalloc(&out, n * 16); cudaMalloc(&idx, n * 4);
    cudaMemset(base, 1, n * 16);
    if (argc > 2) { size_t g = atoi(argv[2]); cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("set L2 fetch granularity %zu -> %s, now %zu\n", g, cudaGetErrorString(e), got); }
    else { size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", got); }
    for (int wl = 0; wl <= 22; wl += (wl == 0 ? 16 : 3)) {
        make_idx<<<(unsigned) ((n + 255) / 256), 256>>>(idx, n, 2654435761u, 12345u, wl);
        cudaDeviceSynchronize();
        printf("window 2^%d records:", wl);
        float ms;
        ms = run<0>(base, idx, out, n); printf(" plain %.3f", ms);
        if (wl == 0) {
        ms = run<1>(base, idx, out, n); printf(" nc %.3f", ms);
        ms = run<2>(base, idx, out, n); printf(" cg %.3f", ms);
        ms = run<3>(base, idx, out, n); printf(" cs %.3f", ms);
        ms = run<4>(base, idx, out, n); printf(" lu %.3f", ms);
        ms = run<5>(base, idx, out, n); printf(" nc.noalloc %.3f", ms);
        ms = run<6>(base, idx, out, n); printf(" L2::64B %.3f", ms);
        ms = run<7>(base, idx, out, n); printf(" volatile %.3f", ms);
        ms = run<8>(base, idx, out, n); printf(" relaxed %.3f", ms);
        ms = run<9>(base, idx, out, n); printf(" evict_first %.3f", ms);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const size_t blocks = (n + 1023) / 1024;
        scatter16<4><<<(unsigned) blocks, 256>>>(base, idx, out, n);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; r++) scatter16<4><<<(unsigned) blocks, 256>>>(base, idx, out, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf(" | scatter %.3f", ms / 3);
        }
        printf("  (ms for %zu records; ideal %.3f ms at 6.5 TB/s)\n", n, n * 36.0 / 6.5e9);
    }
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
