/* kernels/common.cuh -- shared helpers: device count, launch counter, lane mask.
 * Part of the single translation unit mpsort_kernels.cu (included there, first). */

#include "mpsort_kernels.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#define FULL_MASK 0xffffffffu

static int g_num_sms = 0;
static int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

/* every kernel launch of this library passes through here: counted for bench.py's
 * "gpu_launches" claim */
static unsigned long long g_launches = 0;
#define CUDA_LAUNCH_CHECK() do { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); \
    cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int) e__; } while (0)

extern "C" uint64_t mpsk_launch_count(int reset)
{
    const unsigned long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, 0ULL, __ATOMIC_RELAXED);
    return (uint64_t) v;
}

/* splitmix64 finaliser: the hash of the predictor's table */
__host__ __device__ __forceinline__ u64 mset_hash64(u64 x)
{
    u64 z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
