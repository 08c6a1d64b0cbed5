/* kernels/checksum.cuh -- K8: checksum.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
/* ========================================================================= */
/* K8: checksum                                                              */
/* ========================================================================= */

__global__ void __launch_bounds__(256)
checksum_kernel(const unsigned char * __restrict__ base, size_t nbytes, u64 * sum)
{
    /* head bytes up to 16-byte alignment, body as uint4 with dp4a, tail bytes */
    const uintptr_t addr = (uintptr_t) base;
    size_t head = (16 - (addr & 15)) & 15;
    if (head > nbytes) head = nbytes;
    const size_t nvec = (nbytes - head) / 16;
    const size_t tail_start = head + nvec * 16;
    const uint4 * body = (const uint4 *) (base + head);

    long long acc = 0;
    const size_t gtid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    int part = 0;
    int since = 0;
    for (size_t i = gtid; i < nvec; i += nthreads) {
        const uint4 v = body[i];
        part = __dp4a((int) v.x, 0x01010101, part);
        part = __dp4a((int) v.y, 0x01010101, part);
        part = __dp4a((int) v.z, 0x01010101, part);
        part = __dp4a((int) v.w, 0x01010101, part);
        if (++since == 65536) { acc += part; part = 0; since = 0; }
    }
    acc += part;
    if (gtid < head) acc += (signed char) base[gtid];
    if (gtid < nbytes - tail_start) acc += (signed char) base[tail_start + gtid];

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(sum, (u64) acc);
}

extern "C" int mpsk_checksum(const void * base, size_t nbytes, uint64_t * sum, mpsk_stream_t stream)
{
    if (nbytes == 0) return 0;
    size_t blocks = (nbytes / 16 + 255) / 256;
    const size_t maxb = (size_t) num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    if (blocks == 0) blocks = 1;
    checksum_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (const unsigned char *) base, nbytes, (u64 *) sum);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ------------------------------------------------------------------------- */
/* order-independent 64-bit multiset hash of whole records (bench / test support)  */
/*
 * The reference's integrity guard is a signed-byte sum (above): it cannot see two payloads
 * swapped between keys or bytes permuted inside a record. The property checks of the tests
 * and of bench.py therefore compare, before and after a sort, the SUM and the XOR over all
 * records of h(record), h = the record's 8-byte little-endian words (the last one
 * zero-padded) folded through mix64: equal multisets of records give equal pairs, and any
 * change of a single record changes both words with overwhelming probability.
 * out[0] += sum, out[1] ^= xor (device u64[2], zeroed by the caller).
 * Restated in oracle/mpsort_oracle.py:multiset_hash.
 */
__host__ __device__ __forceinline__ u64 mset_mix64(u64 x)
{
    u64 z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define MPSK_MSET_SEED 0x243F6A8885A308D3ULL

__global__ void __launch_bounds__(256)
multiset_hash_kernel(const unsigned char * __restrict__ base, size_t n, size_t elsize, int mode, u64 * out)
{
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    u64 sum = 0, x = 0;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        u64 h = MPSK_MSET_SEED;
        if (mode == 16) {
            const uint4 v = ((const uint4 *) base)[i];
            h = mset_mix64(h ^ (((u64) v.y << 32) | v.x));
            h = mset_mix64(h ^ (((u64) v.w << 32) | v.z));
        } else if (mode == 8) {
            const u64 * w = (const u64 *) (base + i * elsize);
            for (size_t k = 0; k < elsize / 8; k++) h = mset_mix64(h ^ w[k]);
        } else {
            const unsigned char * r = base + i * elsize;
            for (size_t b = 0; b < elsize; b += 8) {
                u64 w = 0;
                for (size_t k = 0; k < 8 && b + k < elsize; k++) w |= (u64) r[b + k] << (8 * k);
                h = mset_mix64(h ^ w);
            }
        }
        sum += h;
        x ^= h;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(FULL_MASK, sum, o);
        x ^= __shfl_xor_sync(FULL_MASK, x, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *) out, (unsigned long long) sum);
        atomicXor((unsigned long long *) out + 1, (unsigned long long) x);
    }
}

extern "C" int mpsk_multiset_hash(const void * base, size_t n, size_t elsize, uint64_t * out, mpsk_stream_t stream)
{
    if (n == 0 || elsize == 0) return 0;
    int mode = 1;
    if (elsize == 16 && (((uintptr_t) base) & 15) == 0) mode = 16;
    else if (elsize % 8 == 0 && (((uintptr_t) base) & 7) == 0) mode = 8;
    size_t blocks = (n + 255) / 256;
    const size_t maxb = (size_t) num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    multiset_hash_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (const unsigned char *) base, n, elsize, mode, (u64 *) out);
    CUDA_LAUNCH_CHECK();
    return 0;
}
