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
