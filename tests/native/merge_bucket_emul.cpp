/*
 * Host emulation of the bucket path of merge_tile_bucket_kernel (mpsort_kernels.cu): the same
 * arithmetic (mp-sort_b200/csrc/mpsort_merge_bucket.cuh, compiled here for the host) run phase
 * by phase over the kernel's 512 "threads" x 8 items, with the kernel's shared-memory layout.
 * Test infrastructure only (tests/test_merge_bucket_emul.py); never linked into the product.
 *
 * What it proves without a GPU: whenever the spread test passes, the list of source positions the
 * bucket path produces is the stable merge order (key, then source position) -- for any arrival
 * order of the shared-memory atomics (`order` permutes the order in which records are counted).
 * What it cannot prove: barriers, launch configuration, register allocation.
 */
#include <stdint.h>
#include <string.h>
#include <vector>

#include "mpsort_merge_bucket.cuh"

namespace {
const uint32_t THREADS = 512, VT = 8, TILE = THREADS * VT;
}

/* returns 1 when the tile must take the merge-path rounds (spread test failed), 0 when osrc is
 * the merged order of the source positions, -1 on bad arguments */
extern "C" int mbk_emul_tile(const uint64_t * key, const uint32_t * src, uint32_t cnt,
                             uint64_t klo, uint64_t khi, const uint32_t * order, uint32_t * osrc)
{
    using namespace mbk;
    if (cnt > TILE) return -1;
    std::vector<u32> cntp(NBP, 0u), slots(THREADS, 0u), ssrc(TILE, 0u);
    std::vector<u64> skey(TILE, 0ull);
    const u32 sh = shift_for(klo, khi);
    bool bad = false;
    /* count: record i belongs to thread i % THREADS, item i / THREADS; the atomics arrive in `order` */
    for (u32 n = 0; n < cnt; n++) {
        const u32 i = order ? order[n] : n;
        const u32 tid = i % THREADS, k = i / THREADS;
        const bool ok = in_range(key[i], klo, khi);
        bad |= !ok;
        const u32 b = ok ? bucket_of(key[i], klo, sh) : 0u;
        const u32 s = cntp[padc(b < NB ? b : NB - 1)]++;
        slots[tid] |= (s < 15u ? s : 15u) << (4 * k);
    }
    /* scan: thread tid owns the 16 counters at 17 * tid */
    u32 mx = 0;
    std::vector<u32> tot(THREADS, 0u);
    for (u32 tid = 0; tid < THREADS; tid++)
        for (u32 j = 0; j < 16; j++) { const u32 c = cntp[17 * tid + j]; tot[tid] += c; mx = c > mx ? c : mx; }
    if (bad || mx > CMAX) return 1;
    u32 run = 0;
    for (u32 tid = 0; tid < THREADS; tid++) {
        u32 r = run;
        for (u32 j = 0; j < 16; j++) { const u32 c = cntp[17 * tid + j]; cntp[17 * tid + j] = r; r += c; }
        run += tot[tid];
    }
    /* scatter to bucket order */
    for (u32 i = 0; i < cnt; i++) {
        const u32 tid = i % THREADS, k = i / THREADS;
        const u32 pos = cntp[padc(bucket_of(key[i], klo, sh))] + ((slots[tid] >> (4 * k)) & 15u);
        if (pos >= cnt) return -1;
        skey[pos] = key[i];
        ssrc[pos] = src[i];
    }
    /* rank inside the bucket */
    for (u32 pos = 0; pos < cnt; pos++) {
        const u32 at = merged_position(cntp.data(), skey.data(), ssrc.data(), pos, cnt, klo, sh);
        if (at >= cnt) return -1;
        osrc[at] = ssrc[pos];
    }
    return 0;
}

extern "C" uint32_t mbk_emul_shift(uint64_t klo, uint64_t khi) { return mbk::shift_for(klo, khi); }
extern "C" uint32_t mbk_emul_bucket(uint64_t key, uint64_t klo, uint32_t sh) { return mbk::bucket_of(key, klo, sh); }
