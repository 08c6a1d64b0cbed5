/*
 * synth.h -- CPU restatement of the synthetic record generator of the device
 * library (mp-sort_b200/csrc/kernels/support.cuh: synth_record), so that the
 * reference and the oracle can be fed byte-identical inputs (TEST INFRASTRUCTURE).
 * Record kinds follow SURVEY.md 8(d):
 *   0 uniform u64 key | 1 mostly sorted (1% perturbed) | 2 skewed signed ids with a
 *   5% run of id 0 | 3 uniform (same as 0, any elsize)
 * Bytes [0,8) key, [8,16) tag = (rank << 40) + i, the rest hashed filler.
 */
#ifndef ORACLE_SYNTH_H
#define ORACLE_SYNTH_H
#include <stddef.h>
#include <stdint.h>

static inline uint64_t synth_mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline void synth_record(unsigned char * rec, size_t elsize, int kind, uint64_t seed,
                                uint64_t rank, uint64_t nranks, uint64_t n, uint64_t i)
{
    const uint64_t h = synth_mix64(seed ^ (rank << 32) ^ i);
    const uint64_t tag = (rank << 40) + i;
    uint64_t key;
    size_t b;
    if (kind == 1) {
        const uint64_t gi = rank * n + i;
        uint64_t src = gi;
        if (synth_mix64(gi ^ 0xA5A5A5A5ULL) % 100 == 0 && n > 0)
            src = (gi + 1 + synth_mix64(gi ^ 0x5A5A5A5AULL) % n) % (nranks * n);
        key = (src << 20) + (synth_mix64(seed ^ src) & 0xFFFFFULL);
    } else if (kind == 2) {
        const double u = (double) (h >> 11) * (1.0 / 9007199254740992.0);
        const double u2 = u * u;
        const double u4 = u2 * u2;
        long long id = (long long) (u4 * 16777216.0) - (1LL << 20);
        if (synth_mix64(h) % 20 == 0) id = 0;
        key = (uint64_t) id;
    } else {
        key = h;
    }
    for (b = 0; b < elsize; b++) {
        unsigned char v;
        if (b < 8) v = (unsigned char) (key >> (8 * b));
        else if (b < 16) v = (unsigned char) (tag >> (8 * (b - 8)));
        else v = (unsigned char) (synth_mix64(h + b / 8) >> (8 * (b & 7)));
        rec[b] = v;
    }
}
#endif
