/*
 * mpsort_layout.c -- host arithmetic of the exchange layout. No CUDA in this file:
 * it is unit-tested on CPU (tests/test_host_logic.py) against the oracle.
 *
 * Replaces _solve_for_layout_mpi (mpsort-mpi.c:663-727) and the three 8-byte
 * Alltoalls around it (mpsort-mpi.c:450-464,487): every rank holds the full
 * p x (p-1) matrices of local counts (they were all-gathered), so each rank solves
 * the whole layout redundantly, the way the reference's OpenMP twin does
 * (mpsort-omp.c:143-218), and no transpose collective is needed.
 */
#include <stddef.h>
#include <stdint.h>

#include "mpsort_internal.h"

/*
 * For boundary b (1 <= b <= p-1; receiver ranks < b get everything below it):
 *   every source j sends its clt[j][b-1] keys that are < P_b below the boundary;
 *   the remaining deficit  C[b] - sum_j clt[j][b-1]  is filled with keys == P_b taken
 *   from source 0 first, then 1, ... each supplying at most cle - clt
 *   (greedy loop of mpsort-mpi.c:704-724, in j = 0..NTask-1 order).
 * cut[j][b] is the resulting first index in rank j's sorted array that goes to
 * rank >= b; cut[j][0] = 0, cut[j][p] = nmemb[j].
 */
/* psrc sources, pdst (virtual) destinations: cut[j * (pdst + 1) + b], clt/cle[j * (pdst - 1) + b - 1].
 * pdst > psrc when every rank's output is split into several consecutive parts so that
 * the exchange and the merge can be pipelined part by part. */
int mpsort_solve_layout2(int psrc, int pdst, const int64_t * C, const int64_t * clt, const int64_t * cle,
        const int64_t * nmemb, int64_t * cut)
{
    const int ns = pdst - 1;
    int j, b;
    for (j = 0; j < psrc; j++) {
        cut[(size_t) j * (pdst + 1) + 0] = 0;
        cut[(size_t) j * (pdst + 1) + pdst] = nmemb[j];
    }
    for (b = 1; b < pdst; b++) {
        int64_t sure = 0;
        for (j = 0; j < psrc; j++) sure += clt[(size_t) j * ns + (b - 1)];
        int64_t deficit = C[b] - sure;
        if (deficit < 0) return -1;   /* "more items than there should be" */
        for (j = 0; j < psrc; j++) {
            const int64_t lt = clt[(size_t) j * ns + (b - 1)];
            const int64_t supply = cle[(size_t) j * ns + (b - 1)] - lt;
            if (supply < 0) return -2; /* "less items than there should be" */
            int64_t take = supply <= deficit ? supply : deficit;
            cut[(size_t) j * (pdst + 1) + b] = lt + take;
            deficit -= take;
        }
        if (deficit != 0) return -3;   /* CLE did not bracket C: splitter was wrong */
    }
    /* sanity: cuts must be monotone per source (SendCount >= 0, mpsort-mpi.c:483-485) */
    for (j = 0; j < psrc; j++) {
        for (b = 0; b < pdst; b++) {
            if (cut[(size_t) j * (pdst + 1) + b] > cut[(size_t) j * (pdst + 1) + b + 1]) return -4;
        }
    }
    return 0;
}

int mpsort_solve_layout(int p, const int64_t * C, const int64_t * clt, const int64_t * cle,
        const int64_t * nmemb, int64_t * cut)
{
    return mpsort_solve_layout2(p, p, C, clt, cle, nmemb, cut);
}

/* Desired cumulative output counts, C[0] = 0, C[i+1] = C[i] + outnmemb[i]
 * (mpsort-mpi.c:643-645). */
void mpsort_cumulative_counts(int p, const int64_t * outnmemb, int64_t * C)
{
    int i;
    C[0] = 0;
    for (i = 0; i < p; i++) C[i + 1] = C[i] + outnmemb[i];
}

/*
 * Global key range over the non-empty ranks (mpsort-mpi.c:646-660) and, from it,
 * the first byte level (counted from the most significant byte of the packed key)
 * at which Pmin and Pmax differ: the byte-wise splitter descent starts there with
 * the common leading bytes as the prefix, the analogue of the reference bisecting
 * inside [Pmin, Pmax] (mpsort-mpi.c:383).
 * kmin/kmax: [p][nw] packed words (word nw-1 most significant). Returns the start
 * level in [0, 8*nw]; prefix[nw] receives the common bytes (others zero).
 */
int mpsort_key_range(int p, uint32_t nw, const int64_t * nmemb,
        const uint64_t * kmin, const uint64_t * kmax,
        uint64_t * Pmin, uint64_t * Pmax, uint64_t * prefix)
{
    int j, w, any = 0;
    for (w = 0; w < (int) nw; w++) { Pmin[w] = 0; Pmax[w] = 0; prefix[w] = 0; }
    for (j = 0; j < p; j++) {
        if (nmemb[j] == 0) continue;   /* "skip the rank, since it has no data" :647 */
        const uint64_t * lo = kmin + (size_t) j * nw;
        const uint64_t * hi = kmax + (size_t) j * nw;
        if (!any) {
            for (w = 0; w < (int) nw; w++) { Pmin[w] = lo[w]; Pmax[w] = hi[w]; }
            any = 1;
            continue;
        }
        int c = 0;
        for (w = (int) nw - 1; w >= 0 && c == 0; w--) c = (lo[w] > Pmin[w]) - (lo[w] < Pmin[w]);
        if (c < 0) for (w = 0; w < (int) nw; w++) Pmin[w] = lo[w];
        c = 0;
        for (w = (int) nw - 1; w >= 0 && c == 0; w--) c = (hi[w] > Pmax[w]) - (hi[w] < Pmax[w]);
        if (c > 0) for (w = 0; w < (int) nw; w++) Pmax[w] = hi[w];
    }
    /* all ranks empty: Pmin = Pmax = 0 (:657-660) -> nothing to descend */
    int level = 0;
    const int nlevels = 8 * (int) nw;
    while (level < nlevels) {
        const int byteidx = nlevels - 1 - level;
        const int wi = byteidx >> 3;
        const int sh = (byteidx & 7) * 8;
        const uint64_t a = (Pmin[wi] >> sh) & 0xff;
        const uint64_t b2 = (Pmax[wi] >> sh) & 0xff;
        if (a != b2) break;
        prefix[wi] |= a << sh;
        level++;
    }
    return level;
}
