/*
 * mpsort_oracle.c -- CPU restatement of MP-sort's distributed histogram sort, p
 * simulated ranks in one process. TEST INFRASTRUCTURE ONLY (see mpsort_oracle.h).
 *
 * Written from the reference's behaviour, function by function; every block names
 * the reference lines it follows (paths relative to MP-sort v0.1.19). It is a
 * restatement, not a copy: MPI calls become loops over the simulated ranks.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpsort_oracle.h"
#include "synth.h"

/* ------------------------------------------------------------------------- */
/* key algebra                                                               */

typedef struct {
    struct oracle_desc d;
    size_t elsize;
    size_t rsize;
} keyctx;

static size_t desc_rsize(const struct oracle_desc * d)
{
    return d->raw ? (size_t) d->width * d->nwords : (size_t) 8 * d->nwords;   /* binding.pyx:209 */
}

/* the radix() callback: binding.pyx:81-121 (u8/i8/u4/i4 widened to 8-byte words,
 * signed ones sign-extended and biased by 2^63; bench-mpi.c:13-15 is the i8 rule) */
static void key_of(const keyctx * k, const void * rec, unsigned char * radix)
{
    const unsigned char * p = (const unsigned char *) rec + k->d.offset;
    uint32_t w;
    if (k->d.raw) { memcpy(radix, p, k->rsize); return; }
    for (w = 0; w < k->d.nwords; w++) {
        uint64_t v = 0;
        memcpy(&v, p + (size_t) w * k->d.width, k->d.width);
        if (k->d.is_signed) {
            const unsigned bits = 8 * k->d.width;
            if (bits < 64 && (v >> (bits - 1))) v |= ~0ULL << bits;
            v += 1ULL << 63;
        }
        memcpy(radix + 8 * (size_t) w, &v, 8);
    }
}

/* _setup_radix_sort's choice of comparator (radixsort.c:146-195): native unsigned
 * compare for rsize 2/4/8 (:47-63), otherwise from the LAST byte (or 8-byte word
 * when rsize % 8 == 0) downwards on little-endian (:65-98). */
static int key_cmp(const keyctx * k, const unsigned char * a, const unsigned char * b)
{
    const size_t r = k->rsize;
    if (r == 2) { uint16_t x, y; memcpy(&x, a, 2); memcpy(&y, b, 2); return (x > y) - (x < y); }
    if (r == 4) { uint32_t x, y; memcpy(&x, a, 4); memcpy(&y, b, 4); return (x > y) - (x < y); }
    if (r == 8) { uint64_t x, y; memcpy(&x, a, 8); memcpy(&y, b, 8); return (x > y) - (x < y); }
    if (r % 8 == 0) {
        size_t i;
        for (i = r; i >= 8; i -= 8) {
            uint64_t x, y;
            memcpy(&x, a + i - 8, 8); memcpy(&y, b + i - 8, 8);
            if (x < y) return -1;
            if (x > y) return 1;
        }
        return 0;
    } else {
        size_t i;
        for (i = r; i >= 1; i--) {
            if (a[i - 1] < b[i - 1]) return -1;
            if (a[i - 1] > b[i - 1]) return 1;
        }
        return 0;
    }
}

/* the midpoint: u1 + ((u2 - u1) >> 1) for native sizes (radixsort.c:55-62), else a
 * multi-precision (a + b) >> 1 over little-endian bytes (:99-138) */
static void key_mid(const keyctx * k, unsigned char * out, const unsigned char * a, const unsigned char * b)
{
    const size_t r = k->rsize;
    if (r == 2) { uint16_t x, y, m; memcpy(&x, a, 2); memcpy(&y, b, 2); m = (uint16_t) (x + ((uint16_t) (y - x) >> 1)); memcpy(out, &m, 2); return; }
    if (r == 4) { uint32_t x, y, m; memcpy(&x, a, 4); memcpy(&y, b, 4); m = x + ((y - x) >> 1); memcpy(out, &m, 4); return; }
    if (r == 8) { uint64_t x, y, m; memcpy(&x, a, 8); memcpy(&y, b, 8); m = x + ((y - x) >> 1); memcpy(out, &m, 8); return; }
    unsigned carry = 0;
    size_t i;
    for (i = 0; i < r; i++) {              /* add, least significant byte first */
        const unsigned t = (unsigned) a[i] + b[i] + carry;
        carry = t >= 256;
        out[i] = (unsigned char) t;
    }
    for (i = r; i >= 1; i--) {             /* halve, most significant byte first */
        const unsigned t = out[i - 1] + carry * 256;
        carry = t & 1;
        out[i - 1] = (unsigned char) (t >> 1);
    }
}

/* ------------------------------------------------------------------------- */
/* local sort: radix_sort (radixsort.c:19-44) = glibc merge sort, stable because  */
/* a merge step takes the LEFT run on compar <= 0 (stdlib/msort.c:78,96,118,137)  */

static void merge_sort_idx(const keyctx * k, const unsigned char * R, size_t * idx, size_t * tmp, size_t n)
{
    if (n <= 1) return;
    const size_t n1 = n / 2, n2 = n - n1;          /* msort.c:60-63 */
    size_t * b1 = idx, * b2 = idx + n1;
    merge_sort_idx(k, R, b1, tmp, n1);
    merge_sort_idx(k, R, b2, tmp, n2);
    size_t i = 0, j = 0, o = 0;
    while (i < n1 && j < n2) {
        if (key_cmp(k, R + b1[i] * k->rsize, R + b2[j] * k->rsize) <= 0) tmp[o++] = b1[i++];
        else tmp[o++] = b2[j++];
    }
    while (i < n1) tmp[o++] = b1[i++];
    memcpy(idx, tmp, o * sizeof(size_t));           /* the rest of b2 is already in place */
}

static void local_sort(const keyctx * k, void * base, size_t n)
{
    if (n <= 1) return;
    unsigned char * R = (unsigned char *) calloc(n, k->rsize);
    size_t * idx = (size_t *) malloc(n * sizeof(size_t));
    size_t * tmp = (size_t *) malloc(n * sizeof(size_t));
    unsigned char * copy = (unsigned char *) malloc(n * k->elsize);
    size_t i;
    for (i = 0; i < n; i++) { key_of(k, (char *) base + i * k->elsize, R + i * k->rsize); idx[i] = i; }
    merge_sort_idx(k, R, idx, tmp, n);
    memcpy(copy, base, n * k->elsize);
    for (i = 0; i < n; i++) memcpy((char *) base + i * k->elsize, copy + idx[i] * k->elsize, k->elsize);
    free(R); free(idx); free(tmp); free(copy);
}

void oracle_radix_sort(void * base, size_t nmemb, size_t elsize, const struct oracle_desc * d)
{
    keyctx k;
    k.d = *d; k.elsize = elsize; k.rsize = desc_rsize(d);
    local_sort(&k, base, nmemb);
}

/* ------------------------------------------------------------------------- */
/* counting against splitters: _bsearch_last_lt/le + _histogram                  */
/* (internal-parallel.h:8-126). Returns counts, i.e. (last index) + 1.           */

static ptrdiff_t count_lt(const keyctx * k, const unsigned char * P, const char * base, size_t n)
{
    unsigned char t[k->rsize];
    size_t lo = 0, hi = n;                     /* first index with key >= P */
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        key_of(k, base + mid * k->elsize, t);
        if (key_cmp(k, t, P) < 0) lo = mid + 1; else hi = mid;
    }
    return (ptrdiff_t) lo;
}

static ptrdiff_t count_le(const keyctx * k, const unsigned char * P, const char * base, size_t n)
{
    unsigned char t[k->rsize];
    size_t lo = 0, hi = n;                     /* first index with key > P */
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        key_of(k, base + mid * k->elsize, t);
        if (key_cmp(k, t, P) <= 0) lo = mid + 1; else hi = mid;
    }
    return (ptrdiff_t) lo;
}

/* myCLT[i+1] = #keys < P[i], myCLE[i+1] = #keys <= P[i]; [0] = 0, [np+1] = n.
 * Like the reference each search starts at the previous splitter's offset
 * (internal-parallel.h:106-110), which assumes P is non-decreasing. */
static void histogram(const keyctx * k, const unsigned char * P, int np, const char * base, size_t n,
                      ptrdiff_t * myCLT, ptrdiff_t * myCLE)
{
    int it;
    myCLT[0] = 0; myCLE[0] = 0;
    for (it = 0; it < np; it++) {
        ptrdiff_t off = myCLT[it];
        myCLT[it + 1] = count_lt(k, P + (size_t) it * k->rsize, base + (size_t) off * k->elsize, n - (size_t) off) + off;
        off = myCLE[it];
        myCLE[it + 1] = count_le(k, P + (size_t) it * k->rsize, base + (size_t) off * k->elsize, n - (size_t) off) + off;
    }
    myCLT[np + 1] = (ptrdiff_t) n;
    myCLE[np + 1] = (ptrdiff_t) n;
}

/* ------------------------------------------------------------------------- */
/* mpsort_mpi_histogram_sort (mpsort-mpi.c:333-604) over L simulated leaders     */

static int histogram_sort(const keyctx * k, int L, char ** base, const size_t * n,
                          char ** out, const size_t * outn, int64_t * sendcount, int * rounds)
{
    const size_t r = k->rsize;
    int i, j, rc = 0;

    /* :369 FirstSort */
    for (i = 0; i < L; i++) local_sort(k, base[i], n[i]);

    /* :606-661 _find_Pmax_Pmin_C */
    unsigned char Pmax[r], Pmin[r];
    ptrdiff_t * C = (ptrdiff_t *) calloc((size_t) L + 1, sizeof(ptrdiff_t));
    size_t total = 0;
    memset(Pmax, 0, r);
    memset(Pmin, 0xff, r);
    for (i = 0; i < L; i++) {
        unsigned char lo[r], hi[r];
        C[i + 1] = C[i] + (ptrdiff_t) outn[i];
        total += n[i];
        if (n[i] == 0) continue;                       /* :647 */
        key_of(k, base[i], lo);
        key_of(k, base[i] + (n[i] - 1) * k->elsize, hi);
        if (key_cmp(k, hi, Pmax) > 0) memcpy(Pmax, hi, r);
        if (key_cmp(k, lo, Pmin) < 0) memcpy(Pmin, lo, r);
    }
    if (total == 0) { memset(Pmin, 0, r); memset(Pmax, 0, r); }   /* :657-660 */

    /* piter (internal-parallel.h:128-248) */
    const int np = L - 1;
    unsigned char * P = (unsigned char *) calloc((size_t) (np > 0 ? np : 1), r);
    unsigned char * Pleft = (unsigned char *) malloc((size_t) (np > 0 ? np : 1) * r);
    unsigned char * Pright = (unsigned char *) malloc((size_t) (np > 0 ? np : 1) * r);
    int * stable = (int *) calloc((size_t) (np > 0 ? np : 1), sizeof(int));
    int * narrow = (int *) calloc((size_t) (np > 0 ? np : 1), sizeof(int));
    ptrdiff_t * myCLT = (ptrdiff_t *) calloc((size_t) L * (L + 1), sizeof(ptrdiff_t));
    ptrdiff_t * myCLE = (ptrdiff_t *) calloc((size_t) L * (L + 1), sizeof(ptrdiff_t));
    ptrdiff_t * CLT = (ptrdiff_t *) calloc((size_t) L + 1, sizeof(ptrdiff_t));
    ptrdiff_t * CLE = (ptrdiff_t *) calloc((size_t) L + 1, sizeof(ptrdiff_t));
    for (i = 0; i < np; i++) { memcpy(Pleft + (size_t) i * r, Pmin, r); memcpy(Pright + (size_t) i * r, Pmax, r); }

    int iter = 0, done = 0;
    while (!done) {                                          /* mpsort-mpi.c:385-437 */
        iter++;
        for (i = 0; i < np; i++) {                           /* piter_bisect :168-198 */
            if (stable[i]) continue;
            if (narrow[i]) {
                memcpy(P + (size_t) i * r, Pright + (size_t) i * r, r);
                stable[i] = 1;
            } else {
                key_mid(k, P + (size_t) i * r, Pleft + (size_t) i * r, Pright + (size_t) i * r);
                if (key_cmp(k, P + (size_t) i * r, Pleft + (size_t) i * r) <= 0) narrow[i] = 1;
            }
        }
        memset(CLT, 0, sizeof(ptrdiff_t) * ((size_t) L + 1));
        memset(CLE, 0, sizeof(ptrdiff_t) * ((size_t) L + 1));
        for (j = 0; j < L; j++) {                            /* _histogram + 2 Allreduce :389-394 */
            histogram(k, P, np, base[j], n[j], myCLT + (size_t) j * (L + 1), myCLE + (size_t) j * (L + 1));
            for (i = 0; i <= L; i++) { CLT[i] += myCLT[(size_t) j * (L + 1) + i]; CLE[i] += myCLE[(size_t) j * (L + 1) + i]; }
        }
        for (i = 0; i < np; i++) {                           /* piter_accept :224-248 */
            if (CLT[i + 1] < C[i + 1] && C[i + 1] <= CLE[i + 1]) { stable[i] = 1; continue; }
            if (CLT[i + 1] >= C[i + 1]) memcpy(Pright + (size_t) i * r, P + (size_t) i * r, r);
            else memcpy(Pleft + (size_t) i * r, P + (size_t) i * r, r);
        }
        done = 1;
        for (i = 0; i < np; i++) if (!stable[i]) { done = 0; break; }
        if (iter > 8 * (int) r + 64) { rc = -10; break; }    /* cannot happen; guards the test suite */
    }
    if (rounds) *rounds = iter;

    /* final _histogram with the accepted P (:441) */
    for (j = 0; j < L; j++)
        histogram(k, P, np, base[j], n[j], myCLT + (size_t) j * (L + 1), myCLE + (size_t) j * (L + 1));

    /* LayDistr + _solve_for_layout_mpi (:450-464, :663-727): receiver t sees column
     * t+1 of every source's myCLT/myCLE, fills its deficit from sources 0..L-1 */
    ptrdiff_t * myC = (ptrdiff_t *) calloc((size_t) L * (L + 1), sizeof(ptrdiff_t));  /* [source][boundary] */
    for (i = 0; i < L && rc == 0; i++) {
        ptrdiff_t sure = 0, deficit;
        ptrdiff_t T_C[L];
        for (j = 0; j < L; j++) { T_C[j] = myCLT[(size_t) j * (L + 1) + i + 1]; sure += T_C[j]; }
        deficit = C[i + 1] - sure;
        for (j = 0; j < L; j++) {
            if (deficit == 0) break;
            if (deficit < 0) { rc = -1; break; }             /* "serious bug: more items" :707 */
            ptrdiff_t supply = myCLE[(size_t) j * (L + 1) + i + 1] - T_C[j];
            if (supply < 0) { rc = -2; break; }              /* "serious bug: less items" :713 */
            if (supply <= deficit) { T_C[j] += supply; deficit -= supply; }
            else { T_C[j] += deficit; deficit = 0; }
        }
        for (j = 0; j < L; j++) myC[(size_t) j * (L + 1) + i + 1] = T_C[j];     /* Alltoall :463 */
    }

    /* SendCount/RecvCount and the exchange (:483-592): receive buffers are laid out
     * in source-rank order */
    if (rc == 0) {
        char ** buf = (char **) calloc((size_t) L, sizeof(char *));
        size_t * fill = (size_t *) calloc((size_t) L, sizeof(size_t));
        for (i = 0; i < L; i++) buf[i] = (char *) malloc(outn[i] * k->elsize + 1);
        for (j = 0; j < L && rc == 0; j++) {                  /* source */
            for (i = 0; i < L; i++) {                         /* destination */
                const ptrdiff_t cnt = myC[(size_t) j * (L + 1) + i + 1] - myC[(size_t) j * (L + 1) + i];
                if (cnt < 0) { rc = -4; break; }
                if (sendcount) sendcount[(size_t) j * L + i] = cnt;
                if (fill[i] + (size_t) cnt > outn[i]) { rc = -5; break; }     /* totrecv mismatch :504 */
                memcpy(buf[i] + fill[i] * k->elsize, base[j] + (size_t) myC[(size_t) j * (L + 1) + i] * k->elsize,
                       (size_t) cnt * k->elsize);
                fill[i] += (size_t) cnt;
            }
        }
        for (i = 0; i < L && rc == 0; i++) if (fill[i] != outn[i]) rc = -5;
        if (rc == 0) {
            for (i = 0; i < L; i++) {
                memcpy(out[i], buf[i], outn[i] * k->elsize);
                local_sort(k, out[i], outn[i]);               /* SecondSort :597 */
            }
        }
        for (i = 0; i < L; i++) free(buf[i]);
        free(buf); free(fill);
    }
    free(C); free(P); free(Pleft); free(Pright); free(stable); free(narrow);
    free(myCLT); free(myCLE); free(CLT); free(CLE); free(myC);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* mpsort_mpi_newarray_impl (mpsort-mpi.c:161-331): segmenter + gather + sort    */

int oracle_mpsort(int p, void ** bases, const size_t * nmemb,
        void ** outs, const size_t * outnmemb, size_t elsize,
        const struct oracle_desc * d, int options,
        int64_t * sendcounts, int * rounds, int * nleaders)
{
    keyctx k;
    int i, g, rc;
    size_t total = 0, totalout = 0;
    k.d = *d; k.elsize = elsize; k.rsize = desc_rsize(d);
    for (i = 0; i < p; i++) { total += nmemb[i]; totalout += outnmemb[i]; }
    if (total != totalout) return -20;                        /* :225-232 */

    /* :234-256 segment size; DISABLE is tested last and wins */
    size_t avgsegsize = (size_t) p;
    if (avgsegsize * elsize > 4 * 1024 * 1024) avgsegsize = 4 * 1024 * 1024 / elsize;
    if (options & ORACLE_REQUIRE_GATHER_SORT) avgsegsize = total;
    if (options & ORACLE_DISABLE_GATHER_SORT) avgsegsize = 0;

    /* _MPIU_Segmenter_assign_colors (mp-mpiu.c:351-392) */
    int * color = (int *) malloc(sizeof(int) * (size_t) p);
    {
        size_t cur = 0, cur2 = 0;
        int current = 0;
        for (i = 0; i < p; i++) {
            cur += nmemb[i]; cur2 += outnmemb[i];
            color[i] = current;
            if (cur > avgsegsize || cur2 > avgsegsize) { cur = 0; cur2 = 0; current++; }
            if (nmemb[i] == 0 && outnmemb[i] == 0) color[i] = -1;   /* all empties: one extra group (:386-388, :446-449) */
        }
    }
    /* groups = runs of equal colour (contiguous rank ranges) + the group of empties;
     * leader = member with the most INPUT, lowest rank on ties (mp-mpiu.c:465, :242-269);
     * the Leaders communicator orders leaders by world rank (:469) */
    int * leader_of = (int *) malloc(sizeof(int) * (size_t) p);   /* world rank -> its leader's world rank */
    for (i = 0; i < p; i++) {
        int best = -1, j;
        for (j = 0; j < p; j++) {
            if (color[j] != color[i]) continue;
            if (best < 0 || nmemb[j] > nmemb[best]) best = j;
        }
        leader_of[i] = best;
    }
    int L = 0;
    int * leaders = (int *) malloc(sizeof(int) * (size_t) p);
    for (i = 0; i < p; i++) if (leader_of[i] == i) leaders[L++] = i;
    if (nleaders) *nleaders = L;

    /* gather each group on its leader (MPIU_Gather, group-rank = world-rank order) */
    char ** sbase = (char **) calloc((size_t) L, sizeof(char *));
    char ** sout = (char **) calloc((size_t) L, sizeof(char *));
    size_t * sn = (size_t *) calloc((size_t) L, sizeof(size_t));
    size_t * soutn = (size_t *) calloc((size_t) L, sizeof(size_t));
    int * gsize = (int *) calloc((size_t) L, sizeof(int));
    for (g = 0; g < L; g++) {
        for (i = 0; i < p; i++) if (leader_of[i] == leaders[g]) { sn[g] += nmemb[i]; soutn[g] += outnmemb[i]; gsize[g]++; }
        if (gsize[g] > 1) {
            size_t off = 0;
            sbase[g] = (char *) malloc(sn[g] * elsize + 1);
            sout[g] = (char *) malloc(soutn[g] * elsize + 1);
            for (i = 0; i < p; i++) if (leader_of[i] == leaders[g]) {
                memcpy(sbase[g] + off * elsize, bases[i], nmemb[i] * elsize);
                off += nmemb[i];
            }
        } else {
            sbase[g] = (char *) bases[leaders[g]];          /* sorted in place, :283-284 */
            sout[g] = (char *) outs[leaders[g]];
        }
    }
    int64_t * sc = sendcounts ? (int64_t *) calloc((size_t) L * L, sizeof(int64_t)) : NULL;
    rc = histogram_sort(&k, L, sbase, sn, sout, soutn, sc, rounds);
    if (sendcounts) {
        int a, b;
        memset(sendcounts, 0, sizeof(int64_t) * (size_t) p * p);
        for (a = 0; a < L; a++) for (b = 0; b < L; b++)
            sendcounts[(size_t) leaders[a] * p + leaders[b]] = sc[(size_t) a * L + b];
        free(sc);
    }
    /* scatter (MPIU_Scatter) */
    for (g = 0; g < L; g++) {
        if (gsize[g] > 1) {
            size_t off = 0;
            for (i = 0; i < p && rc == 0; i++) if (leader_of[i] == leaders[g]) {
                memcpy(outs[i], sout[g] + off * elsize, outnmemb[i] * elsize);
                off += outnmemb[i];
            }
            free(sbase[g]); free(sout[g]);
        }
    }
    free(color); free(leader_of); free(leaders); free(sbase); free(sout); free(sn); free(soutn); free(gsize);
    return rc;
}

/* checksum (mpsort-mpi.c:148-159): plain `char` is signed on x86 */
uint64_t oracle_checksum(const void * base, size_t nbytes)
{
    const signed char * p = (const signed char *) base;
    uint64_t sum = 0;
    size_t i;
    for (i = 0; i < nbytes; i++) sum += (uint64_t) (int64_t) p[i];
    return sum;
}

void oracle_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed, uint64_t rank, uint64_t nranks)
{
    size_t i;
    for (i = 0; i < n; i++) synth_record((unsigned char *) dst + i * elsize, elsize, kind, seed, rank, nranks, n, i);
}
