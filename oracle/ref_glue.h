/*
 * ref_glue.h -- descriptor -> radix() callback glue used by the oracle drivers that
 * call the UNMODIFIED reference (TEST INFRASTRUCTURE). The callback writes what the
 * reference's Cython layer writes (binding.pyx:81-121): every key word widened to
 * 8 bytes, signed words sign-extended and biased by 2^63; rsize = 8 * nwords
 * (binding.pyx:209). With `raw` the key bytes are copied as they are and
 * rsize = width * nwords (the C callers' style, main-mpi.c:10-12).
 */
#ifndef REF_GLUE_H
#define REF_GLUE_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

struct ref_desc { size_t offset; unsigned width, nwords; int is_signed; int raw; };

static inline size_t ref_desc_rsize(const struct ref_desc * d)
{
    return d->raw ? (size_t) d->width * d->nwords : (size_t) 8 * d->nwords;
}

static void ref_desc_radix(const void * ptr, void * radix, void * arg)
{
    const struct ref_desc * d = (const struct ref_desc *) arg;
    const unsigned char * p = (const unsigned char *) ptr + d->offset;
    if (d->raw) { memcpy(radix, p, (size_t) d->width * d->nwords); return; }
    unsigned k;
    for (k = 0; k < d->nwords; k++) {
        uint64_t v = 0;
        memcpy(&v, p + (size_t) k * d->width, d->width);      /* little endian */
        if (d->is_signed) {
            const unsigned bits = 8 * d->width;
            if (bits < 64 && (v >> (bits - 1))) v |= ~0ULL << bits;   /* sign extend */
            v += 1ULL << 63;
        }
        memcpy((unsigned char *) radix + 8 * (size_t) k, &v, 8);
    }
}
#endif
