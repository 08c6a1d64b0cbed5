/*
 * ref_radix_glue.c -- ctypes-callable wrapper around the UNMODIFIED reference
 * radix_sort (radixsort.c:35-44) with a descriptor-driven radix() callback
 * (TEST INFRASTRUCTURE).
 */
#include <stddef.h>
#include "ref_glue.h"

void radix_sort(void * base, size_t nmemb, size_t size,
        void (*radix)(const void * ptr, void * radix, void * arg), size_t rsize, void * arg);

void ref_radix_sort_desc(void * base, size_t nmemb, size_t elsize,
        size_t offset, unsigned width, unsigned nwords, int is_signed, int raw)
{
    struct ref_desc d = { offset, width, nwords, is_signed, raw };
    radix_sort(base, nmemb, elsize, ref_desc_radix, ref_desc_rsize(&d), &d);
}
