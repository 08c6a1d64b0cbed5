/* mpishim_layout.h -- layout of the shared file of the MPI shim (TEST INFRASTRUCTURE). */
#ifndef MPISHIM_LAYOUT_H
#define MPISHIM_LAYOUT_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define SHIM_RING 64
enum { SHIM_EMPTY = 0, SHIM_FULL = 1, SHIM_DONE = 2 };

struct shim_msg { int state; int ctx; int tag; int pad; size_t off; size_t nbytes; };
struct shim_ring {
    uint64_t head; char pad0[56];      /* consumer */
    uint64_t tail; char pad1[56];      /* producer */
    struct shim_msg m[SHIM_RING];
};
struct shim_rankctl { uint64_t consumed; char pad[56]; };
struct shim_hdr {
    int size;
    int abort_flag;
    int next_ctx;
    int pad;
    size_t arena_cap;     /* bytes per rank */
    size_t rings_off, ctl_off, arena_off, total;
};

static inline size_t shim_layout_bytes(int size, size_t arena_cap)
{
    size_t off = 4096;
    off += sizeof(struct shim_ring) * (size_t) size * size;
    off = (off + 4095) & ~(size_t) 4095;
    off += sizeof(struct shim_rankctl) * (size_t) size;
    off = (off + 4095) & ~(size_t) 4095;
    off += arena_cap * (size_t) size;
    return off;
}

/* the memory must be zero-filled (a fresh file or anonymous mapping) */
static inline void shim_layout_init(struct shim_hdr * h, int size, size_t arena_cap)
{
    size_t off = 4096;
    h->size = size;
    h->abort_flag = 0;
    h->next_ctx = 16;
    h->arena_cap = arena_cap;
    h->rings_off = off;
    off += sizeof(struct shim_ring) * (size_t) size * size;
    off = (off + 4095) & ~(size_t) 4095;
    h->ctl_off = off;
    off += sizeof(struct shim_rankctl) * (size_t) size;
    off = (off + 4095) & ~(size_t) 4095;
    h->arena_off = off;
    h->total = off + arena_cap * (size_t) size;
}
#endif
