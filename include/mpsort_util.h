/*
 * mpsort_util.h -- bench / test support exported by libmpsort-b200.so.
 *
 * NOT part of the drop-in boundary (that is mpsort.h). These exist because the
 * product deliberately has no PyTorch/CuPy dependency: benches and tests need a way
 * to allocate device buffers, make synthetic records on the device, time with CUDA
 * events on the communicator's stream and check full-size outputs by properties.
 */
#ifndef MPSORT_B200_UTIL_H
#define MPSORT_B200_UTIL_H

#include <stddef.h>
#include <stdint.h>
#include "mpsort.h"

#ifdef __cplusplus
extern "C" {
#endif

int    mpsort_util_device_count(void);
/* "0000:1b:00.0"-style PCI address of a CUDA device (for NUMA placement of its host buffers); 0 on success */
int    mpsort_util_device_pci_bus_id(int device, char * buf, int len);
void * mpsort_util_dev_malloc(int device, size_t nbytes);
void   mpsort_util_dev_free(int device, void * ptr);
/* page-locked host memory; NULL if it cannot be pinned */
void * mpsort_util_host_malloc_pinned(size_t nbytes);
void   mpsort_util_host_free_pinned(void * ptr);
/* blocking copy in any direction (cudaMemcpyDefault) */
void   mpsort_util_memcpy(int device, void * dst, const void * src, size_t nbytes);
void   mpsort_util_dev_memset(int device, void * dst, int value, size_t nbytes);

/* synthetic records on the device; kind as in mpsk_generate:
 * 0 uniform u64 {key,tag}; 1 mostly sorted; 2 48-byte particle, skewed signed ids;
 * 3 uniform u64 key + tag, any elsize >= 16. Blocking. */
void   mpsort_util_generate(mpsort_comm_t comm, void * dst, size_t n, size_t elsize,
                            int kind, uint64_t seed);

/* the same records as rank `rank` of `nranks` would generate (one GPU can then hold the
 * input of a many-rank CPU run of the reference, chunk by chunk). Blocking. */
void   mpsort_util_generate_as(mpsort_comm_t comm, void * dst, size_t n, size_t elsize,
                               int kind, uint64_t seed, uint64_t rank, uint64_t nranks);

/* Order-independent 64-bit multiset hash of whole records of a device or host buffer (local
 * part): out2[0] = sum, out2[1] = xor over the records of h(record), h = the record's 8-byte
 * little-endian words (the last one zero-padded) folded through mix64. Two buffers hold the
 * same multiset of records iff (with overwhelming probability) both words agree; summed /
 * xored over ranks it is the "records preserved" check of the full-size runs, which the
 * reference's signed-byte sum cannot give (swapped payloads, permuted bytes). Blocking. */
void   mpsort_util_multiset_hash(mpsort_comm_t comm, const void * base, size_t n, size_t elsize,
                                 uint64_t * out2);

/* order check of this rank's sorted output; returns the number of adjacent
 * violations (key order, and with check_ties the tag order inside equal keys) and
 * stores the first/last packed key words into firstlast[2*nw]. Blocking. */
uint64_t mpsort_util_check_sorted(mpsort_comm_t comm, const void * base, size_t n, size_t elsize,
                                  const struct mpsort_radix_desc * desc,
                                  int check_ties, size_t tie_offset, uint64_t * firstlast);

/* the reference's byte checksum (mpsort-mpi.c:148-159) of a device or host buffer,
 * local part only. Blocking. */
uint64_t mpsort_util_checksum(mpsort_comm_t comm, const void * base, size_t nbytes);

/* CUDA-event stopwatch on the communicator's stream */
void * mpsort_util_event_create(mpsort_comm_t comm);
void   mpsort_util_event_record(mpsort_comm_t comm, void * event);
/* milliseconds between two recorded events; synchronises on `stop` */
double mpsort_util_event_elapsed_ms(mpsort_comm_t comm, void * start, void * stop);
void   mpsort_util_event_destroy(void * event);
void   mpsort_util_stream_sync(mpsort_comm_t comm);

/* write a buffer larger than L2 (126 MB) so the next timed kernel starts cold */
void   mpsort_util_flush_l2(mpsort_comm_t comm);

/* how many of this library's kernels were launched since the last reset */
uint64_t mpsort_util_launch_count(int reset);

/* Per-kernel-class device time, for roofline numbers: with timing on, every launch
 * of the library on this communicator is bracketed by CUDA events on its stream and
 * the durations are summed per class ("extract_hist", "onesweep_pass", "gather_keys",
 * "gather_records", "splitter", "checksum", "exchange"). Turning it on (or off)
 * resets the sums. kernel_times returns the number of classes. */
void   mpsort_util_kernel_timing(mpsort_comm_t comm, int on);
int    mpsort_util_kernel_times(mpsort_comm_t comm, const char ** names, double * ms,
                                uint64_t * launches, int max);

/* The SecondSort merge alone (replaces the second radix_sort, mpsort-mpi.c:597): `runs` holds p
 * sorted runs, run r = records [rdispl[r], rdispl[r+1]) with rdispl[0] == 0, all in device memory;
 * their stable p-way merge (ties: lower run first) goes to `out`. One GPU can thus time and
 * profile the merge of the 8-GPU configuration at full size. Blocking. Returns 0 if the merge ran,
 * 1 if these runs would take the radix re-sort fallback (multi-word keys, p > 32, tiny inputs). */
int    mpsort_util_merge_runs(mpsort_comm_t comm, int p, const void * runs, const int64_t * rdispl,
                              void * out, size_t elsize, const struct mpsort_radix_desc * desc);

/* free/total device memory in bytes */
void   mpsort_util_mem_info(int device, size_t * free_bytes, size_t * total_bytes);

#ifdef __cplusplus
}
#endif
#endif
