/*
 * mpsort_kernels.h -- internal thin C ABI between the C host code (mpsort_host.c)
 * and the hand-written sm_100a kernels (mpsort_kernels.cu). Plain pointers and
 * sizes only. Every launcher enqueues on `stream` and returns without syncing;
 * a non-zero return is a cudaError_t.
 */
#ifndef MPSORT_KERNELS_H
#define MPSORT_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void * mpsk_stream_t; /* cudaStream_t */

#define MPSK_RADIX_BITS 8
#define MPSK_RADIX 256
#define MPSK_MAX_ITEMS ((size_t)0x3fffffffu) /* look-back status words carry 30-bit counts */

/* K1: key extraction + all-digit histogram. Replaces the reference's radix()
 * callbacks (binding.pyx:81-121, bench-mpi.c:13-15).
 * Packs key bytes [8*g, 8*g+8) of every record (little-endian over the nwords*width
 * key bytes, signed words sign-flipped) into kout[i] and accumulates the eight
 * 8-bit digit histograms of that word into hist[8][256] (must be zeroed by caller).
 * kout may be NULL (histogram only). `sub` is subtracted from every packed word before it
 * is stored and counted (range compression; kout may then alias base for 8-byte records).
 * minmax (device u64[2], initialised by the caller to {~0, 0}) receives min and max; may be NULL.
 */
int mpsk_extract_keys(const void * base, size_t n, size_t elsize,
        size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t g, uint64_t sub, uint64_t * kout, uint32_t * hist, uint64_t * minmax, mpsk_stream_t stream);

/* Record mode (elsize 8 or 16, key = one aligned u64 in the low or high half): counts of
 * the nh (4 or 8) digits d0 .. d0+nh-1 of (key ^ flip) are ADDED to hist[d][256], and the
 * OR of (key ^ key[0]) over all records is ORed into *diff (device u64, may be NULL):
 * byte b of *diff is zero exactly when digit b is the same in every key. `ref` is the record whose key
 * stands for key[0] (NULL: recs itself): an array counted in several calls, chunk by chunk as it arrives
 * from the host, passes the first record of the WHOLE array every time.
 * mpsk_rec_sample_diff does the OR over s evenly spaced records only (a preview). */
int mpsk_rec_histograms(const void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip,
        uint32_t d0, uint32_t nh, uint32_t * hist, uint64_t * diff, const void * ref, mpsk_stream_t stream);
int mpsk_rec_sample_diff(const void * recs, size_t n, size_t elsize, int key_in_high, uint32_t s,
        uint64_t * diff, mpsk_stream_t stream);

/* Exclusive scan of each of `nhist` 256-bin histograms: hist[h][b] -> bins[h][b]. */
int mpsk_scan_histograms(const uint32_t * hist, uint32_t * bins, int nhist, mpsk_stream_t stream);

/* Geometry of one onesweep pass (so the host can size the look-back buffer). */
size_t mpsk_onesweep_tile_items(void);
/* bytes of scratch needed for a pass over n items (look-back words + ticket) */
size_t mpsk_onesweep_scratch_bytes(size_t n);

/* K2: one stable 8-bit LSD onesweep pass on (u64 key, u32 value) pairs.
 * Replaces mpsort_qsort_r (stdlib/msort.c:177-314).
 * vin == NULL means "values are 0..n-1" (first pass).
 * vout may be NULL and kout may be NULL to drop that output (last pass variants).
 * bins = exclusive-scanned global histogram of this digit (256 words, device).
 * scratch = mpsk_onesweep_scratch_bytes(n) bytes; zeroed by this call.
 */
int mpsk_onesweep_pass(const uint64_t * kin, const uint32_t * vin,
        uint64_t * kout, uint32_t * vout, size_t n, int shift,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream);

/* The same pass over whole records: elsize 16 = {u64 key, u64 other} (key in the low or
 * the high half), elsize 8 = a bare u64 key. `flip` is XORed onto the key before the
 * digit is taken (the sign bit for signed keys). in/out aligned to elsize. */
int mpsk_onesweep_pass_rec(const void * in, void * out, size_t n, size_t elsize, int shift,
        int key_in_high, uint64_t flip, const uint32_t * bins, void * scratch, mpsk_stream_t stream);

/* Hybrid sort support (record mode). After stable passes over the high digits only,
 * records that agree in (key ^ flip) >> lobits form runs that are still in input
 * order; mpsk_fixup_rec orders every run of <= 256 records by the low `lobits`
 * bits in place and appends the start index of every longer run to worklist
 * (*nwork counts them, also beyond cap; both device memory, *nwork zeroed by the
 * caller). mpsk_fixup_extents turns starts into lengths. */
int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip, uint32_t lobits,
        uint32_t * worklist, uint32_t * nwork, uint32_t cap, size_t tile0, size_t ntiles, mpsk_stream_t stream);
/* records per fix-up tile. mpsk_fixup_rec handles the runs whose first record lies in tiles
 * [tile0, tile0 + ntiles) of the n records (ntiles = 0: all of them): once tiles 0 .. t-1 are done, records
 * 0 .. t * tile_items - 1 are final, so the output can leave for the host range by range. */
size_t mpsk_fixup_tile_items(void);
int mpsk_fixup_extents(const void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip, uint32_t lobits,
        const uint32_t * worklist, uint32_t nwork, uint32_t * lengths, mpsk_stream_t stream);
/* predictor of the hybrid sort: pairs[j] += the number of equal PAIRS (sum of k(k-1)/2 over values)
 * among the high parts (key ^ flip) >> lobits[j] of s records at pseudo-random positions, for j < nl <= 2
 * at once; pairs[nl] += the pairs of samples that drew the SAME position (they are in every pairs[j] too:
 * the caller takes them off). table: device u64[nl + 1][2 << log2_tsize], zeroed by the caller,
 * 1 << log2_tsize >= 2 s; pairs: device u64[3], zeroed by the caller. lobits is a host array. */
int mpsk_prefix_pairs(const void * recs, size_t n, size_t elsize, uint32_t s, int key_in_high, uint64_t flip,
        const uint32_t * lobits, uint32_t nl, uint64_t * table, uint32_t log2_tsize, uint64_t * pairs, mpsk_stream_t stream);

/* dst[i] = src[idx[i]] for 64-bit words (key words of multi-word keys). If hist is
 * non-NULL nothing is accumulated (histograms are permutation invariant). */
int mpsk_gather_u64(const uint64_t * src, const uint32_t * idx, uint64_t * dst,
        size_t n, mpsk_stream_t stream);

/* K3: payload gather out[i] = base[idx[i]] for elsize-byte records. Replaces the
 * record moves of the merge sort (msort.c:153-173,270-294). */
int mpsk_gather_records(const void * base, const uint32_t * idx, void * out,
        size_t n, size_t elsize, mpsk_stream_t stream);

/* How the splitter kernels see the locally sorted keys: word w of key i is the u64
 * at base + i*item_stride + w*word_stride, XOR flip, plus `add` for word 0 (keys stored
 * relative to their minimum by the range compression of single-word keys). SoA key arrays of the index
 * sort: {skeys, 8, n*8, 0}; sorted 16-byte records: {recs + offset, 16, 0, signbit}. */
struct mpsk_keyview { const void * base; size_t item_stride; size_t word_stride; uint64_t flip; uint64_t add; };

/* K4: splitter counting on the sorted keys. Replaces _histogram/_bsearch_last_lt/le
 * (internal-parallel.h:8-126).
 *
 * Byte-wise descent state: prefix[b][nw] (u64 words, device) holds the bytes of
 * splitter b decided so far (undecided bytes zero). `level` counts bytes from the
 * most significant one (0 .. 8*nw-1). For every splitter b and digit d the kernel
 * writes counts[b*256+d] = #local keys <= (prefix_b | d at this byte | 0xff below).
 */
int mpsk_splitter_count(struct mpsk_keyview view, size_t n, uint32_t nw,
        const uint64_t * prefix, int nsplit, int level,
        uint64_t * counts, mpsk_stream_t stream);

/* After the counts were summed over ranks: for each splitter pick the smallest d
 * with counts[b][d] >= target[b] and OR it into prefix[b] at byte `level`. */
int mpsk_splitter_select(const uint64_t * counts, const uint64_t * target,
        uint64_t * prefix, uint32_t nw, int nsplit, int level, mpsk_stream_t stream);

/* Final local counts for decided splitters: clt[b] = #keys < P_b, cle[b] = #keys <= P_b
 * written to out[0..nsplit) and out[nsplit..2*nsplit). */
int mpsk_splitter_final(struct mpsk_keyview view, size_t n, uint32_t nw,
        const uint64_t * prefix, int nsplit, uint64_t * out, mpsk_stream_t stream);

/* All levels of the descent in one kernel, the per-level sums over peer memory (the default of one
 * process per GPU). boxes[r] = rank r's mailbox of mpsk_peer_box_bytes() bytes as mapped here (zeroed once
 * at allocation; every rank pushes its counts into every peer's mailbox and polls its own);
 * seq = 256 * (number of earlier calls on this communicator); *err != 0 afterwards means a peer never
 * answered. prefix is updated in place. */
size_t mpsk_peer_box_bytes(void);
int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw,
        uint64_t * prefix, const uint64_t * target, int nsplit, int level0, int nlevels,
        uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err, mpsk_stream_t stream);

/* dst[i] += src_k[i] over nsrc sources (u64), for the in-process transport. */
int mpsk_sum_u64(uint64_t * dst, const uint64_t * const * srcs, int nsrc, size_t count,
        mpsk_stream_t stream);

/* K7: stable p-way merge of the received runs (replaces the second radix_sort,
 * mpsort-mpi.c:597). recv holds p sorted runs, run r = records [rdispl[r], rdispl[r+1]).
 * Single-word keys, p <= 32. S = sample stride, k = samples per tile with
 * (k + p) * S <= mpsk_merge_tile_items(); sstart[r] = first sample id of run r
 * (run r has (len_r / S) samples).
 *   mpsk_merge_samples  writes the sample keys skeys[sstart[p]] in (run, position) order;
 *   mpsk_merge_rank_samples orders them stably by key (sorted_skeys, sorted_sid = permutation);
 *   mpsk_merge_runs     computes cut[(ntiles+1)*p] and merges tile by tile into out.
 * *overflow (device, zeroed by the caller) counts tiles that exceeded the bound
 * (never happens; such a tile is left unwritten instead of corrupting memory).
 * self_recv != NULL: run self_run is NOT in recv -- its record i is at self_recv + (rdispl[self_run] + i) *
 * elsize (the same indexing from another base): a rank's own slice is merged from where the local sort left
 * it and never copied. Needs rdispl[p] < 2^31. For mpsk_merge_tile_items_for pass recv | self_recv (only the
 * alignment of the pointer is looked at). */
size_t mpsk_merge_tile_items(void);
/* tile size the merge will use for these buffers (16-byte records get their own kernel) */
size_t mpsk_merge_tile_items_for(const void * recv, const void * out, size_t elsize, size_t offset,
        uint32_t width, uint32_t nwords, uint32_t p);
int mpsk_merge_samples(const void * recv, size_t elsize, size_t offset, uint32_t width,
        uint32_t nwords, int is_signed, uint32_t p, uint32_t S, uint32_t k,
        const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        uint64_t * skeys, mpsk_stream_t stream);
/* sorted_skeys / sorted_sid = the samples in (key, run, position) order and their sample ids;
 * each run's samples are sorted already, so every sample ranks itself by p - 1 binary searches */
int mpsk_merge_rank_samples(const uint64_t * skeys, uint32_t p, const uint32_t * sstart,
        uint64_t * sorted_skeys, uint32_t * sorted_sid, mpsk_stream_t stream);
int mpsk_merge_runs(const void * recv, void * out, size_t elsize, size_t offset, uint32_t width,
        uint32_t nwords, int is_signed, uint32_t p, uint32_t S, uint32_t k,
        const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles,
        uint32_t * cut, uint32_t * overflow, mpsk_stream_t stream);

/* K6: the record exchange as peer stores (replaces MPI_Alltoallv and its sparse variant,
 * mp-mpiu.c:69-236, inside one box): segment k copies bytes[k] from src[k] (local) to
 * dst[k] (a peer's receive buffer mapped with CUDA IPC, or local memory); empty segments
 * are skipped; remote[k] != 0 marks peer destinations. One kernel, CTAs dealt to segments
 * by their expected time. */
int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes,
        const unsigned char * remote, int nseg, mpsk_stream_t stream);

/* CANDIDATE (MPSORT_FUSED_PACK=1): pack + exchange of index mode in one kernel. Segment k stores
 * the nrec[k] records base[idx[k][0 .. nrec[k])] (elsize bytes each) consecutively at dst[k] (a
 * peer's mapped receive buffer or local memory). */
int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst,
        const uint64_t * nrec, size_t elsize, int nseg, mpsk_stream_t stream);

/* K8: reference checksum (mpsort-mpi.c:148-159): sum of all bytes as SIGNED chars,
 * wrapping in 64 bits; accumulated (atomicAdd) into *sum which the caller zeroes. */
int mpsk_checksum(const void * base, size_t nbytes, uint64_t * sum, mpsk_stream_t stream);

/* ---- bench / test support (not on the product path) ---- */

/* Order-independent multiset hash of n whole records: out[0] += sum, out[1] ^= xor over the
 * records of h(record) (the record's 8-byte little-endian words, last one zero-padded, folded
 * through mix64). out = device u64[2], zeroed by the caller. */
int mpsk_multiset_hash(const void * base, size_t n, size_t elsize, uint64_t * out, mpsk_stream_t stream);

/* Synthetic records, SURVEY.md 8(d). kind:
 *  0 = uniform u64 key, 16-byte {key,payload}; payload = (rank<<40)+i
 *  1 = mostly sorted (1% perturbed) u64 key, 16-byte records
 *  2 = 48-byte particle struct, skewed signed i64 ID with heavy duplicates
 *  3 = uniform u64 key with arbitrary elsize (key at offset 0, payload u64 at 8 if room)
 */
int mpsk_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed,
        uint64_t rank, uint64_t nranks, mpsk_stream_t stream);

/* Order check of a sorted output: counts i with key[i-1] > key[i] (or, when
 * check_ties, key equal and payload u64 at `tie_offset` decreasing) into
 * *violations; also returns first/last packed key words through firstlast[2*nw]. */
int mpsk_check_sorted(const void * base, size_t n, size_t elsize,
        size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        int check_ties, size_t tie_offset,
        uint64_t * violations, uint64_t * firstlast, mpsk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
