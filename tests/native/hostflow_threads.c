/*
 * hostflow_threads.c -- TEST INFRASTRUCTURE ONLY: p rank threads drive the product's C host code
 * (linked against mock_device.c) through the C ABI, once as an in-process group and once as NCCL
 * ranks, and check the output order across ranks. Built with -fsanitize=thread by
 * tests/test_hostflow_mock.py: what is looked for are data races between rank threads in the host
 * code (shared option word, last-run timers, group slots), which would make the rank-thread tests
 * of the GPU suite flaky.
 * usage: hostflow_threads P N ELSIZE
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpsort.h"
#include "synth.h"

static int P;
static size_t N, E;
static mpsort_comm_t comms[64];
static unsigned char * in[64], * out[64];
static unsigned char uid[MPSORT_UNIQUE_ID_BYTES];
static int use_nccl;

static void * rank_main(void * arg)
{
    const int r = (int) (intptr_t) arg;
    struct mpsort_radix_desc d = { 0, 8, 1, 0, 0 };
    mpsort_comm_t c = use_nccl ? mpsort_comm_init_rank(r, P, uid, 0) : comms[r];
    int it;
    for (it = 0; it < 2; it++)
        mpsort_mpi_newarray_desc_impl(in[r], N + (size_t) r, out[r], N + (size_t) (P - 1 - r), E, &d, c, __LINE__, __FILE__);
    mpsort_comm_destroy(c);
    return NULL;
}

static int run(void)
{
    pthread_t th[64];
    int r, bad = 0;
    size_t i;
    uint64_t prev = 0, prevtag = 0;
    for (r = 0; r < P; r++) pthread_create(&th[r], NULL, rank_main, (void *) (intptr_t) r);
    for (r = 0; r < P; r++) pthread_join(th[r], NULL);
    for (r = 0; r < P; r++)
        for (i = 0; i < N + (size_t) (P - 1 - r); i++) {
            uint64_t k, t;
            memcpy(&k, out[r] + i * E, 8);
            memcpy(&t, out[r] + i * E + 8, 8);
            if (k < prev || (k == prev && (r || i) && t <= prevtag)) bad++;
            prev = k; prevtag = t;
        }
    return bad;
}

int main(int argc, char ** argv)
{
    int r, devices[64] = { 0 }, bad;
    size_t i;
    if (argc < 4) return 2;
    P = atoi(argv[1]); N = (size_t) atol(argv[2]); E = (size_t) atol(argv[3]);
    if (P < 1 || P > 64 || E < 16) return 2;
    for (r = 0; r < P; r++) {
        in[r] = (unsigned char *) malloc((N + P) * E);
        out[r] = (unsigned char *) malloc((N + P) * E);
        /* few distinct keys: ties across ranks, so the tag order is checked too */
        for (i = 0; i < N + (size_t) r; i++) {
            synth_record(in[r] + i * E, E, 0, 0x5EED0001, (uint64_t) r, (uint64_t) P, N, i);
            memset(in[r] + i * E + 2, 0, 6);
        }
    }
    mpsort_mpi_set_options(MPSORT_DISABLE_GATHER_SORT);
    if (mpsort_comm_init_local_group(P, devices, comms) != 0) return 3;
    bad = run();
    printf("in-process group: %d order violations\n", bad);
    if (mpsort_comm_get_unique_id(uid) != 0) return 3;
    use_nccl = 1;
    r = run();
    printf("NCCL ranks: %d order violations\n", r);
    bad += r;
    mpsort_mpi_report_last_run();
    printf(bad ? "THREADS FAILED\n" : "THREADS OK\n");
    return bad ? 1 : 0;
}
