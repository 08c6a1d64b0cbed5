/*
 * ref_driver.c -- runs the UNMODIFIED reference mpsort_mpi_newarray on per-rank
 * input files and writes per-rank output files (TEST INFRASTRUCTURE: this is how
 * golden vectors are produced and how the C/numpy restatements are validated).
 *
 *   mpirun-shim -np N ref_driver DIR ELSIZE OFFSET WIDTH NWORDS SIGNED RAW OPTIONS INPLACE
 * rank r reads DIR/in.r (nmemb = file size / ELSIZE) and, unless INPLACE, DIR/outn.r
 * (decimal outnmemb); writes DIR/out.r and, on rank 0, DIR/timers.txt.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "mpsort.h"
#include "ref_glue.h"

static void * slurp(const char * path, size_t * nbytes)
{
    FILE * f = fopen(path, "rb");
    if (!f) { perror(path); MPI_Abort(MPI_COMM_WORLD, 3); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void * p = malloc(n > 0 ? (size_t) n : 1);
    if (n > 0 && fread(p, 1, (size_t) n, f) != (size_t) n) { perror(path); MPI_Abort(MPI_COMM_WORLD, 3); }
    fclose(f);
    *nbytes = (size_t) n;
    return p;
}

int main(int argc, char ** argv)
{
    MPI_Init(&argc, &argv);
    int rank, size;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if (argc != 10) {
        if (rank == 0) fprintf(stderr, "usage: ref_driver DIR ELSIZE OFFSET WIDTH NWORDS SIGNED RAW OPTIONS INPLACE\n");
        MPI_Abort(MPI_COMM_WORLD, 2);
    }
    const char * dir = argv[1];
    const size_t elsize = (size_t) atol(argv[2]);
    struct ref_desc d;
    d.offset = (size_t) atol(argv[3]); d.width = (unsigned) atoi(argv[4]); d.nwords = (unsigned) atoi(argv[5]);
    d.is_signed = atoi(argv[6]); d.raw = atoi(argv[7]);
    const int options = atoi(argv[8]);
    const int inplace = atoi(argv[9]);
    char path[1024];
    size_t nbytes;
    snprintf(path, sizeof(path), "%s/in.%d", dir, rank);
    char * in = (char *) slurp(path, &nbytes);
    const size_t nmemb = nbytes / elsize;
    size_t outn = nmemb;
    char * out = in;
    if (!inplace) {
        snprintf(path, sizeof(path), "%s/outn.%d", dir, rank);
        FILE * f = fopen(path, "r");
        if (!f || fscanf(f, "%zu", &outn) != 1) { perror(path); MPI_Abort(MPI_COMM_WORLD, 3); }
        fclose(f);
        out = (char *) malloc(outn * elsize + 1);
    }
    mpsort_mpi_unset_options(-1);
    if (options) mpsort_mpi_set_options(options);
    mpsort_mpi_newarray(in, nmemb, out, outn, elsize, ref_desc_radix, ref_desc_rsize(&d), &d, MPI_COMM_WORLD);
    snprintf(path, sizeof(path), "%s/out.%d", dir, rank);
    FILE * f = fopen(path, "wb");
    if (!f) { perror(path); MPI_Abort(MPI_COMM_WORLD, 3); }
    if (outn && fwrite(out, elsize, outn, f) != outn) { perror(path); MPI_Abort(MPI_COMM_WORLD, 3); }
    fclose(f);
    if (rank == 0) {
        snprintf(path, sizeof(path), "%s/timers.txt", dir);
        if (freopen(path, "w", stdout)) { mpsort_mpi_report_last_run(); fflush(stdout); }
    }
    MPI_Finalize();
    return 0;
}
