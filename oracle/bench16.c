/*
 * bench16.c -- sibling of the reference's bench-mpi.c for struct records
 * (TEST INFRASTRUCTURE / CPU baseline harness). bench-mpi.c itself sorts bare 8-byte
 * integers (bench-mpi.c:13-15,196-211); BASELINE.json's configs are 16- and 48-byte
 * records, so this driver feeds the UNMODIFIED reference library
 * (mpsort_mpi_newarray, mpsort.h:44-47) the same synthetic records the device
 * bench uses (synth.h) and reports bench-mpi's numbers: total wall time and the
 * per-phase timers of mpsort_mpi_report_last_run.
 *
 *   mpirun-shim -np P bench16 [-k kind] [-e elsize] [-r reps] [-w warmups] [-t seconds] [-o dir]
 *                             [-s seed] [-g|-G] N_per_rank
 *   -w W     the first W repetitions are warm-ups: left out of mean_seconds (not of best_seconds)
 *   -t SEC   stop repeating once SEC seconds of sorting have been spent (at least W + 1 repetitions run)
 *   -T TOTAL the ranks share TOTAL records (rank r gets TOTAL / P, plus one if r < TOTAL % P);
 *            N_per_rank is then ignored
 *   -o DIR   every rank writes its sorted output of the LAST repetition to DIR/out.<rank>
 *            (repetition `it` uses seed + it: with -r 1 the input is exactly the device generator's)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <mpi.h>
#include "mpsort.h"
#include "ref_glue.h"
#include "synth.h"

int main(int argc, char ** argv)
{
    MPI_Init(&argc, &argv);
    int rank, size, opt;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    int kind = 0, reps = 1, warm = 0;
    double budget = 0.0;
    const char * outdir = NULL;
    long long total = 0;
    size_t elsize = 16;
    uint64_t seed = 0x5EED0001ULL;
    while (-1 != (opt = getopt(argc, argv, "k:e:r:s:w:t:o:T:gG"))) {
        switch (opt) {
            case 'k': kind = atoi(optarg); break;
            case 'e': elsize = (size_t) atol(optarg); break;
            case 'r': reps = atoi(optarg); break;
            case 's': seed = strtoull(optarg, NULL, 0); break;
            case 'w': warm = atoi(optarg); break;
            case 't': budget = atof(optarg); break;
            case 'o': outdir = optarg; break;
            case 'T': total = atoll(optarg); break;
            case 'g': mpsort_mpi_set_options(MPSORT_DISABLE_GATHER_SORT); break;
            case 'G': mpsort_mpi_set_options(MPSORT_REQUIRE_GATHER_SORT); break;
            default: MPI_Abort(MPI_COMM_WORLD, 2);
        }
    }
    if (optind >= argc || elsize < 16) {
        if (rank == 0) fprintf(stderr, "usage: bench16 [-k kind] [-e elsize>=16] [-r reps] [-s seed] N_per_rank\n");
        MPI_Abort(MPI_COMM_WORLD, 2);
    }
    const size_t n = total > 0 ? (size_t) (total / size + (rank < total % size ? 1 : 0)) : (size_t) atoll(argv[optind]);
    const double ntot = total > 0 ? (double) total : (double) n * size;
    struct ref_desc d = { 0, 8, 1, kind == 2, 0 };
    unsigned char * src = (unsigned char *) malloc(n * elsize + 1);
    unsigned char * dst = (unsigned char *) malloc(n * elsize + 1);
    double best = 1e30, spent = 0.0, sum_timed = 0.0;
    int it, done = 0, ntimed = 0;
    for (it = 0; it < reps; it++) {
        size_t i;
        for (i = 0; i < n; i++) synth_record(src + i * elsize, elsize, kind, seed + (uint64_t) it, (uint64_t) rank, (uint64_t) size, n, i);
        MPI_Barrier(MPI_COMM_WORLD);
        const double t0 = MPI_Wtime();
        mpsort_mpi_newarray(src, n, dst, n, elsize, ref_desc_radix, ref_desc_rsize(&d), &d, MPI_COMM_WORLD);
        MPI_Barrier(MPI_COMM_WORLD);
        const double t1 = MPI_Wtime();
        /* local order check */
        for (i = 1; i < n; i++) {
            uint64_t a, b;
            ref_desc_radix(dst + (i - 1) * elsize, &a, &d);
            ref_desc_radix(dst + i * elsize, &b, &d);
            if (a > b) { fprintf(stderr, "bench16: local order broken on rank %d\n", rank); MPI_Abort(MPI_COMM_WORLD, 4); }
        }
        if (t1 - t0 < best) best = t1 - t0;
        if (it >= warm) { sum_timed += t1 - t0; ntimed++; }
        done = it + 1;
        if (rank == 0) {
            printf("MPSort total time: %g\n", t1 - t0);
            mpsort_mpi_report_last_run();
        }
        /* the time budget: rank 0's clock decides for everyone */
        spent += t1 - t0;
        int stop = (budget > 0.0 && spent >= budget && it >= warm) ? 1 : 0;
        MPI_Bcast(&stop, 1, MPI_INT, 0, MPI_COMM_WORLD);
        if (stop) break;
    }
    if (outdir) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/out.%d", outdir, rank);
        FILE * f = fopen(path, "wb");
        if (!f || fwrite(dst, elsize, n, f) != n) { fprintf(stderr, "bench16: cannot write %s\n", path); MPI_Abort(MPI_COMM_WORLD, 5); }
        fclose(f);
    }
    if (rank == 0)
        printf("BENCH16 np=%d n_per_rank=%zu elsize=%zu kind=%d reps=%d reps_done=%d reps_timed=%d best_seconds=%.6f "
               "mean_seconds=%.6f records_per_second=%.1f mean_records_per_second=%.1f\n",
               size, n, elsize, kind, reps, done, ntimed, best, ntimed ? sum_timed / ntimed : best,
               ntot / best, ntot / (ntimed ? sum_timed / ntimed : best));
    free(src); free(dst);
    MPI_Finalize();
    return 0;
}
