/*
 * mpirun_shim.c -- launcher of the single-host MPI shim (TEST INFRASTRUCTURE):
 *   mpirun-shim -np N [--] prog args...
 * Creates the shared file (MPISHIM_DIR, default /dev/shm, else /tmp), starts N
 * processes with MPISHIM_FILE / MPISHIM_RANK / MPISHIM_SIZE set, and waits. If a
 * rank fails, the rest are told to stop (abort flag) and then killed.
 * MPISHIM_ARENA_MB (default 2048) is each rank's send arena; the file is sparse.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include "mpishim_layout.h"

static char g_path[512];
static void cleanup(void) { if (g_path[0]) unlink(g_path); }
static void on_signal(int sig) { cleanup(); _exit(128 + sig); }

int main(int argc, char ** argv)
{
    int np = 1, i = 1;
    while (i < argc) {
        if ((!strcmp(argv[i], "-np") || !strcmp(argv[i], "-n")) && i + 1 < argc) { np = atoi(argv[i + 1]); i += 2; }
        else if (!strcmp(argv[i], "--oversubscribe")) { i++; }
        else if (!strcmp(argv[i], "--")) { i++; break; }
        else break;
    }
    if (i >= argc || np < 1 || np > 1024) {
        fprintf(stderr, "usage: mpirun-shim -np N [--] prog args...\n");
        return 2;
    }
    const char * dir = getenv("MPISHIM_DIR");
    if (!dir) dir = (access("/dev/shm", W_OK) == 0) ? "/dev/shm" : "/tmp";
    const char * amb = getenv("MPISHIM_ARENA_MB");
    const size_t arena = (size_t) (amb ? atol(amb) : 2048) << 20;
    snprintf(g_path, sizeof(g_path), "%s/mpishim.%d.XXXXXX", dir, (int) getpid());
    int fd = mkstemp(g_path);
    if (fd < 0) { perror("mpirun-shim: mkstemp"); return 2; }
    atexit(cleanup);
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);
    const size_t total = shim_layout_bytes(np, arena);
    if (ftruncate(fd, (off_t) total) != 0) { perror("mpirun-shim: ftruncate"); return 2; }
    struct shim_hdr * h = (struct shim_hdr *) mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (h == MAP_FAILED) { perror("mpirun-shim: mmap"); return 2; }
    shim_layout_init(h, np, arena);

    pid_t * pids = (pid_t *) calloc((size_t) np, sizeof(pid_t));
    int r;
    for (r = 0; r < np; r++) {
        pid_t pid = fork();
        if (pid < 0) { perror("mpirun-shim: fork"); h->abort_flag = 1; break; }
        if (pid == 0) {
            char buf[64];
            setenv("MPISHIM_FILE", g_path, 1);
            snprintf(buf, sizeof(buf), "%d", r); setenv("MPISHIM_RANK", buf, 1);
            snprintf(buf, sizeof(buf), "%d", np); setenv("MPISHIM_SIZE", buf, 1);
            close(fd);
            execvp(argv[i], argv + i);
            perror("mpirun-shim: exec");
            _exit(127);
        }
        pids[r] = pid;
    }
    int failed = 0, left = r, status;
    while (left > 0) {
        pid_t pid = wait(&status);
        if (pid < 0) { if (errno == EINTR) continue; break; }
        left--;
        for (r = 0; r < np; r++) if (pids[r] == pid) pids[r] = 0;
        const int bad = !(WIFEXITED(status) && WEXITSTATUS(status) == 0);
        if (bad && !failed) {
            failed = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status);
            if (!failed) failed = 1;
            h->abort_flag = 1;
            usleep(200000);
            for (r = 0; r < np; r++) if (pids[r] > 0) kill(pids[r], SIGKILL);
        }
    }
    return failed;
}
