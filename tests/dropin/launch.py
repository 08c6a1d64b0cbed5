"""mpirun stand-in for the drop-in drivers: `launch.py -np N prog args...` starts N processes of
`prog`, one per GPU (RANK / WORLD_SIZE / LOCAL_RANK, and the job name the facade uses for its NCCL-id
and message files). Exit code: the first non-zero one. TEST / INTEGRATION GLUE."""
import os
import subprocess
import sys


def main():
    if len(sys.argv) < 4 or sys.argv[1] != "-np":
        print(__doc__)
        return 2
    n = int(sys.argv[2])
    job = "/dev/shm/mpsort_dropin_%d" % os.getpid()
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(n), LOCAL_RANK=str(r), MPSORT_DROPIN_JOB=job)
        env.pop("MPSORT_DROPIN_THREADS", None)
        procs.append(subprocess.Popen(sys.argv[3:], env=env))
    rc = 0
    for p in procs:
        rc = rc or p.wait()
    for f in os.listdir("/dev/shm"):
        if f.startswith(os.path.basename(job)):
            os.unlink(os.path.join("/dev/shm", f))
    return rc


if __name__ == "__main__":
    sys.exit(main())
