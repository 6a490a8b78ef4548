/*
 * mpirun.c — launcher for the single-node MPI shim: `mpirun -n P prog args...`.
 * Stands in for the reference's `mpirun --oversubscribe -n P bin/main.out ...`
 * (reference scripts/tests.sh:7-8) and `srun` (scripts/run.sh:62-66) on a box
 * without MPI.  Creates the shared segment, forks P ranks with
 * PHPC_MPI_{SHM,RANK,SIZE} (+ LOCAL_RANK for the rank -> GPU mapping), waits,
 * and tears everything down if any rank fails.
 *
 * Options: -n/-np/-c P; --oversubscribe, --allow-run-as-root, --bind-to X,
 * --map-by X are accepted and ignored.  PHPC_GPU_POLICY=visible sets
 * CUDA_VISIBLE_DEVICES=<rank % PHPC_GPUS> per rank (needed by the unmodified
 * reference build, which uses every visible GPU in every rank).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "mpi.h"

static pid_t g_pids[64];
static int g_n = 0;
static char g_path[256];

static void kill_all(int sig) {
  for (int i = 0; i < g_n; ++i)
    if (g_pids[i] > 0) kill(g_pids[i], sig);
}

static void on_signal(int sig) {
  kill_all(SIGTERM);
  phpc_mpi_segment_unlink(g_path);
  _exit(128 + sig);
}

int main(int argc, char **argv) {
  int n = 1, i = 1;
  for (; i < argc; ++i) {
    if (!strcmp(argv[i], "-n") || !strcmp(argv[i], "-np") || !strcmp(argv[i], "-c")) {
      if (i + 1 >= argc) break;
      n = atoi(argv[++i]);
    } else if (!strcmp(argv[i], "--oversubscribe") || !strcmp(argv[i], "--allow-run-as-root")) {
    } else if (!strcmp(argv[i], "--bind-to") || !strcmp(argv[i], "--map-by")) {
      ++i;
    } else
      break;
  }
  if (i >= argc || n < 1 || n > 64) {
    fprintf(stderr, "Usage: %s -n <ranks 1..64> [--oversubscribe] <program> [args...]\n", argv[0]);
    return 2;
  }
  const char *dir = access("/dev/shm", W_OK) == 0 ? "/dev/shm" : "/tmp";
  snprintf(g_path, sizeof g_path, "%s/phpc_mpi_%d_%ld", dir, (int)getpid(), (long)random());
  if (n > 1 && phpc_mpi_segment_create(g_path, n) != 0) {
    fprintf(stderr, "mpirun: cannot create %s: %s\n", g_path, strerror(errno));
    return 1;
  }
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);

  const char *policy = getenv("PHPC_GPU_POLICY");
  const char *gpus_env = getenv("PHPC_GPUS");
  const int gpus = gpus_env ? atoi(gpus_env) : 0;
  g_n = n;
  for (int r = 0; r < n; ++r) {
    pid_t pid = fork();
    if (pid < 0) {
      perror("mpirun: fork");
      kill_all(SIGTERM);
      phpc_mpi_segment_unlink(g_path);
      return 1;
    }
    if (pid == 0) {
      char buf[64];
      if (n > 1) setenv("PHPC_MPI_SHM", g_path, 1);
      snprintf(buf, sizeof buf, "%d", r);
      setenv("PHPC_MPI_RANK", buf, 1);
      setenv("LOCAL_RANK", buf, 1);
      snprintf(buf, sizeof buf, "%d", n);
      setenv("PHPC_MPI_SIZE", buf, 1);
      if (policy && !strcmp(policy, "visible") && gpus > 0) {
        snprintf(buf, sizeof buf, "%d", r % gpus);
        setenv("CUDA_VISIBLE_DEVICES", buf, 1);
      }
      execvp(argv[i], argv + i);
      fprintf(stderr, "mpirun: cannot exec %s: %s\n", argv[i], strerror(errno));
      _exit(127);
    }
    g_pids[r] = pid;
  }
  int rc = 0, left = n;
  while (left > 0) {
    int st = 0;
    pid_t pid = wait(&st);
    if (pid < 0) {
      if (errno == EINTR) continue;
      break;
    }
    --left;
    for (int r = 0; r < n; ++r)
      if (g_pids[r] == pid) g_pids[r] = 0;
    const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
    if (code != 0 && rc == 0) {
      rc = code;
      /* one rank failed: the job is over.  Ranks that abort together (MPI_Abort on every rank,
       * rank 0 printing the reason) get half a second to finish on their own first. */
      for (int spin = 0; spin < 50 && left > 0; ++spin) {
        int st2 = 0;
        pid_t p2 = waitpid(-1, &st2, WNOHANG);
        if (p2 > 0) {
          --left;
          for (int r = 0; r < n; ++r)
            if (g_pids[r] == p2) g_pids[r] = 0;
        } else {
          usleep(10000);
        }
      }
      kill_all(SIGTERM);
    }
  }
  if (n > 1) phpc_mpi_segment_unlink(g_path);
  return rc;
}
