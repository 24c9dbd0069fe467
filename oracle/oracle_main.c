/*
 * oracle_main.c -- CPU ORACLE command line (test infrastructure, NOT product code).
 * Behaves like the reference executable (src/main.f): reads ./start.dat, writes
 * I*.xyz etc. into the cwd, reference-style log on stdout.  The box size, a
 * compile-time constant in the reference (src/modules/defsystem.f:27), is given
 * with --box X Y Z (default 480 480 480 as the regression notebook assumes).
 */
#include "mskc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
  int bx = 480, by = 480, bz = 480, nmc = 0, faithful = 0, files = 1, quiet = 0, la_tree = 0; long long maxiter = -1;
  const char *start = "start.dat";
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--box") && i + 3 < argc) { bx = atoi(argv[i + 1]); by = atoi(argv[i + 2]); bz = atoi(argv[i + 3]); i += 3; }
    else if (!strcmp(argv[i], "--nummcmax") && i + 1 < argc) nmc = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--faithful-cost")) faithful = 1;
    else if (!strcmp(argv[i], "--no-files")) files = 0;
    else if (!strcmp(argv[i], "--quiet")) quiet = 1;
    else if (!strcmp(argv[i], "--la-tree")) la_tree = 1;
    else if (!strcmp(argv[i], "--start") && i + 1 < argc) start = argv[++i];
    else if (!strcmp(argv[i], "--maxiter") && i + 1 < argc) maxiter = atoll(argv[++i]);
    else { fprintf(stderr, "usage: %s [--box X Y Z] [--nummcmax N] [--faithful-cost] [--no-files] [--quiet] [--la-tree] [--start f] [--maxiter n]\n", argv[0]); return 2; }
  }
  Orc *o = orc_create(bx, by, bz, nmc);
  if (!o) { fprintf(stderr, "allocation failed\n"); return 2; }
  orc_set_variant(o, la_tree);        /* follow mulskips-source-LaserAnnealing instead of mulskips-source */
  o->log = quiet ? NULL : stdout; o->write_files = files; o->faithful_cost = faithful;
  if (orc_read_start(o, start)) { fprintf(stderr, "start.dat: %s\n", o->errmsg); return 3; }
  if (maxiter > 0 && o->MaxIter > maxiter) o->MaxIter = maxiter;
  if (orc_setup(o)) return 3;
  double t0 = now();
  int64_t n = orc_run(o, INT64_MAX);
  double t1 = now();
  if (n < 0) return 3;
  orc_finish(o);
  fprintf(stderr, "oracle: %lld events in %.3f s = %.4g events/s; Get_Prob calls %lld (%.4g site-updates/s); ran3 draws %lld; quirks: stale_nbuff=%lld wl_iso=%lld wl_cov=%lld wl_dep=%lld wl_des=%lld oob=%lld anom=%lld\n",
          (long long)n, t1 - t0, n / (t1 - t0), (long long)o->n_getprob, o->n_getprob / (t1 - t0), (long long)o->n_ran3,
          (long long)o->q_stale_nbuff, (long long)o->q_wl_iso, (long long)o->q_wl_cov, (long long)o->q_wl_dep, (long long)o->q_wl_des, (long long)o->q_oob, (long long)o->q_anom);
  orc_destroy(o);
  return 0;
}
