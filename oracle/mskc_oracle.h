/*
 * mskc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, line-faithful restatement of the MulSKIPS KMC event loop
 * (reference: /root/reference/mulskips-source, "src/" below) with a run-time
 * box size.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or execute anything under oracle/.
 * The product (mulskips_b200/, include/mskc.h) never links or calls this.
 *
 * Parity status: PINNED against the reference's own golden vectors
 * (regression-tests/input-and-references-files/{I0000000[01],log}-<case>)
 * by tests/test_oracle_golden.py for the ten regression cases, i.e. the
 * event loop, the rate rules, both random streams, the writers and the
 * builders FS, F, C, S, I, FF, IN, LA; see DESIGN.md section "Oracle".
 * PARITY UNPINNED (the reference holds no golden output for them; they are
 * restated from the sources and exercised only against themselves and the
 * CUDA path): InitSt 'ST' (tabulated-temperature mode), the LaserAnnealing
 * fork variant (orc_set_variant), the builders Fw, Fa, SG with a non-zero Ge
 * fraction (libgfortran's RAND() is not part of the reference sources), A,
 * SS, D, Z, J, K, L, T, E, SH.
 */
#ifndef MSKC_ORACLE_H
#define MSKC_ORACLE_H
#include <stdint.h>
#include <stdio.h>
#include <setjmp.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/modules/defsystem.f:87-91 bit layout of a LattCoo word */
enum { ORC_PosIndex = 0, ORC_LenIndex = 8, ORC_PosT = 9, ORC_LenT = 13,
       ORC_PosOcc = 23, ORC_LenOcc = 1, ORC_PosCoor = 24, ORC_LenCoor = 3,
       ORC_PosSiC = 27, ORC_LenSiC = 3, ORC_PosSi = 27, ORC_PosC = 28 };

/* src/modules/defdertype.f:32-37 (MCParticle); 1-based neighbour slots stored 0-based */
typedef struct { int32_t xyz[3]; int32_t nn[4][3]; int32_t ev; double prob; } OrcMC;
/* src/modules/defdertype.f:26-29,39-48 (BulkParticle / Vacancy) */
typedef struct { int32_t xyz[3]; int32_t nn[4][3]; } OrcSite;

/* one record of the optional event trace */
typedef struct { int64_t iter; int32_t ind; int32_t site[3]; int32_t itrans; int32_t pad; double tree1; double time; } OrcTrace;

typedef struct Orc {
  /* box (src/modules/defsystem.f:27, compile-time there, run-time here) */
  int LenX, LenY, LenZ;
  int NumMcMax, NumVdMax, Levels, SizeTree;
  int32_t *LattCoo, *LattInd;          /* [(LenZ+1)][LenY][LenX], x fastest; +1 scratch cell */
  size_t ncell;                        /* LenX*LenY*(LenZ+1) */
  OrcMC *ListAdAtom;  int NumAdAtom;   /* 1-based, capacity NumMcMax+1, zero-initialised */
  OrcSite *ListAtom;  int NumAtoms;  size_t capAtoms;
  OrcSite *ListVoid;  int NumVoids;
  double *Tree;                        /* 1-based heap order */
  /* species */
  int NCrystal, NCov, ListCrystal[3], ListCov[3];
  int CountCrystal[3], CountCrystalOld[3], CountCov[3];
  /* rate tables, Fortran column-major flattening (src/modules/defsystem.f:60-68) */
  double PtransE[3 * 4096], PtransD[9], PtransAbs[12], PtransDes[12], PtransZig;
  /* start.dat scalars (src/modules/definitions.f) */
  char InitSt[4], RunType[4], SaveFinalState[4], SaveCoo[4], Homogeneous[4], ExitStrategy[8];
  int Len1, Len2, Len3, Len4;
  double sf, alloyfraction, Tm, P0, sigma, Tflex, LenVac, LenNuc;
  int tmax;
  /* LaserAnnealing fork "la/" (mulskips-source-LaserAnnealing): species at bits 28-29, event codes 1-2 / 3 / 4 / 4+k,
   * per-species LA law with the liquid-fraction coupling, occupied-ad-atom counter and the NumOcc0 stop rule
   * (la/modules/defsystem.f:36-37,46,82-83,99-100; la/main.f:634-726; la/mcevent/Get_Prob.f:142-152) */
  int la_tree;
  double Tm_k[3], P0_k[3], sigma_k[3], Tflex_k[3], LenSiGe;
  int Len5, NumAdAtomOcc, NumOcc, NumOcc0, NumOcc1, InitCountCrystal[3], CountLiquid[3];
  /* InitSt 'ST' (src/main.f:109-115, src/modules/defsystem.f:71-81): trench geometry + tabulated rates */
  double angle, Ttop, Tbottom;
  int fixedT, ntemp, nprob, reftempvalues[300];
  double *refPE, *refPD, *refPAbs, *refPDes;   /* [ntemp][12288], [ntemp][9], [ntemp][12], [ntemp][12] */
  int st_nnz; int *st_idx;                      /* PtransE entries that are non-zero in some block (the others interpolate to 0) */
  char cadfilename[2048], tempfilename[2048], coofilename[2048], restartfilename[2048];
  int64_t MaxIter, OutMolMol, Iter, IterOut, framecounter;
  double MaxTime, OutTime, Time;
  int Idum;
  /* ran3 state (src/mcevent/random.f:32-34, SAVEd) */
  int r_iff, r_inext, r_inextp, r_ma[56];
  int64_t counter;                     /* numberslist cursor (src/modules/definitions.f:35) */
  uint64_t alloy_lcg;                  /* state of the documented LCG that stands in for RAND() in the alloy builders */
  int32_t NBuff_evap[3];               /* Evaporation.f's un-initialised NBuff, kept across calls */
  /* bookkeeping added by the oracle */
  int64_t n_getprob, n_ran3, n_events[8];
  int64_t q_stale_nbuff, q_wl_iso, q_wl_cov, q_wl_dep, q_wl_des, q_oob, q_anom, q_far_bond;   /* reference-quirk counters */
  int faithful_cost;                   /* 1: also perform the 96 KiB PE/PD table copy per Get_Prob call */
  double *PEcopy;                      /* target of that copy */
  FILE *log;                           /* reference stdout stream (NULL = silent) */
  int write_files;                     /* write I*.xyz etc. into outdir */
  char outdir[1024];
  OrcTrace *trace; int64_t trace_cap, trace_n;
  int last_site[3];
  int last_arm;                        /* 0: last event made no IF_Arm_Choice, 1: it chose zig-zag, 2: arm-chair (goes to OrcTrace.pad) */
  int done;                            /* 1: loop exited after an event, 2: 'few MC particles' stop */
  int fatal; char errmsg[512]; jmp_buf jb; int jb_set;
  size_t scratch;                      /* index of the scratch cell used for out-of-bounds accesses */
} Orc;

Orc *orc_create(int lenx, int leny, int lenz, int nummcmax);
/* 1: follow the LaserAnnealing fork (la/) instead of mulskips-source (src/); call before orc_read_start */
void orc_set_variant(Orc *o, int la_tree);
void orc_destroy(Orc *o);
/* parse start.dat (src/main.f:83-219 incl. SetProbability); returns 0 ok */
int orc_read_start(Orc *o, const char *path);
/* build initial geometry per InitSt (src/main.f:225-295) and write frame 0 (src/main.f:297-300) */
int orc_setup(Orc *o);
/* one trip through the loop body src/main.f:305-369; returns 0 continue, 1 loop exit, <0 fatal */
int orc_iterate(Orc *o);
/* run until exit or nmax iterations; returns number of iterations executed (or <0 fatal) */
int64_t orc_run(Orc *o, int64_t nmax);
/* final frame + checkpoint + Coor (src/main.f:374-430) */
int orc_finish(Orc *o);
/* frame writer (src/inpout/WriteMolMolSource.f + WriteMolMolXYZFile.f) */
int orc_write_frame(Orc *o, double time, int64_t iter, int64_t iout);
/* helpers exported for tests */
double orc_ran3(Orc *o);
double orc_numberslist(int64_t ind);
void orc_find_nn(const int s1[3], const int s2[3], const int sc[3], int out[2][3], int lx, int ly);
void orc_find_sn(const int site[3], const int sitec[3], const int sn[3][3], int zig[3][3], int arm[3][3], int lx, int ly);
void orc_update_tree(Orc *o, int index, double p);
uint64_t orc_state_hash(const Orc *o);
int orc_sizeof_mc(void);

#ifdef __cplusplus
}
#endif
#endif
