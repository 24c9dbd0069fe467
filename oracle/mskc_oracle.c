/*
 * mskc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Line-faithful plain-C restatement of the MulSKIPS KMC event loop.  Every
 * function cites the reference file:line it follows ("src/" =
 * /root/reference/mulskips-source).  Nothing under mulskips_b200/ or include/
 * links, loads or calls this file; see oracle/mskc_oracle.h for who may.
 *
 * Build with -ffp-contract=off (no FMA contraction) so that the double
 * arithmetic is the plain IEEE sequence gfortran -O emits for the reference.
 */
#include "mskc_oracle.h"
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ util */
static void orc_fatal(Orc *o, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(o->errmsg, sizeof o->errmsg, fmt, ap); va_end(ap);
  o->fatal = 1;
  if (o->log) fprintf(o->log, " %s\n", o->errmsg);
  fprintf(stderr, "oracle STOP: %s\n", o->errmsg);
  if (o->jb_set) longjmp(o->jb, 1);
  exit(3);
}
#define LOG(o, ...) do { if ((o)->log) fprintf((o)->log, __VA_ARGS__); } while (0)

/* species bits and event codes of the two source trees: src/ (modules/defsystem.f:90-91, main.f:335-357) and the older
 * LaserAnnealing fork la/ (la/modules/defsystem.f:99-100, la/main.f:671-699), selected at run time by o->la_tree */
#define PSI(o) ((o)->la_tree ? 28 : 27)
#define PC_(o) ((o)->la_tree ? 29 : 28)
#define PSIC(o) ((o)->la_tree ? 28 : 27)
#define LSIC(o) ((o)->la_tree ? 2 : 3)
#define EV_EVAP(o) ((o)->la_tree ? 3 : 4)
#define EV_DES(o) ((o)->la_tree ? 4 : 5)
static inline int IBITS(int32_t w, int pos, int len) { return (int)(((uint32_t)w >> pos) & ((1u << len) - 1u)); }
static inline void MVBITS(int from, int len, int32_t *to, int pos) {
  uint32_t m = ((1u << len) - 1u);
  *to = (int32_t)(((uint32_t)*to & ~(m << pos)) | (((uint32_t)from & m) << pos));
}
/* cell index with the bounds the Fortran arrays have (0:LenX-1,0:LenY-1,0:LenZ); out of
 * range -> scratch cell + counter (the reference would read foreign memory there) */
static inline size_t CI(Orc *o, int x, int y, int z) {
  if ((unsigned)x >= (unsigned)o->LenX || (unsigned)y >= (unsigned)o->LenY || (unsigned)z > (unsigned)o->LenZ) {
    o->q_oob++; o->LattCoo[o->scratch] = 0; o->LattInd[o->scratch] = 0; return o->scratch;
  }
  return (size_t)x + (size_t)o->LenX * ((size_t)y + (size_t)o->LenY * (size_t)z);
}
#define LC(o, s) ((o)->LattCoo[CI((o), (s)[0], (s)[1], (s)[2])])
#define LI(o, s) ((o)->LattInd[CI((o), (s)[0], (s)[1], (s)[2])])
static inline int eq3(const int32_t a[3], const int32_t b[3]) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
static inline void cp3(int32_t d[3], const int32_t s[3]) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
static inline void cpnn(int32_t d[4][3], const int32_t s[4][3]) { memcpy(d, s, 12 * sizeof(int32_t)); }
static inline int recipr(const int32_t nn[4][3], const int32_t s[3]) {
  int r = 0; for (int j = 0; j < 4; j++) if (eq3(nn[j], s)) r = 1; return r;
}
/* record fetches with the capacity guards the static Fortran arrays provide */
static const int32_t ZERO_NN[4][3] = {{0}};
static const int32_t (*mc_nn(Orc *o, int ind))[3] {
  if (ind < 1 || ind > o->NumMcMax) { o->q_oob++; return ZERO_NN; }
  return (const int32_t (*)[3])o->ListAdAtom[ind].nn;
}
static const int32_t (*bulk_nn(Orc *o, int ind))[3] {
  if (ind < 1 || (size_t)ind >= o->capAtoms) return ZERO_NN;   /* zero-initialised static storage */
  return (const int32_t (*)[3])o->ListAtom[ind].nn;
}
static const int32_t (*void_nn(Orc *o, int ind))[3] {
  if (ind < 1 || ind > o->NumVdMax) { o->q_oob++; return ZERO_NN; }
  return (const int32_t (*)[3])o->ListVoid[ind].nn;
}

/* src/modules/defsystem.f:104-105 (JumpDia, column j = bond j) */
static const int JumpDia[8][3] = {{3, 3, 3}, {3, -3, -3}, {-3, -3, 3}, {-3, 3, -3},
                                  {3, 3, -3}, {3, -3, 3}, {-3, -3, -3}, {-3, 3, 3}};

/* Fortran column-major flattening of PTransE(3,0:3,0:3,0:3,0:3,0:3,0:3), defsystem.f:61-62 */
static inline int IDXE(int sp, int i1, int i2, int i3, int i4, int i5, int i6) {
  return (sp - 1) + 3 * (i1 + 4 * (i2 + 4 * (i3 + 4 * (i4 + 4 * (i5 + 4 * i6)))));
}
#define IDXD(sp, c) (((sp) - 1) + 3 * ((c) - 1))     /* PTransD(3,1:3)   defsystem.f:60 */
#define IDXA(k, c) (((k) - 1) + 3 * (c))             /* PTransAbs/Des(3,0:3) defsystem.f:63-64 */

/* ------------------------------------------------------------------ RNG */
/* src/mcevent/random.f:22-69 (Knuth subtractive ran3; fac is a REAL(4) constant widened) */
double orc_ran3(Orc *o) {
  const int mbig = 1000000000, mseed = 161803398, mz = 0;
  const double fac = (double)(1.0f / 1.0e9f);
  int mj, mk, i, ii, k;
  int *ma = o->r_ma;                       /* 1-based */
  o->n_ran3++;
  if (o->Idum < 0 || o->r_iff == 0) {
    o->r_iff = 1;
    mj = mseed - abs(o->Idum);
    mj = mj % mbig;
    ma[55] = mj;
    mk = 1;
    for (i = 1; i <= 54; i++) {
      ii = (21 * i) % 55;
      ma[ii] = mk;
      mk = mj - mk;
      if (mk < mz) mk += mbig;
      mj = ma[ii];
    }
    for (k = 1; k <= 4; k++)
      for (i = 1; i <= 55; i++) {
        ma[i] = ma[i] - ma[1 + (i + 30) % 55];
        if (ma[i] < mz) ma[i] += mbig;
      }
    o->r_inext = 0;
    o->r_inextp = 31;
    o->Idum = 1;
  }
  o->r_inext++;  if (o->r_inext == 56) o->r_inext = 1;
  o->r_inextp++; if (o->r_inextp == 56) o->r_inextp = 1;
  mj = ma[o->r_inext] - ma[o->r_inextp];
  if (mj < mz) mj += mbig;
  ma[o->r_inext] = mj;
  return mj * fac;
}

/* src/mcevent/PickTreeEvent.f:72-1080: REAL(4) literals stored in REAL(8) */
static const float NUMBERS[1000] = {
#include "numberslist.inc"
};
double orc_numberslist(int64_t ind) { return (double)NUMBERS[(ind - 1) % 1000]; }

/* ------------------------------------------------------------------ tree */
/* src/mcevent/UpdateTree.f:25-66 */
void orc_update_tree(Orc *o, int Index, double Patom) {
  if (Patom == 0 && Index < o->NumAdAtom) orc_fatal(o, "error stabilizzato  da dentro Update! %d", Index);
  else if (Index > o->NumAdAtom) orc_fatal(o, "Index > NumAdAtom %d %d", Index, o->NumAdAtom);
  int Pos = Index + (1 << (o->Levels - 1)) - 1;
  double *Tree = o->Tree;
  Tree[Pos] = Patom;
  for (int Parent = o->Levels - 1; Parent >= 1; Parent--) {
    int NeighPos = (Pos & 1) ? Pos - 1 : Pos + 1;
    int NewPos = Pos >> 1;
    Tree[NewPos] = Tree[Pos] + Tree[NeighPos];
    Pos = NewPos;
  }
}

/* src/mcevent/PickTreeEvent.f:25-70 */
static void pick_tree_event(Orc *o, int *Index, double *Patom) {
  int idx = 1;
  double *Tree = o->Tree, currrand, P = 0;
  for (int Parent = 1; Parent <= o->Levels - 1; Parent++) {
    if (o->RunType[0] == 'T') { o->counter++; currrand = orc_numberslist(o->counter); }
    else currrand = orc_ran3(o);
    P = currrand * Tree[idx];
    idx <<= 1;
    if (Tree[idx] < P) idx++;
  }
  P = Tree[idx];
  if (idx - (1 << (o->Levels - 1)) + 1 > o->NumAdAtom)
    orc_fatal(o, "write Index> NumAdAtom %d %d %.17g", idx - (1 << (o->Levels - 1)) + 1, o->NumAdAtom, Tree[1]);
  *Index = idx - (1 << (o->Levels - 1)) + 1;
  *Patom = P;
}

/* ------------------------------------------------------------------ lists */
/* src/mcevent/Add.f:24-53 */
static void AddMC(Orc *o, const int32_t Site[3], const int32_t NextN[4][3], int ev, double Prob) {
  if (Prob == 0) orc_fatal(o, "provo ad aggiungere prob nulla");
  o->NumAdAtom++;
  if (o->NumAdAtom >= o->NumMcMax) orc_fatal(o, "troppe particelle cinetiche %d %d", o->NumAdAtom, o->NumMcMax);
  LI(o, Site) = o->NumAdAtom;
  OrcMC *r = &o->ListAdAtom[o->NumAdAtom];
  cp3(r->xyz, Site); cpnn(r->nn, NextN); r->ev = ev; r->prob = Prob;
  orc_update_tree(o, o->NumAdAtom, Prob);
}
/* src/mcevent/Add.f:58-76 */
static void AddBulk(Orc *o, const int32_t Site[3], const int32_t NextN[4][3]) {
  o->NumAtoms++;
  if ((size_t)o->NumAtoms + 1 >= o->capAtoms) {
    size_t nc = o->capAtoms * 2;
    o->ListAtom = (OrcSite *)realloc(o->ListAtom, nc * sizeof(OrcSite));
    memset(o->ListAtom + o->capAtoms, 0, (nc - o->capAtoms) * sizeof(OrcSite));
    o->capAtoms = nc;
  }
  LI(o, Site) = o->NumAtoms;
  cp3(o->ListAtom[o->NumAtoms].xyz, Site); cpnn(o->ListAtom[o->NumAtoms].nn, NextN);
}
/* src/mcevent/Add.f:82-100 */
static void AddVoid(Orc *o, const int32_t Site[3], const int32_t NextN[4][3]) {
  o->NumVoids++;
  if (o->NumVoids >= o->NumVdMax) orc_fatal(o, "troppe vacanze %d", o->NumVoids);
  LI(o, Site) = o->NumVoids;
  cp3(o->ListVoid[o->NumVoids].xyz, Site); cpnn(o->ListVoid[o->NumVoids].nn, NextN);
}
/* src/mcevent/Erase.f:24-78 */
static void EraseMC(Orc *o, const int32_t Site[3]) {
  if (Site[2] == -1) orc_fatal(o, "accesso a latt fuori bond");
  int IndNN = LI(o, Site);
  if (IndNN > o->NumMcMax || IndNN == 0) orc_fatal(o, "err: EraseMC %d %d %d %d", IndNN, Site[0], Site[1], Site[2]);
  if (IndNN >= o->NumAdAtom) {
    memset(&o->ListAdAtom[IndNN], 0, sizeof(OrcMC));
    LI(o, Site) = 0;
    orc_update_tree(o, IndNN, 0.0);
    o->NumAdAtom--;
  } else {
    o->ListAdAtom[IndNN] = o->ListAdAtom[o->NumAdAtom];
    int ind = o->NumAdAtom + (1 << (o->Levels - 1)) - 1;
    orc_update_tree(o, IndNN, o->Tree[ind]);
    orc_update_tree(o, o->NumAdAtom, 0.0);
    int32_t SiteTemp[3]; cp3(SiteTemp, o->ListAdAtom[o->NumAdAtom].xyz);
    LI(o, SiteTemp) = IndNN;
    memset(&o->ListAdAtom[o->NumAdAtom], 0, sizeof(OrcMC));
    LI(o, Site) = 0;
    o->NumAdAtom--;
  }
}
/* src/mcevent/Erase.f:84-128 */
static void EraseBulk(Orc *o, const int32_t Site[3]) {
  if (Site[2] == -1) orc_fatal(o, "accesso a latt fuori bond");
  int IndNN = LI(o, Site);
  if (IndNN == 0 || IndNN > o->NumAtoms) orc_fatal(o, "err: EraseBulk %d %d %d %d", IndNN, Site[0], Site[1], Site[2]);
  if (IndNN == o->NumAtoms) {
    memset(&o->ListAtom[o->NumAtoms], 0, sizeof(OrcSite));
    LI(o, Site) = 0;
    o->NumAtoms--;
  } else {
    o->ListAtom[IndNN] = o->ListAtom[o->NumAtoms];
    int32_t SiteTemp[3]; cp3(SiteTemp, o->ListAtom[o->NumAtoms].xyz);
    LI(o, SiteTemp) = IndNN;
    memset(&o->ListAtom[o->NumAtoms], 0, sizeof(OrcSite));
    LI(o, Site) = 0;
    o->NumAtoms--;
  }
}
/* src/mcevent/Erase.f:135-179 */
static void EraseVoid(Orc *o, const int32_t Site[3]) {
  if (Site[2] == -1) orc_fatal(o, "accesso a latt fuori bond");
  int IndNN = LI(o, Site);
  if (IndNN > o->NumVdMax || IndNN == 0) orc_fatal(o, "err: EraseVoid %d %d %d %d", IndNN, Site[0], Site[1], Site[2]);
  if (IndNN == o->NumVoids) {
    memset(&o->ListVoid[o->NumVoids], 0, sizeof(OrcSite));
    LI(o, Site) = 0;
    o->NumVoids--;
  } else {
    o->ListVoid[IndNN] = o->ListVoid[o->NumVoids];
    int32_t SiteTemp[3]; cp3(SiteTemp, o->ListVoid[o->NumVoids].xyz);
    LI(o, SiteTemp) = IndNN;
    memset(&o->ListVoid[o->NumVoids], 0, sizeof(OrcSite));
    LI(o, Site) = 0;
    o->NumVoids--;
  }
}

/* ------------------------------------------------------------------ geometry */
static inline int isign(int a, int b) { return b >= 0 ? abs(a) : -abs(a); }
static inline int nint_d(double x) { return (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); }

/* src/lattice/FIND_NN.f:25-77 */
void orc_find_nn(const int Site1[3], const int Site2[3], const int SiteC[3], int NewSite[2][3], int LX, int LY) {
  const double Pi = (double)3.14159265358979323f;   /* REAL(4) literal widened, FIND_NN.f:27 */
  int Site11[3], Site22[3];
  double ax[3], Rot[3][3] /* Rot[l][k] == Fortran Rot(l+1,k+1) */, snorm, Cth, Sth, R[3];
  for (int k = 0; k < 3; k++) { Site11[k] = Site1[k]; Site22[k] = Site2[k]; }
  if (abs(SiteC[0] - Site11[0]) > LX / 2) Site11[0] += isign(LX, SiteC[0] - Site11[0]);
  if (abs(SiteC[1] - Site11[1]) > LY / 2) Site11[1] += isign(LY, SiteC[1] - Site11[1]);
  if (abs(SiteC[0] - Site22[0]) > LX / 2) Site22[0] += isign(LX, SiteC[0] - Site22[0]);
  if (abs(SiteC[1] - Site22[1]) > LY / 2) Site22[1] += isign(LY, SiteC[1] - Site22[1]);
  for (int k = 0; k < 3; k++) ax[k] = (double)SiteC[k] - (double)Site11[k];
  snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int k = 0; k < 3; k++) ax[k] = ax[k] / snorm;
  for (int irot = 1; irot <= 2; irot++) {
    Cth = cos((double)(2 * irot) * Pi / 3.0);
    Sth = sin((double)(2 * irot) * Pi / 3.0);
    for (int i = 0; i < 3; i++) Rot[i][i] = ax[i] * ax[i] * (1 - Cth) + Cth;
    Rot[1][0] = ax[0] * ax[1] * (1 - Cth) - ax[2] * Sth;   /* Rot(2,1) */
    Rot[0][1] = ax[0] * ax[1] * (1 - Cth) + ax[2] * Sth;   /* Rot(1,2) */
    Rot[2][0] = ax[0] * ax[2] * (1 - Cth) + ax[1] * Sth;   /* Rot(3,1) */
    Rot[0][2] = ax[0] * ax[2] * (1 - Cth) - ax[1] * Sth;   /* Rot(1,3) */
    Rot[2][1] = ax[1] * ax[2] * (1 - Cth) - ax[0] * Sth;   /* Rot(3,2) */
    Rot[1][2] = ax[1] * ax[2] * (1 - Cth) + ax[0] * Sth;   /* Rot(2,3) */
    for (int k = 0; k < 3; k++) {
      R[k] = 0.0;
      for (int l = 0; l < 3; l++) R[k] = R[k] + Rot[l][k] * ((double)Site22[l] - (double)SiteC[l]);
    }
    for (int k = 0; k < 3; k++) NewSite[irot - 1][k] = nint_d(R[k]) + SiteC[k];
    if (NewSite[irot - 1][0] >= LX) NewSite[irot - 1][0] -= LX;
    if (NewSite[irot - 1][0] < 0) NewSite[irot - 1][0] += LX;
    if (NewSite[irot - 1][1] >= LY) NewSite[irot - 1][1] -= LY;
    if (NewSite[irot - 1][1] < 0) NewSite[irot - 1][1] += LY;
  }
}

/* src/lattice/FIND_SN.f:26-84 */
void orc_find_sn(const int Site[3], const int SiteC_inp[3], const int SNSite_inp[3][3], int Zig[3][3], int Arm[3][3], int LX, int LY) {
  int SNSite[3][3], SiteC[3];
  double Cent[3], ax[3], snorm;
  memcpy(SNSite, SNSite_inp, sizeof SNSite);
  for (int k = 0; k < 3; k++) SiteC[k] = SiteC_inp[k];
  if (abs(Site[0] - SiteC[0]) > LX / 2) SiteC[0] += isign(LX, Site[0] - SiteC[0]);
  if (abs(Site[1] - SiteC[1]) > LY / 2) SiteC[1] += isign(LY, Site[1] - SiteC[1]);
  for (int i = 0; i < 3; i++) {
    if (abs(Site[0] - SNSite[i][0]) > LX / 2) SNSite[i][0] += isign(LX, Site[0] - SNSite[i][0]);
    if (abs(Site[1] - SNSite[i][1]) > LY / 2) SNSite[i][1] += isign(LY, Site[1] - SNSite[i][1]);
  }
  for (int k = 0; k < 3; k++) Cent[k] = 0.5 * ((double)SiteC[k] + (double)(float)Site[k]);
  for (int k = 0; k < 3; k++) ax[k] = (double)Site[k] - (double)SiteC[k];
  snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int k = 0; k < 3; k++) ax[k] = ax[k] / snorm;
  const double ARM = (double)8.6602254037f;           /* REAL(4) literal, FIND_SN.f:67 */
  for (int i = 0; i < 3; i++) {
    for (int k = 0; k < 3; k++) {
      double rz = 2 * Cent[k] - (double)SNSite[i][k];
      double ra = (double)SNSite[i][k] + ARM * ax[k];
      Zig[i][k] = nint_d(rz);
      Arm[i][k] = nint_d(ra);
    }
    if (Zig[i][0] >= LX) Zig[i][0] -= LX;
    if (Zig[i][0] < 0) Zig[i][0] += LX;
    if (Zig[i][1] >= LY) Zig[i][1] -= LY;
    if (Zig[i][1] < 0) Zig[i][1] += LY;
    if (Arm[i][0] >= LX) Arm[i][0] -= LX;
    if (Arm[i][0] < 0) Arm[i][0] += LX;
    if (Arm[i][1] >= LY) Arm[i][1] -= LY;
    if (Arm[i][1] < 0) Arm[i][1] += LY;
  }
}

/* src/mcevent/IF_Arm_Choice.f:21-56 */
static int IF_Arm_Choice(Orc *o, int NZig[3][3], int NArm[3][3]) {
  int CoorArm = 0, CoorZig = 0;
  for (int i = 0; i < 3; i++) {
    CoorArm += IBITS(LC(o, NArm[i]), ORC_PosCoor, ORC_LenCoor);
    CoorZig += IBITS(LC(o, NZig[i]), ORC_PosCoor, ORC_LenCoor);
  }
  if (CoorZig > CoorArm) return 0;
  if (CoorZig == CoorArm) return (orc_ran3(o) <= o->PtransZig) ? 0 : 1;
  return 1;
}

/* ------------------------------------------------------------------ rates */
static const double KB = 0.00008617330350;        /* Get_Prob.f:70 */
/* LA melting rate of one table entry, Get_Prob.f:143-144 */
static inline double la_PE(Orc *o, int temperature, double e) {
  return o->P0 * exp(-e / (KB * temperature)) * (e != 0.0 ? 1.0 : 0.0);
}
/* LA solidification rate, Get_Prob.f:146-148 */
static inline double la_PD(Orc *o, int temperature, double d) {
  return 0.5 * o->P0 * exp(-d / (KB * o->Tm)) * (1.0 + erf((temperature - o->Tflex) / o->sigma)) * (d != 0.0 ? 1.0 : 0.0);
}
/* la/ fork, la/mcevent/Get_Prob.f:142-152: per-species parameters, solidification scaled by the liquid fraction of
 * species k (FLOAT() is REAL(4); evaluation left to right) */
static inline double la_PE_k(Orc *o, int temperature, int k, double e) {
  return o->P0_k[k - 1] * exp(-e / (KB * temperature)) * (e != 0.0 ? 1.0 : 0.0);
}
static inline double la_PD_k(Orc *o, int temperature, int k, double d) {
  int sum = o->CountLiquid[0] + o->CountLiquid[1] + o->CountLiquid[2];
  return 0.5 * o->P0_k[k - 1] * (double)(float)o->CountLiquid[k - 1] / (double)sum * exp(-d / (KB * o->Tm_k[k - 1]))
         * (1.0 + erf((temperature - o->Tflex_k[k - 1]) / o->sigma_k[k - 1])) * (d != 0.0 ? 1.0 : 0.0);
}
static int site_temperature(Orc *o, const int32_t Site[3], const char *who) {
  int t = IBITS(LC(o, Site), ORC_PosT, ORC_LenT);
  if (t == 0) orc_fatal(o, "Temperature is zero in %s.", who);
  return t;
}
/* src/mcevent/InterpProbFromField.f:22-57 ('ST' mode): overwrite the four working tables with the blocks of the
 * probability-table file interpolated at an integer temperature, weights are INTEGER distances.  Entries that are
 * zero in every block interpolate to (w*0+w*0)/(w+w) = 0, so only the others are visited. */
static void interp_prob_from_field(Orc *o, int temperature) {
  int iupper = 0, ilower, exact = 0;
  for (int i = 1; i <= o->ntemp; i++) if (o->reftempvalues[i - 1] - temperature == 0) { iupper = i; exact = 1; break; }   /* FINDLOC */
  if (exact) {
    const double *E = o->refPE + (size_t)(iupper - 1) * 12288;
    for (int k = 0; k < o->st_nnz; k++) o->PtransE[o->st_idx[k]] = E[o->st_idx[k]];
    memcpy(o->PtransD, o->refPD + (size_t)(iupper - 1) * 9, 9 * sizeof(double));
    memcpy(o->PtransAbs, o->refPAbs + (size_t)(iupper - 1) * 12, 12 * sizeof(double));
    memcpy(o->PtransDes, o->refPDes + (size_t)(iupper - 1) * 12, 12 * sizeof(double));
    return;
  }
  int best = 0;
  for (int i = 1; i <= o->ntemp; i++) {                                      /* MINLOC(diff, MASK=diff>0): first minimum */
    int d = o->reftempvalues[i - 1] - temperature;
    if (d > 0 && (iupper == 0 || d < best)) { iupper = i; best = d; }
  }
  ilower = iupper - 1;
  if (iupper == 0 || ilower == 0)      /* the reference indexes its arrays out of bounds here */
    orc_fatal(o, "InterpProbFromField: temperature %d K outside the tabulated range %d..%d K", temperature, o->reftempvalues[0], o->reftempvalues[o->ntemp - 1]);
  int wlower = abs(o->reftempvalues[iupper - 1] - temperature), wupper = abs(o->reftempvalues[ilower - 1] - temperature);
  const double *El = o->refPE + (size_t)(ilower - 1) * 12288, *Eu = o->refPE + (size_t)(iupper - 1) * 12288;
#define ST_MIX(lo, up) (((double)wlower * (lo) + (double)wupper * (up)) / (double)(wlower + wupper))
  for (int k = 0; k < o->st_nnz; k++) { int i = o->st_idx[k]; o->PtransE[i] = ST_MIX(El[i], Eu[i]); }
  for (int i = 0; i < 9; i++) o->PtransD[i] = ST_MIX(o->refPD[(size_t)(ilower - 1) * 9 + i], o->refPD[(size_t)(iupper - 1) * 9 + i]);
  for (int i = 0; i < 12; i++) o->PtransAbs[i] = ST_MIX(o->refPAbs[(size_t)(ilower - 1) * 12 + i], o->refPAbs[(size_t)(iupper - 1) * 12 + i]);
  for (int i = 0; i < 12; i++) o->PtransDes[i] = ST_MIX(o->refPDes[(size_t)(ilower - 1) * 12 + i], o->refPDes[(size_t)(iupper - 1) * 12 + i]);
#undef ST_MIX
}
/* the 'ST' hook at the top of Get_Prob.f:76-91 and Get_Prob_Ini.f:41-56 */
static inline void st_hook(Orc *o, const int32_t Site[3], const char *who) {
  if (o->InitSt[0] == 'S' && o->InitSt[1] == 'T' && !o->fixedT) interp_prob_from_field(o, site_temperature(o, Site, who));
}
/* the whole-table work the reference does on every call (Get_Prob.f:143-152); results unused */
static void faithful_table_cost(Orc *o, int la, int temperature) {
  if (!o->PEcopy) o->PEcopy = (double *)malloc((3 * 4096 + 9) * sizeof(double));
  if (la) {
    for (int i = 0; i < 3 * 4096; i++) o->PEcopy[i] = la_PE(o, temperature, o->PtransE[i]);
    for (int i = 0; i < 9; i++) o->PEcopy[3 * 4096 + i] = la_PD(o, temperature, o->PtransD[i]);
  } else {
    memcpy(o->PEcopy, o->PtransE, 3 * 4096 * sizeof(double));
    memcpy(o->PEcopy + 3 * 4096, o->PtransD, 9 * sizeof(double));
  }
  __asm__ volatile("" ::"r"(o->PEcopy) : "memory");
}

/* src/mcevent/Get_Prob.f:22-52 */
static int IndexEvent_Abs(Orc *o, int Coor, double Ptmp, double PDepo) {
  int res = 0;
  if (o->NCov == 1) return EV_DES(o) + 1;
  if (o->NCov >= 2) {
    for (int i = 1; i <= o->NCov; i++) {
      double pmin, pmax, s;
      if (i == 1) { pmin = PDepo; pmax = PDepo + o->PtransAbs[IDXA(1, Coor)]; }
      else {
        s = 0; for (int k = 1; k <= i - 1; k++) s += o->PtransAbs[IDXA(k, Coor)]; pmin = PDepo + s;
        s = 0; for (int k = 1; k <= i; k++) s += o->PtransAbs[IDXA(k, Coor)]; pmax = PDepo + s;
      }
      if (Ptmp > pmin && Ptmp <= pmax) res = EV_DES(o) + i;
    }
  }
  if (res == 0) o->q_anom++;
  return res;
}

/* deposition / absorption choice shared by Coor 1 and Coor 2,3; Get_Prob.f:160-179,186-204 */
static void draw_dep_abs(Orc *o, int Coor, const double PD[4], int *ev, double *prob) {
  double PDep = PD[1] + PD[2] + PD[3];
  double PAbs = 0; for (int k = 1; k <= 3; k++) PAbs += o->PtransAbs[IDXA(k, Coor)];
  double Ptot = PDep + PAbs;
  double Pcurr = orc_ran3(o) * Ptot;
  if (Pcurr <= PDep) {
    if (Pcurr <= PD[1]) *ev = 1;
    else if (Pcurr > (PD[1] + PD[2])) *ev = 3;
    else *ev = 2;
  } else *ev = IndexEvent_Abs(o, Coor, Pcurr, PDep);
  *prob = Ptot;
}

/* src/mcevent/Get_Prob.f:54-331 */
static void Get_Prob(Orc *o, int Occ, int NSiC, int CovInd, int Coor, const int32_t Site[3], const int32_t NextN[4][3], int *Index_Event, double *Prob_Event) {
  o->n_getprob++;
  st_hook(o, Site, "get_prob()");
  if (CovInd > 0 && CovInd < 113) {                       /* desorption, :93-123 */
    if (Occ != 0) orc_fatal(o, "ERROR Covind!=0 with Occ!=0 in get_prob");
    if (Coor <= 3) {
      *Index_Event = EV_DES(o);
      int iCov = 0;
      for (int j = 1; j <= o->NCov; j++) if (o->ListCov[j - 1] == CovInd) { iCov = j; break; }
      if (iCov == 0) orc_fatal(o, "ERROR: could not find CovInd=%d inside ListCov", CovInd);
      *Prob_Event = o->PtransDes[IDXA(iCov, Coor)];
    } else {
      *Index_Event = 0; *Prob_Event = 0.0;
      if (Coor > 4) LOG(o, " anomalous coordination in desorption\n");
    }
  } else {
    int la = (strcmp(o->InitSt, "LA") == 0), temperature = 0;
    if (la) temperature = site_temperature(o, Site, "get_prob()");
    if (o->faithful_cost) faithful_table_cost(o, la, temperature);
    if (Occ == 0) {
      double PD[4] = {0, 0, 0, 0};
      if (Coor >= 1 && Coor <= 3)
        for (int k = 1; k <= 3; k++) PD[k] = !la ? o->PtransD[IDXD(k, Coor)] : !o->la_tree ? la_PD(o, temperature, o->PtransD[IDXD(k, Coor)])
                                             : k <= o->NCrystal ? la_PD_k(o, temperature, k, o->PtransD[IDXD(k, Coor)]) : 0.0;   /* la/: module PD(k>NCrystal,:) stays 0 */
      if (Coor == 1) {
        int CoorNN = IBITS(LC(o, NextN[0]), ORC_PosCoor, ORC_LenCoor);
        if (CoorNN >= 2) draw_dep_abs(o, Coor, PD, Index_Event, Prob_Event);
        else { *Index_Event = 1; *Prob_Event = (double)1.e-8f; }          /* :183 REAL(4) literal */
      } else if (Coor == 2 || Coor == 3) {
        draw_dep_abs(o, Coor, PD, Index_Event, Prob_Event);
      } else if (Coor >= 4) {
        *Index_Event = 0; *Prob_Event = 0.0;
        if (Coor > 4) LOG(o, " anomalous coordination\n");
      } else { LOG(o, " anomalous coordination %d\n", Coor); o->q_anom++; }
    } else {                                                /* Occ=1: evaporation, :213-323 */
      int N_Si = 0, N_C = 0, N_B = 0, N_wall = 0, covNN[3] = {0, 0, 0};
      for (int i = 0; i < 4; i++) {
        int32_t w = LC(o, NextN[i]);
        int OccN = IBITS(w, ORC_PosOcc, ORC_LenOcc), CovIndN = IBITS(w, ORC_PosIndex, ORC_LenIndex);
        int CoorNN = IBITS(w, ORC_PosCoor, ORC_LenCoor);
        if (OccN != 0) {
          if (CovIndN != 113) {
            int IndNN = LI(o, NextN[i]);
            const int32_t (*NextNN)[3];
            if (CoorNN >= 1 && CoorNN <= 3) NextNN = mc_nn(o, IndNN);
            else { NextNN = bulk_nn(o, IndNN); if (CoorNN != 4) o->q_wl_iso++; }
            if (recipr(NextNN, Site)) {
              int SiOrC = IBITS(w, PSIC(o), LSIC(o));
              if (SiOrC == 1) N_Si++;
              if (SiOrC == 2) N_C++;
              if (SiOrC == 3) N_B++;
            }
          } else N_wall++;
        } else if (CovIndN != 0) {
          int IndNN = LI(o, NextN[i]);
          if (CoorNN >= 4) o->q_wl_cov++;
          const int32_t (*NextNN)[3] = mc_nn(o, IndNN);
          if (recipr(NextNN, Site)) {
            int iCov = 0;
            for (int j = 1; j <= o->NCov; j++) if (o->ListCov[j - 1] == CovIndN) { iCov = j; break; }
            if (iCov == 0) orc_fatal(o, "ERROR: could not find CovIndN=%d inside ListCov", CovIndN);
            covNN[iCov - 1]++;
          }
        }
      }
      int ncry = N_Si + N_C + N_B + N_wall;
      if (ncry != Coor)
        orc_fatal(o, "%d %d %d %d %d Get Prob Error N_Si+N_C+N_B+N_wall.NE.Coor site %d %d %d", Coor, N_Si, N_C, N_B, N_wall, Site[0], Site[1], Site[2]);
      if (ncry > 3) { *Index_Event = 0; *Prob_Event = 0.0; }
      else if (ncry == 0) { *Index_Event = EV_EVAP(o); *Prob_Event = (double)100000.f * o->Tree[1]; }
      else {
        *Index_Event = EV_EVAP(o);
        int c1 = 0, c2 = 0, c3 = 0;
        if (o->NCov == 1) c1 = covNN[0];
        else if (o->NCov == 2) { c1 = covNN[0]; c2 = covNN[1]; }
        else if (o->NCov == 3) { c1 = covNN[0]; c2 = covNN[1]; c3 = covNN[2]; }
        double e = o->PtransE[IDXE(NSiC, N_Si + N_wall, N_C, N_B, c1 & 3, c2 & 3, c3 & 3)];
        if ((c1 | c2 | c3) > 3) o->q_oob++;
        *Prob_Event = !la ? e : o->la_tree ? la_PE_k(o, temperature, NSiC, e) : la_PE(o, temperature, e);
      }
    }
  }
  if (*Index_Event != 0 && *Prob_Event < (double)1e-12f)
    LOG(o, " %d %.17g %d\n", *Index_Event, *Prob_Event, Coor);
}

/* src/mcevent/Get_Prob_Ini.f:23-161 */
static void Get_Prob_Ini(Orc *o, int Occ, int NSiC, int Coor, const int32_t Site[3], const int32_t NextN[4][3], int *Index_Event, double *Prob_Event) {
  int la = (strcmp(o->InitSt, "LA") == 0), temperature = 0;
  o->n_getprob++;
  st_hook(o, Site, "get_prob_ini()");
  if (la) temperature = site_temperature(o, Site, "get_prob_ini()");
  if (o->faithful_cost) faithful_table_cost(o, la, temperature);
  if (Occ == 0) {
    double PD[4] = {0, 0, 0, 0};
    if (Coor >= 1 && Coor <= 3)
      for (int k = 1; k <= 3; k++) PD[k] = !la ? o->PtransD[IDXD(k, Coor)] : !o->la_tree ? la_PD(o, temperature, o->PtransD[IDXD(k, Coor)])
                                           : k <= o->NCrystal ? la_PD_k(o, temperature, k, o->PtransD[IDXD(k, Coor)]) : 0.0;
    int draw = 0;
    if (Coor == 1) {
      int CoorNN = IBITS(LC(o, NextN[0]), ORC_PosCoor, ORC_LenCoor);
      if (CoorNN >= 2) draw = 1;
      else { *Index_Event = 1; *Prob_Event = (double)1.e-6f; }
    } else if (Coor == 2 || Coor == 3) draw = 1;
    else if (Coor == 4) { *Index_Event = 0; *Prob_Event = 0.0; LOG(o, " %d Get Prob Vacancy Generation\n", Coor); }
    if (draw) {
      double Ptot = PD[1] + PD[2] + PD[3];
      double Pcurr = orc_ran3(o) * Ptot;
      if (Pcurr <= PD[1]) *Index_Event = 1;
      else if (Pcurr > (PD[1] + PD[2])) *Index_Event = 3;
      else *Index_Event = 2;
      *Prob_Event = Ptot;
    }
  } else {
    int N_Si = 0, N_C = 0, N_B = 0, N_cov = 0;
    for (int i = 0; i < 4; i++) {
      int32_t w = LC(o, NextN[i]);
      int CovInd = IBITS(w, ORC_PosIndex, ORC_LenIndex);
      if (CovInd == 0) {
        int SiOrC = IBITS(w, PSIC(o), LSIC(o));
        if (SiOrC == 1) N_Si++;
        if (SiOrC == 2) N_C++;
        if (SiOrC == 3) N_B++;
      } else N_cov++;
    }
    int n = N_Si + N_C + N_B + N_cov;
    if (n != Coor) orc_fatal(o, "%d %d %d %d %d Get Prob Error N_Si+N_C++N_B+N_cov.NE.Coor", Coor, N_Si, N_C, N_B, N_cov);
    if (n > 3) { *Index_Event = 0; *Prob_Event = 0.0; }
    else if (n == 0) { *Index_Event = EV_EVAP(o); *Prob_Event = (double)100000.f * o->Tree[1]; LOG(o, " %d Get Prob Isolated Atom\n", Coor); }
    else {
      *Index_Event = EV_EVAP(o);
      double e = o->PtransE[IDXE(NSiC, N_Si + N_cov, N_C, N_B, 0, 0, 0)];
      *Prob_Event = !la ? e : o->la_tree ? la_PE_k(o, temperature, NSiC, e) : la_PE(o, temperature, e);
    }
  }
}

/* ------------------------------------------------------------------ events */
static inline void set_rate(Orc *o, int Ind, int ev, double p) {
  if (Ind < 1 || Ind > o->NumMcMax) { o->q_oob++; return; }
  o->ListAdAtom[Ind].ev = ev; o->ListAdAtom[Ind].prob = p;
}

/* src/mcevent/deposition.f:24-251 */
static void Deposition(Orc *o, int *IndVoid, int IAtom) {
  int32_t Site[3], SiteC[3], NextN[4][3], NextNN[4][3];
  int Index_Event = 0; double Prob = 0;
  cp3(Site, o->ListAdAtom[*IndVoid].xyz);
  int Coor = IBITS(LC(o, Site), ORC_PosCoor, ORC_LenCoor), CoorOld = Coor;
  MVBITS(1, ORC_LenOcc, &LC(o, Site), ORC_PosOcc);
  MVBITS(IAtom, LSIC(o), &LC(o, Site), PSIC(o));
  if (Coor == 1) {                                         /* :49-74 */
    cp3(SiteC, o->ListAdAtom[*IndVoid].nn[0]);
    int IndNN = LI(o, SiteC);
    if (IBITS(LC(o, SiteC), ORC_PosCoor, ORC_LenCoor) >= 4) o->q_wl_dep++;
    const int32_t (*cnn)[3] = mc_nn(o, IndNN);
    int SNSite[3][3] = {{0}}, NNZig[3][3], NNArm[3][3], j = 0;
    for (int i = 0; i < 4; i++)
      if (!eq3(cnn[i], Site)) { if (j < 3) cp3(SNSite[j], cnn[i]); else o->q_oob++; j++; }
    orc_find_sn(Site, SiteC, (const int (*)[3])SNSite, NNZig, NNArm, o->LenX, o->LenY);
    int IF_Arm = IF_Arm_Choice(o, NNZig, NNArm);
    o->last_arm = 1 + IF_Arm;
    for (int i = 0; i < 3; i++) cp3(o->ListAdAtom[*IndVoid].nn[i + 1], IF_Arm ? NNArm[i] : NNZig[i]);
  }
  cpnn(NextN, (const int32_t (*)[3])o->ListAdAtom[*IndVoid].nn);
  for (int i = 0; i < 4; i++) {
    int32_t *wN = &LC(o, NextN[i]);
    int Occ = IBITS(*wN, ORC_PosOcc, ORC_LenOcc);
    int IndNN = LI(o, NextN[i]);
    int CoorNN = IBITS(*wN, ORC_PosCoor, ORC_LenCoor);
    int CovInd = IBITS(*wN, ORC_PosIndex, ORC_LenIndex);
    memset(NextNN, 0, sizeof NextNN);
    if (Occ == 0) {
      if (CoorNN == 0) {                                   /* :88-99 new MC particle */
        CoorNN = 1;
        MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
        cp3(NextNN[0], Site);
        Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
        if (Index_Event != 0) AddMC(o, NextN[i], (const int32_t (*)[3])NextNN, Index_Event, Prob);
      } else if (CoorNN == 1) {                            /* :100-160 */
        cpnn(NextNN, mc_nn(o, IndNN));
        int Dx = abs(NextNN[0][0] - Site[0]); if (o->LenX - Dx < Dx) Dx = o->LenX - Dx;
        int Dy = abs(NextNN[0][1] - Site[1]); if (o->LenY - Dy < Dy) Dy = o->LenY - Dy;
        int DisNN2 = Dx * Dx + Dy * Dy + (NextNN[0][2] - Site[2]) * (NextNN[0][2] - Site[2]);
        if (DisNN2 == 72) {
          int NewSite[2][3];
          CoorNN = 2;
          cp3(NextNN[1], Site);
          orc_find_nn(NextNN[0], NextNN[1], NextN[i], NewSite, o->LenX, o->LenY);
          for (int j = 2; j < 4; j++) {
            cp3(NextNN[j], NewSite[j - 2]);
            int32_t w3 = LC(o, NextNN[j]);
            if (IBITS(w3, ORC_PosOcc, ORC_LenOcc) != 0) {
              int CoorNNN = IBITS(w3, ORC_PosCoor, ORC_LenCoor), CovIndNNN = IBITS(w3, ORC_PosIndex, ORC_LenIndex);
              if (CoorNNN <= 3 && CovIndNNN != 113) {
                int IndNNN = LI(o, NextNN[j]);
                if (recipr(mc_nn(o, IndNNN), NextN[i])) CoorNN++;
              }
            }
          }
          MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
          cpnn(o->ListAdAtom[IndNN].nn, (const int32_t (*)[3])NextNN);
          Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
          if (Index_Event == 0) { EraseMC(o, NextN[i]); AddVoid(o, NextN[i], (const int32_t (*)[3])NextNN); }
          else { set_rate(o, IndNN, Index_Event, Prob); orc_update_tree(o, IndNN, Prob); }
        }
      } else if (CoorNN == 2 || CoorNN == 3) {             /* :161-190 */
        cpnn(NextNN, mc_nn(o, IndNN));
        if (recipr((const int32_t (*)[3])NextNN, Site)) {
          CoorNN++;
          MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
          Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
          if (Index_Event == 0) { EraseMC(o, NextN[i]); AddVoid(o, NextN[i], (const int32_t (*)[3])NextNN); }
          else { set_rate(o, IndNN, Index_Event, Prob); orc_update_tree(o, IndNN, Prob); }
        }
      }
    } else {                                               /* full site, :191-234 */
      if (CovInd != 113) {
        if (CovInd != 0) orc_fatal(o, "ERROR in deposition: Site can only have NN with CovInd=0 or 113 when its Occ=1");
        if (CoorNN <= 3) {
          cpnn(NextNN, mc_nn(o, IndNN));
          if (recipr((const int32_t (*)[3])NextNN, Site)) {
            CoorNN++;
            MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
            int SiOrC = IBITS(*wN, PSIC(o), LSIC(o));
            Get_Prob(o, 1, SiOrC, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
            if (Index_Event == 0) { EraseMC(o, NextN[i]); AddBulk(o, NextN[i], (const int32_t (*)[3])NextNN); o->NumAdAtomOcc--; }   /* la/deposition.f:219 */
            else { set_rate(o, IndNN, Index_Event, Prob); orc_update_tree(o, IndNN, Prob); }
            if (CoorOld == 1 && i >= 1) Coor++;
          }
        }
      } else Coor++;
    }
  }
  *IndVoid = LI(o, Site);
  if (Coor != CoorOld) MVBITS(Coor, ORC_LenCoor, &LC(o, Site), ORC_PosCoor);
  Get_Prob(o, 1, IAtom, 0, Coor, Site, (const int32_t (*)[3])NextN, &Index_Event, &Prob);
  if (Index_Event == 0) { EraseMC(o, Site); AddBulk(o, Site, (const int32_t (*)[3])NextN); o->NumAdAtomOcc--; }             /* la/deposition.f:246 */
  else { set_rate(o, *IndVoid, Index_Event, Prob); orc_update_tree(o, *IndVoid, Prob); o->NumAdAtomOcc++; }                /* la/deposition.f:253 */
}

/* the "keep only the occupied reciprocal neighbour" compaction, Evaporation.f:95-122 and :229-258.
 * `rec` is the live list record, `old` the copy taken before, `me` the cell the record belongs to. */
static void compact_to_single(Orc *o, int32_t rec[4][3], const int32_t old[4][3], const int32_t me[3]) {
  int found = 0;
  for (int j = 0; j < 4; j++) {
    int32_t w = LC(o, old[j]);
    if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 0) { rec[j][0] = rec[j][1] = rec[j][2] = 0; }
    else {
      int CoorNNN = IBITS(w, ORC_PosCoor, ORC_LenCoor), CovIndNN = IBITS(w, ORC_PosIndex, ORC_LenIndex);
      if (CoorNNN <= 3 && CovIndNN != 113) {
        int IndNNN = LI(o, old[j]);
        if (recipr(mc_nn(o, IndNNN), me)) { cp3(o->NBuff_evap, rec[j]); found = 1; }
        rec[j][0] = rec[j][1] = rec[j][2] = 0;
      } else { rec[j][0] = rec[j][1] = rec[j][2] = 0; }
    }
  }
  if (!found) o->q_stale_nbuff++;
  cp3(rec[0], o->NBuff_evap);
}

/* src/mcevent/Evaporation.f:25-262 */
static void Evaporation(Orc *o, int *IndAtom) {
  int32_t Site[3], NextN[4][3], NextNN[4][3];
  int Index_Event = 0; double Prob = 0;
  int keepT = (strcmp(o->InitSt, "LA") == 0 || strcmp(o->InitSt, "ST") == 0);
  cp3(Site, o->ListAdAtom[*IndAtom].xyz);
  int IAtom = IBITS(LC(o, Site), PSIC(o), LSIC(o));
  o->NumAdAtomOcc--;                                       /* la/Evaporation.f:47 */
  if (IAtom >= 1 && IAtom <= 3) o->CountCrystal[IAtom - 1]--; else o->q_oob++;
  int Coor = IBITS(LC(o, Site), ORC_PosCoor, ORC_LenCoor);
  MVBITS(0, ORC_LenOcc, &LC(o, Site), ORC_PosOcc);
  MVBITS(0, LSIC(o), &LC(o, Site), PSIC(o));
  cpnn(NextN, (const int32_t (*)[3])o->ListAdAtom[*IndAtom].nn);
  for (int i = 0; i < 4; i++) {
    int32_t *wN = &LC(o, NextN[i]);
    int Occ = IBITS(*wN, ORC_PosOcc, ORC_LenOcc);
    int CoorNN = IBITS(*wN, ORC_PosCoor, ORC_LenCoor);
    int IndNN = LI(o, NextN[i]);
    int CovInd = IBITS(*wN, ORC_PosIndex, ORC_LenIndex);
    memset(NextNN, 0, sizeof NextNN);
    if (Occ == 0) {
      if (CoorNN == 1) {                                   /* :62-79 */
        cpnn(NextNN, mc_nn(o, IndNN));
        if (recipr((const int32_t (*)[3])NextNN, Site)) {
          EraseMC(o, NextN[i]);
          if (keepT) { int t = IBITS(*wN, ORC_PosT, ORC_LenT); *wN = 0; MVBITS(t, ORC_LenT, wN, ORC_PosT); }
          else *wN = 0;
        }
      } else if (CoorNN == 2) {                            /* :80-128 */
        cpnn(NextNN, mc_nn(o, IndNN));
        if (recipr((const int32_t (*)[3])NextNN, Site)) {
          CoorNN = 1;
          MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
          compact_to_single(o, o->ListAdAtom[IndNN].nn, (const int32_t (*)[3])NextNN, NextN[i]);
          Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
          set_rate(o, IndNN, Index_Event, Prob);
          orc_update_tree(o, IndNN, Prob);
        }
      } else if (CoorNN == 3) {                            /* :129-145 */
        cpnn(NextNN, mc_nn(o, IndNN));
        if (recipr((const int32_t (*)[3])NextNN, Site)) {
          CoorNN = 2;
          MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
          Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
          set_rate(o, IndNN, Index_Event, Prob);
          orc_update_tree(o, IndNN, Prob);
        }
      } else if (CoorNN == 4) {                            /* :146-164 vacancy -> MC */
        cpnn(NextNN, void_nn(o, IndNN));
        if (recipr((const int32_t (*)[3])NextNN, Site)) {
          CoorNN = 3;
          MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
          Get_Prob(o, 0, 0, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
          EraseVoid(o, NextN[i]);
          AddMC(o, NextN[i], (const int32_t (*)[3])NextNN, Index_Event, Prob);
        }
      }
    } else {
      if (CovInd != 113) {
        if (CovInd != 0) orc_fatal(o, "ERROR in evaporation: Site can only have NN with CovInd=0 or 113 when its Occ=1");
        if (CoorNN == 4) {                                 /* :173-190 bulk -> MC */
          cpnn(NextNN, bulk_nn(o, IndNN));
          if (recipr((const int32_t (*)[3])NextNN, Site)) {
            CoorNN = 3;
            MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
            int SiOrC = IBITS(*wN, PSIC(o), LSIC(o));
            Get_Prob(o, 1, SiOrC, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
            EraseBulk(o, NextN[i]);
            AddMC(o, NextN[i], (const int32_t (*)[3])NextNN, Index_Event, Prob);
            o->NumAdAtomOcc++;                             /* la/Evaporation.f:192 */
          }
        } else {                                           /* :191-209 */
          cpnn(NextNN, mc_nn(o, IndNN));
          if (recipr((const int32_t (*)[3])NextNN, Site)) {
            CoorNN--;
            MVBITS(CoorNN, ORC_LenCoor, wN, ORC_PosCoor);
            int SiOrC = IBITS(*wN, PSIC(o), LSIC(o));
            Get_Prob(o, 1, SiOrC, CovInd, CoorNN, NextN[i], (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
            set_rate(o, IndNN, Index_Event, Prob);
            orc_update_tree(o, IndNN, Prob);
          }
        }
      } else Coor--;
    }
  }
  *IndAtom = LI(o, Site);
  MVBITS(Coor, ORC_LenCoor, &LC(o, Site), ORC_PosCoor);
  if (Coor == 0) EraseMC(o, Site);
  else {
    Get_Prob(o, 0, 0, 0, Coor, Site, (const int32_t (*)[3])NextN, &Index_Event, &Prob);
    set_rate(o, *IndAtom, Index_Event, Prob);
    if (Coor == 1) {                                       /* :224-258 */
      cpnn(NextN, (const int32_t (*)[3])o->ListAdAtom[*IndAtom].nn);
      compact_to_single(o, o->ListAdAtom[*IndAtom].nn, (const int32_t (*)[3])NextN, Site);
    }
    orc_update_tree(o, *IndAtom, Prob);
  }
}

/* re-rate an occupied reciprocal neighbour; shared by absorption.f:43-72 and desorption.f:54-111 */
static void rerate_occ_nn(Orc *o, const int32_t Site[3], const int32_t N[3], int need_le3) {
  int32_t w = LC(o, N);
  int OccN = IBITS(w, ORC_PosOcc, ORC_LenOcc), IndNN = LI(o, N), CoorN = IBITS(w, ORC_PosCoor, ORC_LenCoor);
  int Index_Event = 0; double Prob = 0;
  if (OccN != 0) {
    int CovIndN = IBITS(w, ORC_PosIndex, ORC_LenIndex);
    if (CovIndN != 113) {
      if (need_le3 && CoorN > 3) return;
      if (!need_le3 && CoorN > 3) o->q_wl_des++;
      int32_t NextNN[4][3]; cpnn(NextNN, mc_nn(o, IndNN));
      if (recipr((const int32_t (*)[3])NextNN, Site)) {
        int SiOrC = IBITS(w, PSIC(o), LSIC(o));
        Get_Prob(o, 1, SiOrC, CovIndN, CoorN, N, (const int32_t (*)[3])NextNN, &Index_Event, &Prob);
        set_rate(o, IndNN, Index_Event, Prob);
        orc_update_tree(o, IndNN, Prob);
      }
    }
  }
}

/* src/mcevent/absorption.f:25-78 */
static void Absorption(Orc *o, int IndVoid, int CovInd) {
  int32_t Site[3], NextN[4][3];
  int Index_Event = 0; double Prob = 0;
  cp3(Site, o->ListAdAtom[IndVoid].xyz);
  int Coor = IBITS(LC(o, Site), ORC_PosCoor, ORC_LenCoor);
  MVBITS(CovInd, ORC_LenIndex, &LC(o, Site), ORC_PosIndex);
  cpnn(NextN, (const int32_t (*)[3])o->ListAdAtom[IndVoid].nn);
  for (int i = 0; i < 4; i++) rerate_occ_nn(o, Site, NextN[i], 1);
  Get_Prob(o, 0, 0, CovInd, Coor, Site, (const int32_t (*)[3])NextN, &Index_Event, &Prob);
  set_rate(o, IndVoid, Index_Event, Prob);
  orc_update_tree(o, IndVoid, Prob);
}

/* src/mcevent/desorption.f:25-117 */
static void Desorption(Orc *o, int IndAtom) {
  int32_t Site[3], NextN[4][3];
  int Index_Event = 0; double Prob = 0;
  cp3(Site, o->ListAdAtom[IndAtom].xyz);
  int CovInd = IBITS(LC(o, Site), ORC_PosIndex, ORC_LenIndex);
  if (CovInd == o->ListCov[0]) o->CountCov[0]--;
  else if (CovInd == o->ListCov[1]) o->CountCov[1]--;
  else if (CovInd == o->ListCov[2]) o->CountCov[2]--;
  int Coor = IBITS(LC(o, Site), ORC_PosCoor, ORC_LenCoor);
  MVBITS(0, ORC_LenIndex, &LC(o, Site), ORC_PosIndex);
  cpnn(NextN, (const int32_t (*)[3])o->ListAdAtom[IndAtom].nn);
  if (Coor == 1) rerate_occ_nn(o, Site, NextN[0], 0);
  else for (int i = 0; i < 4; i++) rerate_occ_nn(o, Site, NextN[i], 0);
  Get_Prob(o, 0, 0, 0, Coor, Site, (const int32_t (*)[3])NextN, &Index_Event, &Prob);
  set_rate(o, IndAtom, Index_Event, Prob);
  orc_update_tree(o, IndAtom, Prob);
}

void orc_set_variant(Orc *o, int la_tree) { o->la_tree = la_tree != 0; }
/* ------------------------------------------------------------------ create */
Orc *orc_create(int lenx, int leny, int lenz, int nummcmax) {
  Orc *o = (Orc *)calloc(1, sizeof(Orc));
  o->LenX = lenx; o->LenY = leny; o->LenZ = lenz;
  o->NumMcMax = nummcmax > 0 ? nummcmax : 500000;       /* defsystem.f:34 */
  o->NumVdMax = 100000;                                  /* defsystem.f:35 */
  o->ncell = (size_t)lenx * leny * ((size_t)lenz + 1);
  o->scratch = o->ncell;
  o->LattCoo = (int32_t *)calloc(o->ncell + 1, sizeof(int32_t));
  o->LattInd = (int32_t *)calloc(o->ncell + 1, sizeof(int32_t));
  o->ListAdAtom = (OrcMC *)calloc((size_t)o->NumMcMax + 2, sizeof(OrcMC));
  o->capAtoms = 1 << 16;
  o->ListAtom = (OrcSite *)calloc(o->capAtoms, sizeof(OrcSite));
  o->ListVoid = (OrcSite *)calloc((size_t)o->NumVdMax + 2, sizeof(OrcSite));
  /* src/mcevent/AllocateArrays.f:30-43 */
  for (int ind = 0; ind <= 31; ind++) if ((o->NumMcMax >> ind) & 1) o->Levels = ind;
  if (o->NumMcMax > (1 << o->Levels)) o->Levels += 2; else o->Levels += 1;
  o->SizeTree = (1 << o->Levels) - 1;
  o->Tree = (double *)calloc((size_t)o->SizeTree + 2, sizeof(double));
  /* src/main.f:31-36 defaults */
  o->Iter = 1; o->IterOut = 0; o->MaxIter = 2100000000; o->OutMolMol = 100000000;
  o->MaxTime = 100000000.0; o->OutTime = 10000000.0; o->Time = 0.0;
  o->Len1 = o->Len2 = o->Len3 = o->Len4 = 120;
  o->alloy_lcg = 88172645463325252ull;                   /* stand-in for libgfortran's RAND() state (one stream per run, as a process has) */
  strcpy(o->outdir, ".");
  if (!o->LattCoo || !o->LattInd || !o->ListAdAtom || !o->Tree) { orc_destroy(o); return NULL; }
  return o;
}
void orc_destroy(Orc *o) {
  if (!o) return;
  free(o->LattCoo); free(o->LattInd); free(o->ListAdAtom); free(o->ListAtom); free(o->ListVoid);
  free(o->Tree); free(o->PEcopy); free(o->trace); free(o->refPE); free(o->refPD); free(o->refPAbs); free(o->refPDes); free(o->st_idx); free(o);
}
int orc_sizeof_mc(void) { return (int)sizeof(OrcMC); }

/* FNV-1a over the lattice words, list sizes and the bits of Time: the parity fingerprint */
uint64_t orc_state_hash(const Orc *o) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < o->ncell; i++) { h ^= (uint32_t)o->LattCoo[i]; h *= 1099511628211ull; }
  uint64_t t; memcpy(&t, &o->Time, 8);
  uint64_t extra[5] = {(uint64_t)o->NumAdAtom, (uint64_t)o->NumAtoms, (uint64_t)o->NumVoids, t, (uint64_t)o->Iter};
  for (int i = 0; i < 5; i++) { h ^= extra[i]; h *= 1099511628211ull; }
  return h;
}

/* ------------------------------------------------------------------ start.dat */
static char *next_line(FILE *f, char *buf, size_t n) {
  if (!fgets(buf, (int)n, f)) return NULL;
  size_t l = strlen(buf);
  while (l && (buf[l - 1] == '\n' || buf[l - 1] == '\r')) buf[--l] = 0;
  return buf;
}
static void trim_trailing(char *s) { size_t l = strlen(s); while (l && (s[l - 1] == ' ' || s[l - 1] == '\t')) s[--l] = 0; }
/* list-directed read of one character token */
static void tok_str(const char *line, char *out, size_t n) {
  while (*line == ' ' || *line == '\t') line++;
  size_t i = 0;
  if (*line == '\'' || *line == '"') { char q = *line++; while (*line && *line != q && i + 1 < n) out[i++] = *line++; }
  else while (*line && *line != ' ' && *line != '\t' && *line != ',' && *line != '/' && i + 1 < n) out[i++] = *line++;
  out[i] = 0;
}
/* Fortran accepts 1.0d0 style exponents */
static double fortran_strtod(const char *s, char **end) {
  char tmp[128]; size_t i = 0;
  while (*s == ' ' || *s == '\t' || *s == ',') s++;
  const char *s0 = s;
  while (*s && *s != ' ' && *s != '\t' && *s != ',' && *s != '!' && i + 1 < sizeof tmp) { char c = *s++; if (c == 'd' || c == 'D') c = 'e'; tmp[i++] = c; }
  tmp[i] = 0;
  char *e2; double v = strtod(tmp, &e2);
  if (end) *end = (char *)(s0 + (e2 - tmp));
  return v;
}

/* src/mcevent/SetProbability.f:21-143 */
static int set_probability(Orc *o, FILE *f) {
  char line[4096];
  static const int tab[3][4] = {{3, 9, 19, 34}, {18, 50, 100, 172}, {57, 150, 288, 480}};
  if (o->NCrystal < 1 || o->NCrystal > 3 || o->NCov < 0 || o->NCov > 3) orc_fatal(o, "ERROR NOT IMPLEMENTED: NCrystal and NCov");
  int maxElines = tab[o->NCrystal - 1][o->NCov];
  LOG(o, " maxElines %d\n", maxElines);
  memset(o->PtransE, 0, sizeof o->PtransE);
  for (int i = 1; i <= maxElines; i++) {
    if (!next_line(f, line, sizeof line)) orc_fatal(o, "start.dat: EOF in evaporation block");
    char *p = line; int indsp = (int)strtol(p, &p, 10); long elind = strtol(p, &p, 10); double v = fortran_strtod(p, &p);
    if (indsp > o->NCrystal) orc_fatal(o, "ERROR: 1st index in evap prob must be <= NCrystal");
    int ind[6] = {0, 0, 0, 0, 0, 0};
    int nd = o->NCrystal + o->NCov;
    for (int j = 1; j <= o->NCrystal; j++) { long d = 1; for (int k = 0; k < nd - j; k++) d *= 10; ind[j - 1] = (int)((elind / d) % 10); }
    for (int j = o->NCrystal + 1; j <= nd; j++) { long d = 1; for (int k = 0; k < nd - j; k++) d *= 10; ind[3 + j - o->NCrystal - 1] = (int)((elind / d) % 10); }
    for (int k = 0; k < 6; k++) if (ind[k] > 3) orc_fatal(o, "evaporation index digit > 3");
    o->PtransE[IDXE(indsp, ind[0], ind[1], ind[2], ind[3], ind[4], ind[5])] = v;
  }
  for (int j = 1; j <= o->NCrystal * 3; j++) {
    if (!next_line(f, line, sizeof line)) orc_fatal(o, "start.dat: EOF in deposition block");
    char *p = line; int indsp = (int)strtol(p, &p, 10); int el = (int)strtol(p, &p, 10); double v = fortran_strtod(p, &p);
    if (indsp > o->NCrystal || el < 1 || el > 3) orc_fatal(o, "ERROR: 1st index in depo prob block must be <= NCrystal");
    o->PtransD[IDXD(indsp, el)] = v;
  }
  for (int pass = 0; pass < 2; pass++)
    for (int j = 1; j <= o->NCov * 3; j++) {
      if (!next_line(f, line, sizeof line)) orc_fatal(o, "start.dat: EOF in abs/des block");
      char *p = line; int indsp = (int)strtol(p, &p, 10); int el = (int)strtol(p, &p, 10); double v = fortran_strtod(p, &p);
      if (indsp <= o->NCrystal || indsp > o->NCrystal + o->NCov || el < 0 || el > 3) orc_fatal(o, "ERROR: 1st index in abs/des prob block");
      (pass == 0 ? o->PtransAbs : o->PtransDes)[IDXA(indsp - o->NCrystal, el)] = v;
    }
  return 0;
}

/* src/mcevent/SetProbabilityFromField.f:21-171: the last line of start.dat names a table file holding ntemp blocks,
 * each = one temperature + the evaporation / deposition / absorption / desorption lines of SetProbability.f */
static int set_probability_from_field(Orc *o, FILE *f) {
  char line[4096];
  static const int tab[3][4] = {{3, 9, 19, 34}, {18, 50, 100, 172}, {57, 150, 288, 480}};
  if (!next_line(f, line, sizeof line)) orc_fatal(o, "start.dat: EOF before the probability-table file name");
  strncpy(o->tempfilename, line, sizeof o->tempfilename - 1); trim_trailing(o->tempfilename);
  LOG(o, " Reading input probability table: %s\n", o->tempfilename);
  FILE *t = fopen(o->tempfilename, "r");
  if (!t) orc_fatal(o, "cannot open %s", o->tempfilename);
  if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF"); o->ntemp = (int)strtol(line, NULL, 10);
  if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF"); o->nprob = (int)strtol(line, NULL, 10);
  if (o->ntemp < 1 || o->ntemp > 300) orc_fatal(o, "probability table: ntemp must be 1..300 (ntempMax, defsystem.f:72)");
  if (o->NCrystal < 1 || o->NCrystal > 3 || o->NCov < 0 || o->NCov > 3) orc_fatal(o, "ERROR: NCrystal and NCov must be <= 3 3");
  int maxElines = tab[o->NCrystal - 1][o->NCov], nd = o->NCrystal + o->NCov;
  free(o->refPE); free(o->refPD); free(o->refPAbs); free(o->refPDes); free(o->st_idx);
  o->refPE = (double *)calloc((size_t)o->ntemp * 12288, sizeof(double));
  o->refPD = (double *)calloc((size_t)o->ntemp * 9, sizeof(double));
  o->refPAbs = (double *)calloc((size_t)o->ntemp * 12, sizeof(double));
  o->refPDes = (double *)calloc((size_t)o->ntemp * 12, sizeof(double));
  for (int tb = 0; tb < o->ntemp; tb++) {
    if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF at block %d", tb + 1);
    o->reftempvalues[tb] = (int)strtol(line, NULL, 10);
    if (tb == 0) LOG(o, " maxElines %d\n", maxElines);
    for (int i = 1; i <= maxElines; i++) {
      if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF in evaporation block");
      char *p = line; int indsp = (int)strtol(p, &p, 10); long elind = strtol(p, &p, 10); double v = fortran_strtod(p, &p);
      if (indsp > o->NCrystal) orc_fatal(o, "ERROR: 1st index in PtransE must be <= NCrystal");
      int ind[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 1; j <= nd; j++) {
        long d = 1; for (int k = 0; k < nd - j; k++) d *= 10;
        int digit = (int)((elind / d) % 10);
        if (digit > 3) orc_fatal(o, "evaporation index digit > 3");
        if (j <= o->NCrystal) ind[j - 1] = digit; else ind[3 + j - o->NCrystal - 1] = digit;
      }
      o->refPE[(size_t)tb * 12288 + IDXE(indsp, ind[0], ind[1], ind[2], ind[3], ind[4], ind[5])] = v;
    }
    for (int j = 1; j <= o->NCrystal * 3; j++) {
      if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF in deposition block");
      char *p = line; int indsp = (int)strtol(p, &p, 10); int el = (int)strtol(p, &p, 10); double v = fortran_strtod(p, &p);
      if (indsp > o->NCrystal || el < 1 || el > 3) orc_fatal(o, "ERROR: 1st index in depo prob block must be <= NCrystal");
      o->refPD[(size_t)tb * 9 + IDXD(indsp, el)] = v;
    }
    for (int pass = 0; pass < 2; pass++)
      for (int j = 1; j <= o->NCov * 3; j++) {
        if (!next_line(t, line, sizeof line)) orc_fatal(o, "probability table: EOF in abs/des block");
        char *p = line; int indsp = (int)strtol(p, &p, 10); int el = (int)strtol(p, &p, 10); double v = fortran_strtod(p, &p);
        if (indsp <= o->NCrystal || indsp > o->NCrystal + o->NCov || el < 0 || el > 3) orc_fatal(o, "ERROR: 1st index in abs/des prob block");
        (pass == 0 ? o->refPAbs : o->refPDes)[(size_t)tb * 12 + IDXA(indsp - o->NCrystal, el)] = v;
      }
  }
  fclose(t);
  o->st_idx = (int *)malloc(12288 * sizeof(int)); o->st_nnz = 0;
  for (int i = 0; i < 12288; i++) for (int tb = 0; tb < o->ntemp; tb++) if (o->refPE[(size_t)tb * 12288 + i] != 0.0) { o->st_idx[o->st_nnz++] = i; break; }
  memset(o->PtransE, 0, sizeof o->PtransE);
  LOG(o, " Done reading input probability table\n");
  return 0;
}

/* src/main.f:83-219 */
int orc_read_start(Orc *o, const char *path) {
  FILE *f = fopen(path, "r");
  char line[4096], tok[64];
  if (!f) { snprintf(o->errmsg, sizeof o->errmsg, "cannot open %s", path); return -1; }
  o->jb_set = 1;
  if (setjmp(o->jb)) { o->jb_set = 0; fclose(f); return -2; }
#define NEXT() do { if (!next_line(f, line, sizeof line)) orc_fatal(o, "start.dat: unexpected EOF"); } while (0)
  LOG(o, " KMC box size:  %d %d %d\n", o->LenX, o->LenY, o->LenZ);
  NEXT(); { char *p = line; o->NCrystal = (int)strtol(p, &p, 10); o->NCov = (int)strtol(p, &p, 10); }
  if (o->NCrystal > 3) orc_fatal(o, "ERR NOT IMPLEMENTED: NCrystal must be <= 3");
  NEXT(); { char *p = line; for (int i = 0; i < o->NCrystal; i++) o->ListCrystal[i] = (int)strtol(p, &p, 10); }
  if (o->NCov > 0 && o->NCov <= 3) {
    LOG(o, "            ***** Simulating CVD *****\n");
    NEXT(); { char *p = line; for (int i = 0; i < o->NCov; i++) o->ListCov[i] = (int)strtol(p, &p, 10); }
  } else if (o->NCov == 0) LOG(o, "            ***** Simulating PVD or LA *****\n");
  else orc_fatal(o, "ERR NOT IMPLEMENTED: NCov must be <= 3");
  LOG(o, " Crystal species (NCrystal= %d ) --> Z:", o->NCrystal);
  for (int i = 0; i < o->NCrystal; i++) LOG(o, " %d", o->ListCrystal[i]);
  LOG(o, "\n");
  if (o->NCov > 0) { LOG(o, " Coverage species (NCov= %d ) --> Z:", o->NCov); for (int i = 0; i < o->NCov; i++) LOG(o, " %d", o->ListCov[i]); LOG(o, "\n"); }
  NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->InitSt, tok, 2); o->InitSt[2] = 0;
  if (!strcmp(o->InitSt, "FF")) { NEXT(); char *p = line; o->Len1 = (int)strtol(p, &p, 10); o->Len2 = (int)strtol(p, &p, 10); o->Len3 = (int)strtol(p, &p, 10); o->Len4 = (int)strtol(p, &p, 10); }
  else if (!strcmp(o->InitSt, "ST")) {                       /* main.f:109-115 */
    NEXT(); char *p = line; o->Len1 = (int)strtol(p, &p, 10); o->Len2 = (int)strtol(p, &p, 10); o->Len3 = (int)strtol(p, &p, 10);
    o->angle = fortran_strtod(p, &p); o->Ttop = fortran_strtod(p, &p); o->Tbottom = fortran_strtod(p, &p);
    LOG(o, " Temperature on top in Kelvin:  %.17g\n Temperature on bottom in Kelvin:  %.17g\n", o->Ttop, o->Tbottom);
    if (o->Ttop == o->Tbottom) o->fixedT = 1;
  }
  else if (!strcmp(o->InitSt, "IN") || !strcmp(o->InitSt, "SH")) { NEXT(); strncpy(o->cadfilename, line, sizeof o->cadfilename - 1); trim_trailing(o->cadfilename); }
  else if (!strcmp(o->InitSt, "LA")) {
    LOG(o, " ****** MulSKIPs run for laser annealing ******\n");
    if (o->NCov > 0) orc_fatal(o, "ERR NOT IMPLEMENTED: NCov=0 for LA");
    NEXT(); strncpy(o->cadfilename, line, sizeof o->cadfilename - 1); trim_trailing(o->cadfilename);
    NEXT(); strncpy(o->tempfilename, line, sizeof o->tempfilename - 1); trim_trailing(o->tempfilename);
    if (o->la_tree) {                                        /* la/main.f:420-424: one value per crystal species */
      double *dst[4] = {o->Tm_k, o->P0_k, o->sigma_k, o->Tflex_k};
      for (int q = 0; q < 4; q++) { NEXT(); char *p = line; for (int k = 0; k < o->NCrystal; k++) dst[q][k] = fortran_strtod(p, &p); }
      o->Tm = o->Tm_k[0]; o->P0 = o->P0_k[0]; o->sigma = o->sigma_k[0]; o->Tflex = o->Tflex_k[0];
      LOG(o, " Melting temperature in Kelvin: "); for (int k = 0; k < o->NCrystal; k++) LOG(o, " %.17g", o->Tm_k[k]); LOG(o, "\n");
    } else {
    NEXT(); o->Tm = fortran_strtod(line, NULL);
    NEXT(); o->P0 = fortran_strtod(line, NULL);
    NEXT(); o->sigma = fortran_strtod(line, NULL);
    NEXT(); o->Tflex = fortran_strtod(line, NULL);
    LOG(o, " Melting temperature in Kelvin:  %.17g\n", o->Tm);
    }
    NEXT(); o->LenVac = fortran_strtod(line, NULL);
    LOG(o, " Top air thickness (ang):  %.17g\n", o->LenVac);
    NEXT(); o->LenNuc = fortran_strtod(line, NULL);
    LOG(o, " Nucleus thickness (ang):  %.17g\n", o->LenNuc);
    NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->Homogeneous, tok, 2);
    LOG(o, " Model homogeneous nucleation: %s\n", o->Homogeneous);
    if (o->la_tree && !strcmp(o->Homogeneous, "D")) orc_fatal(o, "oracle: Homogeneous 'D' (SetSiGeSF stacking-fault seed) not restated");
  } else { NEXT(); char *p = line; o->Len1 = (int)strtol(p, &p, 10); o->Len2 = (int)strtol(p, &p, 10); o->Len3 = (int)strtol(p, &p, 10); }
  /* SiGe alloy fraction, main.f:142-149; trap 5 of SURVEY 8(c) */
  if ((o->NCrystal == 2 && o->ListCrystal[0] == 14 && o->ListCrystal[1] == 32) ||
      (o->NCrystal == 3 && o->ListCrystal[0] == 14 && o->ListCrystal[1] == 32 && o->ListCrystal[2] == 5)) {
    NEXT(); o->alloyfraction = fortran_strtod(line, NULL);
    LOG(o, " Ge fraction in substrate:  %.17g\n", o->alloyfraction);
    if (o->la_tree) { NEXT(); o->LenSiGe = fortran_strtod(line, NULL); LOG(o, " SiGe thickness (ang):  %.17g\n", o->LenSiGe); }   /* la/main.f:441-442 */
  } else o->alloyfraction = 0.0;
  NEXT(); o->PtransZig = fortran_strtod(line, NULL);
  NEXT(); o->sf = fortran_strtod(line, NULL);
  LOG(o, " KMC Super-Lattice parameter (ang):  %.17g\n", o->sf);
  NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->ExitStrategy, tok, 4); o->ExitStrategy[4] = 0;
  LOG(o, " Exit strategy: %s\n", o->ExitStrategy);
  if (!strcmp(o->ExitStrategy, "Iter")) {
    NEXT(); o->MaxIter = strtoll(line, NULL, 10);
    NEXT(); o->OutMolMol = strtoll(line, NULL, 10);
    LOG(o, " MaxIter:  %lld\n OutMolMol:  %lld\n", (long long)o->MaxIter, (long long)o->OutMolMol);
  } else if (!strcmp(o->ExitStrategy, "Time")) {
    NEXT(); o->MaxTime = fortran_strtod(line, NULL);
    NEXT(); o->OutTime = fortran_strtod(line, NULL);
    LOG(o, " Total simulation time in seconds:  %.17g\n Output frequency in seconds:  %.17g\n", o->MaxTime, o->OutTime);
  }
  NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->RunType, tok, 2); o->RunType[2] = 0;
  if (o->RunType[0] == 'T' && o->RunType[1] == 0)
    LOG(o, " ATTENTION! You are running mulskips in test modality,  that is random numbers are not so random!!\n");
  NEXT(); o->Idum = (int)strtol(line, NULL, 10);
  LOG(o, " IDUM %d\n", o->Idum);
  NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->SaveFinalState, tok, 2);
  if (!strcmp(o->SaveFinalState, "T")) {
    NEXT(); strncpy(o->restartfilename, line, sizeof o->restartfilename - 1); trim_trailing(o->restartfilename);
    LOG(o, " ATTENTION! Storing checkpoint file %s\n This will occupy a lot of disk space...\n", o->restartfilename);
  } else LOG(o, " SaveFinalState flag is not \"T\": checkpoint file will  not be written.\n");
  NEXT(); tok_str(line, tok, sizeof tok); strncpy(o->SaveCoo, tok, 2);
  if (!strcmp(o->SaveCoo, "T")) {
    NEXT(); strncpy(o->coofilename, line, sizeof o->coofilename - 1); trim_trailing(o->coofilename);
    if (o->la_tree && !strcmp(o->coofilename, "None")) LOG(o, " Phases passed through sockets...\n");
    else LOG(o, " ATTENTION! Storing final Coor file %s\n This might occupy a a lot of disk space...\n", o->coofilename);
  } else {
    if (!strcmp(o->InitSt, "LA")) orc_fatal(o, "ERR: SaveCoo flag must be \"T\" for LA simulations!");
    LOG(o, " SaveCoo flag is not \"T\": Coor file not stored\n");
  }
  if (!strcmp(o->InitSt, "ST")) set_probability_from_field(o, f);      /* main.f:215-219 */
  else set_probability(o, f);
  LOG(o, " numParticelle= %d Levels = %d , SizeTree = %d\n", o->NumMcMax, o->Levels, o->SizeTree);
#undef NEXT
  o->jb_set = 0;
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------ initial geometries */
static inline int is_cubic_site(int x, int y, int z, int *kind) {
  /* the test used by every builder, e.g. SetSi.f:52-57: kind 0 -> JumpDia 1..4, kind 1 -> JumpDia 5..8 */
  if (x % 3 || y % 3 || z % 3) return 0;
  int x1 = x / 3, y1 = y / 3, z1 = z / 3;
  if ((x1 % 2) != (y1 % 2) || (x1 % 2) != (z1 % 2)) return 0;
  int m = (x1 + y1 + z1) % 4;
  if (m == 0) { *kind = 0; return 1; }
  if (m == 3) { *kind = 1; return 1; }
  return 0;
}
static inline void dia_nn(Orc *o, int x, int y, int z, int kind, int32_t NextN[4][3]) {
  for (int i = 0; i < 4; i++) {
    NextN[i][0] = x + JumpDia[i + 4 * kind][0];
    NextN[i][1] = y + JumpDia[i + 4 * kind][1];
    NextN[i][2] = z + JumpDia[i + 4 * kind][2];
    if (NextN[i][0] >= o->LenX) NextN[i][0] -= o->LenX;
    if (NextN[i][0] < 0) NextN[i][0] += o->LenX;
    if (NextN[i][1] >= o->LenY) NextN[i][1] -= o->LenY;
    if (NextN[i][1] < 0) NextN[i][1] += o->LenY;
  }
}
static void set_wall_planes(Orc *o, int z0, int z1, int x0, int x1) {
  for (int z = z0; z <= z1; z++) for (int y = 0; y < o->LenY; y++) for (int x = x0; x <= x1; x++) {
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    *w |= (1 << ORC_PosOcc);
    MVBITS(113, ORC_LenIndex, w, ORC_PosIndex);
  }
}

/* species of one substrate site of the SiGe alloy builder (SetSiGex.f:60-67): Si if rr > alloyfraction, else the
 * second species.  The reference draws libgfortran's RAND() here, which is not part of the reference sources: a
 * documented 64-bit LCG stands in (alloy initial configurations: parity unpinned, SURVEY 8c trap 6). */
static int alloy_species(Orc *o) {
  o->alloy_lcg = o->alloy_lcg * 6364136223846793005ull + 1442695040888963407ull;
  double rr = (double)(o->alloy_lcg >> 11) * (1.0 / 9007199254740992.0);
  int sp = rr > o->alloyfraction ? 1 : 2;
  o->CountCrystal[sp - 1]++;
  return sp;
}
/* Flat (001) builders: src/initial/SetSi.f:25-221 (walls=1, one species), SetSiC3C.f:25-199 (walls=0, Si/C) and
 * SetSiGex.f:25-256 (walls=1, two_species=2: random alloy, species counted at init) */
static void set_flat(Orc *o, int Len, int walls, int two_species) {
  int32_t Site[3], NextN[4][3]; int kind, ev; double Prob;
  const int alloy = two_species == 2;
  if (alloy) { LOG(o, " Setting up SiGe alloy with fraction =  %.17g\n", o->alloyfraction); two_species = 0; }
  if (walls) set_wall_planes(o, 0, 5, 0, o->LenX - 1);
  for (int z = (walls ? 6 : 0); z <= Len - 4; z++) for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      int32_t *w = &o->LattCoo[CI(o, x, y, z)];
      int sp0 = alloy ? alloy_species(o) : ((kind && two_species) ? 2 : 1);
      *w |= 1 << (sp0 == 2 ? PC_(o) : PSI(o));
      *w |= 1 << ORC_PosOcc;
      MVBITS(4, ORC_LenCoor, w, ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  for (int z = Len - 3; z <= Len - 1; z++) for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      int32_t *w = &o->LattCoo[CI(o, x, y, z)];
      int sp = alloy ? alloy_species(o) : ((kind && two_species) ? 2 : 1);
      *w |= 1 << (sp == 2 ? PC_(o) : PSI(o));
      *w |= 1 << ORC_PosOcc;
      MVBITS(2, ORC_LenCoor, w, ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      Prob = o->PtransE[IDXE(sp, 2, 0, 0, 0, 0, 0)];
      AddMC(o, Site, (const int32_t (*)[3])NextN, EV_EVAP(o), Prob);
    }
  for (int z = Len; z <= Len; z++) for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      int32_t *w = &o->LattCoo[CI(o, x, y, z)];
      MVBITS(2, ORC_LenCoor, w, ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      ev = 0; Prob = 0;
      Get_Prob(o, 0, 0, 0, 2, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
      AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
    }
  if (walls) set_wall_planes(o, o->LenZ - 6, o->LenZ - 1, 0, o->LenX - 1);
}

/* "Set the coordination" pass shared by SetNC.f:60-133, SetNCSp.f, SetInvPOff.f, SetSiFINFET.f:118-214,
 * SetCAD.f:203-296.  wallaware selects the SetCAD/FINFET variant (skip wall cells, empty sites do not
 * count wall neighbours). */
static void pass_coordination(Orc *o, int z0, int z1, int y0, int y1, int x0, int x1, int wallaware) {
  int32_t NextN[4][3]; int kind;
  int la = !strcmp(o->InitSt, "LA");
  for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
    if (wallaware && IBITS(o->LattCoo[CI(o, x, y, z)], ORC_PosIndex, ORC_LenIndex) == 113) continue;
    if (!is_cubic_site(x, y, z, &kind)) continue;
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    int Occ = IBITS(*w, ORC_PosOcc, ORC_LenOcc), Coor = 0;
    dia_nn(o, x, y, z, kind, NextN);
    for (int i = 0; i < 4; i++) {
      int32_t wn = LC(o, NextN[i]);
      int OccN = IBITS(wn, ORC_PosOcc, ORC_LenOcc);
      if (!wallaware || Occ > 0) { if (OccN > 0) Coor++; }
      else if (OccN > 0 && IBITS(wn, ORC_PosIndex, ORC_LenIndex) != 113) Coor++;
    }
    if (Occ != 0 && Coor == 0) {
      if (la) { int t = IBITS(*w, ORC_PosT, ORC_LenT); *w = 0; MVBITS(t, ORC_LenT, w, ORC_PosT); }
      else *w = 0;
      for (int i = 0; i < 4; i++) {
        int32_t *wn = &LC(o, NextN[i]);
        int c = IBITS(*wn, ORC_PosCoor, ORC_LenCoor) - 1; if (c < 0) c = 0;
        MVBITS(c, ORC_LenCoor, wn, ORC_PosCoor);
      }
    } else MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
  }
}
/* "Set the probability" pass: SetNC.f:135-233, SetSiFINFET.f:216-309, SetCAD.f:301-418 */
static void pass_probability(Orc *o, int z0, int z1, int y0, int y1, int x0, int x1, int wallaware, int cadvoid) {
  int32_t Site[3], NextN[4][3], NBuff[3] = {0, 0, 0}; int kind, ev; double Prob;
  for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
    if (wallaware && IBITS(o->LattCoo[CI(o, x, y, z)], ORC_PosIndex, ORC_LenIndex) == 113) continue;
    if (!is_cubic_site(x, y, z, &kind)) continue;
    int32_t w = o->LattCoo[CI(o, x, y, z)];
    Site[0] = x; Site[1] = y; Site[2] = z;
    dia_nn(o, x, y, z, kind, NextN);
    int Occ = IBITS(w, ORC_PosOcc, ORC_LenOcc), Coor = IBITS(w, ORC_PosCoor, ORC_LenCoor);
    if (Occ != 0 && Coor == 0) LOG(o, " error %d %d %d %d %d\n", x, y, z, Occ, Coor);
    ev = 0; Prob = 0;
    if (Occ == 0) {
      if (Coor > 0) {
        if (Coor == 1) {
          for (int i = 0; i < 4; i++) {
            int32_t wn = LC(o, NextN[i]);
            int OccN = IBITS(wn, ORC_PosOcc, ORC_LenOcc), CovIndN = IBITS(wn, ORC_PosIndex, ORC_LenIndex);
            if (OccN == 0 || (wallaware && CovIndN == 113)) { NextN[i][0] = NextN[i][1] = NextN[i][2] = 0; }
            else { cp3(NBuff, NextN[i]); NextN[i][0] = NextN[i][1] = NextN[i][2] = 0; }
          }
          cp3(NextN[0], NBuff);
        }
        if (cadvoid && Coor == 4) { LOG(o, " Initializing a vacancy\n"); AddVoid(o, Site, (const int32_t (*)[3])NextN); }
        else {
          Get_Prob_Ini(o, 0, 0, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
          AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
        }
      }
    } else {
      if (Coor <= 3) {
        int IAt = IBITS(w, PSIC(o), LSIC(o));
        Get_Prob_Ini(o, 1, IAt, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
        AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
        o->NumAdAtomOcc++;                                   /* la/SetCAD.f:544,595 */
      } else AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  }
}
static inline void occupy(Orc *o, int x, int y, int z, int kind, int two_species) {
  int32_t *w = &o->LattCoo[CI(o, x, y, z)];
  *w |= 1 << ((kind && two_species) ? PC_(o) : PSI(o));
  *w |= 1 << ORC_PosOcc;
}

/* src/initial/SetNC.f:25-235 (sphere=0) and SetNCSp.f (sphere=1) */
static void set_nc(Orc *o, int Len1, int Len2, int Len3, int sphere, int two_species) {
  int kind, LX = o->LenX, LY = o->LenY, LZ = o->LenZ;
  if (!sphere) { LOG(o, " %d %d\n", LZ / 2 - Len3 / 2, LZ / 2 + Len3 / 2 + 1); LOG(o, " %d %d\n", LX / 2 - Len1 / 2, LX / 2 + Len1 / 2 + 1); }
  for (int z = LZ / 2 - Len3 / 2; z <= LZ / 2 + Len3 / 2 - 1; z++)
    for (int y = LY / 2 - Len2 / 2; y <= LY / 2 + Len2 / 2 - 1; y++)
      for (int x = LX / 2 - Len1 / 2; x <= LX / 2 + Len1 / 2 - 1; x++) {
        if (sphere) {
          int d2 = (x - LX / 2) * (x - LX / 2) + (y - LY / 2) * (y - LY / 2) + (z - LZ / 2) * (z - LZ / 2);
          if (d2 > Len1 * Len1 / 4) continue;
        }
        if (is_cubic_site(x, y, z, &kind)) occupy(o, x, y, z, kind, two_species);
      }
  pass_coordination(o, LZ / 2 - Len3 / 2 - 10, LZ / 2 + Len3 / 2 + 9, LY / 2 - Len2 / 2 - 10, LY / 2 + Len2 / 2 + 9,
                    LX / 2 - Len1 / 2 - 10, LX / 2 + Len1 / 2 + 9, 0);
  pass_probability(o, LZ / 2 - Len3 / 2 - 10, LZ / 2 + Len3 / 2 + 9, LY / 2 - Len2 / 2 - 10, LY / 2 + Len2 / 2 + 9,
                   LX / 2 - Len1 / 2 - 10, LX / 2 + Len1 / 2 + 9, 0, 0);
}

/* src/initial/SetInvPOff.f:22-286 (only = 0: Si/C sublattices, pyramid faces with the REAL(4) slope 1.10), and its
 * single-species twins SetInvPSi.f (only = 1: Si on both sublattices) and SetInvPC.f (only = 2: C on both), whose
 * pyramid test is the all-integer ABS(x-..)+ABS(y-..)-(z-Len3) <= Len */
/* ... and the antiphase-domain variants (all with the integer pyramid test): SetInvPwOnlySiorC.f (only = 3: Si/C as in the
 * ideal crystal where (x<LenX/2 and y>=LenY/2) or (x>=LenX/2 and y<LenY/2), swapped elsewhere), SetInvPwOnlyCorSi.f
 * (only = 4: the two domains exchanged) and SetInvPwAPB.f (only = 5: ideal where x+y <= LenY, swapped elsewhere) */
static inline int invp_kind(const Orc *o, int only, int x, int y, int kind) {
  if (only < 3) return kind;
  int quad = (x < o->LenX / 2 && y >= o->LenY / 2) || (x >= o->LenX / 2 && y < o->LenY / 2);
  int ideal = only == 3 ? quad : only == 4 ? !quad : (x + y <= o->LenY);
  return ideal ? kind : !kind;
}
static void set_invp_off(Orc *o, int Len, int Len3, int only) {
  int kind, LX = o->LenX, LY = o->LenY; int32_t Site[3], NextN[4][3];
  int Lenh = LX / 2 - 1;
  LOG(o, " %d %d %d\n", Lenh, Len, Len3);
  for (int z = 0; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      occupy(o, x, y, z, only == 2 ? 1 : invp_kind(o, only, x, y, kind), only != 1);
      MVBITS(4, ORC_LenCoor, &o->LattCoo[CI(o, x, y, z)], ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (x % 3 || y % 3 || z % 3) continue;
    int x1 = x / 3, y1 = y / 3, z1 = z / 3;
    if ((x1 % 2) != (y1 % 2) || (x1 % 2) != (z1 % 2)) continue;
    /* mixed INTEGER/REAL(4) expression of SetInvPOff.f:78-82: 1.10 is REAL(4) */
    float zt = 1.10f * (float)(z - Len3);
#define PYR(ax, ay) (only ? (abs(ax) + abs(ay) - (z - Len3)) <= Len : ((float)(abs(ax) + abs(ay)) - zt) <= (float)Len)
    int inside = PYR(x - Lenh, y - Lenh) || PYR(x, y) || PYR(x - LX, y) || PYR(x, y - LY) || PYR(x - LX, y - LY);
#undef PYR
    if (!inside) {
      int m = (x1 + y1 + z1) % 4;
      if (m == 0) occupy(o, x, y, z, only == 2 ? 1 : invp_kind(o, only, x, y, 0), only != 1);
      else if (m == 3) occupy(o, x, y, z, invp_kind(o, only, x, y, 1), only != 1);
    }
  }
  pass_coordination(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0);
  pass_probability(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0, 0);
}

/* src/initial/SetSiFINFET.f:25-312 */
static void set_finfet(Orc *o, int Len, int Lenww, int Lenwwtop, int Laperture) {
  int kind, LX = o->LenX, LY = o->LenY, LZ = o->LenZ;
  set_wall_planes(o, 0, 5, 0, LX - 1);
  set_wall_planes(o, Lenww, LZ - 7, 0, (int)(0.5f * (float)(LX - Laperture)) - 1);
  set_wall_planes(o, Lenww, LZ - 7, (int)(0.5f * (float)(LX + Laperture)), LX - 1);
  set_wall_planes(o, Lenwwtop, LZ - 7, 0, LX - 1);
  set_wall_planes(o, LZ - 6, LZ - 1, 0, LX - 1);
  for (int z = 6; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (IBITS(o->LattCoo[CI(o, x, y, z)], ORC_PosIndex, ORC_LenIndex) == 113) continue;
    if (is_cubic_site(x, y, z, &kind)) occupy(o, x, y, z, kind, 0);
  }
  pass_coordination(o, 6, Lenwwtop + 9, 0, LY - 1, 0, LX - 1, 1);
  pass_probability(o, 6, Lenwwtop + 9, 0, LY - 1, 0, LX - 1, 1, 0);
}

/* src/initial/SetTrench.f:22-283 (InitSt 'T'): 3C-SiC slab up to Len3 with a rectangular trench of half-width Len between
 * Len2 and Len3, no walls, bulk up to z=20 */
static void set_trench(Orc *o, int Len, int Len2, int Len3) {
  int kind, LX = o->LenX, LY = o->LenY; int32_t Site[3], NextN[4][3];
  int Lenh = LX / 2 - 1;
  LOG(o, " %d %d %d\n", Lenh, Len, Len3);
  for (int z = 0; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      occupy(o, x, y, z, kind, 1);
      MVBITS(4, ORC_LenCoor, &o->LattCoo[CI(o, x, y, z)], ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (!is_cubic_site(x, y, z, &kind)) continue;
    if (!(abs(x - Lenh) <= Len && z > Len2 && z < Len3)) occupy(o, x, y, z, kind, 1);
  }
  pass_coordination(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0);
  pass_probability(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0, 0);
}

/* src/initial/SetTrenchSi.f:22-338 (InitSt 'ST'): temperature ramp along z written into the lattice words, wall
 * planes at both ends, Si bulk up to z=20, then a Si slab up to Len3 with a V-shaped trench of half-width Len at
 * the top, bottom at Len2 and side-wall angle `angle` (degrees) */
static void set_trench_si(Orc *o, int Len, int Len2, int Len3, double angle, double Ttop, double Tbottom) {
  const double Pi = (double)3.1415926535897932384626433832795028841971f;       /* REAL(4) literal, SetTrenchSi.f:34 */
  int kind, LX = o->LenX, LY = o->LenY, LZ = o->LenZ; int32_t Site[3], NextN[4][3];
  double arad = angle * Pi / 180;
  int Lenh = LX / 2 - 1;
  LOG(o, " %d %d %d\n", Lenh, Len, Len3);
  for (int z = 0; z <= LZ - 1; z++) {                                            /* :44-62 */
    int temperature = z <= Len3 - 1 ? (int)(Tbottom + (Ttop - Tbottom) * z / (Len3 - 1)) : (int)Ttop;
    for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) MVBITS(temperature, ORC_LenT, &o->LattCoo[CI(o, x, y, z)], ORC_PosT);
  }
  if (o->fixedT) interp_prob_from_field(o, (int)Ttop);                          /* :65-67 */
  set_wall_planes(o, 0, 3, 0, LX - 1);
  set_wall_planes(o, LZ - 4, LZ - 1, 0, LX - 1);
  for (int z = 4; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (is_cubic_site(x, y, z, &kind)) {
      occupy(o, x, y, z, kind, 0);
      MVBITS(4, ORC_LenCoor, &o->LattCoo[CI(o, x, y, z)], ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (!is_cubic_site(x, y, z, &kind)) continue;
    int in_trench = ((double)abs(x - Lenh) <= ((double)Len - (double)(Len3 - z) * cos(arad) / sin(arad))) && (z > Len2 && z < Len3);   /* :138-139 */
    if (!in_trench) occupy(o, x, y, z, kind, 0);
  }
  pass_coordination(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0);
  pass_probability(o, 21, Len3 + 9, 0, LY - 1, 0, LX - 1, 0, 0);
}

/* The three closing passes shared by SetShape.f:356-503 and SetSiCSF.f:325-470: (2) active atoms lose the bonds their
 * partner does not return, (3) bulk atoms (z >= 13) with a non-reciprocal bond become active, (4) the empty sites bonded
 * to the active atoms join the active list */
static void shape_passes(Orc *o) {
  int LX = o->LenX, LY = o->LenY;
  int32_t Site[3], NNSite[3], NextN[4][3]; int NewSites[2][3], ev; double Prob;
  /* the reciprocity census shared by passes 2 and 3, :366-385 / :413-432 */
#define SH_RECIPR_COOR(CoorVar) do { \
    for (int j = 0; j < 4; j++) { \
      int32_t wn = LC(o, NextN[j]); \
      if (IBITS(wn, ORC_PosOcc, ORC_LenOcc) != 0) { \
        int CoorNN = IBITS(wn, ORC_PosCoor, ORC_LenCoor), IndNN = LI(o, NextN[j]); \
        const int32_t (*NextNN)[3] = CoorNN == 4 ? bulk_nn(o, IndNN) : mc_nn(o, IndNN); \
        if (!recipr(NextNN, Site)) CoorVar--; \
      } \
    } } while (0)
  /* pass 2, :358-398: active atoms lose the bonds their partner does not return */
  for (int i = 1; i <= o->NumAdAtom; i++) {
    cp3(Site, o->ListAdAtom[i].xyz); cpnn(NextN, o->ListAdAtom[i].nn);
    int32_t *w = &LC(o, Site);
    int Coor = IBITS(*w, ORC_PosCoor, ORC_LenCoor), Coorold = Coor;
    SH_RECIPR_COOR(Coor);
    if (Coor != Coorold) {
      int IndS = LI(o, Site), IAt = IBITS(*w, PSIC(o), LSIC(o));
      MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
      Get_Prob(o, 1, IAt, 0, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
      set_rate(o, IndS, ev, Prob); orc_update_tree(o, IndS, Prob);
    }
  }
  /* pass 3, :404-449: bulk atoms (z >= 13) with a non-reciprocal bond become active; the loop bound is the list size at
   * entry while EraseBulk swaps the last record into the hole, exactly as the Fortran DO does */
  int N0 = o->NumAtoms;
  for (int i = 1; i <= N0; i++) {
    cp3(Site, o->ListAtom[i].xyz);
    if (Site[2] < 13) continue;
    cpnn(NextN, o->ListAtom[i].nn);
    int32_t *w = &LC(o, Site);
    int Coor = IBITS(*w, ORC_PosCoor, ORC_LenCoor), Coorold = Coor;
    SH_RECIPR_COOR(Coor);
    if (Coor != Coorold) {
      int IAt = IBITS(*w, PSIC(o), LSIC(o));
      MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
      Get_Prob(o, 1, IAt, 0, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
      EraseBulk(o, Site);
      AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
    }
  }
#undef SH_RECIPR_COOR
  /* pass 4, :455-503: the empty sites bonded to the active atoms join the active list */
  int NumAdAtomold = o->NumAdAtom;
  for (int i = 1; i <= NumAdAtomold; i++) {
    cp3(NNSite, o->ListAdAtom[i].xyz);
    for (int j = 0; j < 4; j++) {
      cp3(Site, o->ListAdAtom[i].nn[j]);
      int32_t *w = &LC(o, Site);
      if (IBITS(*w, ORC_PosOcc, ORC_LenOcc) != 0) continue;
      memset(NextN, 0, sizeof NextN);
      int Coor = IBITS(*w, ORC_PosCoor, ORC_LenCoor) + 1;
      if (Coor == 1) {
        cp3(NextN[0], NNSite);
        MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
        Get_Prob_Ini(o, 0, 0, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
        AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
      } else if (Coor == 2 || Coor == 3) {
        int IndS = LI(o, Site);
        cpnn(NextN, o->ListAdAtom[IndS].nn);
        if (Coor == 2) {
          cp3(NextN[1], NNSite);
          orc_find_nn(NextN[0], NextN[1], Site, NewSites, LX, LY);
          for (int k = 0; k < 3; k++) { NextN[2][k] = NewSites[0][k]; NextN[3][k] = NewSites[1][k]; }
        }
        MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
        Get_Prob_Ini(o, 0, 0, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
        if (Coor == 2) cpnn(o->ListAdAtom[IndS].nn, NextN);
        set_rate(o, IndS, ev, Prob); orc_update_tree(o, IndS, Prob);
      } else orc_fatal(o, "Vacancy in the SF setting");
    }
  }
}

/* src/initial/SetShape.f:25-507 (InitSt 'SH') and its caller src/main.f:281-286: an arbitrary structure given as a list of
 * filled sites "sym x y z" (lattice units; planes z<=3 and z>=LenZ-4 become walls), neighbours found by scanning the whole
 * first shell of the super-lattice (the 32 vectors with |v|^2 = 27: cubic and hexagonal bonds), then three passes that
 * restore reciprocity of the bonds and create the empty active sites around the under-coordinated atoms */
static void set_shape(Orc *o) {
  int LX = o->LenX, LY = o->LenY, LZ = o->LenZ;
  char line[4096], sym[8];
  LOG(o, " Reading input geometry file (SHAPE):  %s\n", o->cadfilename);
  FILE *f = fopen(o->cadfilename, "r");
  if (!f) orc_fatal(o, "cannot open %s", o->cadfilename);
  if (!next_line(f, line, sizeof line)) orc_fatal(o, "shape file: EOF");
  int NFilled = (int)strtol(line, NULL, 10);
  if (!next_line(f, line, sizeof line)) orc_fatal(o, "shape file: EOF");          /* second line skipped, :54 */
  for (int ia = 0; ia < NFilled; ia++) {
    int x, y, z;
    if (!next_line(f, line, sizeof line) || sscanf(line, " %7s %d %d %d", sym, &x, &y, &z) != 4) orc_fatal(o, "shape file: bad record %d", ia + 1);
    if (x < 0 || x >= LX || y < 0 || y >= LY || z < 0 || z > LZ) orc_fatal(o, "shape file: site %d %d %d outside the box", x, y, z);
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    *w |= 1 << ORC_PosOcc;
    if (z >= LZ - 4 || z <= 3) MVBITS(113, ORC_LenIndex, w, ORC_PosIndex);
    else if (!strcmp(sym, "Si")) *w |= 1 << PSI(o);
    else if (!strcmp(sym, "C") || !strcmp(sym, "Ge")) *w |= 1 << PC_(o);
  }
  fclose(f);
  LOG(o, " Done reading input geometry file\n");
  int ListNN[32][3], nlist = 0;
  for (int k = -6; k <= 6; k++) for (int j = -6; j <= 6; j++) for (int i = -6; i <= 6; i++)
    if (i * i + j * j + k * k == 27) { ListNN[nlist][0] = i; ListNN[nlist][1] = j; ListNN[nlist][2] = k; nlist++; }
  int32_t Site[3], NNSite[3], NextND[4][3], Buf[3]; int NewSites[2][3], ev; double Prob;
  /* pass 1, :83-181: coordination from the full shell, bonds = the occupied shell sites (+ FIND_NN for the missing two) */
  for (int z = 4; z <= LZ - 5; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    if (IBITS(*w, ORC_PosOcc, ORC_LenOcc) == 0) continue;
    int Coord = 0, j = 0;
    Site[0] = x; Site[1] = y; Site[2] = z;
    memset(NextND, 0, sizeof NextND);
    for (int i = 0; i < 32; i++) {
      for (int k = 0; k < 3; k++) NNSite[k] = Site[k] + ListNN[i][k];
      if (NNSite[0] >= LX) NNSite[0] -= LX;
      if (NNSite[0] < 0) NNSite[0] += LX;
      if (NNSite[1] >= LY) NNSite[1] -= LY;
      if (NNSite[1] < 0) NNSite[1] += LY;
      if (IBITS(LC(o, NNSite), ORC_PosOcc, ORC_LenOcc) > 0) {
        Coord++;
        if (j >= 4) orc_fatal(o, "SetShape: site %d %d %d has more than four occupied shell sites (the reference overruns NextND)", x, y, z);
        cp3(NextND[j++], NNSite);
      }
    }
    MVBITS(Coord, ORC_LenCoor, w, ORC_PosCoor);
    if (Coord <= 3) {
      cp3(Buf, NextND[2]);
      orc_find_nn(NextND[0], NextND[1], Site, NewSites, LX, LY);
      for (int k = 0; k < 3; k++) { NextND[2][k] = NewSites[0][k]; NextND[3][k] = NewSites[1][k]; }
      if (Coord == 3) {
        int Check = 1;
        for (int i = 0; i < 4; i++) if (eq3(NextND[i], Buf)) Check = 0;
        if (Check) LOG(o, " Coord.eq.3 wrong check Buf %d %d %d\n", Buf[0], Buf[1], Buf[2]);
      }
      int IAt = IBITS(*w, PSIC(o), LSIC(o));
      Get_Prob_Ini(o, 1, IAt, Coord, Site, (const int32_t (*)[3])NextND, &ev, &Prob);
      AddMC(o, Site, (const int32_t (*)[3])NextND, ev, Prob);
    } else AddBulk(o, Site, (const int32_t (*)[3])NextND);
  }
  shape_passes(o);
}

/* src/initial/SetSiCSF.f:25-473 (InitSt 'E'; main.f:267-272 passes Len = Len1 and the vertex VSF = (Len2, Len3, Len1-Len2)):
 * a 3C-SiC (001) slab with a triple stacking fault -- the atoms of two {111} planes through the vertex are shifted off the
 * cubic sub-lattice by (2,2,4) and (-2,-2,-4) -- then the neighbour search of SetShape.f for the atoms the cubic bonds do
 * not explain, and the same three closing passes */
static void set_sic_sf(Orc *o, int Len, const int VSF[3]) {
  int LX = o->LenX, LY = o->LenY, kind;
  int32_t Site[3], NNSite[3], NextN[4][3], NextND[4][3], Buf[3]; int NewSites[2][3], ev; double Prob;
  int BoxSF = Len - VSF[2];
  int ListNN[32][3], nlist = 0;
  for (int k = -6; k <= 6; k++) for (int j = -6; j <= 6; j++) for (int i = -6; i <= 6; i++)
    if (i * i + j * j + k * k == 27) { ListNN[nlist][0] = i; ListNN[nlist][1] = j; ListNN[nlist][2] = k; nlist++; }
  if (VSF[0] + BoxSF > LX || VSF[1] + BoxSF > LY || BoxSF <= 0) orc_fatal(o, "error SF dimension is not correctly set");
  if (!(is_cubic_site(VSF[0], VSF[1], VSF[2], &kind) && kind == 0)) { LOG(o, " SF Vertex position in wrong Site %d %d %d %d\n", VSF[0], VSF[1], VSF[2], Len); orc_fatal(o, "SF Vertex position in wrong Site"); }
  LOG(o, " SF Vertex position in a Si Site %d %d %d %d\n", VSF[0], VSF[1], VSF[2], Len);
  for (int z = 0; z <= 12; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)               /* :64-111 */
    if (is_cubic_site(x, y, z, &kind)) {
      occupy(o, x, y, z, kind, 1);
      MVBITS(4, ORC_LenCoor, &o->LattCoo[CI(o, x, y, z)], ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  for (int z = 13; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          /* :114-168 */
    if (!is_cubic_site(x, y, z, &kind)) continue;
    int p1 = -1, p2 = -1;
    if (x >= VSF[0] && y >= VSF[1] && z >= VSF[2]) {
      int d = (x - VSF[0]) + (y - VSF[1]) - (z - VSF[2]);
      p1 = d - (kind ? 3 : 12); p2 = d - (kind ? 15 : 24);
    }
    int tx = x, ty = y, tz = z, put = 1;
    if (p1 == 0) { put = (x + 2 < LX && y + 2 < LY && z + 4 < Len); tx = x + 2; ty = y + 2; tz = z + 4; }
    else if (p2 == 0) { put = (x - 2 > 0 && y - 2 > 0 && z - 4 > 0); tx = x - 2; ty = y - 2; tz = z - 4; }
    if (put) occupy(o, tx, ty, tz, kind, 1);
  }
  /* neighbours of an occupied cell: the cubic bonds by sub-lattice (x/3 etc. are integer divisions, also for the shifted
   * atoms), and, when fewer than two of them are filled, the whole first shell (:173-211, :233-270) */
#define SF_NEIGHBOURS() do { \
    Coor = 0; Coord = 0; nd = 0; \
    Site[0] = x; Site[1] = y; Site[2] = z; \
    dia_nn(o, x, y, z, ((x / 3 + y / 3 + z / 3) % 4 == 3) ? 1 : 0, NextN); \
    for (int i = 0; i < 4; i++) if (IBITS(LC(o, NextN[i]), ORC_PosOcc, ORC_LenOcc) > 0) Coor++; \
    if (Coor < 2) { \
      memset(NextND, 0, sizeof NextND); \
      for (int i = 0; i < 32; i++) { \
        for (int k = 0; k < 3; k++) NNSite[k] = Site[k] + ListNN[i][k]; \
        if (NNSite[0] >= LX) NNSite[0] -= LX; \
        if (NNSite[0] < 0) NNSite[0] += LX; \
        if (NNSite[1] >= LY) NNSite[1] -= LY; \
        if (NNSite[1] < 0) NNSite[1] += LY; \
        if (IBITS(LC(o, NNSite), ORC_PosOcc, ORC_LenOcc) > 0) { \
          Coord++; \
          if (nd >= 4) orc_fatal(o, "SetSiCSF: site %d %d %d has more than four occupied shell sites (the reference overruns NextND)", x, y, z); \
          cp3(NextND[nd++], NNSite); \
        } \
      } \
    } } while (0)
#define SF_BOND_CHECK(n) do { \
    for (int i = 0; i < (n) - 1; i++) for (int k = i + 1; k < (n); k++) { \
      int DSN = (NextND[i][0] - NextND[k][0]) * (NextND[i][0] - NextND[k][0]) + (NextND[i][1] - NextND[k][1]) * (NextND[i][1] - NextND[k][1]) + \
                (NextND[i][2] - NextND[k][2]) * (NextND[i][2] - NextND[k][2]); \
      if (DSN != 72) LOG(o, " not correct bonds direction\n %d %d %d %d %d %d\n", NextND[i][0], NextND[i][1], NextND[i][2], NextND[k][0], NextND[k][1], NextND[k][2]); \
    } } while (0)
  int Coor, Coord, nd;
  for (int z = 13; z <= Len + 6; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          /* :171-229 peel off what stays <= 1-coordinated */
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    if (IBITS(*w, ORC_PosOcc, ORC_LenOcc) == 0) continue;
    SF_NEIGHBOURS();
    if (Coord > Coor) { SF_BOND_CHECK(nd); if (Coord <= 1) *w = 0; }
    else if (Coor <= 1) *w = 0;
  }
  for (int z = 13; z <= Len + 6; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          /* :231-323 lists */
    int32_t *w = &o->LattCoo[CI(o, x, y, z)];
    if (IBITS(*w, ORC_PosOcc, ORC_LenOcc) == 0) continue;
    SF_NEIGHBOURS();
    int IAt = IBITS(*w, PSIC(o), LSIC(o));
    if (Coord > Coor) {
      MVBITS(Coord, ORC_LenCoor, w, ORC_PosCoor);
      if (Coord <= 3) {
        cp3(Buf, NextND[2]);
        orc_find_nn(NextND[0], NextND[1], Site, NewSites, LX, LY);
        for (int k = 0; k < 3; k++) { NextND[2][k] = NewSites[0][k]; NextND[3][k] = NewSites[1][k]; }
        if (Coord == 3) {
          int Check = 1;
          for (int i = 0; i < 4; i++) if (eq3(NextND[i], Buf)) Check = 0;
          if (Check) LOG(o, " Coord.eq.3 wrong check Buf %d %d %d\n", Buf[0], Buf[1], Buf[2]);
        }
        SF_BOND_CHECK(4);
        Get_Prob_Ini(o, 1, IAt, Coord, Site, (const int32_t (*)[3])NextND, &ev, &Prob);
        AddMC(o, Site, (const int32_t (*)[3])NextND, ev, Prob);
      } else AddBulk(o, Site, (const int32_t (*)[3])NextND);
    } else {
      MVBITS(Coor, ORC_LenCoor, w, ORC_PosCoor);
      if (Coor <= 3) {
        Get_Prob_Ini(o, 1, IAt, Coor, Site, (const int32_t (*)[3])NextN, &ev, &Prob);
        AddMC(o, Site, (const int32_t (*)[3])NextN, ev, Prob);
      } else AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  }
#undef SF_NEIGHBOURS
#undef SF_BOND_CHECK
  shape_passes(o);
}

/* src/initial/SetAPBSymZimb1.f:25-497 (SUBROUTINE SetAFBSymZimb1, InitSt 'A'): 3C-SiC (001) slab whose upper half carries two
 * antiphase domains -- the ideal Si/C assignment inside the four corner diamonds |x|+|y| < Lenh (periodic images), the
 * swapped one outside |x|+|y| > Lenh+3, an empty seam in between.  Faithful to the routine's bookkeeping: every cubic site
 * of z = 0..Len-4 goes into the bulk list whether occupied or not (:42-178), coordination 4 is written only at z = 0, and
 * the closing "set the probability" pass (z = 21..Len) lists the 4-coordinated atoms a second time (:474) */
static void set_afb_sym_zimb1(Orc *o, int Len) {
  int kind, LX = o->LenX, LY = o->LenY; int32_t Site[3], NextN[4][3];
  int Lenh = LX / 2 - 1 + 24;
  for (int z = 0; z <= Len / 2 - 4; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)          /* epi substrate */
    if (is_cubic_site(x, y, z, &kind)) {
      occupy(o, x, y, z, kind, 1);
      if (z == 0) MVBITS(4, ORC_LenCoor, &o->LattCoo[CI(o, x, y, z)], ORC_PosCoor);
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
#define AFB_IN(cmp, d) ((abs(x) + abs(y) cmp (d)) || (abs(x - LX) + abs(y) cmp (d)) || (abs(x) + abs(y - LY) cmp (d)) || (abs(x - LX) + abs(y - LY) cmp (d)))
#define AFB_OUT(d) ((abs(x) + abs(y) > (d)) && (abs(x - LX) + abs(y) > (d)) && (abs(x) + abs(y - LY) > (d)) && (abs(x - LX) + abs(y - LY) > (d)))
  for (int z = Len / 2 - 3; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {  /* :99-178 (bulk list) and :181-259 (top three planes, no list) */
    if (!is_cubic_site(x, y, z, &kind)) continue;
    if (AFB_IN(<, Lenh)) occupy(o, x, y, z, kind, 1);
    else if (AFB_OUT(Lenh + 3)) occupy(o, x, y, z, !kind, 1);
    if (z <= Len - 4) {
      Site[0] = x; Site[1] = y; Site[2] = z;
      dia_nn(o, x, y, z, kind, NextN);
      AddBulk(o, Site, (const int32_t (*)[3])NextN);
    }
  }
#undef AFB_IN
#undef AFB_OUT
  pass_coordination(o, 1, Len, 0, LY - 1, 0, LX - 1, 0);
  pass_probability(o, 21, Len, 0, LY - 1, 0, LX - 1, 0, 0);
}

/* list-directed READ of one plane (LattGeo / TempStat files): take n values starting on a fresh
 * record, continue onto following records only if the record is short, and DISCARD whatever is left
 * on the last record read -- the regression LattGeo file has a line with one value too many. */
static int read_plane(Orc *o, FILE *f, double *buf, size_t n) {
  static char *line = NULL; static size_t cap = 0;
  size_t got = 0;
  while (got < n) {
    if (getline(&line, &cap, f) < 0) { orc_fatal(o, "short read in geometry/temperature file"); return -1; }
    char *p = line, *e;
    while (got < n) {
      double v = strtod(p, &e);
      if (e == p) break;
      buf[got++] = v; p = e;
      while (*p == ',') p++;
    }
  }
  return 0;
}
/* src/initial/SetT.f:25-50 */
static void set_T(Orc *o) {
  LOG(o, " Reading input temperature file: %s\n", o->tempfilename);
  FILE *f = fopen(o->tempfilename, "r");
  if (!f) orc_fatal(o, "cannot open %s", o->tempfilename);
  size_t n = (size_t)o->LenX * o->LenY; double *buf = (double *)malloc(n * sizeof(double));
  for (int z = 0; z < o->LenZ; z++) {
    read_plane(o, f, buf, n);
    for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++) {
      int t = (int)buf[x + (size_t)o->LenX * y];
      if (t != 10) MVBITS(t, ORC_LenT, &o->LattCoo[CI(o, x, y, z)], ORC_PosT);
    }
  }
  free(buf); fclose(f);
  LOG(o, " Done reading input temperatures file\n");
}
/* src/initial/SetCAD.f:25-421 (RunType 'C' checkpoint branch :54-79 included) */
static void set_cad(Orc *o) {
  int kind, LX = o->LenX, LY = o->LenY, LZ = o->LenZ, la = !strcmp(o->InitSt, "LA");
  size_t n = (size_t)LX * LY; double *buf = (double *)malloc(n * sizeof(double));
  double LenSiGe = o->la_tree ? o->LenSiGe : 0.0;     /* src/: un-initialised local (SetCAD.f:38), irrelevant while alloyfraction==0; la/: read from start.dat */
  if (o->alloyfraction > 0) LOG(o, " Setting up a SiGe layer with fraction = %.17g\n", o->alloyfraction);
  else LOG(o, " Setting up Silicon\n");
  if (o->RunType[0] == 'C') {
    LOG(o, " Reading checkpoint file %s\n", o->cadfilename);
    FILE *f = fopen(o->cadfilename, "r"); if (!f) orc_fatal(o, "cannot open %s", o->cadfilename);
    for (int z = 0; z < LZ; z++) {
      read_plane(o, f, buf, n);
      for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        int32_t *w = &o->LattCoo[CI(o, x, y, z)];
        if (la) { int t = IBITS(*w, ORC_PosT, ORC_LenT); *w = (int32_t)buf[x + (size_t)LX * y]; MVBITS(t, ORC_LenT, w, ORC_PosT); if (t > o->tmax) o->tmax = t; }
        else *w = (int32_t)buf[x + (size_t)LX * y];
      }
    }
    fclose(f);
    LOG(o, " Done reading checkpoint file\n");
  } else {
    set_wall_planes(o, 0, 3, 0, LX - 1);
    set_wall_planes(o, LZ - 4, LZ - 1, 0, LX - 1);
    LOG(o, " Reading input geometry file: %s\n", o->cadfilename);
    FILE *f = fopen(o->cadfilename, "r"); if (!f) orc_fatal(o, "cannot open %s", o->cadfilename);
    /* SetCAD.f:99-102.  ORC_CAD_NOSKIP=1 emulates the OLDER build that produced the CAD/LA goldens
     * (it did not skip the four wall planes, so file line k lands on z=k+4); see DESIGN.md. */
    if (!getenv("ORC_CAD_NOSKIP")) for (int z = 0; z <= 3; z++) read_plane(o, f, buf, n);
    for (int z = 4; z <= LZ - 5; z++) {
      read_plane(o, f, buf, n);
      for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        int Id = (int)buf[x + (size_t)LX * y];
        int32_t *w = &o->LattCoo[CI(o, x, y, z)];
        if (Id == 10) { *w |= 1 << ORC_PosOcc; MVBITS(113, ORC_LenIndex, w, ORC_PosIndex); }
        else if (Id == 1 && is_cubic_site(x, y, z, &kind)) {
          int fill = 0;
          if (la) {
            /* la/SetCAD.f:140,163,190 count the solid sites below the air gap before the nucleus is carved out.  NOTE: only
             * the socket branch of la/SetCAD.f (:108-255) does this; its file branch (:257-336) leaves NumOcc0 = 0, which
             * stops the run after the first event -- the oracle counts in both (the la/ fork is driven through sockets) */
            if (o->la_tree && z <= (LZ - (int)(o->LenVac / o->sf))) o->NumOcc0++;
            if (!strcmp(o->Homogeneous, "T")) fill = (z <= (LZ - (int)((o->LenVac + o->LenNuc) / o->sf)));
            else if (!strcmp(o->Homogeneous, "F")) {
              int x0 = LX / 2, y0 = LY / 2, z0 = LZ - (int)(o->LenVac / o->sf);
              fill = ((x - x0) * (x - x0) + (y - y0) * (y - y0) + (z - z0) * (z - z0)) > (int)((o->LenNuc / o->sf) * (o->LenNuc / o->sf));
            } else if (o->la_tree && !strcmp(o->Homogeneous, "R")) {      /* la/SetCAD.f:186-207: semi-cylindrical nucleus along y */
              int x0 = LX / 2, z0 = LZ - (int)(o->LenVac / o->sf);
              fill = ((x - x0) * (x - x0) + (z - z0) * (z - z0)) > (int)((o->LenNuc / o->sf) * (o->LenNuc / o->sf));
            } else LOG(o, " ERROR: Homogeneous needs to be either T or F\n");
          } else fill = 1;
          if (fill) {
            *w |= 1 << ORC_PosOcc;
            /* libgfortran RAND() is not restated (SURVEY 8c trap 6): alloy init is "parity unpinned";
             * with alloyfraction==0 every draw selects Si, which is all the goldens exercise. */
            double rr = 1.0;
            if (o->alloyfraction > 0 && !(la && z <= (LZ - (int)((o->LenVac + LenSiGe) / o->sf)))) {
              o->alloy_lcg = o->alloy_lcg * 6364136223846793005ull + 1442695040888963407ull;
              rr = (double)(o->alloy_lcg >> 11) * (1.0 / 9007199254740992.0);
            }
            *w |= 1 << (rr >= o->alloyfraction ? PSI(o) : PC_(o));
            if (o->la_tree) o->CountCrystal[rr >= o->alloyfraction ? 0 : 1]++;                 /* la/SetCAD.f:150-155 */
          }
        }
      }
    }
    fclose(f);
    LOG(o, " Done reading input geometry file\n");
    if (o->la_tree) {                                          /* la/SetCAD.f:343-351 */
      int sum = o->CountCrystal[0] + o->CountCrystal[1] + o->CountCrystal[2];
      for (int k = 0; k < o->NCrystal; k++) {
        if (o->ListCrystal[k] == 14) o->InitCountCrystal[k] = o->CountCrystal[k] + (int)((o->NumOcc0 - sum) * (1 - o->alloyfraction));
        else if (o->ListCrystal[k] == 32) o->InitCountCrystal[k] = o->CountCrystal[k] + (int)((o->NumOcc0 - sum) * o->alloyfraction);
      }
      for (int k = 0; k < 3; k++) o->CountLiquid[k] = o->InitCountCrystal[k] - o->CountCrystal[k];
    }
    pass_coordination(o, 0, LZ - 1, 0, LY - 1, 0, LX - 1, 1);
  }
  free(buf);
  pass_probability(o, 0, LZ - 1, 0, LY - 1, 0, LX - 1, 1, 1);
}

/* src/main.f:225-300 */
int orc_setup(Orc *o) {
  o->jb_set = 1;
  if (setjmp(o->jb)) { o->jb_set = 0; return -2; }
  const char *s = o->InitSt;
  if (o->la_tree && strcmp(s, "LA") && strcmp(s, "IN")) orc_fatal(o, "oracle: the la/ tree variant restates InitSt LA and IN only");
  if (!strcmp(s, "F")) { o->Len1 -= o->Len1 % 24; set_flat(o, o->Len1, 0, 1); }
  else if (!strcmp(s, "Fw")) {                                 /* SetSiC3Cww.f: SetSiC3C.f between wall planes, plus a wall block from Len2 up */
    o->Len1 -= o->Len1 % 24; o->Len2 -= o->Len2 % 24;
    set_flat(o, o->Len1, 1, 1);
    set_wall_planes(o, o->Len2, o->LenZ - 7, 0, o->LenX - 1);
  } else if (!strcmp(s, "Fa")) {                               /* SetSiC3Cwwaperture.f: the wall block leaves an aperture of Len3 along x */
    o->Len1 -= o->Len1 % 24; o->Len2 -= o->Len2 % 24; o->Len3 -= o->Len3 % 24;
    set_flat(o, o->Len1, 1, 1);
    set_wall_planes(o, o->Len2, o->LenZ - 7, 0, (int)(0.5f * (float)(o->LenX - o->Len3)) - 1);
    set_wall_planes(o, o->Len2, o->LenZ - 7, (int)(0.5f * (float)(o->LenX + o->Len3)), o->LenX - 1);
  }
  else if (!strcmp(s, "FS")) { o->Len1 -= o->Len1 % 24; set_flat(o, o->Len1, 1, 0); }
  else if (!strcmp(s, "SG")) { o->Len1 -= o->Len1 % 24; set_flat(o, o->Len1, 1, 2); }
  else if (!strcmp(s, "C")) set_nc(o, o->Len1, o->Len2, o->Len3, 0, 1);
  else if (!strcmp(s, "S")) set_nc(o, o->Len1, o->Len1, o->Len1, 1, 1);
  else if (!strcmp(s, "SS")) set_nc(o, o->Len1, o->Len1, o->Len1, 1, 0);      /* SetNCSpSi.f: the sphere of SetNCSp.f in pure Si */
  else if (!strcmp(s, "I")) set_invp_off(o, o->Len1, o->Len2, 0);
  else if (!strcmp(s, "D")) set_invp_off(o, o->Len1, o->Len2, 2);             /* SetInvPC.f */
  else if (!strcmp(s, "Z")) set_invp_off(o, o->Len1, o->Len2, 1);             /* SetInvPSi.f */
  else if (!strcmp(s, "K")) set_invp_off(o, o->Len1, o->Len2, 3);             /* SetInvPwOnlySiorC.f */
  else if (!strcmp(s, "L")) set_invp_off(o, o->Len1, o->Len2, 4);             /* SetInvPwOnlyCorSi.f */
  else if (!strcmp(s, "J")) set_invp_off(o, o->Len1, o->Len2, 5);             /* SetInvPwAPB.f */
  else if (!strcmp(s, "T")) set_trench(o, o->Len1, o->Len2, o->Len3);         /* SetTrench.f */
  else if (!strcmp(s, "FF")) {
    o->Len1 -= o->Len1 % 24; o->Len2 -= o->Len2 % 24; o->Len3 -= o->Len3 % 24; o->Len4 -= o->Len4 % 24;
    set_finfet(o, o->Len1, o->Len2, o->Len3, o->Len4);
  } else if (!strcmp(s, "ST")) set_trench_si(o, o->Len1, o->Len2, o->Len3, o->angle, o->Ttop, o->Tbottom);
  else if (!strcmp(s, "A")) set_afb_sym_zimb1(o, o->Len1);
  else if (!strcmp(s, "SH")) set_shape(o);
  else if (!strcmp(s, "E")) {                                  /* main.f:267-272 */
    int VSF[3] = {o->Len2, o->Len3, o->Len1 - o->Len2};
    LOG(o, " %d %d %d %d\n", o->Len1, VSF[0], VSF[1], VSF[2]);
    set_sic_sf(o, o->Len1, VSF);
  }
  else if (!strcmp(s, "IN")) set_cad(o);
  else if (!strcmp(s, "LA")) { o->tmax = 0; set_T(o); set_cad(o); LOG(o, " Max temperature is:  %d K\n", o->tmax); }
  else orc_fatal(o, "initialization not_implemented in oracle: %s", s);
  if (o->la_tree) {                                            /* la/main.f:592-611,628 */
    o->NumOcc1 = o->NumAtoms + o->NumAdAtomOcc;
    for (int k = 1; k <= o->NCrystal; k++) {
      if (k == 1) o->InitCountCrystal[0] = o->CountCrystal[0] + (int)((o->NumOcc0 - o->NumOcc1) * (1 - o->alloyfraction));
      else if (k == 2) o->InitCountCrystal[1] = o->CountCrystal[1] + (int)((o->NumOcc0 - o->NumOcc1) * o->alloyfraction);
    }
    LOG(o, " \n Initial number of occupied sites (before nucleation):  %d\n", o->NumOcc0);
    LOG(o, " Initial CountCrystal (before nucleation):  %d %d %d\n", o->InitCountCrystal[0], o->InitCountCrystal[1], o->InitCountCrystal[2]);
    LOG(o, " Initial number of occupied sites (after nucleation):  %d\n", o->NumOcc1);
    LOG(o, " Initial CountCrystal (after nucleation):  %d %d %d\n", o->CountCrystal[0], o->CountCrystal[1], o->CountCrystal[2]);
    LOG(o, " Initial number of liquid sites (after nucleation):  %d\n", o->NumOcc0 - o->NumOcc1);
    LOG(o, " Initial CountLiquid (after nucleation):  %d %d %d\n \n", o->CountLiquid[0], o->CountLiquid[1], o->CountLiquid[2]);
    LOG(o, " Atoms,Voids,AdAtoms,AdAtomsOcc %d %d %d %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom, o->NumAdAtomOcc);
  } else
  LOG(o, " Atoms,Voids,AdAtoms %d %d %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom);
  for (int i = 0; i < 3; i++) o->CountCrystalOld[i] = o->CountCrystal[i];
  orc_write_frame(o, o->Time, o->IterOut, o->framecounter);
  o->jb_set = 0;
  return 0;
}

/* ------------------------------------------------------------------ main loop */
/* src/main.f:304-371, one trip through the DO body */
static int iterate_inner(Orc *o) {
  int Ind; double Patom;
  o->IterOut = o->Iter;
  if (o->la_tree) for (int k = 0; k < 3; k++) o->CountLiquid[k] = o->InitCountCrystal[k] - o->CountCrystal[k];   /* la/main.f:636 */
  if (o->NumAdAtom <= 20) {
    if (o->la_tree) LOG(o, " few MC particles: stop MC %lld %d %d\n", (long long)o->Iter, o->NumAdAtom, o->NumAdAtomOcc);
    else LOG(o, " few MC particles: stop MC %lld %d\n", (long long)o->Iter, o->NumAdAtom);
    o->done = 2; return 1;
  }
  pick_tree_event(o, &Ind, &Patom);
  o->Time = o->Time + 1.0 / o->Tree[1];
  int Itrans = o->ListAdAtom[Ind].ev;
  cp3(o->last_site, o->ListAdAtom[Ind].xyz);
  if (!strcmp(o->ExitStrategy, "Iter")) {
    if (o->Iter % o->OutMolMol == 0) {
      o->framecounter = o->IterOut / o->OutMolMol;
      orc_write_frame(o, o->Time, o->IterOut, o->framecounter);
      const int32_t *S = o->ListAdAtom[Ind].xyz;
      LOG(o, " Iter, Time, Site %lld %.17g %d %d %d\n", (long long)o->Iter, o->Time, S[0], S[1], S[2]);
      if (o->la_tree) LOG(o, " NatNvoidNadNadocc %d %d %d %d\n Number of occupied sites:  %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom, o->NumAdAtomOcc, o->NumAtoms + o->NumAdAtomOcc);
      else LOG(o, " NatNvoidNad %d %d %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom);
    }
  } else if (!strcmp(o->ExitStrategy, "Time")) {
    if (o->Time >= ((double)(o->framecounter + 1) * o->OutTime)) {
      orc_write_frame(o, o->Time, o->IterOut, o->framecounter + 1);
      const int32_t *S = o->ListAdAtom[Ind].xyz;
      LOG(o, " Iter, Time, Site %lld %.17g %d %d %d\n", (long long)o->Iter, o->Time, S[0], S[1], S[2]);
      if (o->la_tree) LOG(o, " NatNvoidNadNadocc %d %d %d %d\n Number of occupied sites:  %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom, o->NumAdAtomOcc, o->NumAtoms + o->NumAdAtomOcc);
      else LOG(o, " NatNvoidNad %d %d %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom);
      o->framecounter++;
    }
  }
  OrcTrace *t = NULL;
  o->last_arm = 0;
  if (o->trace && o->trace_n < o->trace_cap) {
    t = &o->trace[o->trace_n++];
    t->iter = o->Iter; t->ind = Ind; cp3(t->site, o->ListAdAtom[Ind].xyz); t->itrans = Itrans; t->pad = 0; t->tree1 = o->Tree[1]; t->time = o->Time;
  }
  if (Itrans >= 0 && Itrans < 8) o->n_events[Itrans]++; else o->n_events[7]++;
  if (Itrans < EV_EVAP(o)) {
    /* la/main.f:675-681: SiGe melt pool -- when species Itrans has been fully re-solidified, deposit the other one */
    if (o->la_tree && !strcmp(o->InitSt, "LA") && o->NCrystal == 2 && o->ListCrystal[0] == 14 && o->ListCrystal[1] == 32 && Itrans >= 1)
      if (o->CountCrystal[Itrans - 1] >= o->InitCountCrystal[Itrans - 1]) { Itrans = Itrans % 2 + 1; if (t) t->itrans = Itrans; }
    Deposition(o, &Ind, Itrans);
    if (Itrans >= 1) o->CountCrystal[Itrans - 1]++;
  } else if (Itrans == EV_EVAP(o)) Evaporation(o, &Ind);
  else if (Itrans == EV_DES(o)) Desorption(o, Ind);
  else {
    int k = Itrans - EV_DES(o); if (k > 3) { o->q_oob++; k = 3; }
    Absorption(o, Ind, o->ListCov[k - 1]);
    o->CountCov[k - 1]++;
  }
  if (t) t->pad = o->last_arm;
  o->Iter++;
  if (!strcmp(o->ExitStrategy, "Iter")) { if (o->Iter > o->MaxIter) { o->done = 1; return 1; } }
  else if (!strcmp(o->ExitStrategy, "Time")) { if (o->Time > o->MaxTime) { o->done = 1; return 1; } }
  if (o->la_tree) {                                            /* la/main.f:714-723: the melt pool has re-solidified */
    o->NumOcc = o->NumAtoms + o->NumAdAtomOcc;
    if (o->NumOcc > o->NumOcc0) {
      LOG(o, " \n Current number of occupied sites:  %d\n    > initial number of occupied sites:  %d\n ... STOPPING MulSKIPS ...\n \n", o->NumOcc, o->NumOcc0);
      o->done = 1; return 1;
    }
  }
  return 0;
}
int orc_iterate(Orc *o) {
  o->jb_set = 1;
  if (setjmp(o->jb)) { o->jb_set = 0; return -2; }
  int r = iterate_inner(o);
  o->jb_set = 0;
  return r;
}
int64_t orc_run(Orc *o, int64_t nmax) {
  int64_t n = 0;
  o->jb_set = 1;
  if (setjmp(o->jb)) { o->jb_set = 0; return -2; }
  while (n < nmax && !o->done) { int r = iterate_inner(o); if (o->done == 2) break; n++; if (r) break; }
  o->jb_set = 0;
  return n;
}

/* src/main.f:374-430 */
int orc_finish(Orc *o) {
  o->jb_set = 1;
  if (setjmp(o->jb)) { o->jb_set = 0; return -2; }
  if (!strcmp(o->ExitStrategy, "Iter")) orc_write_frame(o, o->Time, o->IterOut, o->IterOut / o->OutMolMol);
  else if (!strcmp(o->ExitStrategy, "Time")) orc_write_frame(o, o->Time, o->IterOut, o->framecounter);
  char path[4096];
  if (!strcmp(o->SaveFinalState, "T") && o->write_files) {
    LOG(o, " Writing checkpoint file for a restart run %s\n", o->restartfilename);
    snprintf(path, sizeof path, "%s%s%s", o->restartfilename[0] == '/' ? "" : o->outdir, o->restartfilename[0] == '/' ? "" : "/", o->restartfilename);
    FILE *f = fopen(path, "w");
    if (f) {
      for (int z = 0; z < o->LenZ; z++) {
        for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++) fprintf(f, " %11d", o->LattCoo[CI(o, x, y, z)]);
        fputc('\n', f);
      }
      fclose(f);
    }
    LOG(o, " Done writing checkpoint file for a restart run %s\n", o->restartfilename);
  }
  if (!strcmp(o->SaveCoo, "T") && o->write_files) {
    LOG(o, " Writing Coor (LA phases) file %s\n", o->coofilename);
    snprintf(path, sizeof path, "%s%s%s", o->coofilename[0] == '/' ? "" : o->outdir, o->coofilename[0] == '/' ? "" : "/", o->coofilename);
    FILE *f = fopen(path, "w");
    if (f) {
      for (int z = 0; z < o->LenZ; z++) {
        for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++) {
          int32_t w = o->LattCoo[CI(o, x, y, z)]; int v = 0;
          if (IBITS(w, ORC_PosIndex, ORC_LenIndex) != 113) { if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 1) v = IBITS(w, ORC_PosCoor, ORC_LenCoor); }
          else v = 10;
          fprintf(f, "%4d", v);
        }
        fputc('\n', f);
      }
      fclose(f);
    }
    LOG(o, " Done writing Coor (LA phases) file %s\n", o->coofilename);
  }
  o->jb_set = 0;
  return 0;
}

/* ------------------------------------------------------------------ writers */
/* Fortran E15.8: 0.dddddddd E+xx */
static void fmt_e15_8(double v, char out[32]) {
  if (v == 0.0) { strcpy(out, " 0.00000000E+00"); return; }
  char t[40]; snprintf(t, sizeof t, "%.7e", fabs(v));
  /* t = d.ddddddde[+-]xx */
  char dig[16]; dig[0] = t[0]; memcpy(dig + 1, t + 2, 7); dig[8] = 0;
  int ex = atoi(strchr(t, 'e') + 1) + 1;
  char body[32];
  if (abs(ex) < 100) snprintf(body, sizeof body, "%s0.%sE%c%02d", v < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  else snprintf(body, sizeof body, "%s0.%s%c%03d", v < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  snprintf(out, 32, "%15s", body);
}
static void atom_line(FILE *f, const char *el, double sf, int x, int y, int z, const char *tail) {
  fprintf(f, "%s%10.5f %10.5f %10.5f%s\n", el, sf * x, sf * y, sf * z, tail);
}
static int noepi_bulk(int x, int y, int z) { int k; return !is_cubic_site(x, y, z, &k); }
static const char *el_of_Z(int Z) { return Z == 6 ? "C " : Z == 32 ? "Ge " : Z == 5 ? "B " : NULL; }

/* src/inpout/WriteMolMolXYZFile.f:22-407 (+ the name logic of WriteMolMolSource.f:22-113 and
 * GetOutputFileName.f:22-83, which reduces to I%08d of Iout) */
int orc_write_frame(Orc *o, double Time, int64_t Iter, int64_t Iout) {
  int LX = o->LenX, LY = o->LenY, LZ = o->LenZ;
  double sf = o->sf;
  char e1[32], e2[32], e3[32], et[32], base[64], path[2048];
  int sic = (o->NCrystal == 2 && (o->ListCrystal[0] == 6 || o->ListCrystal[1] == 6 || o->ListCrystal[2] == 6));
  if (o->la_tree) {                                       /* la/inpout/WriteMolMolSource.f:42-67 */
    int s = 0, so = 0, sl = 0; for (int i = 0; i < 3; i++) { s += o->CountCrystal[i]; so += o->CountCrystalOld[i]; sl += o->CountLiquid[i]; }
    if (o->NCrystal > 1) LOG(o, " SUM(CountCrystal) %d\n", s);
    LOG(o, " CountCrystal %d %d %d\n xCrystal %.9g %.9g %.9g\n", o->CountCrystal[0], o->CountCrystal[1], o->CountCrystal[2],
        (float)o->CountCrystal[0] / s, (float)o->CountCrystal[1] / s, (float)o->CountCrystal[2] / s);
    LOG(o, " CountLiquid %d %d %d\n xLiquid %.9g %.9g %.9g\n", o->CountLiquid[0], o->CountLiquid[1], o->CountLiquid[2],
        (float)o->CountLiquid[0] / sl, (float)o->CountLiquid[1] / sl, (float)o->CountLiquid[2] / sl);
    if (Iter > 1) {
      float d[3]; for (int i = 0; i < 3; i++) d[i] = (float)(o->CountCrystal[i] - o->CountCrystalOld[i]);
      LOG(o, " diffCrystal %.9g %.9g %.9g\n", d[0], d[1], d[2]);
      LOG(o, " Atom fraction from last check: %.17g %.17g %.17g\n", (double)d[0] / (s - so), (double)d[1] / (s - so), (double)d[2] / (s - so));
      for (int i = 0; i < 3; i++) o->CountCrystalOld[i] = o->CountCrystal[i];
    }
    if (o->NCov != 0) LOG(o, " CountCov %d %d %d\n", o->CountCov[0], o->CountCov[1], o->CountCov[2]);
  } else
  if (Iter != 0) {                                        /* WriteMolMolSource.f:42-56 */
    int s = 0, so = 0; for (int i = 0; i < 3; i++) { s += o->CountCrystal[i]; so += o->CountCrystalOld[i]; }
    LOG(o, " SUM(CountCrystal) %d\n CountCrystal %d %d %d\n CountCrystalOld %d %d %d\n", s, o->CountCrystal[0], o->CountCrystal[1], o->CountCrystal[2], o->CountCrystalOld[0], o->CountCrystalOld[1], o->CountCrystalOld[2]);
    float d[3]; for (int i = 0; i < 3; i++) d[i] = (float)(o->CountCrystal[i] - o->CountCrystalOld[i]);
    LOG(o, " diff %.9g %.9g %.9g\n", d[0], d[1], d[2]);
    LOG(o, " xGesolid %.17g %.17g %.17g\n", (double)d[0] / (s - so), (double)d[1] / (s - so), (double)d[2] / (s - so));
    for (int i = 0; i < 3; i++) o->CountCrystalOld[i] = o->CountCrystal[i];
    if (o->NCov != 0) LOG(o, " CountCov %d %d %d\n", o->CountCov[0], o->CountCov[1], o->CountCov[2]);
  }
  if (Iout * 10 >= 1000000000) orc_fatal(o, "GetOutputFileName:: Iter can not be represented");
  snprintf(base, sizeof base, "I%08lld", (long long)Iout);
  LOG(o, " %s\n", base);

  /* census 1, :77-92 */
  int NumAtTotal = 0, NumCovTotal = 0, NoEpiBulk = 0;
  for (int z = 1; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    int32_t w = o->LattCoo[CI(o, x, y, z)];
    int CovInd = IBITS(w, ORC_PosIndex, ORC_LenIndex);
    if (CovInd != 113) {
      int IfOcc = IBITS(w, ORC_PosOcc, ORC_LenOcc), Coo = IBITS(w, ORC_PosCoor, ORC_LenCoor);
      if (IfOcc == 1 && Coo <= 3) NumAtTotal++;
      else if (IfOcc == 0 && CovInd != 0) NumCovTotal++;
    }
  }
  /* census 2, :95-119 */
  for (int z = 0; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    int32_t w = o->LattCoo[CI(o, x, y, z)];
    if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 1 && IBITS(w, ORC_PosCoor, ORC_LenCoor) == 4 && noepi_bulk(x, y, z)) NoEpiBulk++;
  }
  LOG(o, " %d\n", NoEpiBulk);
  if (!o->write_files) {
    /* the _w pass still mutates LattCoo for SiC (WriteMolMolXYZFile.f:267-398); fall through with no files */
  }
  FILE *fx = NULL, *fd = NULL, *fv = NULL, *fw = NULL;
  if (o->write_files) {
    snprintf(path, sizeof path, "%s/%s.xyz", o->outdir, base); fx = fopen(path, "w");
    snprintf(path, sizeof path, "%s/%s_d.xyz", o->outdir, base); fd = fopen(path, "w");
    snprintf(path, sizeof path, "%s/%s_v.xyz", o->outdir, base); fv = fopen(path, "w");
    if (sic) { snprintf(path, sizeof path, "%s/%s_w.xyz", o->outdir, base); fw = fopen(path, "w"); }
    if (!fx || !fd || !fv) orc_fatal(o, "cannot open output files in %s", o->outdir);
  }
  fmt_e15_8(sf * (double)LX, e1); fmt_e15_8(sf * (double)LY, e2); fmt_e15_8(sf * (double)LZ, e3); fmt_e15_8(Time, et);
  if (fd) {                                               /* :125-159 */
    fprintf(fd, " %8d angstroem KMC-time: %s Iter: %12lld\n", NoEpiBulk, et, (long long)Iter);
    fprintf(fd, " periodic %s %s %s\n", e1, e2, e3);
    for (int z = 0; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      int32_t w = o->LattCoo[CI(o, x, y, z)];
      if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 1 && IBITS(w, ORC_PosCoor, ORC_LenCoor) == 4 && noepi_bulk(x, y, z)) {
        int IAt = IBITS(w, PSIC(o), LSIC(o));
        if (IAt == 1) atom_line(fd, "Ge ", sf, x, y, z, "  ");
        if (IAt == 2) atom_line(fd, "B ", sf, x, y, z, "  ");
        if (IAt == 3) atom_line(fd, "Cl ", sf, x, y, z, "  ");
      }
    }
  }
  if (fx) {                                               /* :163-221 */
    fprintf(fx, " %8d angstroem KMC-time: %s Iter: %12lld\n", NumAtTotal + o->NumVoids + NumCovTotal, et, (long long)Iter);
    fprintf(fx, " periodic %s %s %s\n", e1, e2, e3);
    for (int z = 0; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      int32_t w = o->LattCoo[CI(o, x, y, z)];
      if (w == 0) continue;
      int IAt = IBITS(w, PSIC(o), LSIC(o)), Coo = IBITS(w, ORC_PosCoor, ORC_LenCoor), CovInd = IBITS(w, ORC_PosIndex, ORC_LenIndex);
      if (CovInd == 1) atom_line(fx, "H ", sf, x, y, z, "  ");
      if (CovInd == 17) atom_line(fx, "Cl ", sf, x, y, z, "  ");
      if (IAt == 1 && Coo <= 3) atom_line(fx, "Si ", sf, x, y, z, "  ");
      if ((IAt == 2 || IAt == 3) && Coo <= 3) {
        if (o->NCrystal < IAt) orc_fatal(o, "ERR: IAt=%d should not happen, NCrystal=%d", IAt, o->NCrystal);
        const char *el = el_of_Z(o->ListCrystal[IAt - 1]);
        if (!el) orc_fatal(o, "ERROR: 2nd crystal specie can only be C/Ge/B");
        atom_line(fx, el, sf, x, y, z, "  ");
      }
    }
    for (int i = 1; i <= o->NumVoids; i++) atom_line(fx, "O ", sf, o->ListVoid[i].xyz[0], o->ListVoid[i].xyz[1], o->ListVoid[i].xyz[2], "  ");
  }
  if (fv) {                                               /* :228-260 */
    fprintf(fv, " %8d angstroem\n", o->NumVoids);
    fprintf(fv, " periodic %s %s %s\n", e1, e2, e3);
    for (int i = 1; i <= o->NumVoids; i++) {
      int x = o->ListVoid[i].xyz[0], y = o->ListVoid[i].xyz[1], z = o->ListVoid[i].xyz[2];
      if (o->NCov == 0 && o->NCrystal == 2 && o->ListCrystal[1] == 6) {
        int IAt = 0;
        for (int j = 0; j < 4; j++) IAt += IBITS(LC(o, o->ListVoid[i].nn[j]), PSIC(o), LSIC(o));
        if (IAt == 4) atom_line(fv, "Mg ", sf, x, y, z, o->la_tree ? " # SV (Si vacancy)" : " # CV (C vacancy)");       /* la/WriteMolMolXYZFile.f:66-67 swaps the two comments */
        else if (IAt == 8) atom_line(fv, "He ", sf, x, y, z, o->la_tree ? " # CV (C vacancy)" : " # SV (Si vacancy)");
        else if (IAt == 7) atom_line(fv, "Ne ", sf, x, y, z, " # SAV (Si antisite vac)");
        else if (IAt == 5) atom_line(fv, "Be ", sf, x, y, z, " # CAV (C antisite vac)");
        else if (IAt == 6) atom_line(fv, "Ar ", sf, x, y, z, o->la_tree ? " # surrounded by 2C-2Si" : " # XV (near 2C and 2Si)");
        else orc_fatal(o, "ERROR: Weird Defect type was formed...ABORTING");
      } else atom_line(fv, "O ", sf, x, y, z, "  ");
    }
  }
  if (sic) {                                              /* :267-398, including the restore quirk at :386 */
    int NumIsWrong = 0, NumAtWrong = 0, cap = 1024;
    int32_t (*buf)[4] = (int32_t (*)[4])malloc((size_t)cap * sizeof *buf);
#define NCORR(w, x, y, z, out) do { int ia_ = IBITS(w, PSIC(o), LSIC(o)); const int32_t (*nn_)[3] = bulk_nn(o, o->LattInd[CI(o, x, y, z)]); out = 0; \
      if (ia_ == 1) for (int i_ = 0; i_ < 4; i_++) out += IBITS(LC(o, nn_[i_]), PC_(o), 1); \
      else if (ia_ == 2) for (int i_ = 0; i_ < 4; i_++) out += IBITS(LC(o, nn_[i_]), PSI(o), 1); } while (0)
    for (int z = 1; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      int32_t *w = &o->LattCoo[CI(o, x, y, z)];
      if (*w == 0 || IBITS(*w, ORC_PosIndex, ORC_LenIndex) == 113) continue;
      if (IBITS(*w, ORC_PosOcc, ORC_LenOcc) == 1 && IBITS(*w, ORC_PosCoor, ORC_LenCoor) == 4) {
        int IAt = IBITS(*w, PSIC(o), LSIC(o)), n; NCORR(*w, x, y, z, n);
        if (n == 0) {
          if (NumIsWrong == cap) { cap *= 2; buf = (int32_t (*)[4])realloc(buf, (size_t)cap * sizeof *buf); }
          buf[NumIsWrong][0] = x; buf[NumIsWrong][1] = y; buf[NumIsWrong][2] = z;
          if (IAt == 1) { *w &= ~(1 << PSI(o)); *w |= 1 << PC_(o); buf[NumIsWrong][3] = 1; }
          else { *w &= ~(1 << PC_(o)); *w |= 1 << PSI(o); buf[NumIsWrong][3] = 2; }
          NumIsWrong++;
        }
      }
    }
    for (int z = 1; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      int32_t w = o->LattCoo[CI(o, x, y, z)];
      if (w == 0) continue;
      if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 1 && IBITS(w, ORC_PosCoor, ORC_LenCoor) == 4) { int n; NCORR(w, x, y, z, n); if (n == 3) NumAtWrong++; }
    }
    LOG(o, " NumAtWrong:  %d\n", NumAtWrong);
    if (fw) {
      fprintf(fw, " %8d angstroem\n", NumAtWrong);
      fprintf(fw, " periodic %s %s %s\n", e1, e2, e3);
      for (int z = 1; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        int32_t w = o->LattCoo[CI(o, x, y, z)];
        if (w == 0) continue;
        if (IBITS(w, ORC_PosOcc, ORC_LenOcc) == 1 && IBITS(w, ORC_PosCoor, ORC_LenCoor) == 4) {
          int n; NCORR(w, x, y, z, n);
          if (n == 3) { int IAt = IBITS(w, PSIC(o), LSIC(o)); if (IAt == 1) atom_line(fw, "Si ", sf, x, y, z, "  "); if (IAt == 2) atom_line(fw, "C ", sf, x, y, z, "  "); }
        }
      }
    }
    for (int i = 0; i < NumIsWrong; i++) {
      int32_t *w = &o->LattCoo[CI(o, buf[i][0], buf[i][1], buf[i][2])];
      int ty = buf[NumIsWrong - 1][3];                    /* :386 uses the LAST entry's type for all */
      if (ty == 1) { *w &= ~(1 << PC_(o)); *w |= 1 << PSI(o); }
      else { *w &= ~(1 << PSI(o)); *w |= 1 << PC_(o); }
    }
    free(buf);
#undef NCORR
  }
  if (fx) fclose(fx);
  if (fd) fclose(fd);
  if (fv) fclose(fv);
  if (fw) fclose(fw);
  if (o->write_files) {                                   /* WriteMolMolSource.f:77-112, side files */
    snprintf(path, sizeof path, "%s/fort.30", o->outdir); FILE *f = fopen(path, "a"); if (f) { fprintf(f, " %.17g\n", Time); fclose(f); }
    snprintf(path, sizeof path, "%s/%s.src", o->outdir, base); f = fopen(path, "w");
    if (f) {
      fprintf(f, " load xyz %s.xyz\n color temperature\n set background white\n set boundbox true\n zoom 140\n rotate x 41\n rotate y 4\n spacefill 2\n color axes black\n wireframe 2\n color atom green\n wireframe 12\n color atom purple\n set bondmode and\n", base);
      fclose(f);
    }
    snprintf(path, sizeof path, "%s/movie.mos", o->outdir); f = fopen(path, Iter == 0 ? "w" : "a");
    if (f) { fprintf(f, " source    %s.src\n write gif %s.gif\n zap\n", base, base); fclose(f); }
    snprintf(path, sizeof path, "%s/DefAtEnergy.dat", o->outdir); f = fopen(path, Iter == 0 ? "w" : "a");
    if (f) { fprintf(f, " \n"); fclose(f); }
  }
  return 0;
}

/* ------------------------------------------------------------------ accessors for the test harness */
int32_t *orc_lattcoo(Orc *o) { return o->LattCoo; }
void orc_get_sizes(Orc *o, int64_t out[8]) {
  out[0] = o->NumAdAtom; out[1] = o->NumAtoms; out[2] = o->NumVoids; out[3] = o->Iter; out[4] = o->IterOut; out[5] = (int64_t)o->ncell; out[6] = o->Levels; out[7] = o->done;
}
double orc_get_time(Orc *o) { return o->Time; }
double orc_get_tree_root(Orc *o) { return o->Tree[1]; }
void orc_get_last_site(Orc *o, int32_t s[3]) { s[0] = o->last_site[0]; s[1] = o->last_site[1]; s[2] = o->last_site[2]; }
void orc_export_lists(Orc *o, int32_t *mc_xyz, int32_t *mc_nn, int32_t *mc_ev, double *mc_rate,
                      int32_t *bulk_xyz, int32_t *bulk_nn, int32_t *void_xyz, int32_t *void_nn) {
  for (int i = 1; i <= o->NumAdAtom; i++) {
    if (mc_xyz) memcpy(mc_xyz + 3 * (i - 1), o->ListAdAtom[i].xyz, 12);
    if (mc_nn) memcpy(mc_nn + 12 * (i - 1), o->ListAdAtom[i].nn, 48);
    if (mc_ev) mc_ev[i - 1] = o->ListAdAtom[i].ev;
    if (mc_rate) mc_rate[i - 1] = o->Tree[i + (1 << (o->Levels - 1)) - 1];
  }
  for (int i = 1; i <= o->NumAtoms; i++) {
    if (bulk_xyz) memcpy(bulk_xyz + 3 * (i - 1), o->ListAtom[i].xyz, 12);
    if (bulk_nn) memcpy(bulk_nn + 12 * (i - 1), o->ListAtom[i].nn, 48);
  }
  for (int i = 1; i <= o->NumVoids; i++) {
    if (void_xyz) memcpy(void_xyz + 3 * (i - 1), o->ListVoid[i].xyz, 12);
    if (void_nn) memcpy(void_nn + 12 * (i - 1), o->ListVoid[i].nn, 48);
  }
}
void orc_get_tables(Orc *o, double *PE, double *PD, double *PAbs, double *PDes, double *scal /*[8]: PtransZig,sf,Tm,P0,sigma,Tflex,MaxTime,OutTime*/) {
  memcpy(PE, o->PtransE, sizeof o->PtransE); memcpy(PD, o->PtransD, sizeof o->PtransD);
  memcpy(PAbs, o->PtransAbs, sizeof o->PtransAbs); memcpy(PDes, o->PtransDes, sizeof o->PtransDes);
  scal[0] = o->PtransZig; scal[1] = o->sf; scal[2] = o->Tm; scal[3] = o->P0; scal[4] = o->sigma; scal[5] = o->Tflex; scal[6] = o->MaxTime; scal[7] = o->OutTime;
}
/* 'ST' mode: sizes {ntemp, fixedT, INT(Ttop)} and the tabulated blocks (NULL pointers are skipped) */
void orc_get_field(Orc *o, int32_t info[4], int32_t *reftemp, double *refPE, double *refPD, double *refPAbs, double *refPDes) {
  info[0] = o->ntemp; info[1] = o->fixedT; info[2] = (int)o->Ttop; info[3] = 0;
  if (reftemp) memcpy(reftemp, o->reftempvalues, (size_t)o->ntemp * sizeof(int32_t));
  if (refPE) memcpy(refPE, o->refPE, (size_t)o->ntemp * 12288 * sizeof(double));
  if (refPD) memcpy(refPD, o->refPD, (size_t)o->ntemp * 9 * sizeof(double));
  if (refPAbs) memcpy(refPAbs, o->refPAbs, (size_t)o->ntemp * 12 * sizeof(double));
  if (refPDes) memcpy(refPDes, o->refPDes, (size_t)o->ntemp * 12 * sizeof(double));
}
/* LaserAnnealing-fork state: ints {la_tree, NumAdAtomOcc, NumOcc0, NumOcc1, InitCountCrystal[3], CountLiquid[3]}, doubles {Tm[3], P0[3], sigma[3], Tflex[3]} */
void orc_get_fork(Orc *o, int32_t iv[10], double dv[12]) {
  iv[0] = o->la_tree; iv[1] = o->NumAdAtomOcc; iv[2] = o->NumOcc0; iv[3] = o->NumOcc1;
  for (int k = 0; k < 3; k++) { iv[4 + k] = o->InitCountCrystal[k]; iv[7 + k] = o->CountLiquid[k]; dv[k] = o->Tm_k[k]; dv[3 + k] = o->P0_k[k]; dv[6 + k] = o->sigma_k[k]; dv[9 + k] = o->Tflex_k[k]; }
}
/* la/main.f:616-631 -- the second and later legs of a socket-driven run: a new temperature map is written into the
 * lattice words (la/initial/SetT.f:36-50, T[z][x+LenX*y], 10 = wall left alone), nothing is re-rated, the clock and the
 * iteration / frame counters restart and frame 0 is written again */
void orc_next_leg(Orc *o, const int32_t *T) {
  for (int z = 0; z < o->LenZ; z++) for (int y = 0; y < o->LenY; y++) for (int x = 0; x < o->LenX; x++) {
    int t = T[(size_t)z * o->LenX * o->LenY + x + (size_t)o->LenX * y];
    if (t != 10) MVBITS(t, ORC_LenT, &o->LattCoo[CI(o, x, y, z)], ORC_PosT);
  }
  o->Time = 0.0; o->Iter = 1; o->IterOut = 0; o->framecounter = 0; o->done = 0;
  LOG(o, " Atoms,Voids,AdAtoms,AdAtomsOcc %d %d %d %d\n", o->NumAtoms, o->NumVoids, o->NumAdAtom, o->NumAdAtomOcc);
  for (int i = 0; i < 3; i++) o->CountCrystalOld[i] = o->CountCrystal[i];
  orc_write_frame(o, o->Time, o->IterOut, o->framecounter);
}
/* test hook: lower the stop threshold of la/main.f:714-723 so that the rule fires while the melt is still rich */
void orc_set_numocc0(Orc *o, int v) { o->NumOcc0 = v; }
void orc_get_species(Orc *o, int32_t out[8]) {
  out[0] = o->NCrystal; out[1] = o->NCov; for (int i = 0; i < 3; i++) { out[2 + i] = o->ListCrystal[i]; out[5 + i] = o->ListCov[i]; }
}
void orc_get_run(Orc *o, int64_t out[4], char initst[4], char runtype[4], char exitstrategy[8]) {
  out[0] = o->MaxIter; out[1] = o->OutMolMol; out[2] = o->Idum; out[3] = o->counter;
  strcpy(initst, o->InitSt); strcpy(runtype, o->RunType); strcpy(exitstrategy, o->ExitStrategy);
}
void orc_get_rng(Orc *o, int32_t st[58], int64_t *counter, int32_t *idum) {
  st[0] = o->r_iff; st[1] = o->r_inext; st[2] = o->r_inextp; for (int i = 1; i <= 55; i++) st[2 + i] = o->r_ma[i];
  *counter = o->counter; *idum = o->Idum;
}
void orc_get_counts(Orc *o, int32_t cc[3], int32_t cv[3]) { for (int i = 0; i < 3; i++) { cc[i] = o->CountCrystal[i]; cv[i] = o->CountCov[i]; } }
void orc_get_quirks(Orc *o, int64_t q[8]) { q[0] = o->q_stale_nbuff; q[1] = o->q_wl_iso; q[2] = o->q_wl_cov; q[3] = o->q_wl_dep; q[4] = o->q_wl_des; q[5] = o->q_oob; q[6] = o->q_anom; q[7] = o->n_getprob; }
void orc_set_io(Orc *o, const char *outdir, const char *logpath, int write_files) {
  if (outdir) { strncpy(o->outdir, outdir, sizeof o->outdir - 1); }
  o->write_files = write_files;
  if (o->log && o->log != stdout) fclose(o->log);
  o->log = logpath ? fopen(logpath, "w") : NULL;
}
void orc_set_run(Orc *o, int64_t maxiter, int64_t outmolmol, int runtype_T, int idum, int faithful_cost) {
  if (maxiter > 0) o->MaxIter = maxiter;
  if (outmolmol > 0) o->OutMolMol = outmolmol;
  if (runtype_T >= 0) { o->RunType[0] = runtype_T ? 'T' : 'R'; o->RunType[1] = 0; }
  if (idum != 0) o->Idum = idum;
  o->faithful_cost = faithful_cost;
}
void orc_set_trace(Orc *o, int64_t cap) { free(o->trace); o->trace = cap > 0 ? (OrcTrace *)calloc((size_t)cap, sizeof(OrcTrace)) : NULL; o->trace_cap = cap; o->trace_n = 0; }
OrcTrace *orc_get_trace(Orc *o, int64_t *n) { *n = o->trace_n; return o->trace; }
const char *orc_errmsg(Orc *o) { return o->errmsg; }
