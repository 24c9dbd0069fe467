// mulskips_host.cpp -- the host program `mulskips.e` over the C ABI of libmskc.so.
//
// Drop-in for the reference executable (mulskips-source/main.f, PROGRAM KLMCSiC3C): run in a directory that
// holds start.dat, reference-style log on stdout, I########.xyz / _d / _v / _w / .src, movie.mos, fort.30,
// DefAtEnergy.dat and the optional checkpoint / Coor files in the cwd.  The reference is Fortran; no Fortran
// compiler exists in this image, so the host side is C++ and keeps main.f's structure: parse start.dat
// (main.f:83-212, mcevent/SetProbability.f), build the initial geometry with the reference's builders
// (initial/SetSi.f, SetSiC3C.f, SetSiGex.f, SetNC.f, SetNCSp.f, SetInvPOff.f, SetSiFINFET.f, SetCAD.f, SetT.f), hand LattCoo + the three particle lists to the GPU library,
// and drive the KMC loop (main.f:304-371) through mskc_run_parity -- which returns at every output frame so the
// files are written at the reference's cadence -- or, with --parallel, through the sublattice sampler.
// The loop body itself (everything under mcevent/ and lattice/) runs in CUDA; nothing here computes events.
//
// Box size: a compile-time constant in the reference (modules/defsystem.f:27, rewritten by pymulskips before
// `make`).  Here: --box X Y Z, else $MSKC_BOX, else the LenX line of <exe dir>/modules/defsystem.f if such a
// file exists (the file pymulskips edits), else 480 480 480 (the regression suite's size).
#include "../../include/mskc.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

namespace {

enum { PosIndex = 0, LenIndex = 8, PosT = 9, LenT = 13, PosOcc = 23, PosCoor = 24, LenCoor = 3 };
// species bits and the evaporation code: mulskips-source (modules/defsystem.f:90-91, main.f:335-357) by default; --la-tree
// selects the LaserAnnealing fork's (mulskips-source-LaserAnnealing/modules/defsystem.f:99-100, main.f:671-699)
int PosSi = 27, PosC = 28, PosSiC = 27, LenSiC = 3, EvEvap = 4;
const int JumpDia[8][3] = {{3, 3, 3}, {3, -3, -3}, {-3, -3, 3}, {-3, 3, -3}, {3, 3, -3}, {3, -3, 3}, {-3, -3, -3}, {-3, 3, 3}};   // defsystem.f:104-105

[[noreturn]] void die(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stdout, fmt, ap); va_end(ap); fputc('\n', stdout); fflush(stdout);
  exit(1);                                                                   // Fortran STOP after the message
}
inline int ibits(int32_t w, int pos, int len) { return (int)(((uint32_t)w >> pos) & ((1u << len) - 1u)); }
inline void mvbits(int from, int len, int32_t &to, int pos) { uint32_t m = ((1u << len) - 1u) << pos; to = (int32_t)(((uint32_t)to & ~m) | (((uint32_t)from << pos) & m)); }
inline int idxE(int sp, int i1, int i2, int i3, int i4, int i5, int i6) { return (sp - 1) + 3 * (i1 + 4 * (i2 + 4 * (i3 + 4 * (i4 + 4 * (i5 + 4 * i6))))); }

struct Site { int32_t xyz[3]; int32_t nn[4][3]; };
struct Active : Site { int32_t ev; double p; };

struct Host {
  int LX = 480, LY = 480, LZ = 480;
  // start.dat (modules/definitions.f)
  int NCrystal = 0, NCov = 0, ListCrystal[3] = {0, 0, 0}, ListCov[3] = {0, 0, 0};
  std::string InitSt, RunType, ExitStrategy, SaveFinalState, SaveCoo, restartfile, coofile;
  int Len1 = 0, Len2 = 0, Len3 = 0, Len4 = 0, Idum = 0;
  int Idum0 = 0;                                   // the seed as start.dat gives it (ran3 sets IDUM to 1 at its first call, random.f:58)
  double PtransZig = 1.0, sf = 1.0, MaxTime = 0, OutTime = 0, alloy = 0;
  // InitSt IN / LA (main.f:112-141): geometry and temperature files, LA rate law, nucleus geometry
  std::string cadfile, tempfile, Homogeneous;
  double Tm = 0, P0 = 0, sigma = 0, Tflex = 0, LenVac = 0, LenNuc = 0;
  int tmax = 0;                   // main.f:288-291: only the checkpoint branch of SetCAD.f updates it (the regression log prints 0)
  int t_lo = 0, t_hi = 0;         // range of the temperature field, for the device's per-kelvin rate tables
  bool la() const { return InitSt == "LA"; }
  // InitSt ST (main.f:109-115, defsystem.f:71-81): trench geometry, temperature ramp, tabulated rate blocks
  double angle = 0, Ttop = 0, Tbottom = 0;
  bool fixedT = false;
  std::string tablefile;
  int ntemp = 0;
  std::vector<int32_t> reftemp;
  std::vector<double> refPE, refPD, refPA, refPS;         // [ntemp][12288], [ntemp][9], [ntemp][12], [ntemp][12]
  std::vector<int> st_idx;                                // PtransE entries that are non-zero in some block
  bool st() const { return InitSt == "ST"; }
  // socket-driven run (la/main.f:75-345): geometry a, temperature b (received), phases c, Ge fraction d (sent back); all
  // INTEGER(LenZ, LenX*LenY) in Fortran order, i.e. element (z, x+LenX*y) at z + LenZ*(x+LenX*y) (la/modules/defsystem.f:88-91)
  std::vector<int32_t> sock_a, sock_b, sock_c, sock_d;
  // --la-tree: the LaserAnnealing fork (la/main.f:380-520, 592-611, 634-726; la/modules/defsystem.f:36-37,46,82-83)
  bool la_tree = false;
  double Tm_k[3] = {0, 0, 0}, P0_k[3] = {0, 0, 0}, sigma_k[3] = {0, 0, 0}, Tflex_k[3] = {0, 0, 0}, LenSiGe = 0;
  int NumAdAtomOcc = 0, NumOcc0 = 0, NumOcc1 = 0, InitCountCrystal[3] = {0, 0, 0}, CountLiquid[3] = {0, 0, 0};
  double la_PE_k(int T, int k, double e) const { return P0_k[k - 1] * exp(-e / (0.00008617330350 * T)) * (e != 0.0 ? 1.0 : 0.0); }   // la/mcevent/Get_Prob_Ini.f:76-78
  double la_PD_k(int T, int k, double d) const {                                                                                    // la/mcevent/Get_Prob_Ini.f:80-83
    const int sum = CountLiquid[0] + CountLiquid[1] + CountLiquid[2];
    return 0.5 * P0_k[k - 1] * (double)(float)CountLiquid[k - 1] / (double)sum * exp(-d / (0.00008617330350 * Tm_k[k - 1]))
           * (1.0 + erf((T - Tflex_k[k - 1]) / sigma_k[k - 1])) * (d != 0.0 ? 1.0 : 0.0);
  }
  long long MaxIter = 0, OutMolMol = 0;
  std::vector<double> PE = std::vector<double>(12288, 0.0), PD = std::vector<double>(9, 0.0), PA = std::vector<double>(12, 0.0), PS = std::vector<double>(12, 0.0);
  // state built by the initialiser (module globals of defsystem.f)
  std::vector<int32_t> L;                                  // LattCoo(0:LenX-1,0:LenY-1,0:LenZ)
  std::vector<int32_t> LI;                                 // LattInd (defsystem.f:50), 1-based list positions; kept for the builders that look records up (SetShape.f)
  std::vector<Active> mc; std::vector<Site> bulk, voids;
  std::vector<double> Tree; int Levels = 20;               // init-time rate tree: only Tree(1) is consulted (Get_Prob_Ini.f "isolated atom")
  int CountCrystal[3] = {0, 0, 0}, CountCrystalOld[3] = {0, 0, 0}, CountCov[3] = {0, 0, 0};
  // ran3 (mcevent/random.f), SAVEd state
  int r_iff = 0, r_inext = 0, r_inextp = 0, r_ma[56] = {0};

  size_t ci(int x, int y, int z) const { return (size_t)x + (size_t)LX * ((size_t)y + (size_t)LY * (size_t)z); }
  int32_t &at(const int32_t p[3]) { return L[ci(p[0], p[1], p[2])]; }

  double ran3() {                                          // random.f:22-69
    const int mbig = 1000000000, mseed = 161803398, mz = 0;
    const double fac = (double)(1.0f / 1.0e9f);
    int mj, mk;
    if (Idum < 0 || r_iff == 0) {
      r_iff = 1;
      mj = mseed - std::abs(Idum); mj %= mbig; r_ma[55] = mj; mk = 1;
      for (int i = 1; i <= 54; i++) { int ii = (21 * i) % 55; r_ma[ii] = mk; mk = mj - mk; if (mk < mz) mk += mbig; mj = r_ma[ii]; }
      for (int k = 1; k <= 4; k++) for (int i = 1; i <= 55; i++) { r_ma[i] -= r_ma[1 + (i + 30) % 55]; if (r_ma[i] < mz) r_ma[i] += mbig; }
      r_inext = 0; r_inextp = 31; Idum = 1;
    }
    if (++r_inext == 56) r_inext = 1;
    if (++r_inextp == 56) r_inextp = 1;
    mj = r_ma[r_inext] - r_ma[r_inextp];
    if (mj < mz) mj += mbig;
    r_ma[r_inext] = mj;
    return mj * fac;
  }
  void update_tree(int index, double p) {                 // UpdateTree.f:49-65
    size_t pos = (size_t)index + ((size_t)1 << (Levels - 1)) - 1;
    Tree[pos] = p;
    for (int l = 1; l < Levels; l++) { size_t nb = pos ^ 1; Tree[pos >> 1] = Tree[pos] + Tree[nb]; pos >>= 1; }
  }
  void AddMC(const int32_t s[3], const int32_t nn[4][3], int ev, double p) {    // Add.f:24-53
    if (p == 0) die(" provo ad aggiungere prob nulla %d %d %d", s[0], s[1], s[2]);
    Active a; memcpy(a.xyz, s, 12); memcpy(a.nn, nn, 48); a.ev = ev; a.p = p;
    mc.push_back(a);
    if (!LI.empty() && in_box(s)) LI[ci(s[0], s[1], s[2])] = (int32_t)mc.size();
    if ((int)mc.size() >= 500000) die(" troppe particelle cinetiche");
    update_tree((int)mc.size(), p);
  }
  void AddBulk(const int32_t s[3], const int32_t nn[4][3]) {
    Site a; memcpy(a.xyz, s, 12); memcpy(a.nn, nn, 48); bulk.push_back(a);
    if (!LI.empty() && in_box(s)) LI[ci(s[0], s[1], s[2])] = (int32_t)bulk.size();
  }
  bool in_box(const int32_t s[3]) const { return s[0] >= 0 && s[0] < LX && s[1] >= 0 && s[1] < LY && s[2] >= 0 && s[2] <= LZ; }
  int32_t word(const int32_t p[3]) const { return in_box(p) ? L[ci(p[0], p[1], p[2])] : 0; }        // (the reference reads outside its array there)
  void EraseBulk(const int32_t s[3]) {                                           // Erase.f:84-128: the last record moves into the hole
    const int ind = LI[ci(s[0], s[1], s[2])];
    if (ind == 0 || ind > (int)bulk.size()) die(" err: EraseBulk %d %d %d %d", ind, s[0], s[1], s[2]);
    if (ind != (int)bulk.size()) { bulk[ind - 1] = bulk.back(); LI[ci(bulk[ind - 1].xyz[0], bulk[ind - 1].xyz[1], bulk[ind - 1].xyz[2])] = ind; }
    bulk.pop_back();
    LI[ci(s[0], s[1], s[2])] = 0;
  }
  static bool lists(const int32_t nn[4][3], const int32_t s[3]) { for (int j = 0; j < 4; j++) if (nn[j][0] == s[0] && nn[j][1] == s[1] && nn[j][2] == s[2]) return true; return false; }
  // mcevent/Get_Prob.f:213-323 for an occupied site without coverage of its own: census of the four neighbours with the
  // reciprocity test against the neighbour's record (ListAdAtom or ListAtom by its coordination), then the table look-up
  void get_prob_occupied(int NSiC, int Coor, const int32_t s[3], const int32_t nn[4][3], int &ev, double &prob) {
    int N_Si = 0, N_C = 0, N_B = 0, N_wall = 0, cov[3] = {0, 0, 0};
    for (int i = 0; i < 4; i++) {
      const int32_t w = word(nn[i]);
      const int OccN = ibits(w, PosOcc, 1), CovN = ibits(w, PosIndex, LenIndex), CoorNN = ibits(w, PosCoor, LenCoor);
      if (OccN != 0) {
        if (CovN == 113) { N_wall++; continue; }
        const int ind = LI[ci(nn[i][0], nn[i][1], nn[i][2])];
        const bool in_mc = CoorNN >= 1 && CoorNN <= 3;
        if (ind < 1 || ind > (int)(in_mc ? mc.size() : bulk.size())) die(" Get_Prob: neighbour %d %d %d of %d %d %d is in no list", nn[i][0], nn[i][1], nn[i][2], s[0], s[1], s[2]);
        const int32_t (*rec)[3] = in_mc ? mc[ind - 1].nn : bulk[ind - 1].nn;
        if (lists(rec, s)) { const int k = ibits(w, PosSiC, LenSiC); N_Si += (k == 1); N_C += (k == 2); N_B += (k == 3); }
      } else if (CovN != 0) {
        const int ind = LI[ci(nn[i][0], nn[i][1], nn[i][2])];
        if (ind >= 1 && ind <= (int)mc.size() && lists(mc[ind - 1].nn, s)) for (int j = 0; j < NCov; j++) if (ListCov[j] == CovN) cov[j]++;
      }
    }
    const int ncry = N_Si + N_C + N_B + N_wall;
    if (ncry != Coor) die(" %d %d %d %d %d Get Prob Error N_Si+N_C+N_B+N_wall.NE.Coor", Coor, N_Si, N_C, N_B, N_wall);
    ev = 0; prob = 0.0;
    if (ncry > 3) return;
    ev = EvEvap;
    if (ncry == 0) prob = (double)100000.f * Tree[1];
    else prob = PE[idxE(NSiC, N_Si + N_wall, N_C, N_B, NCov >= 1 ? cov[0] & 3 : 0, NCov >= 2 ? cov[1] & 3 : 0, NCov >= 3 ? cov[2] & 3 : 0)];
  }
  void AddVoid(const int32_t s[3], const int32_t nn[4][3]) {                    // Add.f:83-101
    Site a; memcpy(a.xyz, s, 12); memcpy(a.nn, nn, 48); voids.push_back(a);
    if (!LI.empty() && in_box(s)) LI[ci(s[0], s[1], s[2])] = (int32_t)voids.size();
    if ((int)voids.size() >= 100000) die(" troppi vuoti");
  }
  // LA rate law, Get_Prob.f:143-148 / Get_Prob_Ini.f:75-79 (glibc exp/erf, as the gfortran build)
  double la_PE(int T, double e) const { return P0 * exp(-e / (0.00008617330350 * T)) * (e != 0.0 ? 1.0 : 0.0); }
  double la_PD(int T, double d) const { return 0.5 * P0 * exp(-d / (0.00008617330350 * Tm)) * (1.0 + erf((T - Tflex) / sigma)) * (d != 0.0 ? 1.0 : 0.0); }
  int site_T(const int32_t s[3]) { int t = ibits(at(s), PosT, LenT); if (t == 0) die(" Temperature is zero in get_prob_ini(). %d %d %d", s[0], s[1], s[2]); return t; }

  // deposition / adsorption draw of Get_Prob.f:160-179 (pre-drawn species; one ran3 call)
  void draw_dep_abs(int Coor, bool with_abs, int &ev, double &prob, int T = 0) {
    double PD1 = PD[0 + 3 * (Coor - 1)], PD2 = PD[1 + 3 * (Coor - 1)], PD3 = PD[2 + 3 * (Coor - 1)];
    if (T > 0 && la_tree) { PD1 = la_PD_k(T, 1, PD1); PD2 = NCrystal >= 2 ? la_PD_k(T, 2, PD2) : 0.0; PD3 = 0.0; }
    else if (T > 0) { PD1 = la_PD(T, PD1); PD2 = la_PD(T, PD2); PD3 = la_PD(T, PD3); }
    double PDep = PD1 + PD2 + PD3, PAbs = 0;
    if (with_abs) for (int k = 1; k <= 3; k++) PAbs += PA[(k - 1) + 3 * Coor];
    double Ptot = PDep + PAbs, Pcurr = ran3() * Ptot;
    if (Pcurr <= PDep) ev = Pcurr <= PD1 ? 1 : (Pcurr > (PD1 + PD2) ? 3 : 2);
    else if (NCov == 1) ev = 6;
    else {
      ev = 0;
      for (int i = 1; i <= NCov; i++) {
        double pmin = PDep, pmax = PDep;
        for (int k = 1; k <= i - 1; k++) pmin += PA[(k - 1) + 3 * Coor];
        for (int k = 1; k <= i; k++) pmax += PA[(k - 1) + 3 * Coor];
        if (Pcurr > pmin && Pcurr <= pmax) ev = 5 + i;
      }
    }
    prob = Ptot;
  }
  // mcevent/InterpProbFromField.f:22-57: the working tables := the tabulated blocks mixed at an integer temperature
  // with INTEGER distance weights (the device library builds its per-kelvin tables with the same expression)
  void interp_field(int T) {
    int iu = -1, il, best = 0;
    for (int i = 0; i < ntemp; i++) if (reftemp[i] == T) { iu = i; break; }
    if (iu >= 0) {
      for (int k : st_idx) PE[k] = refPE[(size_t)iu * 12288 + k];
      for (int i = 0; i < 9; i++) PD[i] = refPD[(size_t)iu * 9 + i];
      for (int i = 0; i < 12; i++) { PA[i] = refPA[(size_t)iu * 12 + i]; PS[i] = refPS[(size_t)iu * 12 + i]; }
      return;
    }
    for (int i = 0; i < ntemp; i++) { int d = reftemp[i] - T; if (d > 0 && (iu < 0 || d < best)) { iu = i; best = d; } }
    if (iu <= 0) die(" InterpProbFromField: temperature %d K outside the tabulated range %d..%d K", T, reftemp[0], reftemp[ntemp - 1]);
    il = iu - 1;
    const int wl = std::abs(reftemp[iu] - T), wu = std::abs(reftemp[il] - T);
    auto mix = [&](double lo, double up) { return ((double)wl * lo + (double)wu * up) / (double)(wl + wu); };
    for (int k : st_idx) PE[k] = mix(refPE[(size_t)il * 12288 + k], refPE[(size_t)iu * 12288 + k]);
    for (int i = 0; i < 9; i++) PD[i] = mix(refPD[(size_t)il * 9 + i], refPD[(size_t)iu * 9 + i]);
    for (int i = 0; i < 12; i++) { PA[i] = mix(refPA[(size_t)il * 12 + i], refPA[(size_t)iu * 12 + i]); PS[i] = mix(refPS[(size_t)il * 12 + i], refPS[(size_t)iu * 12 + i]); }
  }
  // mcevent/Get_Prob_Ini.f:23-161 (init-time rating: no reciprocity test, walls / coverage count as N_cov)
  void get_prob_ini(int Occ, int NSiC, int Coor, const int32_t s[3], const int32_t nn[4][3], int &ev, double &prob) {
    ev = 0; prob = 0.0;
    if (st() && !fixedT) interp_field(site_T(s));           // Get_Prob_Ini.f:41-56
    const int T = la() ? site_T(s) : 0;
    if (Occ == 0) {
      bool draw = false;
      if (Coor == 1) { if (ibits(at(nn[0]), PosCoor, LenCoor) >= 2) draw = true; else { ev = 1; prob = (double)1.e-6f; } }
      else if (Coor == 2 || Coor == 3) draw = true;
      else if (Coor == 4) printf(" %d Get Prob Vacancy Generation\n", Coor);
      if (draw) draw_dep_abs(Coor, false, ev, prob, T);
      return;
    }
    int N_Si = 0, N_C = 0, N_B = 0, N_cov = 0;
    for (int i = 0; i < 4; i++) {
      int32_t w = at(nn[i]);
      if (ibits(w, PosIndex, LenIndex) == 0) { int k = ibits(w, PosSiC, LenSiC); N_Si += (k == 1); N_C += (k == 2); N_B += (k == 3); }
      else N_cov++;
    }
    int n = N_Si + N_C + N_B + N_cov;
    if (n != Coor) die(" %d %d %d %d %d Get Prob Error N_Si+N_C++N_B+N_cov.NE.Coor", Coor, N_Si, N_C, N_B, N_cov);
    if (n > 3) return;
    ev = EvEvap;
    if (n == 0) { prob = (double)100000.f * Tree[1]; printf(" %d Get Prob Isolated Atom\n", Coor); }
    else { prob = PE[idxE(NSiC, N_Si + N_cov, N_C, N_B, 0, 0, 0)]; if (T > 0) prob = la_tree ? la_PE_k(T, NSiC, prob) : la_PE(T, prob); }
  }
};

// ---- start.dat ----------------------------------------------------------------------------------------------------
struct Lines {
  FILE *f; char buf[4096];
  const char *next() {
    if (!fgets(buf, sizeof buf, f)) die(" start.dat: unexpected end of file");
    size_t l = strlen(buf); while (l && (buf[l - 1] == '\n' || buf[l - 1] == '\r')) buf[--l] = 0;
    return buf;
  }
};
std::string first_token(const char *s) {
  while (*s == ' ' || *s == '\t') s++;
  std::string t;
  if (*s == '\'' || *s == '"') { char q = *s++; while (*s && *s != q) t += *s++; }
  else while (*s && *s != ' ' && *s != '\t' && *s != ',' && *s != '/') t += *s++;
  return t;
}
std::string whole_line(const char *s) { std::string t(s); while (!t.empty() && (t.back() == ' ' || t.back() == '\t')) t.pop_back(); return t; }
double fnum(const char *s, char **end = nullptr) {          // list-directed REAL: accepts 1.0d0
  char tmp[128]; size_t i = 0;
  while (*s == ' ' || *s == '\t' || *s == ',') s++;
  const char *s0 = s;
  while (*s && *s != ' ' && *s != '\t' && *s != ',' && *s != '!' && i + 1 < sizeof tmp) { char c = *s++; tmp[i++] = (c == 'd' || c == 'D') ? 'e' : c; }
  tmp[i] = 0;
  char *e2; double v = strtod(tmp, &e2);
  if (end) *end = const_cast<char *>(s0 + (e2 - tmp));
  return v;
}

void set_probability(Host &H, Lines &in) {                   // mcevent/SetProbability.f:21-143
  static const int tab[3][4] = {{3, 9, 19, 34}, {18, 50, 100, 172}, {57, 150, 288, 480}};
  const int nE = tab[H.NCrystal - 1][H.NCov], nd = H.NCrystal + H.NCov;
  printf(" maxElines %d\n", nE);
  for (int i = 0; i < nE; i++) {
    char *p = const_cast<char *>(in.next());
    int sp = (int)strtol(p, &p, 10); long el = strtol(p, &p, 10); double v = fnum(p);
    if (sp > H.NCrystal) die(" ERROR: 1st index in evap prob must be <= NCrystal");
    int ind[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 1; j <= nd; j++) {
      long d = 1; for (int k = 0; k < nd - j; k++) d *= 10;
      int digit = (int)((el / d) % 10);
      if (digit > 3) die(" ERROR: evaporation index digit > 3");
      if (j <= H.NCrystal) ind[j - 1] = digit; else ind[3 + j - H.NCrystal - 1] = digit;
    }
    H.PE[idxE(sp, ind[0], ind[1], ind[2], ind[3], ind[4], ind[5])] = v;
  }
  for (int j = 0; j < 3 * H.NCrystal; j++) {
    char *p = const_cast<char *>(in.next());
    int sp = (int)strtol(p, &p, 10), el = (int)strtol(p, &p, 10); double v = fnum(p);
    if (sp > H.NCrystal || el < 1 || el > 3) die(" ERROR: 1st index in depo prob block must be <= NCrystal");
    H.PD[(sp - 1) + 3 * (el - 1)] = v;
  }
  for (int pass = 0; pass < 2; pass++)
    for (int j = 0; j < 3 * H.NCov; j++) {
      char *p = const_cast<char *>(in.next());
      int sp = (int)strtol(p, &p, 10), el = (int)strtol(p, &p, 10); double v = fnum(p);
      if (sp <= H.NCrystal || sp > H.NCrystal + H.NCov || el < 0 || el > 3) die(" ERROR: 1st index in abs/des prob block");
      (pass == 0 ? H.PA : H.PS)[(sp - H.NCrystal - 1) + 3 * el] = v;
    }
}

void set_probability_from_field(Host &H, Lines &in) {        // mcevent/SetProbabilityFromField.f:21-171
  static const int tab[3][4] = {{3, 9, 19, 34}, {18, 50, 100, 172}, {57, 150, 288, 480}};
  H.tablefile = whole_line(in.next());
  printf(" Reading input probability table: %s\n", H.tablefile.c_str());
  Lines t; t.f = fopen(H.tablefile.c_str(), "r");
  if (!t.f) die(" cannot open %s", H.tablefile.c_str());
  H.ntemp = (int)strtol(t.next(), nullptr, 10);
  (void)strtol(t.next(), nullptr, 10);                       // nprob: read and never used by the reference
  if (H.ntemp < 1 || H.ntemp > 300) die(" probability table: ntemp must be 1..300 (ntempMax, defsystem.f:72)");
  const int nE = tab[H.NCrystal - 1][H.NCov], nd = H.NCrystal + H.NCov;
  H.reftemp.assign(H.ntemp, 0);
  H.refPE.assign((size_t)H.ntemp * 12288, 0.0); H.refPD.assign((size_t)H.ntemp * 9, 0.0);
  H.refPA.assign((size_t)H.ntemp * 12, 0.0); H.refPS.assign((size_t)H.ntemp * 12, 0.0);
  for (int tb = 0; tb < H.ntemp; tb++) {
    H.reftemp[tb] = (int)strtol(t.next(), nullptr, 10);
    if (tb == 0) printf(" maxElines %d\n", nE);
    for (int i = 0; i < nE; i++) {
      char *p = const_cast<char *>(t.next());
      int sp = (int)strtol(p, &p, 10); long el = strtol(p, &p, 10); double v = fnum(p);
      if (sp > H.NCrystal) die(" ERROR: 1st index in PtransE must be <= NCrystal");
      int ind[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 1; j <= nd; j++) {
        long d = 1; for (int k = 0; k < nd - j; k++) d *= 10;
        int digit = (int)((el / d) % 10);
        if (digit > 3) die(" ERROR: evaporation index digit > 3");
        if (j <= H.NCrystal) ind[j - 1] = digit; else ind[3 + j - H.NCrystal - 1] = digit;
      }
      H.refPE[(size_t)tb * 12288 + idxE(sp, ind[0], ind[1], ind[2], ind[3], ind[4], ind[5])] = v;
    }
    for (int j = 0; j < 3 * H.NCrystal; j++) {
      char *p = const_cast<char *>(t.next());
      int sp = (int)strtol(p, &p, 10), el = (int)strtol(p, &p, 10); double v = fnum(p);
      if (sp > H.NCrystal || el < 1 || el > 3) die(" ERROR: 1st index in depo prob block must be <= NCrystal");
      H.refPD[(size_t)tb * 9 + (sp - 1) + 3 * (el - 1)] = v;
    }
    for (int pass = 0; pass < 2; pass++)
      for (int j = 0; j < 3 * H.NCov; j++) {
        char *p = const_cast<char *>(t.next());
        int sp = (int)strtol(p, &p, 10), el = (int)strtol(p, &p, 10); double v = fnum(p);
        if (sp <= H.NCrystal || sp > H.NCrystal + H.NCov || el < 0 || el > 3) die(" ERROR: 1st index in abs/des prob block");
        (pass == 0 ? H.refPA : H.refPS)[(size_t)tb * 12 + (sp - H.NCrystal - 1) + 3 * el] = v;
      }
  }
  fclose(t.f);
  for (int i = 0; i < 12288; i++) for (int tb = 0; tb < H.ntemp; tb++) if (H.refPE[(size_t)tb * 12288 + i] != 0.0) { H.st_idx.push_back(i); break; }
  printf(" Done reading input probability table\n");
}

void read_start(Host &H, const char *path) {                 // main.f:83-219
  Lines in; in.f = fopen(path, "r");
  if (!in.f) die(" cannot open %s", path);
  printf(" KMC box size:  %d %d %d\n", H.LX, H.LY, H.LZ);
  { char *p = const_cast<char *>(in.next()); H.NCrystal = (int)strtol(p, &p, 10); H.NCov = (int)strtol(p, &p, 10); }
  if (H.NCrystal < 1 || H.NCrystal > 3) die(" ERR NOT IMPLEMENTED: NCrystal must be <= 3");
  if (H.NCov < 0 || H.NCov > 3) die(" ERR NOT IMPLEMENTED: NCov must be <= 3");
  { char *p = const_cast<char *>(in.next()); for (int i = 0; i < H.NCrystal; i++) H.ListCrystal[i] = (int)strtol(p, &p, 10); }
  if (H.NCov > 0) { printf("            ***** Simulating CVD *****\n"); char *p = const_cast<char *>(in.next()); for (int i = 0; i < H.NCov; i++) H.ListCov[i] = (int)strtol(p, &p, 10); }
  else printf("            ***** Simulating PVD or LA *****\n");
  printf(" Crystal species (NCrystal= %d ) --> Z:", H.NCrystal); for (int i = 0; i < H.NCrystal; i++) printf(" %d", H.ListCrystal[i]); printf("\n");
  if (H.NCov > 0) { printf(" Coverage species (NCov= %d ) --> Z:", H.NCov); for (int i = 0; i < H.NCov; i++) printf(" %d", H.ListCov[i]); printf("\n"); }
  H.InitSt = first_token(in.next()).substr(0, 2);
  if (H.InitSt == "FF") { char *p = const_cast<char *>(in.next()); H.Len1 = (int)strtol(p, &p, 10); H.Len2 = (int)strtol(p, &p, 10); H.Len3 = (int)strtol(p, &p, 10); H.Len4 = (int)strtol(p, &p, 10); }
  else if (H.InitSt == "SH") H.cadfile = whole_line(in.next());                         // main.f:116-117
  else if (H.InitSt == "ST") {                                                          // main.f:109-115
    char *p = const_cast<char *>(in.next());
    H.Len1 = (int)strtol(p, &p, 10); H.Len2 = (int)strtol(p, &p, 10); H.Len3 = (int)strtol(p, &p, 10);
    H.angle = fnum(p, &p); H.Ttop = fnum(p, &p); H.Tbottom = fnum(p, &p);
    printf(" Temperature on top in Kelvin:  %.17g\n Temperature on bottom in Kelvin:  %.17g\n", H.Ttop, H.Tbottom);
    H.fixedT = H.Ttop == H.Tbottom;
  }
  else if (H.InitSt == "IN") H.cadfile = whole_line(in.next());                         // main.f:108-111, read with '(A)'
  else if (H.InitSt == "LA") {                                                          // main.f:112-141
    printf(" ****** MulSKIPs run for laser annealing ******\n");
    if (H.NCov > 0) die(" ERR NOT IMPLEMENTED: NCov=0 for LA");
    H.cadfile = whole_line(in.next()); H.tempfile = whole_line(in.next());
    if (H.la_tree) {                                                                    // la/main.f:420-424: one value per crystal species
      double *dst[4] = {H.Tm_k, H.P0_k, H.sigma_k, H.Tflex_k};
      for (int q = 0; q < 4; q++) { char *p = const_cast<char *>(in.next()); for (int k = 0; k < H.NCrystal; k++) dst[q][k] = fnum(p, &p); }
      H.Tm = H.Tm_k[0]; H.P0 = H.P0_k[0]; H.sigma = H.sigma_k[0]; H.Tflex = H.Tflex_k[0];
      printf(" Melting temperature in Kelvin: "); for (int k = 0; k < H.NCrystal; k++) printf(" %.17g", H.Tm_k[k]); printf("\n");
    } else {
    H.Tm = fnum(in.next()); H.P0 = fnum(in.next()); H.sigma = fnum(in.next()); H.Tflex = fnum(in.next());
    printf(" Melting temperature in Kelvin:  %.17g\n", H.Tm);
    }
    H.LenVac = fnum(in.next()); printf(" Top air thickness (ang):  %.17g\n", H.LenVac);
    H.LenNuc = fnum(in.next()); printf(" Nucleus thickness (ang):  %.17g\n", H.LenNuc);
    H.Homogeneous = first_token(in.next()).substr(0, 2); printf(" Model homogeneous nucleation: %s\n", H.Homogeneous.c_str());
    if (H.la_tree && H.Homogeneous == "D") die(" mulskips.e (B200 host): Homogeneous 'D' (SetSiGeSF stacking-fault seed) is not built");
  }
  else { char *p = const_cast<char *>(in.next()); H.Len1 = (int)strtol(p, &p, 10); H.Len2 = (int)strtol(p, &p, 10); H.Len3 = (int)strtol(p, &p, 10); }
  // main.f:142-149 (the ALL(...) test with NCrystal=1 behaves as "not SiGe", SURVEY 8c trap 5)
  if ((H.NCrystal == 2 && H.ListCrystal[0] == 14 && H.ListCrystal[1] == 32) || (H.NCrystal == 3 && H.ListCrystal[0] == 14 && H.ListCrystal[1] == 32 && H.ListCrystal[2] == 5)) {
    H.alloy = fnum(in.next()); printf(" Ge fraction in substrate:  %.17g\n", H.alloy);
    if (H.la_tree) { H.LenSiGe = fnum(in.next()); printf(" SiGe thickness (ang):  %.17g\n", H.LenSiGe); }        // la/main.f:441-442
  }
  H.PtransZig = fnum(in.next());
  H.sf = fnum(in.next());
  printf(" KMC Super-Lattice parameter (ang):  %.17g\n", H.sf);
  H.ExitStrategy = first_token(in.next()).substr(0, 4);
  printf(" Exit strategy: %s\n", H.ExitStrategy.c_str());
  if (H.ExitStrategy == "Iter") { H.MaxIter = strtoll(in.next(), nullptr, 10); H.OutMolMol = strtoll(in.next(), nullptr, 10); printf(" MaxIter:  %lld\n OutMolMol:  %lld\n", H.MaxIter, H.OutMolMol); }
  else if (H.ExitStrategy == "Time") { H.MaxTime = fnum(in.next()); H.OutTime = fnum(in.next()); printf(" Total simulation time in seconds:  %.17g\n Output frequency in seconds:  %.17g\n", H.MaxTime, H.OutTime); }
  else die(" ExitStrategy must be Iter or Time");
  H.RunType = first_token(in.next()).substr(0, 2);
  if (H.RunType == "T") printf(" ATTENTION! You are running mulskips in test modality,  that is random numbers are not so random!!\n");
  H.Idum = (int)strtol(in.next(), nullptr, 10);
  H.Idum0 = H.Idum;
  printf(" IDUM %d\n", H.Idum);
  H.SaveFinalState = first_token(in.next()).substr(0, 2);
  if (H.SaveFinalState == "T") { H.restartfile = whole_line(in.next()); printf(" ATTENTION! Storing checkpoint file %s\n This will occupy a lot of disk space...\n", H.restartfile.c_str()); }
  else printf(" SaveFinalState flag is not \"T\": checkpoint file will  not be written.\n");
  H.SaveCoo = first_token(in.next()).substr(0, 2);
  if (H.SaveCoo == "T") {
    H.coofile = whole_line(in.next());
    if (H.la_tree && H.coofile == "None") printf(" Phases passed through sockets...\n");
    else printf(" ATTENTION! Storing final Coor file %s\n This might occupy a a lot of disk space...\n", H.coofile.c_str());
  }
  else {
    if (H.la()) die(" ERR: SaveCoo flag must be \"T\" for LA simulations!");
    printf(" SaveCoo flag is not \"T\": Coor file not stored\n");
  }
  if (H.st()) set_probability_from_field(H, in); else set_probability(H, in);          // main.f:215-219
  fclose(in.f);
  printf(" numParticelle= %d Levels = %d , SizeTree = %d\n", 500000, 20, (1 << 20) - 1);
}

// ---- initial geometries -------------------------------------------------------------------------------------------
bool cubic_site(int x, int y, int z, int &kind) {            // the site test of every builder, e.g. SetSi.f:52-57
  if (x % 3 || y % 3 || z % 3) return false;
  int x1 = x / 3, y1 = y / 3, z1 = z / 3;
  if ((x1 % 2) != (y1 % 2) || (x1 % 2) != (z1 % 2)) return false;
  int m = (x1 + y1 + z1) % 4;
  if (m == 0) { kind = 0; return true; }
  if (m == 3) { kind = 1; return true; }
  return false;
}
void dia_nn(const Host &H, int x, int y, int z, int kind, int32_t nn[4][3]) {
  for (int i = 0; i < 4; i++) {
    nn[i][0] = x + JumpDia[i + 4 * kind][0]; nn[i][1] = y + JumpDia[i + 4 * kind][1]; nn[i][2] = z + JumpDia[i + 4 * kind][2];
    if (nn[i][0] >= H.LX) nn[i][0] -= H.LX;
    if (nn[i][0] < 0) nn[i][0] += H.LX;
    if (nn[i][1] >= H.LY) nn[i][1] -= H.LY;
    if (nn[i][1] < 0) nn[i][1] += H.LY;
  }
}
void wall_planes(Host &H, int z0, int z1, int x0 = 0, int x1 = -1) {
  if (x1 < 0) x1 = H.LX - 1;
  for (int z = z0; z <= z1; z++) for (int y = 0; y < H.LY; y++) for (int x = x0; x <= x1; x++) { int32_t &w = H.L[H.ci(x, y, z)]; w |= 1 << PosOcc; mvbits(113, LenIndex, w, PosIndex); }
}
// initial/SetSi.f:25-221 (walls, one species), initial/SetSiC3C.f:25-199 (no walls, Si/C) and initial/SetSiGex.f:25-256
// (walls, random Si/Ge alloy, species counted at init): flat (001) slab
void set_flat(Host &H, int Len, bool walls, bool two_species, bool alloy = false) {
  int32_t s[3], nn[4][3]; int kind;
  if (alloy) printf(" Setting up SiGe alloy with fraction =  %.17g\n", H.alloy);
  // SetSiGex.f:60-67 draws libgfortran's RAND() per site; that generator is not part of the reference sources, a
  // documented LCG stands in (alloy initial configurations are statistically equivalent only)
  unsigned long long lcg = 88172645463325252ull;
  auto species = [&]() {
    if (!alloy) return (kind && two_species) ? 2 : 1;
    lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
    const int sp = (double)(lcg >> 11) * (1.0 / 9007199254740992.0) > H.alloy ? 1 : 2;
    H.CountCrystal[sp - 1]++;
    return sp;
  };
  if (walls) wall_planes(H, 0, 5);
  auto scan = [&](int z0, int z1, auto &&fn) {
    for (int z = z0; z <= z1; z++) for (int y = 0; y < H.LY; y++) for (int x = 0; x < H.LX; x++)
      if (cubic_site(x, y, z, kind)) { s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); fn(H.L[H.ci(x, y, z)]); }
  };
  scan(walls ? 6 : 0, Len - 4, [&](int32_t &w) { const int sp = species(); w |= 1 << (sp == 2 ? PosC : PosSi); w |= 1 << PosOcc; mvbits(4, LenCoor, w, PosCoor); H.AddBulk(s, nn); });
  scan(Len - 3, Len - 1, [&](int32_t &w) { const int sp = species(); w |= 1 << (sp == 2 ? PosC : PosSi); w |= 1 << PosOcc; mvbits(2, LenCoor, w, PosCoor); H.AddMC(s, nn, 4, H.PE[idxE(sp, 2, 0, 0, 0, 0, 0)]); });
  scan(Len, Len, [&](int32_t &w) { mvbits(2, LenCoor, w, PosCoor); int ev; double p; H.draw_dep_abs(2, true, ev, p); H.AddMC(s, nn, ev, p); });   // Get_Prob(0,0,0,2,..)
  if (walls) wall_planes(H, H.LZ - 6, H.LZ - 1);
}
// "Set the coordination" / "Set the probability" passes shared by the centre-of-box and file-driven builders:
// SetNC.f:60-233, SetInvPOff.f:105-284, SetSiFINFET.f:120-309 (wall aware), SetCAD.f:196-418 (wall aware, vacancies)
void pass_coordination(Host &H, int z0, int z1, int y0, int y1, int x0, int x1, bool wallaware = false) {
  int32_t nn[4][3]; int kind;
  for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
    int32_t &w = H.L[H.ci(x, y, z)];
    if (wallaware && ibits(w, PosIndex, LenIndex) == 113) continue;
    if (!cubic_site(x, y, z, kind)) continue;
    int Occ = ibits(w, PosOcc, 1), Coor = 0;
    dia_nn(H, x, y, z, kind, nn);
    for (int i = 0; i < 4; i++) {
      int32_t wn = H.at(nn[i]);
      if (ibits(wn, PosOcc, 1) == 0) continue;
      if (!wallaware || Occ > 0 || ibits(wn, PosIndex, LenIndex) != 113) Coor++;       // an empty site does not count walls (SetSiFINFET.f:150-160)
    }
    if (Occ != 0 && Coor == 0) {                            // isolated atom: removed, its neighbours forget it
      int t = ibits(w, PosT, LenT); w = 0; if (H.la()) mvbits(t, LenT, w, PosT);       // (LA keeps the temperature bits, SetCAD.f:228-232)
      for (int i = 0; i < 4; i++) { int32_t &wn = H.at(nn[i]); int c = ibits(wn, PosCoor, LenCoor) - 1; if (c < 0) c = 0; mvbits(c, LenCoor, wn, PosCoor); }
    } else mvbits(Coor, LenCoor, w, PosCoor);
  }
}
void pass_probability(Host &H, int z0, int z1, int y0, int y1, int x0, int x1, bool wallaware = false, bool cadvoid = false) {
  int32_t s[3], nn[4][3], keep[3] = {0, 0, 0}; int kind, ev; double p;
  for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
    int32_t w = H.L[H.ci(x, y, z)];
    if (wallaware && ibits(w, PosIndex, LenIndex) == 113) continue;
    if (!cubic_site(x, y, z, kind)) continue;
    s[0] = x; s[1] = y; s[2] = z;
    dia_nn(H, x, y, z, kind, nn);
    int Occ = ibits(w, PosOcc, 1), Coor = ibits(w, PosCoor, LenCoor);
    if (Occ != 0 && Coor == 0) printf(" error %d %d %d %d %d\n", x, y, z, Occ, Coor);
    if (Occ == 0) {
      if (Coor > 0) {
        if (Coor == 1) {                                     // keep the single occupied (non-wall) neighbour in slot 1 (SetNC.f:166-185)
          for (int i = 0; i < 4; i++) {
            int32_t wn = H.at(nn[i]);
            if (ibits(wn, PosOcc, 1) != 0 && !(wallaware && ibits(wn, PosIndex, LenIndex) == 113)) memcpy(keep, nn[i], 12);
            nn[i][0] = nn[i][1] = nn[i][2] = 0;
          }
          memcpy(nn[0], keep, 12);
        }
        if (cadvoid && Coor == 4) { printf(" Initializing a vacancy\n"); H.AddVoid(s, nn); }      // SetCAD.f:340-346
        else { H.get_prob_ini(0, 0, Coor, s, nn, ev, p); H.AddMC(s, nn, ev, p); }
      }
    } else if (Coor <= 3) { H.get_prob_ini(1, ibits(w, PosSiC, LenSiC), Coor, s, nn, ev, p); H.AddMC(s, nn, ev, p); H.NumAdAtomOcc++; }      // (la/initial/SetCAD.f:544,595)
    else H.AddBulk(s, nn);
  }
}
// initial/SetNC.f:25-235 (parallelepiped) and initial/SetNCSp.f (sphere) of 3C-SiC in the box centre
void set_nc(Host &H, int Len1, int Len2, int Len3, bool sphere, bool two_species = true) {      // two_species = false: SetNCSpSi.f
  int kind, LX = H.LX, LY = H.LY, LZ = H.LZ;
  if (!sphere) { printf(" %d %d\n", LZ / 2 - Len3 / 2, LZ / 2 + Len3 / 2 + 1); printf(" %d %d\n", LX / 2 - Len1 / 2, LX / 2 + Len1 / 2 + 1); }
  for (int z = LZ / 2 - Len3 / 2; z <= LZ / 2 + Len3 / 2 - 1; z++)
    for (int y = LY / 2 - Len2 / 2; y <= LY / 2 + Len2 / 2 - 1; y++)
      for (int x = LX / 2 - Len1 / 2; x <= LX / 2 + Len1 / 2 - 1; x++) {
        if (sphere && (x - LX / 2) * (x - LX / 2) + (y - LY / 2) * (y - LY / 2) + (z - LZ / 2) * (z - LZ / 2) > Len1 * Len1 / 4) continue;
        if (cubic_site(x, y, z, kind)) { int32_t &w = H.L[H.ci(x, y, z)]; w |= 1 << ((kind && two_species) ? PosC : PosSi); w |= 1 << PosOcc; }
      }
  const int z0 = LZ / 2 - Len3 / 2 - 10, z1 = LZ / 2 + Len3 / 2 + 9, y0 = LY / 2 - Len2 / 2 - 10, y1 = LY / 2 + Len2 / 2 + 9, x0 = LX / 2 - Len1 / 2 - 10, x1 = LX / 2 + Len1 / 2 + 9;
  pass_coordination(H, z0, z1, y0, y1, x0, x1);
  pass_probability(H, z0, z1, y0, y1, x0, x1);
}

void occupy(Host &H, int x, int y, int z, int kind, bool two_species) { int32_t &w = H.L[H.ci(x, y, z)]; w |= 1 << ((kind && two_species) ? PosC : PosSi); w |= 1 << PosOcc; }
// initial/SetInvPOff.f:22-286: 3C-SiC substrate with an array of inverted pyramids (one in the box centre, one on the corners)
// initial/SetInvPOff.f (only = 0: Si/C sublattices, REAL(4) slope 1.10) and its single-species twins initial/SetInvPSi.f
// (only = 1) and initial/SetInvPC.f (only = 2), whose pyramid test is all-integer with slope 1
void set_invp_off(Host &H, int Len, int Len3, int only = 0) {
  int kind, LX = H.LX, LY = H.LY; int32_t s[3], nn[4][3];
  // antiphase-domain variants (integer pyramid test): SetInvPwOnlySiorC.f (only = 3), SetInvPwOnlyCorSi.f (4), SetInvPwAPB.f (5):
  // the ideal Si/C assignment inside one domain, the swapped one outside
  auto fill = [&](int x, int y, int z, int k) {
    if (only >= 3) {
      const bool quad = (x < LX / 2 && y >= LY / 2) || (x >= LX / 2 && y < LY / 2);
      const bool ideal = only == 3 ? quad : only == 4 ? !quad : (x + y <= LY);
      if (!ideal) k = !k;
    }
    int32_t &w = H.L[H.ci(x, y, z)]; w |= 1 << (only == 2 ? PosC : (only == 1 ? PosSi : (k ? PosC : PosSi))); w |= 1 << PosOcc;
  };
  const int Lenh = LX / 2 - 1;
  printf(" %d %d %d\n", Lenh, Len, Len3);
  for (int z = 0; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind)) {
      fill(x, y, z, kind); mvbits(4, LenCoor, H.L[H.ci(x, y, z)], PosCoor);
      s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); H.AddBulk(s, nn);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (!cubic_site(x, y, z, kind)) continue;
    const float zt = 1.10f * (float)(z - Len3);              // REAL(4) arithmetic of SetInvPOff.f:78-82
    auto pyr = [&](int ax, int ay) { return only ? (std::abs(ax) + std::abs(ay) - (z - Len3)) <= Len : ((float)(std::abs(ax) + std::abs(ay)) - zt) <= (float)Len; };
    if (!(pyr(x - Lenh, y - Lenh) || pyr(x, y) || pyr(x - LX, y) || pyr(x, y - LY) || pyr(x - LX, y - LY))) fill(x, y, z, kind);
  }
  pass_coordination(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
  pass_probability(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
}
// initial/SetSiFINFET.f:25-312: Si fin between two wall blocks, wall cap above
void set_finfet(Host &H, int Len, int Lenww, int Lenwwtop, int Laperture) {
  int kind, LX = H.LX, LY = H.LY, LZ = H.LZ;
  wall_planes(H, 0, 5);
  wall_planes(H, Lenww, LZ - 7, 0, (int)(0.5f * (float)(LX - Laperture)) - 1);
  wall_planes(H, Lenww, LZ - 7, (int)(0.5f * (float)(LX + Laperture)), LX - 1);
  wall_planes(H, Lenwwtop, LZ - 7);
  wall_planes(H, LZ - 6, LZ - 1);
  for (int z = 6; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (ibits(H.L[H.ci(x, y, z)], PosIndex, LenIndex) != 113 && cubic_site(x, y, z, kind)) occupy(H, x, y, z, kind, false);
  pass_coordination(H, 6, Lenwwtop + 9, 0, LY - 1, 0, LX - 1, true);
  pass_probability(H, 6, Lenwwtop + 9, 0, LY - 1, 0, LX - 1, true);
}
// initial/SetTrench.f:22-283 (InitSt T): 3C-SiC slab up to Len3 with a rectangular trench (half-width Len, between Len2 and Len3)
void set_trench(Host &H) {
  int kind, LX = H.LX, LY = H.LY; int32_t s[3], nn[4][3];
  const int Len = H.Len1, Len2 = H.Len2, Len3 = H.Len3, Lenh = LX / 2 - 1;
  printf(" %d %d %d\n", Lenh, Len, Len3);
  for (int z = 0; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind)) {
      occupy(H, x, y, z, kind, true); mvbits(4, LenCoor, H.L[H.ci(x, y, z)], PosCoor);
      s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); H.AddBulk(s, nn);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind) && !(std::abs(x - Lenh) <= Len && z > Len2 && z < Len3)) occupy(H, x, y, z, kind, true);
  pass_coordination(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
  pass_probability(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
}
// initial/SetTrenchSi.f:22-338 (InitSt ST): temperature ramp along z in the lattice words, wall planes at both ends,
// Si bulk up to z=20, a Si slab up to Len3 with a V-shaped trench (half-width Len at the top, bottom at Len2, side-wall
// angle in degrees)
void set_trench_si(Host &H) {
  const double Pi = (double)3.1415926535897932384626433832795028841971f;      // REAL(4) literal (SetTrenchSi.f:34)
  int kind, LX = H.LX, LY = H.LY, LZ = H.LZ; int32_t s[3], nn[4][3];
  const int Len = H.Len1, Len2 = H.Len2, Len3 = H.Len3;
  const double arad = H.angle * Pi / 180;
  const int Lenh = LX / 2 - 1;
  printf(" %d %d %d\n", Lenh, Len, Len3);
  H.t_lo = 8191; H.t_hi = 0;
  for (int z = 0; z <= LZ - 1; z++) {
    const int T = z <= Len3 - 1 ? (int)(H.Tbottom + (H.Ttop - H.Tbottom) * z / (Len3 - 1)) : (int)H.Ttop;
    H.t_lo = std::min(H.t_lo, T); H.t_hi = std::max(H.t_hi, T);
    for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) mvbits(T, LenT, H.L[H.ci(x, y, z)], PosT);
  }
  if (H.fixedT) H.interp_field((int)H.Ttop);                                  // :65-67
  wall_planes(H, 0, 3);
  wall_planes(H, LZ - 4, LZ - 1);
  for (int z = 4; z <= 20; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind)) {
      occupy(H, x, y, z, kind, false);
      mvbits(4, LenCoor, H.L[H.ci(x, y, z)], PosCoor);
      s[0] = x; s[1] = y; s[2] = z;
      dia_nn(H, x, y, z, kind, nn);
      H.AddBulk(s, nn);
    }
  for (int z = 21; z <= Len3 - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (!cubic_site(x, y, z, kind)) continue;
    const bool in_trench = ((double)std::abs(x - Lenh) <= ((double)Len - (double)(Len3 - z) * cos(arad) / sin(arad))) && (z > Len2 && z < Len3);
    if (!in_trench) occupy(H, x, y, z, kind, false);
  }
  pass_coordination(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
  pass_probability(H, 21, Len3 + 9, 0, LY - 1, 0, LX - 1);
}
// lattice/FIND_NN.f:25-77: the two missing tetrahedral partners of SiteC given two of them (rotation by 120 and 240 degrees
// about the first bond; Pi is a REAL(4) literal widened, libm cos/sin/sqrt as the gfortran build)
void find_nn(const Host &H, const int32_t S1[3], const int32_t S2[3], const int32_t SC[3], int32_t out[2][3]) {
  const double Pi = (double)3.14159265358979323f;
  auto isign = [](int a, int b) { return b >= 0 ? std::abs(a) : -std::abs(a); };
  auto nint = [](double x) { return (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); };
  int a1[3] = {S1[0], S1[1], S1[2]}, a2[3] = {S2[0], S2[1], S2[2]};
  if (std::abs(SC[0] - a1[0]) > H.LX / 2) a1[0] += isign(H.LX, SC[0] - a1[0]);
  if (std::abs(SC[1] - a1[1]) > H.LY / 2) a1[1] += isign(H.LY, SC[1] - a1[1]);
  if (std::abs(SC[0] - a2[0]) > H.LX / 2) a2[0] += isign(H.LX, SC[0] - a2[0]);
  if (std::abs(SC[1] - a2[1]) > H.LY / 2) a2[1] += isign(H.LY, SC[1] - a2[1]);
  double ax[3], Rot[3][3], R[3];
  for (int k = 0; k < 3; k++) ax[k] = (double)SC[k] - (double)a1[k];
  const double snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int k = 0; k < 3; k++) ax[k] = ax[k] / snorm;
  for (int irot = 1; irot <= 2; irot++) {
    const double Cth = cos((double)(2 * irot) * Pi / 3.0), Sth = sin((double)(2 * irot) * Pi / 3.0);
    for (int i = 0; i < 3; i++) Rot[i][i] = ax[i] * ax[i] * (1 - Cth) + Cth;
    Rot[1][0] = ax[0] * ax[1] * (1 - Cth) - ax[2] * Sth; Rot[0][1] = ax[0] * ax[1] * (1 - Cth) + ax[2] * Sth;
    Rot[2][0] = ax[0] * ax[2] * (1 - Cth) + ax[1] * Sth; Rot[0][2] = ax[0] * ax[2] * (1 - Cth) - ax[1] * Sth;
    Rot[2][1] = ax[1] * ax[2] * (1 - Cth) - ax[0] * Sth; Rot[1][2] = ax[1] * ax[2] * (1 - Cth) + ax[0] * Sth;
    for (int k = 0; k < 3; k++) { R[k] = 0.0; for (int l = 0; l < 3; l++) R[k] = R[k] + Rot[l][k] * ((double)a2[l] - (double)SC[l]); }
    for (int k = 0; k < 3; k++) out[irot - 1][k] = nint(R[k]) + SC[k];
    if (out[irot - 1][0] >= H.LX) out[irot - 1][0] -= H.LX;
    if (out[irot - 1][0] < 0) out[irot - 1][0] += H.LX;
    if (out[irot - 1][1] >= H.LY) out[irot - 1][1] -= H.LY;
    if (out[irot - 1][1] < 0) out[irot - 1][1] += H.LY;
  }
}
// The three closing passes shared by SetShape.f:356-503 and SetSiCSF.f:325-470: (2) active atoms lose the bonds their partner does
// not return, (3) bulk atoms (z >= 13) with a non-reciprocal bond become active, (4) the empty sites bonded to the active atoms
// join the active list
void shape_passes(Host &H) {
  int32_t s[3], nn[4][3], fresh[2][3]; int ev; double p;
  auto recipr_coor = [&](const int32_t site[3], const int32_t bonds[4][3], int Coor) {                 // :366-385 / :413-432
    for (int j = 0; j < 4; j++) {
      const int32_t wn = H.word(bonds[j]);
      if (ibits(wn, PosOcc, 1) == 0) continue;
      const int ind = H.LI[H.ci(bonds[j][0], bonds[j][1], bonds[j][2])];
      const bool in_bulk = ibits(wn, PosCoor, LenCoor) == 4;
      if (ind < 1 || ind > (int)(in_bulk ? H.bulk.size() : H.mc.size())) die(" SetShape: occupied site %d %d %d is in no list (a wall or a site outside z = 4..LenZ-5 bonded to the shape)", bonds[j][0], bonds[j][1], bonds[j][2]);
      const int32_t (*rec)[3] = in_bulk ? H.bulk[ind - 1].nn : H.mc[ind - 1].nn;
      if (!Host::lists(rec, site)) Coor--;
    }
    return Coor;
  };
  for (size_t i = 0; i < H.mc.size(); i++) {                                                            // pass 2, :358-398
    memcpy(s, H.mc[i].xyz, 12); memcpy(nn, H.mc[i].nn, 48);
    int32_t &w = H.at(s);
    const int old = ibits(w, PosCoor, LenCoor), Coor = recipr_coor(s, nn, old);
    if (Coor == old) continue;
    mvbits(Coor, LenCoor, w, PosCoor);
    H.get_prob_occupied(ibits(w, PosSiC, LenSiC), Coor, s, nn, ev, p);
    H.mc[i].ev = ev; H.mc[i].p = p; H.update_tree((int)i + 1, p);
  }
  const size_t N0 = H.bulk.size();                                                                      // pass 3, :404-449 (loop bound fixed at entry)
  for (size_t i = 0; i < N0; i++) {
    if (i >= H.bulk.size()) break;                          // (the reference reads zeroed records here: z = 0 < 13, skipped)
    memcpy(s, H.bulk[i].xyz, 12);
    if (s[2] < 13) continue;
    memcpy(nn, H.bulk[i].nn, 48);
    int32_t &w = H.at(s);
    const int old = ibits(w, PosCoor, LenCoor), Coor = recipr_coor(s, nn, old);
    if (Coor == old) continue;
    mvbits(Coor, LenCoor, w, PosCoor);
    H.get_prob_occupied(ibits(w, PosSiC, LenSiC), Coor, s, nn, ev, p);
    H.EraseBulk(s);
    H.AddMC(s, nn, ev, p);
  }
  const size_t nold = H.mc.size();                                                                      // pass 4, :455-503
  for (size_t i = 0; i < nold; i++) {
    int32_t from[3]; memcpy(from, H.mc[i].xyz, 12);
    for (int j = 0; j < 4; j++) {
      memcpy(s, H.mc[i].nn[j], 12);
      if (!H.in_box(s)) die(" SetShape: bond of %d %d %d leaves the box", from[0], from[1], from[2]);
      int32_t &w = H.at(s);
      if (ibits(w, PosOcc, 1) != 0) continue;
      memset(nn, 0, sizeof nn);
      const int Coor = ibits(w, PosCoor, LenCoor) + 1;
      if (Coor == 1) {
        memcpy(nn[0], from, 12);
        mvbits(Coor, LenCoor, w, PosCoor);
        H.get_prob_ini(0, 0, Coor, s, nn, ev, p);
        H.AddMC(s, nn, ev, p);
      } else if (Coor == 2 || Coor == 3) {
        const int ind = H.LI[H.ci(s[0], s[1], s[2])];
        memcpy(nn, H.mc[ind - 1].nn, 48);
        if (Coor == 2) { memcpy(nn[1], from, 12); find_nn(H, nn[0], nn[1], s, fresh); memcpy(nn[2], fresh[0], 12); memcpy(nn[3], fresh[1], 12); }
        mvbits(Coor, LenCoor, w, PosCoor);
        H.get_prob_ini(0, 0, Coor, s, nn, ev, p);
        if (Coor == 2) memcpy(H.mc[ind - 1].nn, nn, 48);
        H.mc[ind - 1].ev = ev; H.mc[ind - 1].p = p; H.update_tree(ind, p);
      } else die(" Vacancy in the SF setting");
    }
  }
}
// initial/SetShape.f:25-507 (InitSt SH; main.f:281-286): an arbitrary structure from a list "sym x y z" of filled sites.
// Neighbours are found on the whole first shell (32 vectors, cubic + hexagonal bonds); three passes then restore the
// reciprocity of the bonds and list the empty sites around the under-coordinated atoms.
void set_shape(Host &H) {
  const int LX = H.LX, LY = H.LY, LZ = H.LZ;
  H.LI.assign(H.L.size(), 0);
  printf(" Reading input geometry file (SHAPE):  %s\n", H.cadfile.c_str());
  Lines in; in.f = fopen(H.cadfile.c_str(), "r"); if (!in.f) die(" cannot open %s", H.cadfile.c_str());
  const int NFilled = (int)strtol(in.next(), nullptr, 10);
  in.next();                                                 // the second line is skipped (:54)
  for (int ia = 0; ia < NFilled; ia++) {
    char sym[8]; int x, y, z;
    if (sscanf(in.next(), " %7s %d %d %d", sym, &x, &y, &z) != 4) die(" shape file: bad record %d", ia + 1);
    if (x < 0 || x >= LX || y < 0 || y >= LY || z < 0 || z > LZ) die(" shape file: site %d %d %d outside the box", x, y, z);
    int32_t &w = H.L[H.ci(x, y, z)];
    w |= 1 << PosOcc;
    if (z >= LZ - 4 || z <= 3) mvbits(113, LenIndex, w, PosIndex);
    else if (!strcmp(sym, "Si")) w |= 1 << PosSi;
    else if (!strcmp(sym, "C") || !strcmp(sym, "Ge")) w |= 1 << PosC;
  }
  fclose(in.f);
  printf(" Done reading input geometry file\n");
  int shell[32][3], n = 0;
  for (int k = -6; k <= 6; k++) for (int j = -6; j <= 6; j++) for (int i = -6; i <= 6; i++) if (i * i + j * j + k * k == 27) { shell[n][0] = i; shell[n][1] = j; shell[n][2] = k; n++; }
  int32_t s[3], q[3], nd[4][3], fresh[2][3]; int ev; double p;
  for (int z = 4; z <= LZ - 5; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {       // pass 1, :83-181
    int32_t &w = H.L[H.ci(x, y, z)];
    if (ibits(w, PosOcc, 1) == 0) continue;
    s[0] = x; s[1] = y; s[2] = z;
    memset(nd, 0, sizeof nd);
    int Coord = 0;
    for (int i = 0; i < 32; i++) {
      for (int k = 0; k < 3; k++) q[k] = s[k] + shell[i][k];
      if (q[0] >= LX) q[0] -= LX;
      if (q[0] < 0) q[0] += LX;
      if (q[1] >= LY) q[1] -= LY;
      if (q[1] < 0) q[1] += LY;
      if (ibits(H.word(q), PosOcc, 1) > 0) {
        if (Coord >= 4) die(" SetShape: site %d %d %d has more than four occupied shell sites", x, y, z);
        memcpy(nd[Coord++], q, 12);
      }
    }
    mvbits(Coord, LenCoor, w, PosCoor);
    if (Coord <= 3) {
      int32_t keep[3]; memcpy(keep, nd[2], 12);
      find_nn(H, nd[0], nd[1], s, fresh);
      memcpy(nd[2], fresh[0], 12); memcpy(nd[3], fresh[1], 12);
      if (Coord == 3 && !Host::lists(nd, keep)) printf(" Coord.eq.3 wrong check Buf %d %d %d\n", keep[0], keep[1], keep[2]);
      H.get_prob_ini(1, ibits(w, PosSiC, LenSiC), Coord, s, nd, ev, p);
      H.AddMC(s, nd, ev, p);
    } else H.AddBulk(s, nd);
  }
  shape_passes(H);
}
// initial/SetSiCSF.f:25-473 (InitSt E; main.f:267-272: Len = Len1, vertex VSF = (Len2, Len3, Len1-Len2)): a 3C-SiC (001) slab with a
// triple stacking fault -- the atoms of two {111} planes through the vertex are shifted off the cubic sub-lattice by (2,2,4) and
// (-2,-2,-4) -- then the shell neighbour search of SetShape.f for the atoms the cubic bonds do not explain, and the same closing passes
void set_sic_sf(Host &H, int Len, const int VSF[3]) {
  const int LX = H.LX, LY = H.LY;
  H.LI.assign(H.L.size(), 0);
  int kind; int32_t s[3], q[3], nn[4][3], nd[4][3], fresh[2][3]; int ev; double p;
  const int BoxSF = Len - VSF[2];
  int shell[32][3], n = 0;
  for (int k = -6; k <= 6; k++) for (int j = -6; j <= 6; j++) for (int i = -6; i <= 6; i++) if (i * i + j * j + k * k == 27) { shell[n][0] = i; shell[n][1] = j; shell[n][2] = k; n++; }
  if (VSF[0] + BoxSF > LX || VSF[1] + BoxSF > LY || BoxSF <= 0) die(" error SF dimension is not correctly set");
  if (!(cubic_site(VSF[0], VSF[1], VSF[2], kind) && kind == 0)) die(" SF Vertex position in wrong Site %d %d %d %d", VSF[0], VSF[1], VSF[2], Len);
  printf(" SF Vertex position in a Si Site %d %d %d %d\n", VSF[0], VSF[1], VSF[2], Len);
  for (int z = 0; z <= 12; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind)) {
      occupy(H, x, y, z, kind, true); mvbits(4, LenCoor, H.L[H.ci(x, y, z)], PosCoor);
      s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); H.AddBulk(s, nn);
    }
  for (int z = 13; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          // :114-168
    if (!cubic_site(x, y, z, kind)) continue;
    int p1 = -1, p2 = -1;
    if (x >= VSF[0] && y >= VSF[1] && z >= VSF[2]) { const int d = (x - VSF[0]) + (y - VSF[1]) - (z - VSF[2]); p1 = d - (kind ? 3 : 12); p2 = d - (kind ? 15 : 24); }
    if (p1 == 0) { if (x + 2 < LX && y + 2 < LY && z + 4 < Len) occupy(H, x + 2, y + 2, z + 4, kind, true); }
    else if (p2 == 0) { if (x - 2 > 0 && y - 2 > 0 && z - 4 > 0) occupy(H, x - 2, y - 2, z - 4, kind, true); }
    else occupy(H, x, y, z, kind, true);
  }
  // cubic bonds by sub-lattice (integer divisions, also for the shifted atoms); the whole first shell when fewer than two are filled
  int Coor, Coord;
  auto neighbours = [&](int x, int y, int z) {
    Coor = 0; Coord = 0;
    s[0] = x; s[1] = y; s[2] = z;
    dia_nn(H, x, y, z, ((x / 3 + y / 3 + z / 3) % 4 == 3) ? 1 : 0, nn);
    for (int i = 0; i < 4; i++) if (ibits(H.word(nn[i]), PosOcc, 1) > 0) Coor++;
    if (Coor >= 2) return;
    memset(nd, 0, sizeof nd);
    for (int i = 0; i < 32; i++) {
      for (int k = 0; k < 3; k++) q[k] = s[k] + shell[i][k];
      if (q[0] >= LX) q[0] -= LX;
      if (q[0] < 0) q[0] += LX;
      if (q[1] >= LY) q[1] -= LY;
      if (q[1] < 0) q[1] += LY;
      if (ibits(H.word(q), PosOcc, 1) > 0) {
        if (Coord >= 4) die(" SetSiCSF: site %d %d %d has more than four occupied shell sites", x, y, z);
        memcpy(nd[Coord++], q, 12);
      }
    }
  };
  auto bond_check = [&](int m) {
    for (int i = 0; i < m - 1; i++) for (int k = i + 1; k < m; k++) {
      const int d = (nd[i][0] - nd[k][0]) * (nd[i][0] - nd[k][0]) + (nd[i][1] - nd[k][1]) * (nd[i][1] - nd[k][1]) + (nd[i][2] - nd[k][2]) * (nd[i][2] - nd[k][2]);
      if (d != 72) printf(" not correct bonds direction\n %d %d %d %d %d %d\n", nd[i][0], nd[i][1], nd[i][2], nd[k][0], nd[k][1], nd[k][2]);
    }
  };
  for (int z = 13; z <= Len + 6; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          // :171-229
    int32_t &w = H.L[H.ci(x, y, z)];
    if (ibits(w, PosOcc, 1) == 0) continue;
    neighbours(x, y, z);
    if (Coord > Coor) { bond_check(Coord); if (Coord <= 1) w = 0; }
    else if (Coor <= 1) w = 0;
  }
  for (int z = 13; z <= Len + 6; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {          // :231-323
    int32_t &w = H.L[H.ci(x, y, z)];
    if (ibits(w, PosOcc, 1) == 0) continue;
    neighbours(x, y, z);
    const int IAt = ibits(w, PosSiC, LenSiC);
    if (Coord > Coor) {
      mvbits(Coord, LenCoor, w, PosCoor);
      if (Coord <= 3) {
        int32_t keep[3]; memcpy(keep, nd[2], 12);
        find_nn(H, nd[0], nd[1], s, fresh);
        memcpy(nd[2], fresh[0], 12); memcpy(nd[3], fresh[1], 12);
        if (Coord == 3 && !Host::lists(nd, keep)) printf(" Coord.eq.3 wrong check Buf %d %d %d\n", keep[0], keep[1], keep[2]);
        bond_check(4);
        H.get_prob_ini(1, IAt, Coord, s, nd, ev, p);
        H.AddMC(s, nd, ev, p);
      } else H.AddBulk(s, nd);
    } else {
      mvbits(Coor, LenCoor, w, PosCoor);
      if (Coor <= 3) { H.get_prob_ini(1, IAt, Coor, s, nn, ev, p); H.AddMC(s, nn, ev, p); }
      else H.AddBulk(s, nn);
    }
  }
  shape_passes(H);
}
// initial/SetAPBSymZimb1.f:25-497 (SUBROUTINE SetAFBSymZimb1, InitSt A): 3C-SiC (001) slab whose upper half carries two antiphase
// domains (ideal Si/C inside the four corner diamonds |x|+|y| < Lenh, swapped outside Lenh+3, an empty seam between).  Faithful to
// the routine's bookkeeping: every cubic site of z = 0..Len-4 enters the bulk list, occupied or not; coordination 4 is written only
// at z = 0; the closing probability pass (z = 21..Len) lists the 4-coordinated atoms a second time.
void set_afb_sym_zimb1(Host &H, int Len) {
  int kind, LX = H.LX, LY = H.LY; int32_t s[3], nn[4][3];
  const int Lenh = LX / 2 - 1 + 24;
  for (int z = 0; z <= Len / 2 - 4; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++)
    if (cubic_site(x, y, z, kind)) {
      occupy(H, x, y, z, kind, true);
      if (z == 0) mvbits(4, LenCoor, H.L[H.ci(x, y, z)], PosCoor);
      s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); H.AddBulk(s, nn);
    }
  for (int z = Len / 2 - 3; z <= Len - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
    if (!cubic_site(x, y, z, kind)) continue;
    const int d[4] = {std::abs(x) + std::abs(y), std::abs(x - LX) + std::abs(y), std::abs(x) + std::abs(y - LY), std::abs(x - LX) + std::abs(y - LY)};
    const bool inside = d[0] < Lenh || d[1] < Lenh || d[2] < Lenh || d[3] < Lenh;
    const bool outside = d[0] > Lenh + 3 && d[1] > Lenh + 3 && d[2] > Lenh + 3 && d[3] > Lenh + 3;
    if (inside) occupy(H, x, y, z, kind, true);
    else if (outside) occupy(H, x, y, z, !kind, true);
    if (z <= Len - 4) { s[0] = x; s[1] = y; s[2] = z; dia_nn(H, x, y, z, kind, nn); H.AddBulk(s, nn); }
  }
  pass_coordination(H, 1, Len, 0, LY - 1, 0, LX - 1);
  pass_probability(H, 21, Len, 0, LY - 1, 0, LX - 1);
}
// One list-directed READ of a plane of LenX*LenY numbers (LattGeo / TempStat / checkpoint files): starts on a fresh
// record, continues onto following records while values are missing, and drops what is left of the last record read.
void read_plane(FILE *f, std::vector<double> &buf, const char *what) {
  static char *line = nullptr; static size_t cap = 0;
  size_t got = 0;
  while (got < buf.size()) {
    if (getline(&line, &cap, f) < 0) die(" short read in %s", what);
    char *p = line, *e;
    while (got < buf.size()) { double v = strtod(p, &e); if (e == p) break; buf[got++] = v; p = e; while (*p == ',') p++; }
  }
}
// initial/SetT.f:25-50: integer kelvin of every cell into bits 9-21 (10 marks a wall cell: left alone)
// plane z of an array received through the socket (la/initial/SetCAD.f:117, la/initial/SetT.f:42)
void socket_plane(const Host &H, const std::vector<int32_t> &arr, int z, std::vector<double> &buf) {
  for (size_t i = 0; i < buf.size(); i++) buf[i] = (double)arr[(size_t)z + (size_t)H.LZ * i];
}
void set_T(Host &H) {
  const bool from_socket = H.tempfile == "None";                // la/initial/SetT.f:34-50
  FILE *f = nullptr;
  if (from_socket) { printf(" b(1,1) %d\n Setting temperatures in LattCoo from array received via sockets\n", H.sock_b.empty() ? 0 : H.sock_b[0]); if (H.sock_b.empty()) die(" no temperature array received"); }
  else { printf(" Reading input temperature file: %s\n", H.tempfile.c_str()); f = fopen(H.tempfile.c_str(), "r"); if (!f) die(" cannot open %s", H.tempfile.c_str()); }
  std::vector<double> buf((size_t)H.LX * H.LY);
  H.t_lo = 1 << 30; H.t_hi = 0;
  for (int z = 0; z < H.LZ; z++) {
    if (from_socket) socket_plane(H, H.sock_b, z, buf); else
    read_plane(f, buf, "temperature file");
    for (int y = 0; y < H.LY; y++) for (int x = 0; x < H.LX; x++) {
      int t = (int)buf[x + (size_t)H.LX * y];
      if (t != 10) { mvbits(t, LenT, H.L[H.ci(x, y, z)], PosT); if (t > 0) { H.t_lo = std::min(H.t_lo, t); H.t_hi = std::max(H.t_hi, t); } }
    }
  }
  if (f) fclose(f);
  printf(from_socket ? " Done setting temperatures in LattCoo\n" : " Done reading input temperatures file\n");
}
// initial/SetCAD.f:25-421: geometry voxelised offline (0 empty, 1 crystal, 10 wall), one z plane per record; RunType 'C'
// reads a checkpoint (the lattice words themselves) instead.
void set_cad(Host &H) {
  int kind, LX = H.LX, LY = H.LY, LZ = H.LZ; const bool la = H.la();
  std::vector<double> buf((size_t)LX * LY);
  if (H.alloy > 0) printf(" Setting up a SiGe layer with fraction = %.17g\n", H.alloy); else printf(" Setting up Silicon\n");
  const bool from_socket = H.cadfile == "None" && H.RunType != "C";      // la/initial/SetCAD.f:109
  FILE *f = from_socket ? nullptr : fopen(H.cadfile.c_str(), "r"); if (!f && !from_socket) die(" cannot open %s", H.cadfile.c_str());
  if (from_socket && H.sock_a.empty()) die(" no geometry array received");
  if (H.RunType == "C") {                                     // SetCAD.f:54-79
    printf(" Reading checkpoint file %s\n", H.cadfile.c_str());
    for (int z = 0; z < LZ; z++) {
      read_plane(f, buf, "checkpoint file");
      for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        int32_t &w = H.L[H.ci(x, y, z)]; const int t = ibits(w, PosT, LenT);
        w = (int32_t)buf[x + (size_t)LX * y];
        if (la) { mvbits(t, LenT, w, PosT); H.tmax = std::max(H.tmax, t); }
      }
    }
    printf(" Done reading checkpoint file\n");
  } else {
    wall_planes(H, 0, 3); wall_planes(H, LZ - 4, LZ - 1);
    if (from_socket) printf(" a(1,1) %d\n Setting input geometry from sockets\n", H.sock_a[0]); else
    printf(" Reading input geometry file: %s\n", H.cadfile.c_str());
    // SetCAD.f:99-102 skips the records of the four bottom wall planes.  The build that produced the regression
    // goldens did not (file record k lands on z = k+4): MSKC_CAD_NOSKIP=1 selects that legacy offset.
    if (!from_socket && !getenv("MSKC_CAD_NOSKIP")) for (int z = 0; z <= 3; z++) read_plane(f, buf, "geometry file");
    unsigned long long lcg = 88172645463325252ull;
    for (int z = 4; z <= LZ - 5; z++) {
      if (from_socket) socket_plane(H, H.sock_a, z, buf); else
      read_plane(f, buf, "geometry file");
      for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
        const int Id = (int)buf[x + (size_t)LX * y];
        int32_t &w = H.L[H.ci(x, y, z)];
        if (Id == 10) { w |= 1 << PosOcc; mvbits(113, LenIndex, w, PosIndex); continue; }
        if (Id != 1 || !cubic_site(x, y, z, kind)) continue;
        bool fill = true;
        if (la) {                                             // solid everywhere but the liquid nucleus (:118-141)
          // la/initial/SetCAD.f:140,161,185 (the branch fed by the socket array): solid sites below the air gap, before nucleation
          if (H.la_tree && z <= LZ - (int)(H.LenVac / H.sf)) H.NumOcc0++;
          if (H.la_tree && H.Homogeneous == "R") {             // la/initial/SetCAD.f:186-207: semi-cylindrical nucleus along y
            const int xc = LX / 2, zc = LZ - (int)(H.LenVac / H.sf);
            fill = (x - xc) * (x - xc) + (z - zc) * (z - zc) > (int)((H.LenNuc / H.sf) * (H.LenNuc / H.sf));
          } else
          if (H.Homogeneous == "T") fill = z <= LZ - (int)((H.LenVac + H.LenNuc) / H.sf);
          else if (H.Homogeneous == "F") {
            const int xc = LX / 2, yc = LY / 2, zc = LZ - (int)(H.LenVac / H.sf);
            fill = (x - xc) * (x - xc) + (y - yc) * (y - yc) + (z - zc) * (z - zc) > (int)((H.LenNuc / H.sf) * (H.LenNuc / H.sf));
          } else printf(" ERROR: Homogeneous needs to be either T or F\n");
        }
        if (!fill) continue;
        w |= 1 << PosOcc;
        // alloy species: the reference draws libgfortran's RAND() here (SetCAD.f:145,164,178), which is not part of the
        // reference sources -- alloy initial configurations are statistically equivalent only
        double rr = 1.0;
        // (SetCAD.f:141-145: Si below the SiGe layer.  mulskips-source never sets LenSiGe; the fork reads it from start.dat)
        if (H.alloy > 0 && !(la && z <= LZ - (int)((H.LenVac + (H.la_tree ? H.LenSiGe : 0.0)) / H.sf))) {
          lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; rr = (double)(lcg >> 11) * (1.0 / 9007199254740992.0);
        }
        w |= 1 << (rr >= H.alloy ? PosSi : PosC);
        if (H.la_tree) H.CountCrystal[rr >= H.alloy ? 0 : 1]++;                                                     // la/initial/SetCAD.f:150-155
      }
    }
    printf(from_socket ? " Done setting input geometry from sockets\n" : " Done reading input geometry file\n");
    if (H.la_tree) {                                          // la/initial/SetCAD.f:343-351
      const int sum = H.CountCrystal[0] + H.CountCrystal[1] + H.CountCrystal[2];
      for (int k = 0; k < H.NCrystal; k++) {
        if (H.ListCrystal[k] == 14) H.InitCountCrystal[k] = H.CountCrystal[k] + (int)((H.NumOcc0 - sum) * (1 - H.alloy));
        else if (H.ListCrystal[k] == 32) H.InitCountCrystal[k] = H.CountCrystal[k] + (int)((H.NumOcc0 - sum) * H.alloy);
      }
      for (int k = 0; k < 3; k++) H.CountLiquid[k] = H.InitCountCrystal[k] - H.CountCrystal[k];
    }
    pass_coordination(H, 0, LZ - 1, 0, LY - 1, 0, LX - 1, true);
  }
  if (f) fclose(f);
  pass_probability(H, 0, LZ - 1, 0, LY - 1, 0, LX - 1, true, true);
}

// ---- output (inpout/WriteMolMolSource.f, WriteMolMolXYZFile.f, GetOutputFileName.f) ---------------------------------
void e15_8(double v, char out[32]) {                         // Fortran E15.8: 0.dddddddd E+xx
  if (v == 0.0) { strcpy(out, " 0.00000000E+00"); return; }
  char t[40]; snprintf(t, sizeof t, "%.7e", fabs(v));
  char dig[16]; dig[0] = t[0]; memcpy(dig + 1, t + 2, 7); dig[8] = 0;
  int ex = atoi(strchr(t, 'e') + 1) + 1;
  char body[32];
  if (abs(ex) < 100) snprintf(body, sizeof body, "%s0.%sE%c%02d", v < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  else snprintf(body, sizeof body, "%s0.%s%c%03d", v < 0 ? "-" : "", dig, ex < 0 ? '-' : '+', abs(ex));
  snprintf(out, 32, "%15s", body);
}
void atom_line(FILE *f, const char *el, double sf, int x, int y, int z, const char *tail) { fprintf(f, "%s%10.5f %10.5f %10.5f%s\n", el, sf * x, sf * y, sf * z, tail); }

struct Run {
  mskc_ctx *h = nullptr; long long Iter = 1, IterOut = 0, framecounter = 0; double Time = 0; int NumAtoms = 0, NumVoids = 0, NumAdAtom = 0; bool sampler = false;
  // frames from the device (mskc_extract_frame_all / mskc_extract_antisites / mskc_gather_lattice): the writers below then
  // never see the lattice -- a frame costs streaming passes on the GPU and a download of the printed cells only
  bool device_frames = false, lattice_stale = false;
  bool crosscheck = false; long long crosschecked = 0;      // MSKC_FRAME_CROSSCHECK (test hook): every device-side frame pass against a whole-lattice scan of the same state
  std::vector<uint32_t> fcell, fword;                        // receive buffers, grown on demand
};
struct Rec { uint32_t cell; int32_t w; };
// the host copy of the lattice, allocated when something first needs it (a device-built sampler run may never)
inline int32_t *host_lattice(Host &H) { const size_t n = (size_t)H.LX * H.LY * (H.LZ + 1); if (H.L.size() != n) H.L.assign(n, 0); return H.L.data(); }
// every cell one of the frame writers prints or counts, vacancies included: the union of the tests of
// WriteMolMolXYZFile.f:77-260 (the device pass frame_extract_kernel<1> applies the same union to the device's words)
inline bool frame_cell(int32_t w, int x, int y, int z) {
  if (w == 0) return false;
  const int cov = ibits(w, PosIndex, LenIndex), co = ibits(w, PosCoor, LenCoor), a = ibits(w, PosSiC, LenSiC), occ = ibits(w, PosOcc, 1);
  if (cov == 1 || cov == 17) return true;
  if (a >= 1 && a <= 3 && co <= 3) return true;
  if (cov != 113 && (occ ? co <= 3 : (cov != 0 || co == 4))) return true;
  int kind;
  return occ && co == 4 && !cubic_site(x, y, z, kind);
}

#define MS(call) do { if ((call) != 0) die(" %s", mskc_last_error(R.h)); } while (0)

// (cell, word) lists from one of the device's extraction passes, in the reference's z,y,x scan order
template <class F>
void fetch_cells(Run &R, F call, std::vector<Rec> &out) {
  if (R.fcell.size() < (1u << 16)) { R.fcell.resize(1u << 16); R.fword.resize(1u << 16); }
  int64_t n = 0;
  MS(call((int64_t)R.fcell.size(), R.fcell.data(), R.fword.data(), &n));
  if ((size_t)n > R.fcell.size()) {                          // more cells than the buffers hold: once more with room
    R.fcell.resize((size_t)n + (size_t)n / 8); R.fword.resize(R.fcell.size());
    MS(call((int64_t)R.fcell.size(), R.fcell.data(), R.fword.data(), &n));
  }
  out.resize((size_t)n);
  for (size_t i = 0; i < (size_t)n; i++) { out[i].cell = R.fcell[i]; out[i].w = (int32_t)R.fword[i]; }
  std::sort(out.begin(), out.end(), [](const Rec &p, const Rec &q) { return p.cell < q.cell; });
}

void write_frame(Host &H, Run &R, double Time, long long Iter, long long Iout, bool on_device) {
  const int LX = H.LX, LY = H.LY, LZ = H.LZ; const double sf = H.sf;
  const bool sic = H.NCrystal == 2 && (H.ListCrystal[0] == 6 || H.ListCrystal[1] == 6);
  const bool dev = on_device && R.device_frames;
  const size_t plane = (size_t)LX * LY;
  std::vector<Site> voids;
  std::vector<Rec> recs;                                     // the cells the writers look at, z,y,x order
  if (dev) {
    fetch_cells(R, [&](int64_t cap, uint32_t *c, uint32_t *w, int64_t *n) { return mskc_extract_frame_all(R.h, cap, c, w, n, nullptr); }, recs);
    R.lattice_stale = true;
    if (R.crosscheck) {                                       // the same list from a host scan of the downloaded lattice
      MS(mskc_download_lattice(R.h, host_lattice(H)));
      size_t k = 0; bool same = true;
      for (int z = 0; z <= LZ && same; z++) for (int y = 0; y < LY && same; y++) for (int x = 0; x < LX; x++) {
        const int32_t w = H.L[H.ci(x, y, z)];
        if (!frame_cell(w, x, y, z)) continue;
        if (k >= recs.size() || recs[k].cell != (uint32_t)H.ci(x, y, z) || recs[k].w != w) { same = false; break; }
        k++;
      }
      if (!same || k != recs.size()) die(" frame cross-check: mskc_extract_frame_all differs from the scan of the lattice (frame %lld, entry %zu of %zu)", Iout, k, recs.size());
    }
  } else {
    if (on_device) { MS(mskc_download_lattice(R.h, host_lattice(H))); R.lattice_stale = false; }
    for (int z = 0; z <= LZ; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      const int32_t w = H.L[H.ci(x, y, z)];
      if (frame_cell(w, x, y, z)) recs.push_back(Rec{(uint32_t)H.ci(x, y, z), w});
    }
  }
  auto rz = [&](const Rec &r) { return (int)(r.cell / plane); };
  auto ry = [&](const Rec &r) { return (int)((r.cell % plane) / (size_t)LX); };
  auto rx = [&](const Rec &r) { return (int)(r.cell % (size_t)LX); };
  // lattice words of given cells: from the device when the lattice stayed there
  auto words_at = [&](const std::vector<int64_t> &cells, std::vector<int32_t> &out) {
    out.resize(cells.size());
    if (dev) { if (!cells.empty()) MS(mskc_gather_lattice(R.h, (int64_t)cells.size(), cells.data(), out.data())); }
    else for (size_t i = 0; i < cells.size(); i++) out[i] = H.L[(size_t)cells[i]];
  };
  if (on_device) {                                           // the module globals the writers read, from the device
    mskc_counters c; MS(mskc_get_counters(R.h, &c));
    for (int i = 0; i < 3; i++) { H.CountCrystal[i] = c.CountCrystal[i]; H.CountCov[i] = c.CountCov[i]; H.CountLiquid[i] = H.InitCountCrystal[i] - c.CountCrystal[i]; }
    H.NumAdAtomOcc = c.num_adatom_occ;
    if (R.sampler) {
      // the sampler keeps no reference-ordered vacancy list (mskc_download_lists reads the serial loop's): the vacancies are
      // the empty 4-coordinated sites of the lattice (Add.f:82-100 is called exactly when an empty site reaches four
      // neighbours), here in z,y,x scan order, with their bond lists from the device
      std::vector<int32_t> vx;
      for (const Rec &r : recs)
        if (ibits(r.w, PosOcc, 1) == 0 && ibits(r.w, PosCoor, LenCoor) == 4 && ibits(r.w, PosIndex, LenIndex) != 113) { vx.push_back(rx(r)); vx.push_back(ry(r)); vx.push_back(rz(r)); }
      voids.resize(vx.size() / 3);
      if (!voids.empty()) {
        std::vector<int32_t> vn(12 * voids.size());
        MS(mskc_download_bonds(R.h, (int64_t)voids.size(), vx.data(), vn.data()));
        for (size_t i = 0; i < voids.size(); i++) { memcpy(voids[i].xyz, &vx[3 * i], 12); memcpy(voids[i].nn, &vn[12 * i], 48); }
      }
    } else {
      voids.resize((size_t)std::max(c.num_voids, 0));
      if (!voids.empty()) {
        std::vector<int32_t> vx(3 * voids.size()), vn(12 * voids.size());
        MS(mskc_download_lists(R.h, nullptr, nullptr, nullptr, nullptr, vx.data(), vn.data(), nullptr, nullptr, 0, nullptr));
        for (size_t i = 0; i < voids.size(); i++) { memcpy(voids[i].xyz, &vx[3 * i], 12); memcpy(voids[i].nn, &vn[12 * i], 48); }
      }
    }
  } else voids = H.voids;
  if (H.la_tree) {                                          // la/inpout/WriteMolMolSource.f:42-67
    int s = 0, so = 0, sl = 0; for (int i = 0; i < 3; i++) { s += H.CountCrystal[i]; so += H.CountCrystalOld[i]; sl += H.CountLiquid[i]; }
    if (H.NCrystal > 1) printf(" SUM(CountCrystal) %d\n", s);
    printf(" CountCrystal %d %d %d\n xCrystal %.9g %.9g %.9g\n", H.CountCrystal[0], H.CountCrystal[1], H.CountCrystal[2],
           (float)H.CountCrystal[0] / s, (float)H.CountCrystal[1] / s, (float)H.CountCrystal[2] / s);
    printf(" CountLiquid %d %d %d\n xLiquid %.9g %.9g %.9g\n", H.CountLiquid[0], H.CountLiquid[1], H.CountLiquid[2],
           (float)H.CountLiquid[0] / sl, (float)H.CountLiquid[1] / sl, (float)H.CountLiquid[2] / sl);
    if (Iter > 1) {
      float d[3]; for (int i = 0; i < 3; i++) d[i] = (float)(H.CountCrystal[i] - H.CountCrystalOld[i]);
      printf(" diffCrystal %.9g %.9g %.9g\n", d[0], d[1], d[2]);
      printf(" Atom fraction from last check: %.17g %.17g %.17g\n", (double)d[0] / (s - so), (double)d[1] / (s - so), (double)d[2] / (s - so));
      for (int i = 0; i < 3; i++) H.CountCrystalOld[i] = H.CountCrystal[i];
    }
    if (H.NCov != 0) printf(" CountCov %d %d %d\n", H.CountCov[0], H.CountCov[1], H.CountCov[2]);
  } else
  if (Iter != 0) {                                           // WriteMolMolSource.f:42-56
    int s = 0, so = 0; for (int i = 0; i < 3; i++) { s += H.CountCrystal[i]; so += H.CountCrystalOld[i]; }
    printf(" SUM(CountCrystal) %d\n CountCrystal %d %d %d\n CountCrystalOld %d %d %d\n", s, H.CountCrystal[0], H.CountCrystal[1], H.CountCrystal[2], H.CountCrystalOld[0], H.CountCrystalOld[1], H.CountCrystalOld[2]);
    float d[3]; for (int i = 0; i < 3; i++) d[i] = (float)(H.CountCrystal[i] - H.CountCrystalOld[i]);
    printf(" diff %.9g %.9g %.9g\n", d[0], d[1], d[2]);
    printf(" xGesolid %.17g %.17g %.17g\n", (double)d[0] / (s - so), (double)d[1] / (s - so), (double)d[2] / (s - so));
    for (int i = 0; i < 3; i++) H.CountCrystalOld[i] = H.CountCrystal[i];
    if (H.NCov != 0) printf(" CountCov %d %d %d\n", H.CountCov[0], H.CountCov[1], H.CountCov[2]);
  }
  if (Iout * 10 >= 1000000000) die(" GetOutputFileName:: Iter can not be represented");
  char base[64]; snprintf(base, sizeof base, "I%08lld", Iout);
  printf(" %s\n", base);
  int NumAtTotal = 0, NumCovTotal = 0, NoEpiBulk = 0, kind;
  for (const Rec &r : recs) {                                 // WriteMolMolXYZFile.f:77-92 (planes 1..LenZ-1)
    const int z = rz(r); if (z < 1 || z > LZ - 1) continue;
    const int32_t w = r.w;
    int cov = ibits(w, PosIndex, LenIndex);
    if (cov != 113) { int occ = ibits(w, PosOcc, 1), coo = ibits(w, PosCoor, LenCoor); if (occ == 1 && coo <= 3) NumAtTotal++; else if (occ == 0 && cov != 0) NumCovTotal++; }
  }
  auto is_noepi = [&](int32_t w, int x, int y, int z) { return ibits(w, PosOcc, 1) == 1 && ibits(w, PosCoor, LenCoor) == 4 && !cubic_site(x, y, z, kind); };
  for (const Rec &r : recs) { if (rz(r) <= LZ - 1 && is_noepi(r.w, rx(r), ry(r), rz(r))) NoEpiBulk++; }   // :95-119 (planes 0..LenZ-1)
  printf(" %d\n", NoEpiBulk);
  char e1[32], e2[32], e3[32], et[32], path[128];
  e15_8(sf * (double)LX, e1); e15_8(sf * (double)LY, e2); e15_8(sf * (double)LZ, e3); e15_8(Time, et);
  snprintf(path, sizeof path, "%s_d.xyz", base); FILE *fd = fopen(path, "w");
  snprintf(path, sizeof path, "%s.xyz", base); FILE *fx = fopen(path, "w");
  snprintf(path, sizeof path, "%s_v.xyz", base); FILE *fv = fopen(path, "w");
  if (!fd || !fx || !fv) die(" cannot open output files in the working directory");
  fprintf(fd, " %8d angstroem KMC-time: %s Iter: %12lld\n periodic %s %s %s\n", NoEpiBulk, et, Iter, e1, e2, e3);         // :125-159
  for (const Rec &r : recs) {
    const int x = rx(r), y = ry(r), z = rz(r); if (z > LZ - 1) continue;
    const int32_t w = r.w;
    if (is_noepi(w, x, y, z)) { int a = ibits(w, PosSiC, LenSiC); if (a == 1) atom_line(fd, "Ge ", sf, x, y, z, "  "); if (a == 2) atom_line(fd, "B ", sf, x, y, z, "  "); if (a == 3) atom_line(fd, "Cl ", sf, x, y, z, "  "); }
  }
  fprintf(fx, " %8d angstroem KMC-time: %s Iter: %12lld\n periodic %s %s %s\n", NumAtTotal + (int)voids.size() + NumCovTotal, et, Iter, e1, e2, e3);   // :163-221
  for (const Rec &r : recs) {
    const int x = rx(r), y = ry(r), z = rz(r); if (z > LZ - 1) continue;
    const int32_t w = r.w;
    int a = ibits(w, PosSiC, LenSiC), coo = ibits(w, PosCoor, LenCoor), cov = ibits(w, PosIndex, LenIndex);
    if (cov == 1) atom_line(fx, "H ", sf, x, y, z, "  ");
    if (cov == 17) atom_line(fx, "Cl ", sf, x, y, z, "  ");
    if (a == 1 && coo <= 3) atom_line(fx, "Si ", sf, x, y, z, "  ");
    if ((a == 2 || a == 3) && coo <= 3) {
      if (H.NCrystal < a) die(" ERR: IAt=%d should not happen, NCrystal=%d", a, H.NCrystal);
      int Z = H.ListCrystal[a - 1]; const char *el = Z == 6 ? "C " : Z == 32 ? "Ge " : Z == 5 ? "B " : nullptr;
      if (!el) die(" ERROR: 2nd crystal specie can only be C/Ge/B");
      atom_line(fx, el, sf, x, y, z, "  ");
    }
  }
  for (auto &v : voids) atom_line(fx, "O ", sf, v.xyz[0], v.xyz[1], v.xyz[2], "  ");
  fprintf(fv, " %8d angstroem\n periodic %s %s %s\n", (int)voids.size(), e1, e2, e3);                                     // :228-260
  const bool vac_kinds = H.NCov == 0 && H.NCrystal == 2 && H.ListCrystal[1] == 6;
  std::vector<int32_t> vw;                                    // the words of the vacancies' bonded neighbours
  if (vac_kinds && !voids.empty()) {
    std::vector<int64_t> vc(4 * voids.size());
    for (size_t i = 0; i < voids.size(); i++) for (int j = 0; j < 4; j++) vc[4 * i + j] = (int64_t)H.ci(voids[i].nn[j][0], voids[i].nn[j][1], voids[i].nn[j][2]);
    words_at(vc, vw);
    if (dev && R.crosscheck) for (size_t i = 0; i < vc.size(); i++) if (vw[i] != H.L[(size_t)vc[i]]) die(" frame cross-check: mskc_gather_lattice differs from the lattice (frame %lld)", Iout);
  }
  for (size_t i = 0; i < voids.size(); i++) {
    const Site &v = voids[i];
    if (vac_kinds) {
      int a = 0; for (int j = 0; j < 4; j++) a += ibits(vw[4 * i + j], PosSiC, LenSiC);
      const char *el = a == 4 ? "Mg " : a == 8 ? "He " : a == 7 ? "Ne " : a == 5 ? "Be " : a == 6 ? "Ar " : nullptr;
      const char *tail = a == 4 ? (H.la_tree ? " # SV (Si vacancy)" : " # CV (C vacancy)") : a == 8 ? (H.la_tree ? " # CV (C vacancy)" : " # SV (Si vacancy)")      // la/inpout/WriteMolMolXYZFile.f:66-70
                         : a == 7 ? " # SAV (Si antisite vac)" : a == 5 ? " # CAV (C antisite vac)" : (H.la_tree ? " # surrounded by 2C-2Si" : " # XV (near 2C and 2Si)");
      if (!el) die(" ERROR: Weird Defect type was formed...ABORTING");
      atom_line(fv, el, sf, v.xyz[0], v.xyz[1], v.xyz[2], tail);
    } else atom_line(fv, "O ", sf, v.xyz[0], v.xyz[1], v.xyz[2], "  ");
  }
  fclose(fd); fclose(fx); fclose(fv);
  // ---- :267-398 antisite pass (SiC).  The reference walks the 4-coordinated atoms of planes 1..LenZ-1 in scan order and flips
  // the species of every atom that has, AT THAT MOMENT, no neighbour of the other species (:330-349) -- a flip is seen by the
  // atoms after it -- then prints the atoms with three (:352-372) and restores the flipped ones with the LAST entry's type
  // (:386).  Both forms below return the printed atoms (scan order, words as they are while the flips stand) and the patch
  // that leaves the lattice as the reference's LattCoo would be.
  struct WPass { std::vector<Rec> w3; std::vector<int64_t> pc; std::vector<int32_t> pw; };
  auto restore_word = [&](int32_t w, int ty) { if (ty == 1) { w &= ~(1 << PosC); w |= 1 << PosSi; } else { w &= ~(1 << PosSi); w |= 1 << PosC; } return w; };
  auto flip_word = [&](int32_t w, std::vector<int> &wtype) {
    if (ibits(w, PosSiC, LenSiC) == 1) { w &= ~(1 << PosSi); w |= 1 << PosC; wtype.push_back(1); } else { w &= ~(1 << PosC); w |= 1 << PosSi; wtype.push_back(2); }
    return w;
  };
  // the whole-lattice form, on H.L (which it leaves in the reference's final state)
  auto antisite_host = [&](WPass &out) {
    std::vector<int32_t> bx; std::vector<size_t> bcell;
    for (int z = 1; z <= LZ - 1; z++) for (int y = 0; y < LY; y++) for (int x = 0; x < LX; x++) {
      int32_t w = H.L[H.ci(x, y, z)];
      if (w == 0 || ibits(w, PosIndex, LenIndex) == 113) continue;
      if (ibits(w, PosOcc, 1) == 1 && ibits(w, PosCoor, LenCoor) == 4) { bx.push_back(x); bx.push_back(y); bx.push_back(z); bcell.push_back(H.ci(x, y, z)); }
    }
    std::vector<int32_t> bnn(12 * bcell.size() + 12);
    if (on_device) { if (!bcell.empty()) MS(mskc_download_bonds(R.h, (int64_t)bcell.size(), bx.data(), bnn.data())); }
    else {                                        // frame 0: bonds of the bulk list the builder made (z,y,x scan order is not its order: look up by cell)
      std::vector<uint32_t> idx(H.L.size(), 0xFFFFFFFFu);
      for (size_t i = 0; i < H.bulk.size(); i++) idx[H.ci(H.bulk[i].xyz[0], H.bulk[i].xyz[1], H.bulk[i].xyz[2])] = (uint32_t)i;
      for (size_t i = 0; i < bcell.size(); i++) { uint32_t k = idx[bcell[i]]; if (k != 0xFFFFFFFFu) memcpy(&bnn[12 * i], H.bulk[k].nn, 48); else memset(&bnn[12 * i], 0, 48); }
    }
    auto ncorr = [&](size_t i) {                              // neighbours of the OTHER species
      int32_t w = H.L[bcell[i]]; int a = ibits(w, PosSiC, LenSiC), n = 0;
      const int32_t *nn = &bnn[12 * i];
      if (a == 1) for (int j = 0; j < 4; j++) n += ibits(H.L[H.ci(nn[3 * j], nn[3 * j + 1], nn[3 * j + 2])], PosC, 1);
      else if (a == 2) for (int j = 0; j < 4; j++) n += ibits(H.L[H.ci(nn[3 * j], nn[3 * j + 1], nn[3 * j + 2])], PosSi, 1);
      return n;
    };
    std::vector<size_t> wrong; std::vector<int> wtype;
    for (size_t i = 0; i < bcell.size(); i++)
      if (ncorr(i) == 0) { H.L[bcell[i]] = flip_word(H.L[bcell[i]], wtype); wrong.push_back(bcell[i]); }
    for (size_t i = 0; i < bcell.size(); i++) if (ncorr(i) == 3) out.w3.push_back(Rec{(uint32_t)bcell[i], H.L[bcell[i]]});
    for (size_t i = 0; i < wrong.size(); i++) {
      int32_t &w = H.L[wrong[i]]; w = restore_word(w, wtype.back());
      out.pc.push_back((int64_t)wrong[i]); out.pw.push_back(w);
    }
  };
  // The device form.  Only atoms without an unlike neighbour BEFORE the pass can flip (a flip turns a like neighbour into an
  // unlike one, so counts only grow), and an atom of that set flips exactly when none of its bonded neighbours flipped before
  // it: the device finds the set (target 0), the host replays the order on that set alone with the set's bond lists, the
  // flips travel as a patch, and the second scan (target 3) runs on the device again.  Returns false -- with the device
  // lattice as it was -- when the check described inside fails.
  auto antisite_device = [&](WPass &out) -> bool {
    std::vector<Rec> w0;
    std::vector<int> wtype; std::vector<size_t> pidx;
    fetch_cells(R, [&](int64_t cap, uint32_t *c, uint32_t *w, int64_t *n) { return mskc_extract_antisites(R.h, 0, cap, c, w, n, nullptr); }, w0);
    if (!w0.empty()) {
      std::vector<int32_t> wx(3 * w0.size()), wn(12 * w0.size());
      for (size_t i = 0; i < w0.size(); i++) { wx[3 * i] = rx(w0[i]); wx[3 * i + 1] = ry(w0[i]); wx[3 * i + 2] = rz(w0[i]); }
      MS(mskc_download_bonds(R.h, (int64_t)w0.size(), wx.data(), wn.data()));
      auto find = [&](size_t cell) -> long long {              // position of a cell in the (sorted) set, -1 = not in it
        auto it = std::lower_bound(w0.begin(), w0.end(), cell, [](const Rec &r, size_t c) { return (size_t)r.cell < c; });
        return (it != w0.end() && (size_t)it->cell == cell) ? (long long)(it - w0.begin()) : -1;
      };
      for (size_t i = 0; i < w0.size(); i++) {
        const int32_t w = w0[i].w; const int a = ibits(w, PosSiC, LenSiC);
        int n = 0;
        if (a == 1 || a == 2) for (int j = 0; j < 4; j++) {   // cells outside the set held no unlike neighbour before the pass and never change
          const int32_t *q = &wn[12 * i + 3 * j];
          if (q[2] < 0 || q[2] > LZ) continue;
          const long long k = find(H.ci(q[0], q[1], q[2]));
          if (k >= 0) n += ibits(w0[(size_t)k].w, a == 1 ? PosC : PosSi, 1);
        }
        if (n != 0) continue;
        w0[i].w = flip_word(w, wtype); out.pc.push_back((int64_t)w0[i].cell); out.pw.push_back(w0[i].w); pidx.push_back(i);
      }
    }
    if (!out.pc.empty()) {
      MS(mskc_patch_lattice(R.h, (int)out.pc.size(), out.pc.data(), out.pw.data()));
      // The argument needs the bond lists of 4-coordinated atoms to be reciprocal (they are wherever the event code fixed them).
      // Its visible consequence is checked: after the flips no atom is left without an unlike neighbour.  If one is, the
      // flips are undone and this frame's pass runs over the whole lattice on the host, as the serial loop's writer does.
      std::vector<Rec> left;
      fetch_cells(R, [&](int64_t cap, uint32_t *c, uint32_t *w, int64_t *n) { return mskc_extract_antisites(R.h, 0, cap, c, w, n, nullptr); }, left);
      if (!left.empty() || getenv("MSKC_TEST_W_FALLBACK")) {     // (the variable: a test hook that takes this branch)
        for (size_t i = 0; i < pidx.size(); i++) out.pw[i] = restore_word(out.pw[i], wtype[i]);
        MS(mskc_patch_lattice(R.h, (int)out.pc.size(), out.pc.data(), out.pw.data()));
        out.pc.clear(); out.pw.clear();
        return false;
      }
    }
    fetch_cells(R, [&](int64_t cap, uint32_t *c, uint32_t *w, int64_t *n) { return mskc_extract_antisites(R.h, 3, cap, c, w, n, nullptr); }, out.w3);
    for (size_t i = 0; i < pidx.size(); i++) out.pw[i] = restore_word(w0[pidx[i]].w, wtype.back());
    return true;
  };
  if (sic) {
    WPass wp, wh;
    const bool check = dev && R.crosscheck;                   // test hook: the whole-lattice pass on the snapshot taken above, beside the device's
    if (check) antisite_host(wh);
    bool done = false;
    if (dev) {
      done = antisite_device(wp);
      if (!done) { MS(mskc_download_lattice(R.h, host_lattice(H))); R.lattice_stale = false; }
    }
    if (!done) { wp = WPass(); antisite_host(wp); }
    if (check) {
      bool same = wp.w3.size() == wh.w3.size() && wp.pc == wh.pc && wp.pw == wh.pw;
      for (size_t i = 0; same && i < wp.w3.size(); i++) same = wp.w3[i].cell == wh.w3[i].cell && wp.w3[i].w == wh.w3[i].w;
      if (!same) die(" frame cross-check: the antisite pass on the device differs from the whole-lattice pass (frame %lld)", Iout);
    }
    const int NumAtWrong = (int)wp.w3.size();
    printf(" NumAtWrong:  %d\n", NumAtWrong);
    snprintf(path, sizeof path, "%s_w.xyz", base); FILE *fw = fopen(path, "w");
    if (fw) {
      fprintf(fw, " %8d angstroem\n periodic %s %s %s\n", NumAtWrong, e1, e2, e3);
      for (const Rec &r : wp.w3) { int a = ibits(r.w, PosSiC, LenSiC); if (a == 1) atom_line(fw, "Si ", sf, rx(r), ry(r), rz(r), "  "); if (a == 2) atom_line(fw, "C ", sf, rx(r), ry(r), rz(r), "  "); }
      fclose(fw);
    }
    if (on_device && !wp.pc.empty()) MS(mskc_patch_lattice(R.h, (int)wp.pc.size(), wp.pc.data(), wp.pw.data()));      // the device lattice sees what the reference's LattCoo would hold
  }
  if (dev && R.crosscheck) R.crosschecked++;
  FILE *f;                                                   // WriteMolMolSource.f:77-112 side files
  if ((f = fopen("fort.30", "a"))) { fprintf(f, " %.17g\n", Time); fclose(f); }
  snprintf(path, sizeof path, "%s.src", base);
  if ((f = fopen(path, "w"))) { fprintf(f, " load xyz %s.xyz\n color temperature\n set background white\n set boundbox true\n zoom 140\n rotate x 41\n rotate y 4\n spacefill 2\n color axes black\n wireframe 2\n color atom green\n wireframe 12\n color atom purple\n set bondmode and\n", base); fclose(f); }
  if ((f = fopen("movie.mos", Iter == 0 ? "w" : "a"))) { fprintf(f, " source    %s.src\n write gif %s.gif\n zap\n", base, base); fclose(f); }
  if ((f = fopen("DefAtEnergy.dat", Iter == 0 ? "w" : "a"))) { fprintf(f, " \n"); fclose(f); }
  fflush(stdout);
}

// ---- complete checkpoint ("persist bond lists") ------------------------------------------------------
// The reference's checkpoint (main.f:385-403) holds LattCoo only; its restart (SetCAD.f:54-79,307-322) rebuilds the particle
// lists by scanning the lattice and re-deriving every bond list from ideal positions, so list order, the bonds of
// hexagonal / stacking-fault atoms, the rates in the tree, the generator and the clock are lost: a restarted run is a
// different trajectory.  Beside that file the host writes `<name>.mskc`: lattice words, the three lists in list order with
// their bond lists, event codes and rates as raw doubles, the ran3 state / numberslist cursor and the clock, so that
// `--exact-restart <name>.mskc` continues the SAME trajectory bit for bit (tests/test_host_gpu.py).  Serial-parity mode only:
// the sampler keeps no reference-ordered lists.
struct CkptHeader {
  char magic[8]; int32_t LX, LY, LZ, nmc, nbulk, nvoid; int64_t iter, framecounter, counter; double time;
  int32_t CC[3], CV[3], CCold[3], rng[58], idum, runtype;
};
void write_exact_checkpoint(Host &H, Run &R, const std::string &path) {
  mskc_counters c; MS(mskc_get_counters(R.h, &c));
  CkptHeader h{}; memcpy(h.magic, "MSKCKPT1", 8);
  h.LX = H.LX; h.LY = H.LY; h.LZ = H.LZ; h.nmc = c.num_adatom; h.nvoid = c.num_voids; h.iter = c.iter; h.time = c.time; h.framecounter = R.framecounter;
  for (int i = 0; i < 3; i++) { h.CC[i] = c.CountCrystal[i]; h.CV[i] = c.CountCov[i]; h.CCold[i] = H.CountCrystalOld[i]; }
  MS(mskc_get_rng(R.h, h.rng, &h.counter, &h.idum)); h.runtype = H.RunType[0];
  const size_t nb_cap = (size_t)std::max(c.num_atoms, 0) + 16;
  std::vector<int32_t> mx(3 * (size_t)h.nmc + 3), mn(12 * (size_t)h.nmc + 12), me((size_t)h.nmc + 1), vx(3 * (size_t)h.nvoid + 3), vn(12 * (size_t)h.nvoid + 12), bx(3 * nb_cap), bn(12 * nb_cap);
  std::vector<double> mr((size_t)h.nmc + 1);
  int64_t nb = 0;
  MS(mskc_download_lists(R.h, mx.data(), mn.data(), me.data(), mr.data(), vx.data(), vn.data(), bx.data(), bn.data(), (int64_t)nb_cap, &nb));
  if (nb > (int64_t)nb_cap) die(" exact checkpoint: %lld bulk atoms found, %zu expected", (long long)nb, nb_cap);
  h.nbulk = (int32_t)nb;
  MS(mskc_download_lattice(R.h, host_lattice(H)));
  FILE *f = fopen(path.c_str(), "wb"); if (!f) die(" cannot write %s", path.c_str());
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(H.L.data(), 4, H.L.size(), f) == H.L.size();
  ok = ok && fwrite(mx.data(), 4, 3 * (size_t)h.nmc, f) == 3 * (size_t)h.nmc && fwrite(mn.data(), 4, 12 * (size_t)h.nmc, f) == 12 * (size_t)h.nmc;
  ok = ok && fwrite(me.data(), 4, (size_t)h.nmc, f) == (size_t)h.nmc && fwrite(mr.data(), 8, (size_t)h.nmc, f) == (size_t)h.nmc;
  ok = ok && fwrite(vx.data(), 4, 3 * (size_t)h.nvoid, f) == 3 * (size_t)h.nvoid && fwrite(vn.data(), 4, 12 * (size_t)h.nvoid, f) == 12 * (size_t)h.nvoid;
  ok = ok && fwrite(bx.data(), 4, 3 * (size_t)h.nbulk, f) == 3 * (size_t)h.nbulk && fwrite(bn.data(), 4, 12 * (size_t)h.nbulk, f) == 12 * (size_t)h.nbulk;
  if (fclose(f) != 0 || !ok) die(" error writing %s", path.c_str());
}
// fills H.L / H.mc / H.bulk / H.voids / the generator / the counters from the file; returns the clock
void read_exact_checkpoint(Host &H, Run &R, const char *path, CkptHeader &h) {
  FILE *f = fopen(path, "rb"); if (!f) die(" cannot open %s", path);
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "MSKCKPT1", 8)) die(" %s is not a complete checkpoint of this program", path);
  if (h.LX != H.LX || h.LY != H.LY || h.LZ != H.LZ) die(" %s was written for a %d x %d x %d box, this run is %d x %d x %d", path, h.LX, h.LY, h.LZ, H.LX, H.LY, H.LZ);
  auto rd = [&](void *p, size_t sz, size_t n) { if (n && fread(p, sz, n, f) != n) die(" %s is truncated", path); };
  rd(H.L.data(), 4, H.L.size());
  std::vector<int32_t> x, n, e; std::vector<double> r;
  auto sites = [&](size_t cnt, auto &out) {
    x.resize(3 * cnt); n.resize(12 * cnt); rd(x.data(), 4, 3 * cnt); rd(n.data(), 4, 12 * cnt);
    out.resize(cnt);
    for (size_t i = 0; i < cnt; i++) { memcpy(out[i].xyz, &x[3 * i], 12); memcpy(out[i].nn, &n[12 * i], 48); }
  };
  {
    const size_t cnt = (size_t)h.nmc;
    x.resize(3 * cnt); n.resize(12 * cnt); e.resize(cnt); r.resize(cnt);
    rd(x.data(), 4, 3 * cnt); rd(n.data(), 4, 12 * cnt); rd(e.data(), 4, cnt); rd(r.data(), 8, cnt);
    H.mc.resize(cnt);
    for (size_t i = 0; i < cnt; i++) { memcpy(H.mc[i].xyz, &x[3 * i], 12); memcpy(H.mc[i].nn, &n[12 * i], 48); H.mc[i].ev = e[i]; H.mc[i].p = r[i]; }
  }
  sites((size_t)h.nvoid, H.voids); sites((size_t)h.nbulk, H.bulk);
  fclose(f);
  H.r_iff = h.rng[0]; H.r_inext = h.rng[1]; H.r_inextp = h.rng[2]; for (int i = 1; i <= 55; i++) H.r_ma[i] = h.rng[2 + i];
  H.Idum = h.idum;
  for (int i = 0; i < 3; i++) { H.CountCrystal[i] = h.CC[i]; H.CountCov[i] = h.CV[i]; H.CountCrystalOld[i] = h.CCold[i]; }
  R.Iter = h.iter; R.IterOut = h.iter - 1; R.Time = h.time; R.framecounter = h.framecounter;
}

bool box_from_defsystem(const char *exe, int box[3]) {      // the line pymulskips rewrites: "INTEGER, PARAMETER :: LenX=480, LenY=480, LenZ=480"
  std::string p(exe); size_t k = p.find_last_of('/'); p = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/modules/defsystem.f";
  FILE *f = fopen(p.c_str(), "r"); if (!f) return false;
  char line[1024]; bool ok = false;
  while (fgets(line, sizeof line, f)) {
    const char *a = strstr(line, "LenX="), *b = strstr(line, "LenY="), *c = strstr(line, "LenZ=");
    if (strstr(line, "PARAMETER") && a && b && c) { box[0] = atoi(a + 5); box[1] = atoi(b + 5); box[2] = atoi(c + 5); ok = true; break; }
  }
  fclose(f); return ok;
}

// the line pymulskips rewrites in main.f when it is given a socket port ("      port = 31415", setuprun.py:524-545)
int port_from_mainf(const char *exe) {
  std::string p(exe); size_t k = p.find_last_of('/'); p = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/main.f";
  FILE *f = fopen(p.c_str(), "r"); if (!f) return -1;
  char line[1024]; int port = -1;
  while (fgets(line, sizeof line, f)) {
    if (line[0] == '!' || line[0] == 'c' || line[0] == 'C') continue;
    const char *a = strstr(line, "port =");
    if (a) { port = atoi(a + 6); break; }
  }
  fclose(f); return port > 0 ? port : -1;
}

}  // namespace

// la/sockets/sockets.c + fsockets.f: 12-byte headers and INTEGER(4) arrays over a TCP connection to the python server
struct Wire {
  int fd = -1;
  void connect_to(const char *host, int port) {
    printf("  driver - connecting to host %s\n  on port %d  using an internet socket.\n", host, port);
    char ps[16]; snprintf(ps, sizeof ps, "%d", port);
    addrinfo hints{}, *res = nullptr; hints.ai_socktype = SOCK_STREAM; hints.ai_family = AF_INET;
    if (getaddrinfo(host, ps, &hints, &res) != 0 || !res) die(" Error fetching host data. Wrong host name?");
    fd = socket(res->ai_family, res->ai_socktype, res->ai_protocol);
    if (fd < 0 || connect(fd, res->ai_addr, res->ai_addrlen) < 0) die(" Error opening INET socket: wrong port or server unreachable");
    freeaddrinfo(res);
  }
  bool rd(void *p, size_t n) { char *c = (char *)p; while (n) { ssize_t k = read(fd, c, n); if (k <= 0) return false; c += k; n -= (size_t)k; } return true; }
  void wr(const void *p, size_t n) { const char *c = (const char *)p; while (n) { ssize_t k = write(fd, c, n); if (k <= 0) die(" Error writing to socket: server has quit or connection broke"); c += k; n -= (size_t)k; } }
  bool get_header(std::string &h) { char b[13]; if (!rd(b, 12)) return false; b[12] = 0; h = b; while (!h.empty() && h.back() == ' ') h.pop_back(); return true; }
  void put_header(const char *m) { char b[13]; snprintf(b, sizeof b, "%-12s", m); wr(b, 12); }
};

int main(int argc, char **argv) {
  Host H; Run R;
  int box[3] = {480, 480, 480}; bool parallel = false; double ept = 4.0; int device = 0; const char *start = "start.dat";
  int mode = -1;                                                // -1: by RunType (below), 0: serial parity, 1: sublattice sampler
  bool have_box = false;
  const char *exact_restart = nullptr;                          // --exact-restart file.mskc: continue from the complete state (lists, bonds, rates, generator), see write_exact_checkpoint
  const char *dump_state = nullptr;                             // --dump-state file: what is about to be uploaded, as text (tests compare it with the oracle's state)
  const char *sock_host = "localhost"; int sock_port = -1;     // la/main.f:102-131: -h hostname -p port => socket-driven run
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--box") && i + 3 < argc) { box[0] = atoi(argv[i + 1]); box[1] = atoi(argv[i + 2]); box[2] = atoi(argv[i + 3]); i += 3; have_box = true; }
    else if (!strcmp(argv[i], "--parallel")) mode = 1;
    else if (!strcmp(argv[i], "--parity") || !strcmp(argv[i], "--serial")) mode = 0;
    else if (!strcmp(argv[i], "--la-tree")) { H.la_tree = true; PosSi = 28; PosC = 29; PosSiC = 28; LenSiC = 2; EvEvap = 3; }
    else if (!strcmp(argv[i], "--events-per-tile") && i + 1 < argc) ept = atof(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--start") && i + 1 < argc) start = argv[++i];
    else if (!strcmp(argv[i], "--dump-state") && i + 1 < argc) dump_state = argv[++i];
    else if (!strcmp(argv[i], "--exact-restart") && i + 1 < argc) exact_restart = argv[++i];
    else if (!strcmp(argv[i], "-h") && i + 1 < argc) sock_host = argv[++i];
    else if (!strcmp(argv[i], "-p") && i + 1 < argc) sock_port = atoi(argv[++i]);
    else { fprintf(stderr, "usage: %s [--box X Y Z] [--parity | --parallel [--events-per-tile e]] [--la-tree [-h host] -p port] [--device d] [--start file] [--exact-restart file.mskc]\n", argv[0]); return 2; }
  }
  if (!have_box) { if (const char *e = getenv("MSKC_BOX")) have_box = sscanf(e, "%d %d %d", &box[0], &box[1], &box[2]) == 3; }
  if (!have_box) box_from_defsystem(argv[0], box);
  if (sock_port < 0 && H.la_tree) sock_port = port_from_mainf(argv[0]);
  H.LX = box[0]; H.LY = box[1]; H.LZ = box[2];
  bool by_time = false;
  float dev_ms = 0; long long events = 0;
  // ---- first cycle: input, initial structure, upload (main.f:83-300; la/main.f:348-631 with LAcounter = 0) ----
  // Which loop runs when the command line does not say (pymulskips launches `mulskips.e` without arguments): RunType 'T'
  // ("not so random: use this for regression tests", py/setuprun.py:344) and 'C' ask for the reference's own event sequence
  // -> serial parity mode, bit-exact; RunType 'R' ("truly random", the production default of pymulskips) asks for a
  // statistically faithful trajectory -> the synchronous-sublattice sampler, the mode that uses the GPU.  --parity /
  // --parallel (or MSKC_MODE=parity|parallel) override; socket-driven runs and boxes that are not multiples of 48 stay serial.
  auto pick_mode = [&]() {
    if (mode < 0) { if (const char *e = getenv("MSKC_MODE")) mode = !strcmp(e, "parallel") ? 1 : (!strcmp(e, "parity") ? 0 : -1); }
    if (exact_restart) mode = 0;                              // (the same trajectory: the serial loop)
    if (mode < 0) mode = (H.RunType == "R" && sock_port < 0 && !H.la_tree && H.LX % 48 == 0 && H.LY % 48 == 0) ? 1 : 0;
    parallel = mode == 1; R.sampler = parallel;
    // frames: the sampler's come from the device's extraction passes (a production box is 10^9 cells: no lattice download, no
    // host scans); the serial loop keeps the whole-lattice writer its byte-for-byte tests were pinned with.  MSKC_DEVICE_FRAMES=1 / 0
    // forces one or the other (the two writers produce identical files: tests/test_zzz_frames_gpu.py)
    R.device_frames = parallel;
    if (const char *e = getenv("MSKC_DEVICE_FRAMES")) R.device_frames = atoi(e) != 0;
    R.crosscheck = getenv("MSKC_FRAME_CROSSCHECK") != nullptr;
  };
  auto init_run = [&]() {
  read_start(H, start);
  pick_mode();
  H.Tree.assign((size_t)1 << H.Levels, 0.0);
  // Sampler mode on a flat slab (InitSt F / FS, the production PVD / CVD start): the structure is built ON the device
  // (mskc_init_flat: the geometry of initial/SetSiC3C.f / SetSi.f, checked word for word against the oracle's builders,
  // tests/test_dense_gpu.py) and frame 0 comes from the device's frame passes -- no 4 B/cell lattice, no particle lists and no
  // scan on the host, which at a 10^9-cell box is minutes of set-up and 8+ GB.  MSKC_HOST_BUILD=1 keeps the host builder.
  if (parallel && !exact_restart && !dump_state && ((H.InitSt == "F" && H.NCrystal >= 2) || H.InitSt == "FS") && H.LX % 48 == 0 && H.LY % 48 == 0 && !getenv("MSKC_HOST_BUILD")) {
    H.Len1 -= H.Len1 % 24;                                    // main.f:226,240
    if (mskc_create(H.LX, H.LY, H.LZ, H.NCrystal, H.NCov, H.ListCrystal, H.ListCov, device, 0, &R.h) != 0) die(" %s", mskc_last_error(nullptr));
    MS(mskc_set_rates(R.h, H.PE.data(), H.PD.data(), H.PA.data(), H.PS.data(), H.PtransZig));
    MS(mskc_init_flat(R.h, H.Len1, H.InitSt == "FS" ? 1 : 0, H.InitSt == "F" ? 1 : 0));
    mskc_counters c; MS(mskc_get_counters(R.h, &c));
    int64_t nact = 0; MS(mskc_download_active(R.h, 0, nullptr, nullptr, nullptr, &nact));
    R.NumAtoms = c.num_atoms; R.NumVoids = c.num_voids; R.NumAdAtom = (int)nact;
    printf(" Atoms,Voids,AdAtoms %d %d %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom);
    for (int i = 0; i < 3; i++) H.CountCrystalOld[i] = H.CountCrystal[i];
    write_frame(H, R, 0.0, 0, 0, true);                       // main.f:297-300
    by_time = H.ExitStrategy == "Time";
    return;
  }
  H.L.assign((size_t)H.LX * H.LY * (H.LZ + 1), 0);
  // main.f:225-295
  if (H.la_tree && H.InitSt != "LA" && H.InitSt != "IN") die(" mulskips.e (B200 host): --la-tree (the LaserAnnealing fork) is built for InitSt LA and IN");
  CkptHeader ck{};
  if (exact_restart) {                                        // the complete state of an earlier run instead of a builder
    if (H.la_tree || sock_port >= 0) die(" mulskips.e (B200 host): --exact-restart is built for the mulskips-source tree without the socket loop");
    read_exact_checkpoint(H, R, exact_restart, ck);
    printf(" Continuing from the complete checkpoint %s: Iter %lld Time %.17g\n", exact_restart, (long long)ck.iter, ck.time);
  }
  else if (H.InitSt == "F") { H.Len1 -= H.Len1 % 24; set_flat(H, H.Len1, false, true); }
  else if (H.InitSt == "Fw") {                               // SetSiC3Cww.f: SetSiC3C.f between wall planes, plus a wall block from Len2 up
    H.Len1 -= H.Len1 % 24; H.Len2 -= H.Len2 % 24;
    set_flat(H, H.Len1, true, true); wall_planes(H, H.Len2, H.LZ - 7);
  } else if (H.InitSt == "Fa") {                             // SetSiC3Cwwaperture.f: the wall block leaves an aperture of Len3 along x
    H.Len1 -= H.Len1 % 24; H.Len2 -= H.Len2 % 24; H.Len3 -= H.Len3 % 24;
    set_flat(H, H.Len1, true, true);
    if ((int)(0.5f * (float)(H.LX - H.Len3)) - 1 >= 0) wall_planes(H, H.Len2, H.LZ - 7, 0, (int)(0.5f * (float)(H.LX - H.Len3)) - 1);
    wall_planes(H, H.Len2, H.LZ - 7, (int)(0.5f * (float)(H.LX + H.Len3)), H.LX - 1);
  }
  else if (H.InitSt == "FS") { H.Len1 -= H.Len1 % 24; set_flat(H, H.Len1, true, false); }
  else if (H.InitSt == "SG") { H.Len1 -= H.Len1 % 24; set_flat(H, H.Len1, true, false, true); }
  else if (H.InitSt == "C") set_nc(H, H.Len1, H.Len2, H.Len3, false);
  else if (H.InitSt == "S") set_nc(H, H.Len1, H.Len1, H.Len1, true);
  else if (H.InitSt == "SS") set_nc(H, H.Len1, H.Len1, H.Len1, true, false);
  else if (H.InitSt == "I") set_invp_off(H, H.Len1, H.Len2);
  else if (H.InitSt == "D") set_invp_off(H, H.Len1, H.Len2, 2);
  else if (H.InitSt == "Z") set_invp_off(H, H.Len1, H.Len2, 1);
  else if (H.InitSt == "K") set_invp_off(H, H.Len1, H.Len2, 3);
  else if (H.InitSt == "L") set_invp_off(H, H.Len1, H.Len2, 4);
  else if (H.InitSt == "J") set_invp_off(H, H.Len1, H.Len2, 5);
  else if (H.InitSt == "T") set_trench(H);
  else if (H.InitSt == "FF") { H.Len1 -= H.Len1 % 24; H.Len2 -= H.Len2 % 24; H.Len3 -= H.Len3 % 24; H.Len4 -= H.Len4 % 24; set_finfet(H, H.Len1, H.Len2, H.Len3, H.Len4); }
  else if (H.InitSt == "ST") set_trench_si(H);
  else if (H.InitSt == "A") set_afb_sym_zimb1(H, H.Len1);
  else if (H.InitSt == "SH") set_shape(H);
  else if (H.InitSt == "E") {                                 // main.f:267-272
    const int VSF[3] = {H.Len2, H.Len3, H.Len1 - H.Len2};
    printf(" %d %d %d %d\n", H.Len1, VSF[0], VSF[1], VSF[2]);
    set_sic_sf(H, H.Len1, VSF);
  }
  else if (H.InitSt == "IN") set_cad(H);
  else if (H.InitSt == "LA") { set_T(H); set_cad(H); printf(" Max temperature is:  %d K\n", H.tmax); }
  else die(" mulskips.e (B200 host): InitSt %s is not built (all 22 codes of main.f:225-295 are built)", H.InitSt.c_str());
  R.NumAtoms = (int)H.bulk.size(); R.NumVoids = (int)H.voids.size(); R.NumAdAtom = (int)H.mc.size();
  if (H.la_tree) {                                            // la/main.f:592-611,628
    H.NumOcc1 = R.NumAtoms + H.NumAdAtomOcc;
    for (int k = 1; k <= H.NCrystal; k++) {
      if (k == 1) H.InitCountCrystal[0] = H.CountCrystal[0] + (int)((H.NumOcc0 - H.NumOcc1) * (1 - H.alloy));
      else if (k == 2) H.InitCountCrystal[1] = H.CountCrystal[1] + (int)((H.NumOcc0 - H.NumOcc1) * H.alloy);
    }
    printf(" \n Initial number of occupied sites (before nucleation):  %d\n", H.NumOcc0);
    printf(" Initial CountCrystal (before nucleation):  %d %d %d\n", H.InitCountCrystal[0], H.InitCountCrystal[1], H.InitCountCrystal[2]);
    printf(" Initial number of occupied sites (after nucleation):  %d\n", H.NumOcc1);
    printf(" Initial CountCrystal (after nucleation):  %d %d %d\n", H.CountCrystal[0], H.CountCrystal[1], H.CountCrystal[2]);
    printf(" Initial number of liquid sites (after nucleation):  %d\n", H.NumOcc0 - H.NumOcc1);
    printf(" Initial CountLiquid (after nucleation):  %d %d %d\n \n", H.CountLiquid[0], H.CountLiquid[1], H.CountLiquid[2]);
    printf(" Atoms,Voids,AdAtoms,AdAtomsOcc %d %d %d %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom, H.NumAdAtomOcc);
  } else
  printf(" Atoms,Voids,AdAtoms %d %d %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom);
  if (!exact_restart) {
  for (int i = 0; i < 3; i++) H.CountCrystalOld[i] = H.CountCrystal[i];
  write_frame(H, R, 0.0, 0, 0, false);                       // main.f:297-300
  }
  if (dump_state) {
    // the module globals the builder left, in list order: rates as raw bits, the ran3 stream position, a 64-bit FNV-1a of LattCoo
    FILE *f = fopen(dump_state, "w"); if (!f) die(" cannot write %s", dump_state);
    unsigned long long hsh = 1469598103934665603ull;
    for (int32_t w : H.L) for (int b = 0; b < 4; b++) { hsh ^= (unsigned long long)(((uint32_t)w >> (8 * b)) & 255u); hsh *= 1099511628211ull; }
    fprintf(f, "sizes %zu %zu %zu\nlattice_fnv %016llx\ncounts %d %d %d %d %d %d\nran3 %d %d %d %d", H.mc.size(), H.bulk.size(), H.voids.size(), hsh,
            H.CountCrystal[0], H.CountCrystal[1], H.CountCrystal[2], H.CountCov[0], H.CountCov[1], H.CountCov[2], H.Idum, H.r_iff, H.r_inext, H.r_inextp);
    for (int i = 1; i <= 55; i++) fprintf(f, " %d", H.r_ma[i]);
    fputc('\n', f);
    auto put_site = [&](const char *tag, const Site &a) { fprintf(f, "%s %d %d %d", tag, a.xyz[0], a.xyz[1], a.xyz[2]); for (int j = 0; j < 4; j++) fprintf(f, " %d %d %d", a.nn[j][0], a.nn[j][1], a.nn[j][2]); };
    for (const Active &a : H.mc) { put_site("mc", a); unsigned long long bits; memcpy(&bits, &a.p, 8); fprintf(f, " %d %016llx\n", a.ev, bits); }
    for (const Site &a : H.bulk) { put_site("bulk", a); fputc('\n', f); }
    for (const Site &a : H.voids) { put_site("void", a); fputc('\n', f); }
    fclose(f);
  }
  // ---- hand the state to the GPU (include/mskc.h) ----
  if (mskc_create(H.LX, H.LY, H.LZ, H.NCrystal, H.NCov, H.ListCrystal, H.ListCov, device, 0, &R.h) != 0) die(" %s", mskc_last_error(nullptr));
  if (H.st()) MS(mskc_set_rate_field(R.h, H.ntemp, H.reftemp.data(), H.refPE.data(), H.refPD.data(), H.refPA.data(), H.refPS.data(), H.PtransZig,
                                     H.t_lo, H.t_hi, H.fixedT ? (int)H.Ttop : 0));
  else MS(mskc_set_rates(R.h, H.PE.data(), H.PD.data(), H.PA.data(), H.PS.data(), H.PtransZig));
  if (H.la_tree) MS(mskc_set_tree_variant(R.h, MSKC_TREE_LA));
  if (H.la() && H.la_tree) MS(mskc_set_la_fork(R.h, 1, H.Tm_k, H.P0_k, H.sigma_k, H.Tflex_k, H.t_lo, H.t_hi, H.InitCountCrystal, H.NumOcc0, H.NumAdAtomOcc));
  else if (H.la()) MS(mskc_set_la(R.h, 1, H.Tm, H.P0, H.sigma, H.Tflex, H.t_lo, H.t_hi));
  {
    std::vector<int32_t> mx(3 * H.mc.size() + 3), mn(12 * H.mc.size() + 12), me(H.mc.size() + 1), bx(3 * H.bulk.size() + 3), bn(12 * H.bulk.size() + 12), vx(3 * H.voids.size() + 3), vn(12 * H.voids.size() + 12);
    for (size_t i = 0; i < H.voids.size(); i++) { memcpy(&vx[3 * i], H.voids[i].xyz, 12); memcpy(&vn[12 * i], H.voids[i].nn, 48); }
    std::vector<double> mr(H.mc.size() + 1);
    for (size_t i = 0; i < H.mc.size(); i++) { memcpy(&mx[3 * i], H.mc[i].xyz, 12); memcpy(&mn[12 * i], H.mc[i].nn, 48); me[i] = H.mc[i].ev; mr[i] = H.mc[i].p; }
    for (size_t i = 0; i < H.bulk.size(); i++) { memcpy(&bx[3 * i], H.bulk[i].xyz, 12); memcpy(&bn[12 * i], H.bulk[i].nn, 48); }
    int32_t st[58]; st[0] = H.r_iff; st[1] = H.r_inext; st[2] = H.r_inextp; for (int i = 1; i <= 55; i++) st[2 + i] = H.r_ma[i];
    MS(mskc_seed(R.h, H.Idum, H.RunType[0], exact_restart ? ck.counter : 0, st));
    MS(mskc_upload_state(R.h, H.L.data(), (int)H.mc.size(), mx.data(), mn.data(), me.data(), mr.data(), (int)H.bulk.size(), bx.data(), bn.data(), (int)H.voids.size(), vx.data(), vn.data()));
    if (exact_restart) MS(mskc_set_clock(R.h, ck.iter, ck.time, H.CountCrystal, H.CountCov));
    else MS(mskc_set_clock(R.h, 1, 0.0, H.CountCrystal, H.CountCov));
  }
  std::vector<Active>().swap(H.mc); std::vector<Site>().swap(H.bulk); std::vector<Site>().swap(H.voids);
  by_time = H.ExitStrategy == "Time";
  };
  // ---- the KMC loop between two calls of the host (main.f:304-371) ----
  auto kmc_loop = [&]() {
  pick_mode();
  if (!parallel) {                                            // main.f:304-371, bit-exact replay on the device
    mskc_parity_args a{}; mskc_parity_result r{};
    a.max_iter = H.MaxIter; a.max_time = H.MaxTime; a.exit_on_time = by_time; a.out_every = H.OutMolMol; a.out_time = H.OutTime; a.event_budget = 0;
    while (true) {
      a.framecounter = R.framecounter;
      if (mskc_run_parity(R.h, &a, &r) != 0) die(" %s", mskc_last_error(R.h));
      R.Iter = r.iter; R.IterOut = r.iter_out; R.Time = r.time; R.NumAtoms = r.num_atoms; R.NumVoids = r.num_voids; R.NumAdAtom = r.num_adatom;
      dev_ms += r.kernel_ms; events += r.events_done;
      if (r.stop_reason == MSKC_STOP_FEWMC) {
        if (H.la_tree) { mskc_counters c; MS(mskc_get_counters(R.h, &c)); printf(" few MC particles: stop MC %lld %d %d\n", R.Iter, R.NumAdAtom, c.num_adatom_occ); }
        else printf(" few MC particles: stop MC %lld %d\n", R.Iter, R.NumAdAtom);
        break;
      }
      if (r.stop_reason == MSKC_STOP_RESOLIDIFIED) {          // la/main.f:714-723
        mskc_counters c; MS(mskc_get_counters(R.h, &c));
        printf(" \n Current number of occupied sites:  %d\n    > initial number of occupied sites:  %d\n ... STOPPING MulSKIPS ...\n \n", c.num_atoms + c.num_adatom_occ, H.NumOcc0);
        break;
      }
      if (r.stop_reason == MSKC_STOP_FRAME) {                 // main.f:316-333
        if (!by_time) R.framecounter = R.IterOut / H.OutMolMol;
        write_frame(H, R, R.Time, R.IterOut, by_time ? R.framecounter + 1 : R.framecounter, true);
        printf(" Iter, Time, Site %lld %.17g %d %d %d\n", R.Iter, R.Time, r.last_site[0], r.last_site[1], r.last_site[2]);
        if (H.la_tree) printf(" NatNvoidNadNadocc %d %d %d %d\n Number of occupied sites:  %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom, H.NumAdAtomOcc, R.NumAtoms + H.NumAdAtomOcc);
        else printf(" NatNvoidNad %d %d %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom);
        if (by_time) R.framecounter++;
        continue;
      }
      break;                                                  // MaxIter / MaxTime
    }
  } else {                                                    // production mode: synchronous-sublattice sampler, frames at the same cadence
    mskc_parallel_args a{}; mskc_parallel_result r{};
    a.events_per_tile = ept; a.seed = (uint64_t)(H.Idum0 < 0 ? -(long long)H.Idum0 : H.Idum0); a.n_sweeps = 8;      // Philox key = the seed of start.dat
    long long next_out = H.OutMolMol; double next_t = H.OutTime;
    // a leg of a socket-driven run starts its own clock and iteration count (la/main.f:621-624); the sampler's clock runs on
    long long leg_events = 0;
    double t_leg0 = 0.0;
    { mskc_counters c0; MS(mskc_get_counters(R.h, &c0)); t_leg0 = c0.time; }
    double tau = 0.0;                                         // the sampler's adaptive quantum as of the last call
    while (true) {
      // frames by time: land on the frame time instead of overshooting it by up to a call's worth of quanta -- whole sweeps
      // while they fit, then one sweep with the remaining time as its quantum (a shorter quantum only lowers the O(tau) bias)
      a.n_sweeps = (by_time && tau == 0.0) ? 1 : 8; a.tau_fixed = 0.0;      // (first call: learn the quantum)
      if (by_time && tau > 0.0) {                             // the next time the loop has to stop at: a frame, or the end of the run
        const double stop_t = H.OutTime > 0 ? std::min(next_t, H.MaxTime) : H.MaxTime;
        const double rem = stop_t - R.Time;
        const double nq = rem / tau;                          // quanta left at the present rate
        if (nq > 12.0) a.n_sweeps = 8;                        // far away: adaptive quanta
        else if (rem > 1e-12 * stop_t) { a.n_sweeps = std::max<long long>(1, (long long)std::ceil(nq)); a.tau_fixed = rem / (double)a.n_sweeps; }   // land on it: equal quanta <= tau
        else { a.n_sweeps = 1; a.tau_fixed = tau; }
      }
      if (mskc_run_parallel(R.h, &a, &r) != 0) die(" %s", mskc_last_error(R.h));
      tau = r.tau;
      events += r.events_done; leg_events += r.events_done; dev_ms += r.kernel_ms; R.Time = r.time - t_leg0; R.Iter = leg_events + 1; R.IterOut = leg_events;
      mskc_counters c; MS(mskc_get_counters(R.h, &c)); R.NumAtoms = c.num_atoms; R.NumVoids = c.num_voids; R.NumAdAtom = c.num_adatom;
      if (r.n_active <= 20) { printf(" few MC particles: stop MC %lld %d\n", R.Iter, (int)r.n_active); break; }
      if (H.la_tree && H.la() && c.num_atoms + c.num_adatom_occ > H.NumOcc0) {            // la/main.f:714-723, checked once per call of the sampler
        printf(" \n Current number of occupied sites:  %d\n    > initial number of occupied sites:  %d\n ... STOPPING MulSKIPS ...\n \n", c.num_atoms + c.num_adatom_occ, H.NumOcc0);
        break;
      }
      bool frame = by_time ? (H.OutTime > 0 && R.Time >= next_t * (1.0 - 1e-12)) : (H.OutMolMol > 0 && leg_events >= next_out);
      if (frame) {
        R.framecounter++;
        write_frame(H, R, R.Time, R.IterOut, R.framecounter, true);
        printf(" Iter, Time, Site %lld %.17g %d %d %d\n", R.Iter, R.Time, 0, 0, 0);
        printf(" NatNvoidNad %d %d %d\n", R.NumAtoms, R.NumVoids, R.NumAdAtom);
        next_out += H.OutMolMol; next_t += H.OutTime;
      }
      if (by_time ? R.Time >= H.MaxTime * (1.0 - 1e-12) : leg_events >= H.MaxIter) break;     // (the serial loop stops one event past MaxTime)
    }
  }
  };
  // main.f:374-430: final frame, checkpoint, Coor
  auto finish_run = [&]() {
  write_frame(H, R, R.Time, R.IterOut, by_time ? R.framecounter : (H.OutMolMol > 0 ? R.IterOut / H.OutMolMol : 0), true);
  if (R.lattice_stale && (H.SaveFinalState == "T" || H.SaveCoo == "T")) { MS(mskc_download_lattice(R.h, host_lattice(H))); R.lattice_stale = false; }   // the files below hold every cell
  if (H.SaveFinalState == "T") {
    printf(" Writing checkpoint file for a restart run %s\n", H.restartfile.c_str());
    if (FILE *f = fopen(H.restartfile.c_str(), "w")) {
      for (int z = 0; z < H.LZ; z++) { for (int y = 0; y < H.LY; y++) for (int x = 0; x < H.LX; x++) fprintf(f, " %11d", H.L[H.ci(x, y, z)]); fputc('\n', f); }
      fclose(f);
    }
    printf(" Done writing checkpoint file for a restart run %s\n", H.restartfile.c_str());
    if (!parallel && !H.la_tree) write_exact_checkpoint(H, R, H.restartfile + ".mskc");      // (stdout stays the reference's)
  }
  if (H.SaveCoo == "T" && H.la_tree && H.coofile == "None") {          // la/main.f:795-817: phases and species go back through the socket
    printf(" Sending phases and Ge fraction back to LA code \n");
    H.sock_c.assign((size_t)H.LZ * H.LX * H.LY, 0); H.sock_d.assign((size_t)H.LZ * H.LX * H.LY, 0);
    for (int z = 0; z < H.LZ; z++) for (int y = 0; y < H.LY; y++) for (int x = 0; x < H.LX; x++) {
      const int32_t w = H.L[H.ci(x, y, z)]; const size_t k = (size_t)z + (size_t)H.LZ * ((size_t)x + (size_t)H.LX * y);
      if (ibits(w, PosIndex, LenIndex) == 113) { H.sock_c[k] = 10; H.sock_d[k] = 10; }
      else if (ibits(w, PosOcc, 1) == 1) { H.sock_c[k] = ibits(w, PosCoor, LenCoor); H.sock_d[k] = ibits(w, PosSiC, LenSiC); }
    }
    printf(" Done sending phases and Ge fraction back to LA code \n");
  } else
  if (H.SaveCoo == "T") {
    printf(" Writing Coor (LA phases) file %s\n", H.coofile.c_str());
    if (FILE *f = fopen(H.coofile.c_str(), "w")) {
      for (int z = 0; z < H.LZ; z++) {
        for (int y = 0; y < H.LY; y++) for (int x = 0; x < H.LX; x++) { int32_t w = H.L[H.ci(x, y, z)]; int v = ibits(w, PosIndex, LenIndex) == 113 ? 10 : (ibits(w, PosOcc, 1) == 1 ? ibits(w, PosCoor, LenCoor) : 0); fprintf(f, "%4d", v); }
        fputc('\n', f);
      }
      fclose(f);
    }
    printf(" Done writing Coor (LA phases) file %s\n", H.coofile.c_str());
  }
  };
  if (sock_port < 0) { init_run(); kmc_loop(); finish_run(); }
  else {
    // ---- socket-driven run: la/main.f:160-345 (message loop) and :346-852 (one KMC leg per temperature map) ----
    if (!H.la_tree) die(" mulskips.e (B200 host): the socket loop is the LaserAnnealing fork's: use --la-tree -p port");
    Wire S; S.connect_to(sock_host, sock_port);
    const size_t narr = (size_t)H.LZ * H.LX * H.LY;
    bool isinit = false, phasesready = false, xgeready = false, geoisinit = false, fieldisinit = false;
    long long LAcounter = 0;
    auto get_array = [&](bool have, const std::vector<int32_t> &arr, const char *what) {
      if (!isinit) printf(" Driver not iniliasied.\n");
      else if (!have) printf(" Driver does not have %s to send\n", what);
      else { S.put_header("DATAREADY"); printf(" @ message to server: DATAREADY\n"); S.wr(arr.data(), narr * 4); }
    };
    while (true) {
      printf(" Current status: geoisinit, fieldisinit, phasesready, XGeready %c %c %c %c\n", geoisinit ? 'T' : 'F', fieldisinit ? 'T' : 'F', phasesready ? 'T' : 'F', xgeready ? 'T' : 'F');
      fflush(stdout);
      std::string header;
      if (!S.get_header(header)) { printf(" server has quit or connection broke\n"); break; }
      printf(" @ message from server: %s\n", header.c_str());
      if (header == "STATUS") {
        const char *reply = !isinit ? "NEEDINIT" : (phasesready || xgeready) ? "HAVEDATA" : "READY";
        S.put_header(reply); printf(" @ message to server: %s\n", reply);
      } else if (header == "INIT") { printf("  Initializing system from server\n"); isinit = true; }
      else if (header == "SENDGEO" || header == "SENDFIELD") {
        if (!isinit) printf(" Driver not iniliasied.\n");
        else if (phasesready || xgeready) printf(" Driver has data to send back to server\n");
        else {
          std::vector<int32_t> &dst = header == "SENDGEO" ? H.sock_a : H.sock_b;
          dst.resize(narr);
          if (!S.rd(dst.data(), narr * 4)) die(" Error reading from socket: server has quit or connection broke");
          printf(" Received array from server:\n");
          (header == "SENDGEO" ? geoisinit : fieldisinit) = true;
        }
      }
      else if (header == "GETGEO") get_array(geoisinit, H.sock_a, "data");
      else if (header == "GETFIELD") get_array(fieldisinit, H.sock_b, "data");
      else if (header == "GETPHASES") { get_array(phasesready, H.sock_c, "phases"); if (isinit && phasesready) phasesready = false; }
      else if (header == "GETXGE") { get_array(xgeready, H.sock_d, "Ge fraction"); if (isinit && xgeready) xgeready = false; }
      else if (header == "EXIT") break;                        // pymulskips/sockets.py:196-200 (the Fortran driver stops on it as an unexpected header)
      else { printf("  unexpected header %s\n", header.c_str()); break; }
      if (!geoisinit || !fieldisinit) continue;                // la/main.f:343
      printf(" KMC ACTUALLY STARTING NOW: geoisinit, fieldisinit, phasesready, XGeready %c %c %c %c\n", geoisinit ? 'T' : 'F', fieldisinit ? 'T' : 'F', phasesready ? 'T' : 'F', xgeready ? 'T' : 'F');
      if (LAcounter == 0) init_run();
      else {
        // la/main.f:616-631: the new temperature map goes into the lattice words (initial/SetT.f), the rates of the sites
        // already listed stay as they are until events touch them; clock, iteration and frame counters restart
        std::vector<int32_t> T(narr); int lo = H.t_lo, hi = H.t_hi;
        for (size_t i = 0; i < (size_t)H.LX * H.LY; i++) for (int z = 0; z < H.LZ; z++) { const int t = H.sock_b[(size_t)z + (size_t)H.LZ * i]; T[(size_t)z * H.LX * H.LY + i] = t; if (t != 10 && t > 0) { lo = std::min(lo, t); hi = std::max(hi, t); } }
        printf(" b(1,1) %d\n Setting temperatures in LattCoo from array received via sockets\n", H.sock_b[0]);
        MS(mskc_set_temperature(R.h, T.data()));
        printf(" Done setting temperatures in LattCoo\n");
        mskc_counters c; MS(mskc_get_counters(R.h, &c));
        if (lo < H.t_lo || hi > H.t_hi) { H.t_lo = lo; H.t_hi = hi; MS(mskc_set_la_fork(R.h, 1, H.Tm_k, H.P0_k, H.sigma_k, H.Tflex_k, H.t_lo, H.t_hi, H.InitCountCrystal, H.NumOcc0, c.num_adatom_occ)); }
        MS(mskc_set_clock(R.h, 1, 0.0, c.CountCrystal, c.CountCov));
        R.Iter = 1; R.IterOut = 0; R.framecounter = 0; R.Time = 0.0;
        R.NumAtoms = c.num_atoms; R.NumVoids = c.num_voids; R.NumAdAtom = c.num_adatom;
        printf(" Atoms,Voids,AdAtoms,AdAtomsOcc %d %d %d %d\n", c.num_atoms, c.num_voids, c.num_adatom, c.num_adatom_occ);
        for (int i = 0; i < 3; i++) H.CountCrystalOld[i] = c.CountCrystal[i];
        write_frame(H, R, 0.0, 0, 0, true);
      }
      kmc_loop();
      finish_run();
      fieldisinit = false; phasesready = true; xgeready = true; LAcounter++;
    }
  }
  if (R.crosscheck) fprintf(stderr, "frame cross-check: %lld frames, device passes == whole-lattice scans\n", R.crosschecked);
  fprintf(stderr, "mulskips.e (B200): %lld events, %.1f ms on the device = %.4g events/s (%s mode)\n", events, dev_ms, dev_ms > 0 ? events / (dev_ms * 1e-3) : 0.0, parallel ? "sublattice-sampler" : "serial-parity");
  mskc_destroy(R.h);
  return 0;
}
