// mskc_device.cuh -- device-side core of the KMC event loop (sm_100a).
//
// Data layout in HBM (one copy per GPU):
//   state[cell]  u32  the reference's LattCoo word, bit-for-bit (modules/defsystem.f:87-91):
//                     bits 0-7 coverage/wall id, 9-21 T[K], 23 occupied, 24-26 coordination, 27-29 species
//   bonds[cell]  u64  the four bond vectors of the site, 16 bits each: 5-bit biased dx,dy,dz
//                     (raw 0 = "not fixed", decodes to the absolute coordinate (0,0,0) exactly as
//                     the reference's NextNXYZ=0 does).  Replaces the NextNXYZ(3,4) member of the
//                     three AoS particle lists (modules/defdertype.f:26-48): one record per CELL,
//                     so bulk/void/MC migration moves no data.
// Cell index = x + LenX*(y + LenY*z), z in 0..LenZ (the reference's (0:LenX-1,0:LenY-1,0:LenZ)).
//
// The transition rules below follow mcevent/{Get_Prob,deposition,Evaporation,absorption,desorption,
// IF_Arm_Choice}.f and lattice/{FIND_NN,FIND_SN}.f; they are written once and instantiated with two
// bookkeeping policies: ParityPol (mskc_parity.cuh: the reference's lists + rate tree, bit-exact) and
// TilePol (mskc_parallel.cuh: per-tile active lists for the synchronous-sublattice sampler).
#pragma once
#include <cstdint>
#ifndef MSKC_EMU
#include <cuda_runtime.h>
// kernel launch / dynamic shared memory, spelled through macros so that tests/emu/cuda_emu.h (a functional SIMT
// emulator used by the GPU-less test container, test infrastructure only) can compile these sources with g++
#define MSKC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MSKC_DYN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace mskc {

enum { PosIndex = 0, LenIndex = 8, PosT = 9, LenT = 13, PosOcc = 23, PosCoor = 24, LenCoor = 3, PosSiC = 27, LenSiC = 3 };
enum { EV_DRAW = 100 };   // parallel mode: "deposit or adsorb, species drawn when the event fires"

// error codes (mirroring the reference's STOP sites)
enum {
  E_NONE = 0, E_OOB = 1 /* z out of range */, E_FARBOND = 2, E_TREE_ZERO = 3 /* UpdateTree.f:38 */, E_TREE_IDX = 4 /* UpdateTree.f:44 */,
  E_PICK = 5 /* PickTreeEvent.f:58 */, E_ADD_ZERO = 6 /* Add.f:38 */, E_MC_FULL = 7 /* Add.f:42 */, E_VOID_FULL = 8 /* Add.f:94 */,
  E_ERASE = 9 /* Erase.f:47 */, E_COVOCC = 10 /* Get_Prob.f:94 */, E_COVLIST = 11 /* Get_Prob.f:108 */, E_CENSUS = 12 /* Get_Prob.f:293 */,
  E_TEMP0 = 13 /* Get_Prob.f:133 */, E_NNCOV = 14 /* deposition.f:194 */, E_TILE_FULL = 15, E_STALE_NBUFF = 16, E_TRANGE = 17, E_BADEV = 18,
  E_FLAGS = 19 /* incremental census flags disagree with the full census (mskc_coop.cuh) */,
  E_HALO_TIMEOUT = 20 /* slab mode: a neighbour rank never published its halo records */
};

struct P3 { int x, y, z; };
__device__ __forceinline__ bool eq(P3 a, P3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct DevCtx {
  int LX, LY, LZ;
  long long ncell;                  // LX*LY*(LZ+1)
  uint32_t *state;
  unsigned long long *bonds;
  uint32_t *aux;                    // sampler: list slot (bits 0-15) + census flags (bits 16-23) per cell, see mskc_coop.cuh
  const double *PE, *PD, *PAbs, *PDes;   // Fortran-flattened tables (defsystem.f:60-65)
  double PZig;
  double PDepAbs[4];                // [coor]: sum of the deposition + adsorption rates of an empty site (Get_Prob.f:160-163), plain tables
  double cPD[9];                    // PtransD(sp, coor) again, in the constant bank (species draw of the sampler)
  int NCrystal, NCov, ListCov[3], ListCrystal[3];
  // temperature-dependent rates: host-built tables per integer kelvin.  la = 1: the LA rate law (Get_Prob.f:143-148,
  // evaporation + deposition); la = 2: the 'ST' tabulated field (InterpProbFromField.f: all four tables, and the
  // site temperature is consulted for coverage sites too, Get_Prob.f:76-91)
  int la, keepT, la_tmin, la_tmax, la_nnz;
  const short *la_slot;             // [12288] -> column in la_E, -1 where PtransE==0
  const double *la_E;               // [(tmax-tmin+1)][la_nnz]
  const double *la_D;               // [(tmax-tmin+1)][9]
  const double *la_A, *la_S;        // la = 2 only: [(tmax-tmin+1)][12] absorption / desorption
  // la = 3: the LaserAnnealing fork's law (mulskips-source-LaserAnnealing/mcevent/Get_Prob.f:142-152): per-species
  // P0 / Tm / sigma / Tflex, melting from la_E as above, solidification = 0.5*P0(k) * FLOAT(CountLiquid(k))/SUM(CountLiquid)
  // * EXP(-PtransD/(kb*Tm(k))) * (1+ERF((T-Tflex(k))/sigma(k))) * merge, multiplied left to right on the device from
  // host-built factors (glibc exp/erf) so that the product is the CPU's bit for bit
  double la3_halfP0[3], la3_expD[9], la3_merge[9];
  const double *la3_F;              // [(tmax-tmin+1)][3]: 1 + erf((T - Tflex(k)) / sigma(k))
  int la3_init[3];                  // InitCountCrystal (la/main.f:596-602)
  int la3_swap;                     // Si/Ge melt: deposit the other species once one is used up (la/main.f:675-681)
  double rot_c[2], rot_s[2];        // cos/sin(2*irot*Pi_f/3), host libm (FIND_NN.f:51-52)
  int nn_exact;                     // 1: the integer form of FIND_NN.f on the ideal bond shell was verified against the double routine
  const unsigned short *arm_tab;    // [32768] arm-chair offset code per bond code (mskc_coop.cuh), 0 = use find_sn()
  const uint32_t *pa_tab;           // [192] what an event does to one neighbour of its site (mskc_coop.cuh: pa_rule)
  int *err;                         // [8]: code, x, y, z, aux...
};

// The context lives in constant memory: every D.field below is a constant-bank operand (no parameter
// copies on the stack of the non-inlined transition functions).  mskc_api.cu uploads the handle's
// DevCtx before launching (one context is "current" per device at a time).
__constant__ DevCtx D;
// la = 3: CountLiquid(1..3) and its sum as of the top of the current iteration (la/main.f:636).  The serial-parity
// kernel rewrites it before every event; the sampler and the trace replay take it from the host per call / per event.
__device__ int g_la_CL[4];

// cold path, kept out of line so the hot code stays small (the kernels are instruction-fetch sensitive)
__device__ __noinline__ void fail(int code, int a = 0, int b = 0, int c = 0, int d = 0) {
  if (atomicCAS(&D.err[0], 0, code) == 0) { D.err[1] = a; D.err[2] = b; D.err[3] = c; D.err[4] = d; }
}

__device__ __forceinline__ int IBITS(uint32_t w, int pos, int len) { return (int)((w >> pos) & ((1u << len) - 1u)); }
__device__ __forceinline__ uint32_t MVBITS(uint32_t to, int from, int len, int pos) {
  uint32_t m = (1u << len) - 1u; return (to & ~(m << pos)) | (((uint32_t)from & m) << pos);
}
__device__ __forceinline__ int w_occ(uint32_t w) { return (w >> PosOcc) & 1; }
__device__ __forceinline__ int w_coor(uint32_t w) { return (w >> PosCoor) & 7; }
__device__ __forceinline__ int w_cov(uint32_t w) { return w & 255; }
__device__ __forceinline__ int w_sp(uint32_t w) { return (w >> PosSiC) & 7; }

// cell index; out-of-range z (the reference would read foreign memory) raises E_OOB and is redirected
// to cell 0 so the kernel can unwind.
__device__ __forceinline__ long long cidx(P3 p) {
  if ((unsigned)p.x >= (unsigned)D.LX || (unsigned)p.y >= (unsigned)D.LY || (unsigned)p.z > (unsigned)D.LZ) {
    fail(E_OOB, p.x, p.y, p.z); return 0;
  }
  return (long long)p.x + (long long)D.LX * ((long long)p.y + (long long)D.LY * (long long)p.z);
}
__device__ __forceinline__ P3 cell_xyz(long long c) {     // ncell < 2^32 (mskc_create checks)
  P3 p; unsigned plane = (unsigned)D.LX * (unsigned)D.LY, cc = (unsigned)c;
  unsigned z = cc / plane, r = cc - z * plane, y = r / (unsigned)D.LX;
  p.z = (int)z; p.y = (int)y; p.x = (int)(r - y * (unsigned)D.LX); return p;
}
// fire-and-forget loads that pull a cell's words into L1 (results unused; volatile asm is not eliminated)
#ifndef MSKC_EMU
__device__ __forceinline__ void touch32(const void *p) { unsigned v; asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); }
__device__ __forceinline__ void touch64(const void *p) { unsigned long long v; asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); }
#else
__device__ __forceinline__ void touch32(const void *) {}
__device__ __forceinline__ void touch64(const void *) {}
#endif
__device__ __forceinline__ bool in_box(P3 p) {
  return (unsigned)p.x < (unsigned)D.LX && (unsigned)p.y < (unsigned)D.LY && (unsigned)p.z <= (unsigned)D.LZ;
}
__device__ __forceinline__ long long cidx_nofail(P3 p) {
  return (long long)p.x + (long long)D.LX * ((long long)p.y + (long long)D.LY * (long long)p.z);
}

// ---- bond codec ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_bond(P3 s, P3 n) {
  if (n.x == 0 && n.y == 0 && n.z == 0) return 0u;      // "not fixed" == absolute origin, as in the reference
  int dx = n.x - s.x, dy = n.y - s.y, dz = n.z - s.z;
  if (dx > D.LX / 2) dx -= D.LX; else if (dx < -(D.LX / 2)) dx += D.LX;
  if (dy > D.LY / 2) dy -= D.LY; else if (dy < -(D.LY / 2)) dy += D.LY;
  if (dx < -15 || dx > 15 || dy < -15 || dy > 15 || dz < -15 || dz > 15) { fail(E_FARBOND, s.x, s.y, s.z); return 0u; }
  return (unsigned)(dx + 16) | ((unsigned)(dy + 16) << 5) | ((unsigned)(dz + 16) << 10);
}
__device__ __forceinline__ P3 dec_bond(P3 s, unsigned raw) {
  P3 n;
  if (raw == 0u) { n.x = n.y = n.z = 0; return n; }
  n.x = s.x + (int)(raw & 31u) - 16; n.y = s.y + (int)((raw >> 5) & 31u) - 16; n.z = s.z + (int)((raw >> 10) & 31u) - 16;
  if (n.x >= D.LX) n.x -= D.LX; else if (n.x < 0) n.x += D.LX;
  if (n.y >= D.LY) n.y -= D.LY; else if (n.y < 0) n.y += D.LY;
  return n;
}
__device__ __noinline__ unsigned long long enc_bonds(P3 s, const P3 nn[4]) {
  return (unsigned long long)enc_bond(s, nn[0]) | ((unsigned long long)enc_bond(s, nn[1]) << 16) |
         ((unsigned long long)enc_bond(s, nn[2]) << 32) | ((unsigned long long)enc_bond(s, nn[3]) << 48);
}
__device__ __noinline__ void dec_bonds(P3 s, unsigned long long b, P3 nn[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) nn[i] = dec_bond(s, (unsigned)((b >> (16 * i)) & 0xFFFFu));
}
// "is `me` one of the four neighbours recorded at cell n?"  (the reciprocity test, Get_Prob.f:236-239).
// Done on the packed codes: the code `me` would have relative to n is compared with the four stored
// 16-bit fields; a stored 0 ("not fixed") stands for the absolute origin, as in the reference.
__device__ __forceinline__ bool recipr_b(P3 n, unsigned long long bn, P3 me) {
  int dx = me.x - n.x, dy = me.y - n.y, dz = me.z - n.z;
  if (dx > D.LX / 2) dx -= D.LX; else if (dx < -(D.LX / 2)) dx += D.LX;
  if (dy > D.LY / 2) dy -= D.LY; else if (dy < -(D.LY / 2)) dy += D.LY;
  unsigned code = 0x10000u;                                   // matches nothing
  if (dx >= -15 && dx <= 15 && dy >= -15 && dy <= 15 && dz >= -15 && dz <= 15)
    code = (unsigned)(dx + 16) | ((unsigned)(dy + 16) << 5) | ((unsigned)(dz + 16) << 10);
  unsigned zero_code = (me.x == 0 && me.y == 0 && me.z == 0) ? 0u : 0x10000u;
  unsigned f0 = (unsigned)(bn & 0xFFFFu), f1 = (unsigned)((bn >> 16) & 0xFFFFu), f2 = (unsigned)((bn >> 32) & 0xFFFFu), f3 = (unsigned)(bn >> 48);
  return (f0 == code) | (f1 == code) | (f2 == code) | (f3 == code) | (f0 == zero_code) | (f1 == zero_code) | (f2 == zero_code) | (f3 == zero_code);
}
__device__ __forceinline__ bool recipr_nn(const P3 nn[4], P3 me) {
  return eq(nn[0], me) | eq(nn[1], me) | eq(nn[2], me) | eq(nn[3], me);
}
__device__ __forceinline__ P3 zero3() { P3 p; p.x = p.y = p.z = 0; return p; }

// ---- geometry (compiled with -fmad=false: same IEEE sequence as the CPU reference) -----------
__device__ __forceinline__ int isign(int a, int b) { return b >= 0 ? abs(a) : -abs(a); }
__device__ __forceinline__ int nint_d(double x) { return (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); }
__device__ __forceinline__ void wrapxy(P3 &p) {
  if (p.x >= D.LX) p.x -= D.LX;
  if (p.x < 0) p.x += D.LX;
  if (p.y >= D.LY) p.y -= D.LY;
  if (p.y < 0) p.y += D.LY;
}
// lattice/FIND_NN.f:25-77: the two missing neighbours of SiteC given two known ones
__device__ __noinline__ void find_nn(P3 Site1, P3 Site2, P3 SiteC, P3 out[2]) {
  int LX = D.LX, LY = D.LY;
  if (abs(SiteC.x - Site1.x) > LX / 2) Site1.x += isign(LX, SiteC.x - Site1.x);
  if (abs(SiteC.y - Site1.y) > LY / 2) Site1.y += isign(LY, SiteC.y - Site1.y);
  if (abs(SiteC.x - Site2.x) > LX / 2) Site2.x += isign(LX, SiteC.x - Site2.x);
  if (abs(SiteC.y - Site2.y) > LY / 2) Site2.y += isign(LY, SiteC.y - Site2.y);
  double ax[3] = {(double)SiteC.x - (double)Site1.x, (double)SiteC.y - (double)Site1.y, (double)SiteC.z - (double)Site1.z};
  double snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  ax[0] = ax[0] / snorm; ax[1] = ax[1] / snorm; ax[2] = ax[2] / snorm;
  double v[3] = {(double)Site2.x - (double)SiteC.x, (double)Site2.y - (double)SiteC.y, (double)Site2.z - (double)SiteC.z};
  for (int irot = 0; irot < 2; irot++) {
    double Cth = D.rot_c[irot], Sth = D.rot_s[irot], Rot[3][3], R[3];
    for (int i = 0; i < 3; i++) Rot[i][i] = ax[i] * ax[i] * (1 - Cth) + Cth;
    Rot[1][0] = ax[0] * ax[1] * (1 - Cth) - ax[2] * Sth;
    Rot[0][1] = ax[0] * ax[1] * (1 - Cth) + ax[2] * Sth;
    Rot[2][0] = ax[0] * ax[2] * (1 - Cth) + ax[1] * Sth;
    Rot[0][2] = ax[0] * ax[2] * (1 - Cth) - ax[1] * Sth;
    Rot[2][1] = ax[1] * ax[2] * (1 - Cth) - ax[0] * Sth;
    Rot[1][2] = ax[1] * ax[2] * (1 - Cth) + ax[0] * Sth;
    for (int k = 0; k < 3; k++) { R[k] = 0.0; for (int l = 0; l < 3; l++) R[k] = R[k] + Rot[l][k] * v[l]; }
    P3 p; p.x = nint_d(R[0]) + SiteC.x; p.y = nint_d(R[1]) + SiteC.y; p.z = nint_d(R[2]) + SiteC.z;
    wrapxy(p);
    out[irot] = p;
  }
}
// lattice/FIND_SN.f:26-84: zig-zag (cubic) and arm-chair (hexagonal) continuations
__device__ __noinline__ void find_sn(P3 Site, P3 SiteC, const P3 SNin[3], P3 Zig[3], P3 Arm[3]) {
  int LX = D.LX, LY = D.LY;
  if (abs(Site.x - SiteC.x) > LX / 2) SiteC.x += isign(LX, Site.x - SiteC.x);
  if (abs(Site.y - SiteC.y) > LY / 2) SiteC.y += isign(LY, Site.y - SiteC.y);
  double ax[3] = {(double)Site.x - (double)SiteC.x, (double)Site.y - (double)SiteC.y, (double)Site.z - (double)SiteC.z};
  double snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  ax[0] = ax[0] / snorm; ax[1] = ax[1] / snorm; ax[2] = ax[2] / snorm;
  const double ARM = (double)8.6602254037f;
  for (int i = 0; i < 3; i++) {
    P3 sn = SNin[i];
    if (abs(Site.x - sn.x) > LX / 2) sn.x += isign(LX, Site.x - sn.x);
    if (abs(Site.y - sn.y) > LY / 2) sn.y += isign(LY, Site.y - sn.y);
    P3 z, a;
    // 2*Cent - SN with Cent = 0.5*(SiteC+Site): exact in double
    z.x = nint_d(2 * (0.5 * ((double)SiteC.x + (double)Site.x)) - (double)sn.x);
    z.y = nint_d(2 * (0.5 * ((double)SiteC.y + (double)Site.y)) - (double)sn.y);
    z.z = nint_d(2 * (0.5 * ((double)SiteC.z + (double)Site.z)) - (double)sn.z);
    a.x = nint_d((double)sn.x + ARM * ax[0]);
    a.y = nint_d((double)sn.y + ARM * ax[1]);
    a.z = nint_d((double)sn.z + ARM * ax[2]);
    wrapxy(z); wrapxy(a);
    Zig[i] = z; Arm[i] = a;
  }
}

// ---- rate tables -------------------------------------------------------------------------------
__device__ __forceinline__ int IDXE(int sp, int i1, int i2, int i3, int i4, int i5, int i6) {
  return (sp - 1) + 3 * (i1 + 4 * (i2 + 4 * (i3 + 4 * (i4 + 4 * (i5 + 4 * i6)))));
}
__device__ __forceinline__ int site_T(uint32_t w, P3 s) {
  int t = IBITS(w, PosT, LenT);
  if (t == 0) { fail(E_TEMP0, s.x, s.y, s.z); return D.la_tmin; }
  if (t < D.la_tmin || t > D.la_tmax) { fail(E_TRANGE, s.x, s.y, s.z, t); return D.la_tmin; }
  return t;
}
__device__ __forceinline__ double rate_E(int idx, int T) {
  if (!D.la) return __ldg(&D.PE[idx]);
  int slot = D.la_slot[idx];
  return slot < 0 ? 0.0 : __ldg(&D.la_E[(long long)(T - D.la_tmin) * D.la_nnz + slot]);
}
__device__ __noinline__ double rate_D_fork(int sp, int i, int T) {
  if (sp > D.NCrystal) return 0.0;                        // the module array PD(k>NCrystal,:) is never written
  const volatile int *cl = g_la_CL;
  double v = D.la3_halfP0[sp - 1] * (double)(float)cl[sp - 1];
  v = v / (double)cl[3];
  v = v * D.la3_expD[i];
  v = v * __ldg(&D.la3_F[(long long)(T - D.la_tmin) * 3 + (sp - 1)]);
  return v * D.la3_merge[i];
}
__device__ __forceinline__ double rate_D(int sp, int coor, int T) {
  int i = (sp - 1) + 3 * (coor - 1);
  if (D.la == 3) return rate_D_fork(sp, i, T);
  return D.la ? __ldg(&D.la_D[(long long)(T - D.la_tmin) * 9 + i]) : __ldg(&D.PD[i]);
}
__device__ __forceinline__ double rate_Abs(int k, int coor, int T) {
  int i = (k - 1) + 3 * coor;
  return D.la == 2 ? __ldg(&D.la_A[(long long)(T - D.la_tmin) * 12 + i]) : __ldg(&D.PAbs[i]);
}
__device__ __forceinline__ double rate_Des(int k, int coor, int T) {
  int i = (k - 1) + 3 * coor;
  return D.la == 2 ? __ldg(&D.la_S[(long long)(T - D.la_tmin) * 12 + i]) : __ldg(&D.PDes[i]);
}

// Get_Prob.f:22-52
__device__ __noinline__ int index_event_abs(int Coor, double Ptmp, double PDepo, int T) {
  if (D.NCov == 1) return 6;
  int res = 0;
  for (int i = 1; i <= D.NCov; i++) {
    double pmin, pmax, s;
    if (i == 1) { pmin = PDepo; pmax = PDepo + rate_Abs(1, Coor, T); }
    else {
      s = 0; for (int k = 1; k <= i - 1; k++) s += rate_Abs(k, Coor, T); pmin = PDepo + s;
      s = 0; for (int k = 1; k <= i; k++) s += rate_Abs(k, Coor, T); pmax = PDepo + s;
    }
    if (Ptmp > pmin && Ptmp <= pmax) res = 5 + i;
  }
  return res;
}
// the species / adsorbate choice of Get_Prob.f:160-179 for a uniform u in [0,1)
__device__ __noinline__ int choose_dep_abs(int Coor, int T, double u, double *ptot) {
  double PD1 = rate_D(1, Coor, T), PD2 = rate_D(2, Coor, T), PD3 = rate_D(3, Coor, T);
  double PDep = PD1 + PD2 + PD3;
  double PAbs = 0; for (int k = 1; k <= 3; k++) PAbs += rate_Abs(k, Coor, T);
  double Ptot = PDep + PAbs;
  if (ptot) *ptot = Ptot;
  double Pcurr = u * Ptot;
  if (Pcurr <= PDep) {
    if (Pcurr <= PD1) return 1;
    if (Pcurr > (PD1 + PD2)) return 3;
    return 2;
  }
  return index_event_abs(Coor, Pcurr, PDep, T);
}
__device__ __noinline__ double total_dep_abs(int Coor, int T) {
  double PDep = rate_D(1, Coor, T) + rate_D(2, Coor, T) + rate_D(3, Coor, T);
  double PAbs = 0; for (int k = 1; k <= 3; k++) PAbs += rate_Abs(k, Coor, T);
  return PDep + PAbs;
}

// mcevent/Get_Prob.f:54-331.  NextN are the absolute neighbour coordinates the CALLER holds (they can
// differ from the stored bonds, exactly as in the reference, e.g. Evaporation.f:123).
template <class Pol>
__device__ __noinline__ void get_prob(Pol &pol, int Occ, int NSiC, int CovInd, int Coor, P3 Site, const P3 NextN[4], int &ev, double &prob) {
  pol.n_getprob++;
  int T = 0;
  if (D.la) T = site_T(D.state[cidx(Site)], Site);      // :76-91 ('ST': before the coverage branch) / :127-141 ('LA': NCov = 0)
  if (CovInd > 0 && CovInd < 113) {                     // desorption, :93-123
    if (Occ != 0) { fail(E_COVOCC, Site.x, Site.y, Site.z); return; }
    if (Coor <= 3) {
      int iCov = 0;
      for (int j = 1; j <= D.NCov; j++) if (D.ListCov[j - 1] == CovInd) { iCov = j; break; }
      if (iCov == 0) { fail(E_COVLIST, Site.x, Site.y, Site.z, CovInd); return; }
      ev = 5; prob = rate_Des(iCov, Coor, T);
    } else { ev = 0; prob = 0.0; }
    return;
  }
  if (Occ == 0) {
    if (Coor == 1) {
      int CoorNN = w_coor(D.state[cidx(NextN[0])]);
      if (CoorNN >= 2) {
        if (pol.predraw()) ev = choose_dep_abs(Coor, T, pol.uniform(), &prob);
        else { ev = EV_DRAW; prob = total_dep_abs(Coor, T); }
      } else { ev = 1; prob = (double)1.e-8f; }             // Get_Prob.f:183
    } else if (Coor == 2 || Coor == 3) {
      if (pol.predraw()) ev = choose_dep_abs(Coor, T, pol.uniform(), &prob);
      else { ev = EV_DRAW; prob = total_dep_abs(Coor, T); }
    } else if (Coor >= 4) { ev = 0; prob = 0.0; }
    return;
  }
  // Occ = 1: census of the four neighbours with reciprocity, :213-291
  int N_Si = 0, N_C = 0, N_B = 0, N_wall = 0, cv0 = 0, cv1 = 0, cv2 = 0;
  uint32_t wn[4]; unsigned long long bn[4]; long long cn[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { cn[i] = cidx(NextN[i]); wn[i] = D.state[cn[i]]; bn[i] = D.bonds[cn[i]]; }
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    uint32_t w = wn[i];
    int CovIndN = w_cov(w);
    if (w_occ(w)) {
      if (CovIndN != 113) {
        if (recipr_b(NextN[i], bn[i], Site)) {
          int s = w_sp(w);
          if (s == 1) N_Si++;
          if (s == 2) N_C++;
          if (s == 3) N_B++;
        }
      } else N_wall++;
    } else if (CovIndN != 0) {
      bool rec;
      if (w_coor(w) >= 4) rec = pol.cov_void_recipr(NextN[i], cn[i], Site);   // reference reads the wrong list here
      else rec = recipr_b(NextN[i], bn[i], Site);
      if (rec) {
        int iCov = 0;
        for (int j = 1; j <= D.NCov; j++) if (D.ListCov[j - 1] == CovIndN) { iCov = j; break; }
        if (iCov == 0) { fail(E_COVLIST, Site.x, Site.y, Site.z, CovIndN); return; }
        if (iCov == 1) cv0++; else if (iCov == 2) cv1++; else cv2++;
      }
    }
  }
  int ncry = N_Si + N_C + N_B + N_wall;
  if (ncry != Coor) { fail(E_CENSUS, Site.x, Site.y, Site.z, Coor * 16 + ncry); return; }
  if (ncry > 3) { ev = 0; prob = 0.0; }
  else if (ncry == 0) { ev = 4; prob = (double)100000.f * pol.tree_root(); }
  else {
    int c1 = 0, c2 = 0, c3 = 0;
    if (D.NCov >= 1) c1 = cv0;
    if (D.NCov >= 2) c2 = cv1;
    if (D.NCov >= 3) c3 = cv2;
    ev = 4;
    prob = rate_E(IDXE(NSiC, N_Si + N_wall, N_C, N_B, c1 & 3, c2 & 3, c3 & 3), T);
  }
}

// mcevent/IF_Arm_Choice.f:21-56
template <class Pol>
__device__ __noinline__ bool arm_choice(Pol &pol, const P3 Zig[3], const P3 Arm[3]) {
  int CoorArm = 0, CoorZig = 0;
  for (int i = 0; i < 3; i++) {
    CoorArm += w_coor(D.state[cidx(Arm[i])]);
    CoorZig += w_coor(D.state[cidx(Zig[i])]);
  }
  if (CoorZig > CoorArm) return false;
  if (CoorZig == CoorArm) return !(pol.uniform() <= D.PZig);
  return true;
}

// ---- deferred re-rating -------------------------------------------------------------------------
// The reference re-rates a site (Get_Prob + UpdateTree / list migration) right where it touches it.
// Here every such request is queued by the transition function and executed afterwards, in the SAME
// order, by rq_flush: one Get_Prob call site for all events, so that SIMT lanes working on different
// tiles converge again for the expensive part, and the hot code stays small.  The order of list / tree
// operations (hence parity mode's bit-exactness) is unchanged; Get_Prob of a neighbour never reads
// what a later neighbour's update writes (neighbours of one site are not bonded to each other).
enum { A_NEW = 0, A_SETVOID, A_SETBULK, A_SET, A_VOID2MC, A_BULK2MC, A_ERASE, A_SETSELF /* the site that was just filled (deposition.f:237-250) */ };
struct RQItem { P3 N; long long c; P3 nn[4]; int Occ, NSiC, Cov, Coor, act, pre; uint32_t pre_w; };
struct RQ { int n; RQItem it[6]; };
__device__ __forceinline__ void rq_push(RQ &q, int act, P3 N, long long c, int Occ, int NSiC, int Cov, int Coor, const P3 nn[4], int pre = 0, uint32_t pre_w = 0) {
  RQItem &it = q.it[q.n++];
  it.N = N; it.c = c; it.Occ = Occ; it.NSiC = NSiC; it.Cov = Cov; it.Coor = Coor; it.act = act; it.pre = pre; it.pre_w = pre_w;
  it.nn[0] = nn[0]; it.nn[1] = nn[1]; it.nn[2] = nn[2]; it.nn[3] = nn[3];
}
template <class Pol>
__device__ __forceinline__ void rq_commit(Pol &pol, RQItem &it, int ev, double p) {
  if (it.pre) D.state[it.c] = it.pre_w;            // the site's own coordination is written after its neighbours were rated
  switch (it.act) {
    case A_NEW: if (ev != 0) pol.add_mc(it.N, it.c, it.nn, ev, p); break;
    case A_SETVOID: if (ev == 0) { pol.erase_mc(it.N, it.c); pol.add_void(it.N, it.c, it.nn); } else pol.set_rate(it.N, it.c, ev, p); break;
    // NumAdAtomOcc follows la/mcevent/deposition.f:219,246,253 and la/mcevent/Evaporation.f:47,192
    case A_SETBULK: if (ev == 0) { pol.erase_mc(it.N, it.c); pol.add_bulk(it.N, it.c, it.nn); pol.NumAdAtomOcc--; } else pol.set_rate(it.N, it.c, ev, p); break;
    case A_SETSELF: if (ev == 0) { pol.erase_mc(it.N, it.c); pol.add_bulk(it.N, it.c, it.nn); pol.NumAdAtomOcc--; } else { pol.set_rate(it.N, it.c, ev, p); pol.NumAdAtomOcc++; } break;
    case A_SET: pol.set_rate(it.N, it.c, ev, p); break;
    case A_VOID2MC: pol.erase_void(it.N, it.c); pol.add_mc(it.N, it.c, it.nn, ev, p); break;
    case A_BULK2MC: pol.erase_bulk(it.N, it.c); pol.add_mc(it.N, it.c, it.nn, ev, p); pol.NumAdAtomOcc++; break;
    default: pol.erase_mc(it.N, it.c); break;
  }
}
// serial flush, in request order (parity mode: the order of tree/list operations is the reference's)
template <class Pol>
__device__ __noinline__ void rq_flush(Pol &pol, RQ &q) {
#pragma unroll 1
  for (int k = 0; k < q.n; k++) {
    RQItem &it = q.it[k];
    if (it.pre) { D.state[it.c] = it.pre_w; it.pre = 0; }
    int ev = 0; double p = 0.0;
    if (it.act != A_ERASE) get_prob(pol, it.Occ, it.NSiC, it.Cov, it.Coor, it.N, it.nn, ev, p);
    rq_commit(pol, it, ev, p);
  }
  q.n = 0;
}

// mcevent/deposition.f:24-251
template <class Pol>
__device__ __noinline__ void deposition(Pol &pol, RQ &q, P3 Site, int IAtom) {
  long long cS = cidx(Site);
  uint32_t wS = D.state[cS];
  int Coor = w_coor(wS), CoorOld = Coor;
  wS = MVBITS(wS, 1, 1, PosOcc);
  wS = MVBITS(wS, IAtom, LenSiC, PosSiC);
  D.state[cS] = wS;
  P3 NextN[4];
  dec_bonds(Site, D.bonds[cS], NextN);
  if (Coor == 1) {                                       // :49-74 fix the other three bonds
    P3 SiteC = NextN[0];
    long long cC = cidx(SiteC);
    P3 cnn[4], SN[3], Zig[3], Arm[3];
    dec_bonds(SiteC, D.bonds[cC], cnn);
    int j = 0;
    SN[0] = SN[1] = SN[2] = zero3();
    for (int i = 0; i < 4; i++) if (!eq(cnn[i], Site)) { if (j < 3) SN[j] = cnn[i]; j++; }
    find_sn(Site, SiteC, SN, Zig, Arm);
    bool IF_Arm = arm_choice(pol, Zig, Arm);
    for (int i = 0; i < 3; i++) NextN[i + 1] = IF_Arm ? Arm[i] : Zig[i];
    D.bonds[cS] = enc_bonds(Site, NextN);
  }
  // the four neighbour words / bond codes are independent loads: issue them together
  long long cNs[4]; uint32_t wNs[4]; unsigned long long bNs[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { cNs[i] = cidx(NextN[i]); wNs[i] = D.state[cNs[i]]; bNs[i] = D.bonds[cNs[i]]; }
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    P3 N = NextN[i];
    long long cN = cNs[i];
    uint32_t wN = wNs[i];
    unsigned long long bN = bNs[i];
    int Occ = w_occ(wN), CoorNN = w_coor(wN), CovInd = w_cov(wN);
    P3 NextNN[4];
    if (Occ == 0) {
      if (CoorNN == 0) {                                 // :88-99 a new MC particle with one fixed bond
        CoorNN = 1;
        D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
        NextNN[0] = Site; NextNN[1] = NextNN[2] = NextNN[3] = zero3();
        rq_push(q, A_NEW, N, cN, 0, 0, CovInd, CoorNN, NextNN);
      } else if (CoorNN == 1) {                          // :100-160 second bond: complete the tetrahedron
        dec_bonds(N, bN, NextNN);
        int Dx = abs(NextNN[0].x - Site.x); Dx = min(Dx, D.LX - Dx);
        int Dy = abs(NextNN[0].y - Site.y); Dy = min(Dy, D.LY - Dy);
        int dz = NextNN[0].z - Site.z;
        if (Dx * Dx + Dy * Dy + dz * dz == 72) {
          P3 NewSite[2];
          CoorNN = 2;
          NextNN[1] = Site;
          find_nn(NextNN[0], NextNN[1], N, NewSite);
          for (int j = 2; j < 4; j++) {
            NextNN[j] = NewSite[j - 2];
            long long c3 = cidx(NextNN[j]);
            uint32_t w3 = D.state[c3];
            if (w_occ(w3) && w_coor(w3) <= 3 && w_cov(w3) != 113)
              if (recipr_b(NextNN[j], D.bonds[c3], N)) CoorNN++;
          }
          D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
          D.bonds[cN] = enc_bonds(N, NextNN);
          rq_push(q, A_SETVOID, N, cN, 0, 0, CovInd, CoorNN, NextNN);
        }
      } else if (CoorNN == 2 || CoorNN == 3) {           // :161-190
        if (recipr_b(N, bN, Site)) {
          dec_bonds(N, bN, NextNN);
          CoorNN++;
          D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
          rq_push(q, A_SETVOID, N, cN, 0, 0, CovInd, CoorNN, NextNN);
        }
      }
    } else {                                             // full neighbour, :191-234
      if (CovInd != 113) {
        if (CovInd != 0) { fail(E_NNCOV, N.x, N.y, N.z); return; }
        if (CoorNN <= 3) {
          if (recipr_b(N, bN, Site)) {
            dec_bonds(N, bN, NextNN);
            CoorNN++;
            D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
            rq_push(q, A_SETBULK, N, cN, 1, w_sp(wN), CovInd, CoorNN, NextNN);
            if (CoorOld == 1 && i >= 1) Coor++;
          }
        }
      } else Coor++;
    }
  }
  uint32_t wS2 = (Coor != CoorOld) ? MVBITS(wS, Coor, LenCoor, PosCoor) : wS;
  rq_push(q, A_SETSELF, Site, cS, 1, IAtom, 0, Coor, NextN, 1, wS2);
}

// Evaporation.f:95-122 / :229-258: keep only the occupied, under-coordinated, reciprocal neighbour and
// move it to slot 1.  (The reference leaves slot 1 to an uninitialised buffer when none qualifies; that
// never happens in any regression case -- here it raises E_STALE_NBUFF.)
__device__ __noinline__ void compact_to_single(P3 me, long long cme, const P3 old[4]) {
  P3 keep = zero3(); bool found = false;
  for (int j = 0; j < 4; j++) {
    long long c = cidx(old[j]);
    uint32_t w = D.state[c];
    if (w_occ(w) && w_coor(w) <= 3 && w_cov(w) != 113 && recipr_b(old[j], D.bonds[c], me)) { keep = old[j]; found = true; }
  }
  if (!found) fail(E_STALE_NBUFF, me.x, me.y, me.z);
  D.bonds[cme] = (unsigned long long)enc_bond(me, keep);
}

// mcevent/Evaporation.f:25-262
template <class Pol>
__device__ __noinline__ void evaporation(Pol &pol, RQ &q, P3 Site) {
  long long cS = cidx(Site);
  uint32_t wS = D.state[cS];
  int IAtom = w_sp(wS);
  if (IAtom >= 1 && IAtom <= 3) pol.CountCrystal[IAtom - 1]--;
  pol.NumAdAtomOcc--;
  int Coor = w_coor(wS);
  wS = MVBITS(wS, 0, 1, PosOcc);
  wS = MVBITS(wS, 0, LenSiC, PosSiC);
  D.state[cS] = wS;
  P3 NextN[4];
  dec_bonds(Site, D.bonds[cS], NextN);
  long long cNs[4]; uint32_t wNs[4]; unsigned long long bNs[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { cNs[i] = cidx(NextN[i]); wNs[i] = D.state[cNs[i]]; bNs[i] = D.bonds[cNs[i]]; }
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    P3 N = NextN[i];
    long long cN = cNs[i];
    uint32_t wN = wNs[i];
    int Occ = w_occ(wN), CoorNN = w_coor(wN), CovInd = w_cov(wN);
    unsigned long long bN = bNs[i];
    P3 NextNN[4];
    if (Occ == 0) {
      if (CoorNN < 1 || CoorNN > 4) continue;
      if (!recipr_b(N, bN, Site)) continue;
      dec_bonds(N, bN, NextNN);
      if (CoorNN == 1) {                                 // :62-79 orphaned empty site leaves the lattice
        D.state[cN] = D.keepT ? (wN & (((1u << LenT) - 1u) << PosT)) : 0u;
        rq_push(q, A_ERASE, N, cN, 0, 0, 0, 0, NextNN);
      } else if (CoorNN == 2) {                          // :80-128
        CoorNN = 1;
        D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
        compact_to_single(N, cN, NextNN);
        rq_push(q, A_SET, N, cN, 0, 0, CovInd, CoorNN, NextNN);     // NB: rated with the OLD neighbour list, as in :123
      } else if (CoorNN == 3) {                          // :129-145
        CoorNN = 2;
        D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
        rq_push(q, A_SET, N, cN, 0, 0, CovInd, CoorNN, NextNN);
      } else {                                           // :146-164 vacancy becomes active again
        CoorNN = 3;
        D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
        rq_push(q, A_VOID2MC, N, cN, 0, 0, CovInd, CoorNN, NextNN);
      }
    } else {
      if (CovInd != 113) {
        if (CovInd != 0) { fail(E_NNCOV, N.x, N.y, N.z); return; }
        if (!recipr_b(N, bN, Site)) continue;
        dec_bonds(N, bN, NextNN);
        if (CoorNN == 4) {                               // :173-190 bulk atom becomes active again
          CoorNN = 3;
          D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
          rq_push(q, A_BULK2MC, N, cN, 1, w_sp(wN), CovInd, CoorNN, NextNN);
        } else {                                         // :191-209
          CoorNN--;
          D.state[cN] = MVBITS(wN, CoorNN, LenCoor, PosCoor);
          rq_push(q, A_SET, N, cN, 1, w_sp(wN), CovInd, CoorNN, NextNN);
        }
      } else Coor--;
    }
  }
  uint32_t wS2 = MVBITS(wS, Coor, LenCoor, PosCoor);
  if (Coor == 0) rq_push(q, A_ERASE, Site, cS, 0, 0, 0, 0, NextN, 1, wS2);
  else {
    // the orphan case above zeroes a neighbour's word in the reference only AFTER EraseMC; nothing reads it in between
    if (Coor == 1) {                                     // :224-258; needs the final coordination words of the neighbours
      rq_push(q, A_SET, Site, cS, 0, 0, 0, Coor, NextN, 1, wS2);
      compact_to_single(Site, cS, NextN);
    } else rq_push(q, A_SET, Site, cS, 0, 0, 0, Coor, NextN, 1, wS2);
  }
}

// absorption.f:43-72 / desorption.f:54-111: re-rate an occupied reciprocal neighbour
__device__ __noinline__ void rerate_occ_nn(RQ &q, P3 Site, P3 N, bool need_le3) {
  long long cN = cidx(N);
  uint32_t wN = D.state[cN];
  if (!w_occ(wN)) return;
  int CovIndN = w_cov(wN), CoorN = w_coor(wN);
  if (CovIndN == 113) return;
  if (CoorN > 3) return;     // desorption.f reads the MC list for a bulk atom here (never hit in the goldens); treated as inert
  (void)need_le3;
  unsigned long long bN = D.bonds[cN];
  if (!recipr_b(N, bN, Site)) return;
  P3 NextNN[4]; dec_bonds(N, bN, NextNN);
  rq_push(q, A_SET, N, cN, 1, w_sp(wN), CovIndN, CoorN, NextNN);
}

// mcevent/absorption.f:25-78
template <class Pol>
__device__ __noinline__ void absorption(Pol &pol, RQ &q, P3 Site, int CovInd) {
  long long cS = cidx(Site);
  uint32_t wS = D.state[cS];
  int Coor = w_coor(wS);
  D.state[cS] = MVBITS(wS, CovInd, LenIndex, PosIndex);
  P3 NextN[4]; dec_bonds(Site, D.bonds[cS], NextN);
  for (int i = 0; i < 4; i++) rerate_occ_nn(q, Site, NextN[i], true);
  rq_push(q, A_SET, Site, cS, 0, 0, CovInd, Coor, NextN);
}

// mcevent/desorption.f:25-117
template <class Pol>
__device__ __noinline__ void desorption(Pol &pol, RQ &q, P3 Site) {
  long long cS = cidx(Site);
  uint32_t wS = D.state[cS];
  int CovInd = w_cov(wS);
  if (CovInd == D.ListCov[0]) pol.CountCov[0]--;
  else if (CovInd == D.ListCov[1]) pol.CountCov[1]--;
  else if (CovInd == D.ListCov[2]) pol.CountCov[2]--;
  int Coor = w_coor(wS);
  D.state[cS] = MVBITS(wS, 0, LenIndex, PosIndex);
  P3 NextN[4]; dec_bonds(Site, D.bonds[cS], NextN);
  if (Coor == 1) rerate_occ_nn(q, Site, NextN[0], false);
  else for (int i = 0; i < 4; i++) rerate_occ_nn(q, Site, NextN[i], false);
  rq_push(q, A_SET, Site, cS, 0, 0, 0, Coor, NextN);
}

// dispatch of main.f:335-357; the re-rate requests are left in q
template <class Pol>
__device__ __forceinline__ void apply_event_q(Pol &pol, RQ &q, P3 Site, int Itrans) {
  q.n = 0;
  if (Itrans <= 3) { deposition(pol, q, Site, Itrans); if (Itrans >= 1) pol.CountCrystal[Itrans - 1]++; }
  else if (Itrans == 4) evaporation(pol, q, Site);
  else if (Itrans == 5) desorption(pol, q, Site);
  else {
    int k = Itrans - 5;
    if (k > 3) { fail(E_BADEV, Site.x, Site.y, Site.z, Itrans); return; }
    absorption(pol, q, Site, D.ListCov[k - 1]);
    pol.CountCov[k - 1]++;
  }
}
template <class Pol>
__device__ __forceinline__ void apply_event(Pol &pol, P3 Site, int Itrans) {
  RQ q; q.n = 0;
  if (Itrans <= 3) { deposition(pol, q, Site, Itrans); if (Itrans >= 1) pol.CountCrystal[Itrans - 1]++; }
  else if (Itrans == 4) evaporation(pol, q, Site);
  else if (Itrans == 5) desorption(pol, q, Site);
  else {
    int k = Itrans - 5;
    if (k > 3) { fail(E_BADEV, Site.x, Site.y, Site.z, Itrans); return; }
    absorption(pol, q, Site, D.ListCov[k - 1]);
    pol.CountCov[k - 1]++;
  }
  rq_flush(pol, q);
}

}  // namespace mskc
