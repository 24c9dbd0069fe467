// mskc_coop.cuh -- warp-cooperative form of the four KMC transitions (sm_100a).
//
// A team of lanes owns one tile of the synchronous-sublattice sampler.  An event at site S touches S and its
// four bonded neighbours N0..N3 and re-rates those five sites.  The serial reference walks these one after the
// other (mcevent/deposition.f, Evaporation.f, absorption.f, desorption.f) and every re-rate is a fresh census of
// that site's own four neighbours (Get_Prob.f:216-291: 4 lattice words + 4 bond lists, i.e. 20 more cells per
// event).  Here
//   lanes 0..3  handle neighbour i  (phase A: coordination / bond / census-flag update of N_i, then its rate)
//   lane  4     handles S itself    (its bonds when they are fixed at deposition, its rate, its final word)
//   lanes 0..5  read the six zig-zag / arm-chair candidate cells of IF_Arm_Choice.f together
// and the census is INCREMENTAL: every cell carries, next to its list slot, two bits per bond that say which
// crystal species the occupied, reciprocal neighbour on that bond has (aux[] below).  An event updates the flag of
// the bond that points back at S in each neighbour it touches -- exactly where the reference changes that
// neighbour's coordination -- so the rate of a re-rated occupied site is a table look-up on its own record and
// the dependent chain of an event is  S -> {N_i}  (+ one more hop only where the reference itself reads further:
// IF_Arm_Choice candidates, the two new bonds of FIND_NN.f).  The census of S itself falls out of phase A (each lane
// holds its neighbour's word and bond list), so S's flags are recomputed, not trusted, at every event, and the
// reference's consistency stop `N_Si+N_C+N_B+N_wall.NE.Coor` (Get_Prob.f:293) is kept for every rated site.
// Runs with coverage species (FEAT & F_COV) still rate occupied sites with the full census (coop_rate): a
// coverage neighbour's contribution depends on ITS coordination (Get_Prob.f:262-291), which changes without the
// reference re-rating the sites around it; there the flags are maintained and cross-checked against the census.
//
// tests/test_replay_gpu.py replays the oracle's event trace through this code (one 32-lane team, and four 8-lane
// teams in lock step) and compares lattice words, bonds, census flags and (event class, rate) of every active site
// bit for bit.
//
// Geometry without floating point: bonds are stored as 15-bit relative codes (mskc_device.cuh), so
//   zig-zag continuation  (FIND_SN.f:60-64)  2*Cent - SN            ==  -(bond of C)          exact integers
//   arm-chair continuation(FIND_SN.f:66-70)  SN + 8.66*unit(S - C)  ==  bond of C + e + armoff(e)
// with armoff(e) = NINT(8.6602254037f * unit(-e)) tabulated on the host for all 31^3 bond codes by the same
// IEEE sequence (entries closer than 1e-6 to a rounding tie are left out and take the double-precision
// fallback find_sn()).  FIND_NN.f's rotation is exact in integers on the ideal bond shell (checked at start-up
// against the double-precision routine, D.nn_exact) and stays in double precision elsewhere.
#pragma once
#include "mskc_device.cuh"

namespace mskc {

enum { NEGC = 0x8420 };            // code(-v) = NEGC - code(v): every 5-bit field c -> 32 - c, no borrows (1 <= c <= 31)

__device__ __forceinline__ unsigned fld(unsigned long long b, int i) { return (unsigned)(b >> (16 * i)) & 0xFFFFu; }
__device__ __forceinline__ bool lists_code(unsigned long long b, unsigned code) {
  return (fld(b, 0) == code) | (fld(b, 1) == code) | (fld(b, 2) == code) | (fld(b, 3) == code);
}
// index of the (last) field of b that equals code, -1 if none
__device__ __forceinline__ int find_code(unsigned long long b, unsigned code) {
  int j = -1;
  j = fld(b, 0) == code ? 0 : j; j = fld(b, 1) == code ? 1 : j; j = fld(b, 2) == code ? 2 : j; j = fld(b, 3) == code ? 3 : j;
  return j;
}
__device__ __noinline__ bool recipr_cold(P3 n, unsigned long long bn, P3 me) { return recipr_b(n, bn, me); }
// "does cell n (bond word bn) list `me`?", given the field f under which me lists n (n = me + v(f)).
// Same outcome as recipr_b(); f == 0 ("not fixed" -> absolute origin) takes the general test.
__device__ __forceinline__ bool recip_f(P3 n, unsigned long long bn, P3 me, unsigned f) {
  if (f == 0u) return recipr_cold(n, bn, me);
  bool r = lists_code(bn, (unsigned)NEGC - f);
  if ((me.x | me.y | me.z) == 0) r |= lists_code(bn, 0u);
  return r;
}
__device__ __forceinline__ int cdx(unsigned f) { return (int)(f & 31u) - 16; }
__device__ __forceinline__ int cdy(unsigned f) { return (int)((f >> 5) & 31u) - 16; }
__device__ __forceinline__ int cdz(unsigned f) { return (int)((f >> 10) & 31u) - 16; }

__device__ __noinline__ unsigned enc_bond_cold(P3 s, P3 n) { return enc_bond(s, n); }
// bond code of `to` seen from `from` (minimum image), 0x10000 (matches no stored field) when it is out of the code's range
__device__ __noinline__ unsigned code_or_none(P3 from, P3 to) {
  int dx = to.x - from.x, dy = to.y - from.y, dz = to.z - from.z;
  if (dx > D.LX / 2) dx -= D.LX; else if (dx < -(D.LX / 2)) dx += D.LX;
  if (dy > D.LY / 2) dy -= D.LY; else if (dy < -(D.LY / 2)) dy += D.LY;
  if (dx < -15 || dx > 15 || dy < -15 || dy > 15 || dz < -15 || dz > 15) return 0x10000u;
  return (unsigned)(dx + 16) | ((unsigned)(dy + 16) << 5) | ((unsigned)(dz + 16) << 10);
}
__device__ __noinline__ void fail_cell(int code, uint32_t c) { P3 p = cell_xyz((long long)c); fail(code, p.x, p.y, p.z); }
// cell index of a site whose x,y are already wrapped (dec_bond output): only z can leave the box
__device__ __forceinline__ uint32_t ccell(P3 p) {
  if ((unsigned)p.z > (unsigned)D.LZ) { fail(E_OOB, p.x, p.y, p.z); return 0u; }
  return (uint32_t)p.x + (uint32_t)D.LX * ((uint32_t)p.y + (uint32_t)D.LY * (uint32_t)p.z);
}
// neighbour of site s along bond code f: coordinates (x, y wrapped, branch-free) and cell index.  f == 0 ("not
// fixed") is the absolute origin, as the reference's NextNXYZ = 0; z outside 0..LenZ raises E_OOB (cell 0 returned).
__device__ __forceinline__ uint32_t nbr(const P3 s, const unsigned f, P3 &n) {
  int x = s.x + cdx(f), y = s.y + cdy(f), z = s.z + cdz(f);
  x += (x < 0 ? D.LX : 0) - (x >= D.LX ? D.LX : 0);
  y += (y < 0 ? D.LY : 0) - (y >= D.LY ? D.LY : 0);
  if (f == 0u) { x = 0; y = 0; z = 0; }
  n.x = x; n.y = y; n.z = z;
  if ((unsigned)z > (unsigned)D.LZ) { fail(E_OOB, x, y, z); return 0u; }
  return (uint32_t)x + (uint32_t)D.LX * ((uint32_t)y + (uint32_t)D.LY * (uint32_t)z);
}
__device__ __forceinline__ int cov_index(int cov) {      // position of a coverage id in ListCov (1-based), 0 = absent
  return cov == D.ListCov[0] ? 1 : (cov == D.ListCov[1] ? 2 : (cov == D.ListCov[2] ? 3 : 0));
}

// ---- aux[cell]: list slot + census flags ---------------------------------------------------------------------
//   bits  0-15  1-based slot of the cell in its tile's active list (0 = not listed)
//   bits 16-23  census flags, 2 bits per bond i at 16+2i: 0 = the neighbour on bond i is not counted, 1..3 = crystal
//               species of the occupied neighbour that lists this cell back; a wall neighbour of an OCCUPIED site is 1
//               (Get_Prob.f:311 adds N_wall to the first index; an empty site's coordination does not count walls)
// The two halves are written with separate 16-bit / 8-bit stores, never read-modify-write.
enum { AUX_FLAG0 = 16 };
__device__ __forceinline__ int aux_slot(uint32_t a) { return (int)(a & 0xFFFFu); }
__device__ __forceinline__ unsigned aux_flags(uint32_t a) { return (a >> AUX_FLAG0) & 0xFFu; }
__device__ __forceinline__ unsigned flag_get(unsigned fl, int j) { return (fl >> (2 * j)) & 3u; }
__device__ __forceinline__ unsigned flag_set(unsigned fl, int j, unsigned v) { return (fl & ~(3u << (2 * j))) | (v << (2 * j)); }
__device__ __forceinline__ void st_slot(uint32_t c, int slot) { reinterpret_cast<unsigned short *>(D.aux)[2 * (size_t)c] = (unsigned short)slot; }
__device__ __forceinline__ void st_flags(uint32_t c, unsigned fl) { reinterpret_cast<unsigned char *>(D.aux)[4 * (size_t)c + 2] = (unsigned char)fl; }
// species counts of a flag byte: n1 | n2 << 4 | n3 << 8
__device__ __forceinline__ unsigned flag_counts(unsigned fl) {
  const unsigned lo = fl & 0x55u, hi = (fl >> 1) & 0x55u;
  return (unsigned)__popc(lo & ~hi) | ((unsigned)__popc(hi & ~lo) << 4) | ((unsigned)__popc(lo & hi) << 8);
}
// census flags of one cell from scratch (the dense pass after an upload, and the checker): Get_Prob.f:216-291 without
// the coverage part.  Q's own word decides whether walls count.
__device__ __forceinline__ unsigned census_flags(const P3 Q, const uint32_t wQ, const unsigned long long bQ) {
  unsigned fl = 0u;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const unsigned f = fld(bQ, i);
    if (f != 0u && (unsigned)(Q.z + cdz(f)) > (unsigned)D.LZ) continue;   // bond leaving the box (bottom layer of a slab without walls)
    P3 n; const uint32_t c = nbr(Q, f, n);               // f == 0: the absolute origin, like the reference's NextNXYZ = 0
    const uint32_t w = D.state[c];
    if (!w_occ(w)) continue;
    if (w_cov(w) == 113) { if (w_occ(wQ)) fl |= 1u << (2 * i); continue; }
    if (recip_f(n, D.bonds[c], Q, f)) fl |= (unsigned)(w_sp(w) & 3) << (2 * i);
  }
  return fl;
}

// ---- rate of one site: mcevent/Get_Prob.f:54-331 ---------------------------------------------------
// Occ/NSiC/Cov/Coor are what the caller holds (not re-read), rb the four bonds to rate with (they can
// differ from the stored ones, Evaporation.f:123).  Deposition/adsorption sites get EV_DRAW: the species
// is drawn when the event fires (Get_Prob.f:160-179 draws it here; same distribution).
// FEAT: bit 0 = coverage species exist (NCov > 0), bit 1 = per-kelvin rate tables (LA rate law or 'ST' field); compiled out otherwise (hot-loop code size)
enum { F_COV = 1, F_LA = 2, F_PAR = 4 };
// F_PAR: the event is applied for the serial-parity loop (mskc_parity_coop.cuh): no census flags / list slots are touched
// (aux[] may not exist), rates come from the full census with the reference's wrong-list quirk (below), and instead of the
// tile-list commit every lane hands its request back to the caller, which commits them in the reference's order.
struct CoopReq {
  int act, ev, Coor, self, needs_root, arm_used; double p; uint32_t c, w;
  int Occ, Sp, Cov; P3 N; unsigned long long rb;       // what the site was rated with (the caller may have to rate it again, see mskc_parity_coop.cuh)
};
// Get_Prob.f:266-269 looks a coverage neighbour on a 4-coordinated empty site up in ListAdAtom at the vacancy's index
// (defined by the parity kernel, which owns the lists)
__device__ bool par_cov_void_recipr(P3 n, uint32_t cn, P3 me);
// full-census form (coverage runs, the dense re-rate pass): reads the four neighbours' words and bond lists
template <int FEAT>
__device__ __forceinline__ void coop_rate(int Occ, int NSiC, int Cov, int Coor, P3 Q, uint32_t cQ, unsigned long long rb, double root, int &ev, double &p, unsigned *flags_out = nullptr) {
  ev = 0; p = 0.0;
  int T = 0;
  if ((FEAT & F_LA) && D.la) T = site_T(D.state[cQ], Q);                   // :76-91 ('ST') / :127-141 ('LA')
  if ((FEAT & F_COV) && Cov > 0 && Cov < 113) {                            // desorption, :93-123
    if (Occ != 0) { fail_cell(E_COVOCC, cQ); return; }
    if (Coor <= 3) {
      const int iCov = cov_index(Cov);
      if (iCov == 0 || iCov > D.NCov) { fail_cell(E_COVLIST, cQ); return; }
      ev = 5; p = (FEAT & F_LA) ? rate_Des(iCov, Coor, T) : __ldg(&D.PDes[(iCov - 1) + 3 * Coor]);
    }
    return;
  }
  if (Occ == 0) {
    if (Coor >= 1 && Coor <= 3) {
      ev = EV_DRAW;
      if (Coor == 1 && w_coor(D.state[ccell(dec_bond(Q, fld(rb, 0)))]) < 2) { ev = 1; p = (double)1.e-8f; }   // Get_Prob.f:183
      else p = (FEAT & F_LA) ? total_dep_abs(Coor, T) : D.PDepAbs[Coor];
    }
    return;
  }
  // occupied: census of the four neighbours with reciprocity, :213-291
  uint32_t wn[4]; unsigned long long bn[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { const uint32_t c = ccell(dec_bond(Q, fld(rb, i))); wn[i] = D.state[c]; bn[i] = D.bonds[c]; }
  const bool qorigin = (Q.x | Q.y | Q.z) == 0;
  int N_Si = 0, N_C = 0, N_B = 0, N_wall = 0, cv0 = 0, cv1 = 0, cv2 = 0;
  unsigned fl = 0u;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t w = wn[i];
    const unsigned f = fld(rb, i);
    const int CovN = w_cov(w);
    bool rec = lists_code(bn[i], (unsigned)NEGC - f);
    if (f == 0u) rec = recipr_cold(dec_bond(Q, 0u), bn[i], Q);
    if (qorigin) rec |= lists_code(bn[i], 0u);
    if (w_occ(w)) {
      if (CovN == 113) { N_wall++; fl |= 1u << (2 * i); }
      else if (rec) { const int s = w_sp(w); N_Si += (s == 1); N_C += (s == 2); N_B += (s == 3); fl |= (unsigned)(s & 3) << (2 * i); }
    } else if ((FEAT & F_COV) && CovN != 0 && (((FEAT & F_PAR) && w_coor(w) >= 4) ? par_cov_void_recipr(dec_bond(Q, f), ccell(dec_bond(Q, f)), Q) : (rec && w_coor(w) < 4))) {
      // (coverage on a 4-coordinated empty site: the reference looks it up in the wrong list -- the serial-parity loop
      // reproduces that lookup, the sampler in effect never counts such a neighbour)
      const int iCov = cov_index(CovN);
      if (iCov == 0 || iCov > D.NCov) { fail_cell(E_COVLIST, cQ); return; }
      cv0 += (iCov == 1); cv1 += (iCov == 2); cv2 += (iCov == 3);
    }
  }
  if (flags_out) *flags_out = fl;
  const int ncry = N_Si + N_C + N_B + N_wall;
  if (ncry != Coor) { fail(E_CENSUS, Q.x, Q.y, Q.z, Coor * 16 + ncry); return; }
  if (ncry > 3) return;
  ev = 4;
  if (ncry == 0) { p = (FEAT & F_PAR) ? -1.0 : (double)100000.f * root; return; }     // (F_PAR: the caller multiplies Tree(1) as of that moment)
  const int idx = IDXE(NSiC, N_Si + N_wall, N_C, N_B, cv0 & 3, cv1 & 3, cv2 & 3);
  p = (FEAT & F_LA) ? rate_E(idx, T) : __ldg(&D.PE[idx]);
}
// incremental form (no coverage species): the census is the site's flag byte; nb1 = coordination of the site's
// first-bond neighbour as the lattice holds it when the reference would rate the site (Get_Prob.f:182-183)
template <int FEAT>
__device__ __forceinline__ void flag_rate(int Occ, int NSiC, int Coor, uint32_t wQ, uint32_t cQ, unsigned fl, int nb1, double root, int &ev, double &p) {
  ev = 0; p = 0.0;
  int T = 0;
  if ((FEAT & F_LA) && D.la) T = site_T(wQ, cell_xyz((long long)cQ));      // :127-141
  if (Occ == 0) {
    if (Coor >= 1 && Coor <= 3) {
      ev = EV_DRAW;
      if (Coor == 1 && nb1 < 2) { ev = 1; p = (double)1.e-8f; }            // Get_Prob.f:183
      else p = (FEAT & F_LA) ? total_dep_abs(Coor, T) : D.PDepAbs[Coor];
    }
    return;
  }
  const unsigned cn = flag_counts(fl);
  const int n1 = (int)(cn & 15u), n2 = (int)((cn >> 4) & 15u), n3 = (int)(cn >> 8);
  const int ncry = n1 + n2 + n3;
  if (ncry != Coor) { P3 q = cell_xyz((long long)cQ); fail(E_CENSUS, q.x, q.y, q.z, Coor * 16 + ncry); return; }   // :293
  if (ncry > 3) return;
  ev = 4;
  if (ncry == 0) { p = (double)100000.f * root; return; }
  const int idx = IDXE(NSiC, n1, n2, n3, 0, 0, 0);
  p = (FEAT & F_LA) ? rate_E(idx, T) : __ldg(&D.PE[idx]);
}

// Evaporation.f:95-122 on bond codes, from the lattice itself (serial-parity path, which keeps no census flags): keep the
// occupied, under-coordinated, reciprocal neighbour (the last one that qualifies), forget the others.
__device__ __noinline__ unsigned long long coop_compact(P3 me, unsigned long long old) {
  unsigned keep = 0u; bool found = false;
  uint32_t w[4]; unsigned long long b[4]; P3 n[4];
#pragma unroll
  for (int j = 0; j < 4; j++) { const uint32_t c = nbr(me, fld(old, j), n[j]); w[j] = D.state[c]; b[j] = D.bonds[c]; }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned f = fld(old, j);
    if (w_occ(w[j]) && w_coor(w[j]) <= 3 && w_cov(w[j]) != 113 && recip_f(n[j], b[j], me, f)) { keep = f; found = true; }
  }
  if (!found) fail(E_STALE_NBUFF, me.x, me.y, me.z);
  return (unsigned long long)keep;
}

// double-precision fallback of the bond fixing (cold): serial find_sn on absolute coordinates
__device__ __noinline__ void slow_find_sn(P3 S, P3 C, const unsigned g[3], unsigned zig[3], unsigned arm[3]) {
  P3 SN[3], Z[3], A[3];
  for (int i = 0; i < 3; i++) SN[i] = dec_bond(C, g[i]);
  find_sn(S, C, SN, Z, A);
  for (int i = 0; i < 3; i++) { zig[i] = enc_bond_cold(S, Z[i]); arm[i] = enc_bond_cold(S, A[i]); }
}
// deposition.f:100-142 off the ideal shell (cold): the two new bonds of FIND_NN.f by the double-precision rotation
__device__ __noinline__ void slow_find_nn(P3 S, P3 N, unsigned f0, unsigned fn[2]) {
  P3 NewSite[2];
  find_nn(dec_bond(N, f0), S, N, NewSite);
  fn[0] = enc_bond_cold(N, NewSite[0]); fn[1] = enc_bond_cold(N, NewSite[1]);
}

// ---- what an event does to ONE neighbour N of its site S: the rules of deposition.f:78-235, Evaporation.f:51-215,
// absorption.f:43-72 and desorption.f:54-111 as a table over (event kind, N occupied, N wall, N lists S, coordination of N).
// Entry layout:
enum {
  PA_ACT = 15,        // bits 0-3: request + 1 (A_NEW ... of mskc_device.cuh), 0 = N is not re-rated
  PA_COOR = 4,        // bits 4-5: coordination of N: 0 unchanged, 1 +1, 2 -1, 3 := 1
  PA_SAME = 1 << 6,   // re-rate at the unchanged coordination (absorption / desorption)
  PA_SUP = 1 << 7,    // N raises S's coordination (a wall next to a deposited atom, deposition.f:231-232)
  PA_SUPFIX = 1 << 8, // ... only when S has just fixed this bond (S was 1-coordinated, bonds 2..4, deposition.f:218-220)
  PA_SDOWN = 1 << 9,  // N lowers S's coordination (a wall next to an evaporated atom, Evaporation.f:211-213)
  PA_KEEP = 1 << 10,  // N may be the one neighbour S keeps (Evaporation.f:229-258)
  PA_SFLAG = 11,      // bits 11-12: N in S's census afterwards: 0 not counted, 1 as a wall, 2 with its species
  PA_FLAG = 13,       // bits 13-15: N's census flags: 0 unchanged, 1 the flag for S := deposited species, 2 the flag for S := 0,
                      //             3 all := (deposited species on bond 1), 4 all := 0
  PA_SPECIAL = 16,    // bits 16-18: PA_NEWSITE ... below
  PA_NEWSITE = 1, PA_SECOND = 2, PA_ORPHAN = 3, PA_COMPACT = 4,
  PA_ENTRIES = 192
};
__host__ __device__ __forceinline__ int pa_index(int kind, int Occ, int wall, int rec, int Coor) {
  return kind | (Occ << 2) | ((wall ? 1 : 0) << 3) | ((rec ? 1 : 0) << 4) | ((Coor > 5 ? 5 : Coor) << 5);
}
// the rules, stated once (host side: mskc_create fills the table with this)
static inline uint32_t pa_rule(int kind, int Occ, int wall, int rec, int C) {
  const uint32_t INC = 1u << PA_COOR, DEC = 2u << PA_COOR, ONE = 3u << PA_COOR;
  const uint32_t F_SET = 1u << PA_FLAG, F_CLR = 2u << PA_FLAG, F_NEW = 3u << PA_FLAG, F_ZERO = 4u << PA_FLAG;
  const uint32_t S_WALL = 1u << PA_SFLAG, S_SPEC = 2u << PA_SFLAG;
  auto ACT = [](int a) { return (uint32_t)(a + 1); };
  if (kind == 0) {                                           // deposition at S (deposition.f:78-235)
    if (Occ && wall) return PA_SUP | S_WALL;                                                   // :231-232
    if (Occ) {                                                                                 // :191-229 full neighbour
      uint32_t e = rec ? S_SPEC : 0u;
      if (C <= 3 && rec) e |= ACT(A_SETBULK) | INC | PA_SUPFIX | F_SET;
      return e;
    }
    if (C == 0) return ACT(A_NEW) | ONE | F_NEW | ((uint32_t)PA_NEWSITE << PA_SPECIAL);        // :88-99
    if (C == 1) return (uint32_t)PA_SECOND << PA_SPECIAL;                                      // :100-160
    if (C <= 3 && rec) return ACT(A_SETVOID) | INC | F_SET;                                    // :161-190
    return 0u;
  }
  if (kind == 1) {                                           // evaporation of S (Evaporation.f:51-215)
    if (Occ && wall) return PA_SDOWN;                                                          // :211-213
    if (!rec) return 0u;
    if (Occ) return ACT(C == 4 ? A_BULK2MC : A_SET) | DEC | PA_KEEP | S_SPEC | F_CLR;           // :173-209
    if (C == 1) return ACT(A_ERASE) | F_ZERO | ((uint32_t)PA_ORPHAN << PA_SPECIAL);            // :62-79
    if (C == 2) return ACT(A_SET) | ONE | ((uint32_t)PA_COMPACT << PA_SPECIAL);                // :80-128
    if (C == 3) return ACT(A_SET) | DEC | F_CLR;                                               // :129-145
    if (C == 4) return ACT(A_VOID2MC) | DEC | F_CLR;                                           // :146-164
    return 0u;
  }
  // adsorption on / desorption from S (absorption.f:43-72, desorption.f:54-111): the occupied reciprocal neighbours are re-rated
  if (Occ && !wall && C <= 3 && rec) return ACT(A_SET) | PA_SAME;
  return 0u;
}

struct CoopCounters {                // per-lane partial sums, reduced over the warp at the end of a step
  int nsu, dAtoms, dVoids, dCrystal[3], dCov[3];
};

// ---- one event per TEAM of lanes (8 or 32; a power of two), all teams of the warp in lock step --------------
// EVERY lane of the warp calls this together (full-mask votes, shuffles and __syncwarp() are the convergence
// points): tl = lane index inside its team, go = "this team has an event" (a team without one only takes part in
// the synchronisation).  Loads are issued in code all teams share, the event-kind specific parts are short
// predicated stretches between two convergence points, so the memory round trips of the teams overlap.
// Itrans: 1..3 deposit species, 4 evaporate, 5 desorb, 6.. adsorb (main.f:335-357).
// arm_u: uniform used when IF_Arm_Choice.f ties; forced_arm >= 0 overrides the choice (trace replay).
// the species / adsorbate of a deposition event, drawn when it fires (Get_Prob.f:160-179 pre-draws it; same distribution):
// plain tables come from the constant bank, per-kelvin tables take the general routine
template <int FEAT>
__device__ __forceinline__ int draw_dep_abs(int Coor, uint32_t wS, uint32_t cS, double u) {
  if ((FEAT & F_LA) && D.la) return choose_dep_abs(Coor, site_T(wS, cell_xyz((long long)cS)), u, nullptr);
  const int c = Coor < 1 ? 1 : (Coor > 3 ? 3 : Coor);
  const double PD1 = D.cPD[3 * (c - 1)], PD2 = D.cPD[3 * (c - 1) + 1], PD3 = D.cPD[3 * (c - 1) + 2];
  const double PDep = PD1 + PD2 + PD3;
  const double Pcurr = u * D.PDepAbs[c];
  if (!(FEAT & F_COV) || Pcurr <= PDep) return Pcurr <= PD1 ? 1 : (Pcurr > (PD1 + PD2) ? 3 : 2);
  return index_event_abs(c, Pcurr, PDep, 0);
}

template <int TEAM, int FEAT, class Lists>
__device__ __forceinline__ int coop_event(Lists &L, const int tl, bool go, const P3 S, const uint32_t cS, int Itrans, const double root,
                                          const double arm_u, const int forced_arm, CoopCounters &K, const double species_u = 0.5, CoopReq *rq = nullptr) {
  enum { DEP = 0, EVAP = 1, DES = 2, ABS = 3 };
  constexpr unsigned FULL = 0xFFFFFFFFu;
  constexpr unsigned tm = TEAM == 32 ? 0xFFFFFFFFu : ((1u << (TEAM & 31)) - 1u);
  const int tshift = (int)(threadIdx.x & 31u) - tl;         // first lane of the team inside the warp
  // ---- S's record (team-uniform: the lanes of a team load the same three words) ------------------------------
  uint32_t wS = 0u, aS = 0u; unsigned long long bS = 0ull;
  if (go) { wS = D.state[cS]; bS = D.bonds[cS]; if (!(FEAT & F_PAR)) aS = D.aux[cS]; }
  const int c0 = w_coor(wS);
  if (Itrans == EV_DRAW) Itrans = go ? draw_dep_abs<FEAT>(c0, wS, cS, species_u) : 4;
  const int kind = Itrans <= 3 ? DEP : (Itrans == 4 ? EVAP : (Itrans == 5 ? DES : ABS));
  if (!(FEAT & F_COV) && go && kind >= DES) { fail_cell(E_BADEV, cS); go = false; }
  if ((FEAT & F_COV) && go && kind == ABS && Itrans - 5 > 3) { fail_cell(E_BADEV, cS); go = false; }
  if (go && kind == DEP && Itrans < 1) { fail_cell(E_BADEV, cS); go = false; }
  int CovNew = 0;
  {                                                        // S's word: occupancy / species / coverage (coordination follows at the end)
    int dsp = -1, dcv = -1, sgn = 1;
    if (kind == DEP) { wS = (wS | (1u << PosOcc)) & ~(7u << PosSiC) | ((uint32_t)Itrans << PosSiC); dsp = Itrans - 1; }
    else if (kind == EVAP) { dsp = w_sp(wS) - 1; sgn = -1; wS &= ~((1u << PosOcc) | (7u << PosSiC)); }
    else if (FEAT & F_COV) {
      if (kind == DES) { dcv = cov_index(w_cov(wS)) - 1; sgn = -1; wS &= ~255u; }
      else { const int k = (Itrans - 5) & 3; CovNew = D.ListCov[k ? k - 1 : 0]; dcv = k - 1; wS = (wS & ~255u) | (uint32_t)CovNew; }
    }
    if (go && tl == 4) {
      K.dCrystal[0] += dsp == 0 ? sgn : 0; K.dCrystal[1] += dsp == 1 ? sgn : 0; K.dCrystal[2] += dsp == 2 ? sgn : 0;
      if (FEAT & F_COV) { K.dCov[0] += dcv == 0 ? sgn : 0; K.dCov[1] += dcv == 1 ? sgn : 0; K.dCov[2] += dcv == 2 ? sgn : 0; }
      if (FEAT & F_COV) D.state[cS] = wS;                  // the census of a neighbour looks back at S (coverage runs only)
    }
  }
  if (FEAT & F_COV) __syncwarp();                          // S's new occupancy is visible before any lane looks back at it

  // ---- deposition.f:49-74: a 1-coordinated site fixes bonds 2..4 (cubic or hexagonal continuation) -------------
  const bool fix = go && kind == DEP && c0 == 1;
  if (__any_sync(FULL, fix)) {
    unsigned zig[3] = {0u, 0u, 0u}, arm[3] = {0u, 0u, 0u};
    const unsigned e = fld(bS, 0);
    int cc = 0;
    if (fix) {
      P3 C; const uint32_t cC = nbr(S, e, C);
      const unsigned long long bC = D.bonds[cC];
      const unsigned back = (unsigned)NEGC - e;
      unsigned g[3] = {0u, 0u, 0u}; int j = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { const unsigned f = fld(bC, i); if (f != back) { if (j < 3) g[j] = f; j++; } }
      const unsigned ao = (unsigned)__ldg(&D.arm_tab[e]);
      bool slow = (ao == 0u) | (g[0] == 0u) | (g[1] == 0u) | (g[2] == 0u);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        zig[i] = (unsigned)NEGC - g[i];
        // componentwise g + e + ao on the biased fields: (a+16)+(b+16)+(c+16) = (a+b+c+16)+32
        const unsigned sx = (g[i] & 31u) + (e & 31u) + (ao & 31u), sy = ((g[i] >> 5) & 31u) + ((e >> 5) & 31u) + ((ao >> 5) & 31u), sz = (g[i] >> 10) + (e >> 10) + (ao >> 10);
        slow |= (sx - 33u > 30u) | (sy - 33u > 30u) | (sz - 33u > 30u);
        arm[i] = (sx - 32u) | ((sy - 32u) << 5) | ((sz - 32u) << 10);
      }
      if (slow) slow_find_sn(S, C, g, zig, arm);
      // IF_Arm_Choice.f:38-54: lanes 0..2 read the arm-chair cells, 3..5 the zig-zag cells
      if (tl < 6) {
        const int k = tl < 3 ? tl : tl - 3;
        const unsigned fa = k == 0 ? arm[0] : (k == 1 ? arm[1] : arm[2]), fz = k == 0 ? zig[0] : (k == 1 ? zig[1] : zig[2]);
        P3 X; cc = w_coor(D.state[nbr(S, tl < 3 ? fa : fz, X)]);
      }
    }
    const int CoorArm = __shfl_sync(FULL, cc, 0, TEAM) + __shfl_sync(FULL, cc, 1, TEAM) + __shfl_sync(FULL, cc, 2, TEAM);
    const int CoorZig = __shfl_sync(FULL, cc, 3, TEAM) + __shfl_sync(FULL, cc, 4, TEAM) + __shfl_sync(FULL, cc, 5, TEAM);
    if (fix) {
      bool IF_Arm = CoorZig < CoorArm || (CoorZig == CoorArm && !(arm_u <= D.PZig));
      if ((FEAT & F_PAR) && rq) rq->arm_used = CoorZig == CoorArm;          // IF_Arm_Choice.f:46-50 drew a uniform
      if (forced_arm >= 0) IF_Arm = forced_arm != 0;
      bS = (unsigned long long)e;
#pragma unroll
      for (int i = 0; i < 3; i++) bS |= (unsigned long long)(IF_Arm ? arm[i] : zig[i]) << (16 * (i + 1));
      if (tl == 4) D.bonds[cS] = bS;
    }
  }

  // ---- phase A: lane i < 4 updates neighbour i ---------------------------------------------------
  // the request this lane will rate and commit (lanes 0..3: neighbour i, lane 4: S)
  int act = -1, qOcc = 0, qSp = 0, qCov = 0, qCoor = 0, nb1 = 4;
  P3 qN = S; uint32_t qc = cS, qw = wS, qaux = aS; unsigned long long qrb = bS;
  unsigned qfl = aux_flags(aS), qfl0 = qfl;               // census flags of the request's site: new, and as stored
  unsigned sflag = 0u;                                    // what this neighbour contributes to S's census after the event
  bool sUp = false, sDown = false;                        // this neighbour raises / lowers S's own coordination
  bool keepable = false;                                  // Evaporation.f:229-258: occupied, under-coordinated, reciprocal neighbour of S
  int fcoor = 0;                                          // coordination of this neighbour as the lattice holds it after phase A
  int xmode = 0;                                          // 1: second bond (two more cells), 2: coordination of the old first neighbour (one cell)
  unsigned xf0 = 0u, xf1 = 0u, xfb = 0u; uint32_t wN0 = 0u;
  if (go && tl < 4 && !(kind == DES && c0 == 1 && tl > 0)) {    // desorption.f:54-81: a 1-coordinated site looks at bond 1 only
    const unsigned f = fld(bS, tl);
    P3 N; const uint32_t cN = nbr(S, f, N);
    const uint32_t wN = D.state[cN], aN = (FEAT & F_PAR) ? 0u : D.aux[cN];
    unsigned long long bN = D.bonds[cN];
    const int Occ = w_occ(wN), CovN = w_cov(wN), CoorNN = w_coor(wN), spN = w_sp(wN) & 3;
    const bool rec = recip_f(N, bN, S, f);
    const bool wall = Occ && CovN == 113;
    // the bond N -> S, and where N lists S (its census flag for S lives there).  Bonds to / from the absolute origin are
    // stored as 0 like the reference's NextNXYZ = 0, hence the two special cases
    const unsigned fb = f ? (unsigned)NEGC - f : code_or_none(N, S);
    const int jS = find_code(bN, (S.x | S.y | S.z) ? fb : 0u);
    unsigned flN = aux_flags(aN);
    qN = N; qc = cN; qCov = CovN; qOcc = Occ; qSp = Occ ? w_sp(wN) : 0; qaux = aN; qfl0 = flN; wN0 = wN;
    const unsigned long long bN0 = bN;
    unsigned long long rbN = bN;                           // the bond list the site is rated with (Evaporation.f:123: the OLD one)
    // what happens to this neighbour is a pure function of (event kind, its occupancy, wall, reciprocity, coordination):
    // one table look-up (pa_rule() below states the reference's rules) instead of a branch per case, so that the
    // sixteen neighbour lanes of a warp -- four events of any kinds -- run one instruction stream
    unsigned e = __ldg(&D.pa_tab[pa_index(kind, Occ, wall, rec, CoorNN)]);
    if (Occ && !wall && CovN != 0 && kind <= EVAP) { fail_cell(E_NNCOV, cN); e = 0u; }    // deposition.f:194-198
    act = (int)(e & PA_ACT) - 1;
    const unsigned cop = (e >> PA_COOR) & 3u;
    const int newCoor = cop == 1u ? CoorNN + 1 : (cop == 2u ? CoorNN - 1 : (cop == 3u ? 1 : ((e & PA_SAME) ? CoorNN : -1)));
    sUp = (e & PA_SUP) || ((e & PA_SUPFIX) && c0 == 1 && tl >= 1);
    sDown = (e & PA_SDOWN) != 0u; keepable = (e & PA_KEEP) != 0u;
    const unsigned sm = (e >> PA_SFLAG) & 3u;
    sflag = sm == 1u ? 1u : (sm == 2u ? (unsigned)spN : 0u);
    const unsigned fop = (e >> PA_FLAG) & 7u, spc = (e >> PA_SPECIAL) & 7u;
    const unsigned jmask = jS >= 0 ? 3u << (2 * jS) : 0u, jval = jS >= 0 ? (unsigned)Itrans << (2 * jS) : 0u;
    flN = fop == 1u ? ((flN & ~jmask) | jval) : (fop == 2u ? (flN & ~jmask) : (fop == 3u ? (unsigned)Itrans : (fop == 4u ? 0u : flN)));
    uint32_t wNew = newCoor >= 0 ? MVBITS(wN, newCoor, LenCoor, PosCoor) : wN;
    if (newCoor >= 0) qCoor = newCoor;
    if (spc == PA_NEWSITE) {                               // deposition.f:88-99: a new active empty site with one fixed bond
      if (fb > 0xFFFFu) fail(E_FARBOND, S.x, S.y, S.z);
      bN = (unsigned long long)(fb & 0xFFFFu); rbN = bN;
      nb1 = c0;                                            // rated against S's OLD coordination (deposition.f:237-250 writes it last)
    } else if (spc == PA_ORPHAN) {                         // Evaporation.f:62-79: an orphaned empty site leaves the lattice
      wNew = D.keepT ? (wN & (((1u << LenT) - 1u) << PosT)) : 0u;
      bN = 0ull; qCov = 0;
    } else if ((FEAT & F_PAR) && spc == PA_COMPACT) {      // (no census flags on the serial-parity path: look at the four neighbours)
      bN = coop_compact(N, bN);
    } else if (spc == PA_COMPACT) {                        // Evaporation.f:80-128: back to one bond: keep the occupied reciprocal neighbour that is left
      const unsigned others = flN & ~jmask;
      int k = -1;
#pragma unroll
      for (int j = 0; j < 4; j++) if (flag_get(others, j)) k = j;
      if (k < 0) { fail(E_STALE_NBUFF, N.x, N.y, N.z); k = 0; }
      flN = flag_get(others, k);
      // (rated with the OLD neighbour list, Evaporation.f:123: what matters is the coordination of the old first neighbour)
      const unsigned f0 = fld(bN, 0);
      if (f0 == fb) nb1 = c0;                              // it is S, whose word still holds its old coordination
      else { xmode = 2; xf0 = f0; }
      bN = (unsigned long long)fld(bN, k);
    } else if (spc == PA_SECOND) {                         // deposition.f:100-160: second bond: complete the tetrahedron (two more cells, below)
      const unsigned f0 = fld(bN, 0);
      const int dx = cdx(f0) + cdx(f), dy = cdy(f0) + cdy(f), dz = cdz(f0) + cdz(f);      // M - S = (M - N) + (N - S)
      int Dx = abs(dx); Dx = min(Dx, D.LX - Dx);
      int Dy = abs(dy); Dy = min(Dy, D.LY - Dy);
      if (f0 != 0u && f != 0u && Dx * Dx + Dy * Dy + dz * dz == 72) {
        // FIND_NN.f:46-76 rotates b2 = S-N by +-120 degrees about b1 = M-N.  For two bonds of the ideal shell (|b|^2 = 27,
        // tetrahedral pair) the rotation is exact in integers: -(b1+b2)/2 -+ (b1 x b2)/6 (D.nn_exact, checked at start-up)
        const int b1x = cdx(f0), b1y = cdy(f0), b1z = cdz(f0), b2x = cdx(fb), b2y = cdy(fb), b2z = cdz(fb);
        unsigned fn[2];
        if (D.nn_exact && b1x * b1x + b1y * b1y + b1z * b1z == 27 && b2x * b2x + b2y * b2y + b2z * b2z == 27) {
          const int cx = b1y * b2z - b1z * b2y, cy = b1z * b2x - b1x * b2z, cz = b1x * b2y - b1y * b2x;
          const int sx = -(b1x + b2x) / 2, sy = -(b1y + b2y) / 2, sz = -(b1z + b2z) / 2;
          fn[0] = (unsigned)(sx - cx / 6 + 16) | ((unsigned)(sy - cy / 6 + 16) << 5) | ((unsigned)(sz - cz / 6 + 16) << 10);
          fn[1] = (unsigned)(sx + cx / 6 + 16) | ((unsigned)(sy + cy / 6 + 16) << 5) | ((unsigned)(sz + cz / 6 + 16) << 10);
        } else slow_find_nn(S, N, f0, fn);
        xmode = 1; xf0 = fn[0]; xf1 = fn[1]; xfb = fb;
      } else if (f0 == 0u || f == 0u) {                    // unfixed first bond: the general distance test on coordinates
        const P3 M = dec_bond(N, f0);
        int Ex = abs(M.x - S.x); Ex = min(Ex, D.LX - Ex);
        int Ey = abs(M.y - S.y); Ey = min(Ey, D.LY - Ey);
        const int ez = M.z - S.z;
        if (Ex * Ex + Ey * Ey + ez * ez == 72) { unsigned fn[2]; slow_find_nn(S, N, f0, fn); xmode = 1; xf0 = fn[0]; xf1 = fn[1]; xfb = fb; }
      }
    }
    if (xmode != 1) {                                      // (a second bond is committed after its two extra cells were seen)
      if (wNew != wN) D.state[cN] = wNew;
      if (bN != bN0) D.bonds[cN] = bN;
    }
    qw = wNew; qrb = rbN; qfl = flN; fcoor = w_coor(wNew);
  }
  // ---- the cells one more bond away that the reference itself reads (all lanes that need some, together) -----------
  if (__any_sync(FULL, xmode != 0)) {
    if (xmode == 1) {                                      // deposition.f:107-142: already-present reciprocal atoms on the two new bonds count
      P3 X0, X1; const uint32_t c0x = nbr(qN, xf0, X0), c1x = nbr(qN, xf1, X1);
      const uint32_t w0 = D.state[c0x], w1 = D.state[c1x];
      const unsigned long long b0 = D.bonds[c0x], b1 = D.bonds[c1x];
      const bool norigin = (qN.x | qN.y | qN.z) == 0;
      bool r0 = xf0 ? lists_code(b0, (unsigned)NEGC - xf0) : recipr_cold(X0, b0, qN);
      bool r1 = xf1 ? lists_code(b1, (unsigned)NEGC - xf1) : recipr_cold(X1, b1, qN);
      if (norigin) { r0 |= lists_code(b0, 0u); r1 |= lists_code(b1, 0u); }
      const bool q0 = w_occ(w0) && w_coor(w0) <= 3 && w_cov(w0) != 113 && r0;
      const bool q1 = w_occ(w1) && w_coor(w1) <= 3 && w_cov(w1) != 113 && r1;
      const unsigned f0 = fld(qrb, 0);
      const unsigned long long bN = (unsigned long long)f0 | ((unsigned long long)xfb << 16) | ((unsigned long long)xf0 << 32) | ((unsigned long long)xf1 << 48);
      const int cnew = 2 + (q0 ? 1 : 0) + (q1 ? 1 : 0);
      qfl = flag_get(qfl, 0) | ((unsigned)Itrans << 2) | ((q0 ? (unsigned)(w_sp(w0) & 3) : 0u) << 4) | ((q1 ? (unsigned)(w_sp(w1) & 3) : 0u) << 6);
      qw = MVBITS(wN0, cnew, LenCoor, PosCoor); qCoor = cnew; fcoor = cnew; act = A_SETVOID;
      D.state[qc] = qw; D.bonds[qc] = bN; qrb = bN;
    } else if (xmode == 2) {
      P3 X; nb1 = w_coor(D.state[nbr(qN, xf0, X)]);
    }
  }
  const unsigned upm = (__ballot_sync(FULL, sUp) >> tshift) & tm, downm = (__ballot_sync(FULL, sDown) >> tshift) & tm;
  const int Coor = c0 + __popc(upm) - __popc(downm);
  const unsigned keepm = (__ballot_sync(FULL, keepable) >> tshift) & 15u;
  const unsigned sf0 = (__ballot_sync(FULL, (sflag & 1u) != 0u) >> tshift) & 15u, sf1 = (__ballot_sync(FULL, (sflag & 2u) != 0u) >> tshift) & 15u;
  const int fcoor0 = __shfl_sync(FULL, fcoor, 0, TEAM);    // the first neighbour's coordination after phase A
  if (FEAT & F_COV) __syncwarp();                          // neighbours' words are final and visible to the whole team
  // ---- S's own request (lane 4) -----------------------------------------------------------------------
  uint32_t wS2 = wS;
  if (go && tl == 4) {
    // S's census, recomputed from what the four lanes saw: bit pairs (sf1,sf0) of lane i -> flag of bond i
    unsigned flS = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) flS |= (((sf0 >> i) & 1u) | (((sf1 >> i) & 1u) << 1)) << (2 * i);
    qCoor = Coor; qCov = CovNew; act = A_SET; qw = wS; qfl = qfl0;
    if (kind <= EVAP) { wS2 = MVBITS(wS, Coor, LenCoor, PosCoor); qfl = flS; }
    if (kind == DEP) { act = A_SETBULK; qOcc = 1; qSp = Itrans; }
    else if (kind == EVAP) {
      if (Coor == 0) { act = A_ERASE; qfl = 0u; }
      else if (Coor == 1) {                                // Evaporation.f:224-258: keep the last qualifying neighbour only
        if (!keepm) fail(E_STALE_NBUFF, S.x, S.y, S.z);
        const int k = 31 - __clz((int)(keepm | 1u));
        D.bonds[cS] = (unsigned long long)fld(bS, k);
        qfl = flag_get(flS, k);
        nb1 = fcoor0;                                      // rated with the OLD list (Evaporation.f:216-223): its first entry is N0
      }
    } else if (c0 == 1) nb1 = fcoor0;                      // desorption / adsorption on a 1-coordinated site
    qw = wS2;
  }
  // ---- phase B: lanes 0..4 rate their site -----------------------------------------------------------
  int ev = 0; double p = 0.0;
  if (act >= 0 && act != A_ERASE) {
    if (FEAT & F_COV) {
      unsigned cf = qfl;
      coop_rate<FEAT>(qOcc, qSp, qCov, qCoor, qN, qc, qrb, root, ev, p, &cf);
      if (!(FEAT & F_PAR) && qOcc && cf != qfl && !(qCov > 0 && qCov < 113)) fail(E_FLAGS, qN.x, qN.y, qN.z, (int)(cf | (qfl << 8)));
    } else flag_rate<FEAT>(qOcc, qSp, qCoor, qw, qc, qfl, nb1, root, ev, p);
    K.nsu++;
  }
  if (FEAT & F_COV) __syncwarp();
  // S's coordination is written after its neighbours were rated (deposition.f:237-250)
  if (go && tl == 4) {
    if (!(FEAT & F_COV) || wS2 != wS) D.state[cS] = wS2;
    if (act == A_ERASE) D.bonds[cS] = 0ull;
  }
  if ((FEAT & F_PAR) && rq) {                              // the caller commits (reference order: neighbours 1..4, then the site itself)
    rq->act = act; rq->ev = ev; rq->p = p; rq->c = qc; rq->w = qw; rq->Coor = qCoor; rq->self = go && tl == 4; rq->needs_root = p < 0.0;
    rq->Occ = qOcc; rq->Sp = qSp; rq->Cov = qCov; rq->N = qN; rq->rb = qrb;
    __syncwarp();
    return Itrans;
  }
  // ---- phase C: census flags, list bookkeeping (mskc_device.cuh: rq_commit) -------------------------------------------
  if (act >= 0) {
    if (qfl != qfl0) st_flags(qc, qfl);
    const bool gone = (ev == 0) && (act == A_SETVOID || act == A_SETBULK);   // the site left the active set
    if (act == A_SETVOID && gone) K.dVoids++;
    if (act == A_SETBULK && gone) K.dAtoms++;
    if (act == A_VOID2MC) K.dVoids--;
    if (act == A_BULK2MC) K.dAtoms--;
    if (act == A_ERASE || ev == 0) L.erase(qN, qc, aux_slot(qaux)); else L.put(qN, qc, ev, p, aux_slot(qaux));
  }
  __syncwarp();
  return Itrans;
}

}  // namespace mskc
