// mskc_coop.cuh -- warp-cooperative form of the four KMC transitions (sm_100a).
//
// A team of lanes owns one tile of the synchronous-sublattice sampler.  An event at site S touches S and its
// four bonded neighbours N0..N3, then re-rates those five sites (each re-rate is a census of that site's
// own four neighbours).  The serial reference walks these one after the other (mcevent/deposition.f,
// Evaporation.f, absorption.f, desorption.f + Get_Prob.f); here
//   lanes 0..3  handle neighbour i  (phase A: coordination / bond update of N_i, phase B: its re-rate)
//   lane  4     handles S itself    (its bonds when they are fixed at deposition, its re-rate, its final word)
//   lanes 0..5  read the six zig-zag / arm-chair candidate cells of IF_Arm_Choice.f together
// so the dependent chain of an event is  S -> {N_i} -> {neighbours of N_i}  = 3..4 memory round trips
// instead of ~15, and every lane runs the same small code (the hot loop is kept near the 32 KB L1.5
// instruction cache: single call sites, cold paths out of line).  The rules themselves are those of
// mskc_device.cuh (the serial templates the parity kernel uses); tests/test_replay_gpu.py replays the
// oracle's event trace through this code and compares lattice words, bonds and (event class, rate) of
// every active site bit for bit.
//
// Geometry without floating point: bonds are stored as 15-bit relative codes (mskc_device.cuh), so
//   zig-zag continuation  (FIND_SN.f:60-64)  2*Cent - SN            ==  -(bond of C)          exact integers
//   arm-chair continuation(FIND_SN.f:66-70)  SN + 8.66*unit(S - C)  ==  bond of C + e + armoff(e)
// with armoff(e) = NINT(8.6602254037f * unit(-e)) tabulated on the host for all 31^3 bond codes by the same
// IEEE sequence (entries closer than 1e-6 to a rounding tie are left out and take the double-precision
// fallback find_sn()).  FIND_NN.f's rotation stays in double precision (one lane, only when an empty site
// receives its second neighbour).
#pragma once
#include "mskc_device.cuh"

namespace mskc {

enum { NEGC = 0x8420 };            // code(-v) = NEGC - code(v): every 5-bit field c -> 32 - c, no borrows (1 <= c <= 31)

__device__ __forceinline__ unsigned fld(unsigned long long b, int i) { return (unsigned)(b >> (16 * i)) & 0xFFFFu; }
__device__ __forceinline__ bool lists_code(unsigned long long b, unsigned code) {
  return (fld(b, 0) == code) | (fld(b, 1) == code) | (fld(b, 2) == code) | (fld(b, 3) == code);
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
__device__ __noinline__ void fail_cell(int code, uint32_t c) { P3 p = cell_xyz((long long)c); fail(code, p.x, p.y, p.z); }
// cell index of a site whose x,y are already wrapped (dec_bond output): only z can leave the box
__device__ __forceinline__ uint32_t ccell(P3 p) {
  if ((unsigned)p.z > (unsigned)D.LZ) { fail(E_OOB, p.x, p.y, p.z); return 0u; }
  return (uint32_t)p.x + (uint32_t)D.LX * ((uint32_t)p.y + (uint32_t)D.LY * (uint32_t)p.z);
}
__device__ __forceinline__ int cov_index(int cov) {      // position of a coverage id in ListCov (1-based), 0 = absent
  return cov == D.ListCov[0] ? 1 : (cov == D.ListCov[1] ? 2 : (cov == D.ListCov[2] ? 3 : 0));
}

// ---- rate of one site: mcevent/Get_Prob.f:54-331 ---------------------------------------------------
// Occ/NSiC/Cov/Coor are what the caller holds (not re-read), rb the four bonds to rate with (they can
// differ from the stored ones, Evaporation.f:123).  Deposition/adsorption sites get EV_DRAW: the species
// is drawn when the event fires (Get_Prob.f:160-179 draws it here; same distribution).
// FEAT: bit 0 = coverage species exist (NCov > 0), bit 1 = per-kelvin rate tables (LA rate law or 'ST' field); compiled out otherwise (hot-loop code size)
enum { F_COV = 1, F_LA = 2 };
template <int FEAT>
__device__ __forceinline__ void coop_rate(int Occ, int NSiC, int Cov, int Coor, P3 Q, uint32_t cQ, unsigned long long rb, double root, int &ev, double &p) {
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
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t w = wn[i];
    const unsigned f = fld(rb, i);
    const int CovN = w_cov(w);
    bool rec = lists_code(bn[i], (unsigned)NEGC - f);
    if (f == 0u) rec = recipr_cold(dec_bond(Q, 0u), bn[i], Q);
    if (qorigin) rec |= lists_code(bn[i], 0u);
    if (w_occ(w)) {
      if (CovN == 113) N_wall++;
      else if (rec) { const int s = w_sp(w); N_Si += (s == 1); N_C += (s == 2); N_B += (s == 3); }
    } else if ((FEAT & F_COV) && CovN != 0 && rec && w_coor(w) < 4) {
      // (coverage on a 4-coordinated empty site: the reference looks it up in the wrong list, see
      // ParityPol::cov_void_recipr, and in effect never counts it)
      const int iCov = cov_index(CovN);
      if (iCov == 0 || iCov > D.NCov) { fail_cell(E_COVLIST, cQ); return; }
      cv0 += (iCov == 1); cv1 += (iCov == 2); cv2 += (iCov == 3);
    }
  }
  const int ncry = N_Si + N_C + N_B + N_wall;
  if (ncry != Coor) { fail(E_CENSUS, Q.x, Q.y, Q.z, Coor * 16 + ncry); return; }
  if (ncry > 3) return;
  ev = 4;
  if (ncry == 0) { p = (double)100000.f * root; return; }
  const int idx = IDXE(NSiC, N_Si + N_wall, N_C, N_B, cv0 & 3, cv1 & 3, cv2 & 3);
  p = (FEAT & F_LA) ? rate_E(idx, T) : __ldg(&D.PE[idx]);
}

// Evaporation.f:95-122 / :229-258 on bond codes: keep the occupied, under-coordinated, reciprocal
// neighbour (the last one that qualifies) in slot 1, forget the others.
__device__ __noinline__ unsigned long long coop_compact(P3 me, unsigned long long old) {
  unsigned keep = 0u; bool found = false;
  uint32_t w[4]; unsigned long long b[4];                 // all eight words in flight together (one memory round trip)
#pragma unroll
  for (int j = 0; j < 4; j++) { const uint32_t c = ccell(dec_bond(me, fld(old, j))); w[j] = D.state[c]; b[j] = D.bonds[c]; }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned f = fld(old, j);
    if (w_occ(w[j]) && w_coor(w[j]) <= 3 && w_cov(w[j]) != 113 && recip_f(dec_bond(me, f), b[j], me, f)) { keep = f; found = true; }
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

// deposition.f:100-142: an empty site N with one fixed bond (to M) receives its second neighbour S: the
// other two bonds follow by rotation (FIND_NN.f), and already-present reciprocal atoms there count.
// Returns the new coordination (0 = the two known neighbours are not a tetrahedral pair: nothing happens).
__device__ __noinline__ int second_bond(P3 S, P3 N, unsigned fNS /* bond code N -> S */, unsigned long long &bN) {
  const unsigned f0 = fld(bN, 0);
  const P3 M = dec_bond(N, f0);
  int Dx = abs(M.x - S.x); Dx = min(Dx, D.LX - Dx);
  int Dy = abs(M.y - S.y); Dy = min(Dy, D.LY - Dy);
  const int dz = M.z - S.z;
  if (Dx * Dx + Dy * Dy + dz * dz != 72) return 0;
  P3 NewSite[2]; unsigned fn[2]; int cnew = 2;
  // FIND_NN.f:46-76 rotates b2 = S-N by +-120 degrees about b1 = M-N.  For two bonds of the ideal shell (|b|^2 = 27,
  // tetrahedral pair) the rotation is exact in integers: -(b1+b2)/2 -+ (b1 x b2)/6; mskc_create checks this
  // against the double-precision routine for all 120 such pairs (D.nn_exact).  Anything else rotates in double.
  const int b1x = cdx(f0), b1y = cdy(f0), b1z = cdz(f0), b2x = cdx(fNS), b2y = cdy(fNS), b2z = cdz(fNS);
  if (D.nn_exact && f0 != 0u && b1x * b1x + b1y * b1y + b1z * b1z == 27 && b2x * b2x + b2y * b2y + b2z * b2z == 27) {
    const int cx = b1y * b2z - b1z * b2y, cy = b1z * b2x - b1x * b2z, cz = b1x * b2y - b1y * b2x;
    const int sx = -(b1x + b2x) / 2, sy = -(b1y + b2y) / 2, sz = -(b1z + b2z) / 2;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int sg = k == 0 ? -1 : 1;
      const int ox = sx + sg * (cx / 6), oy = sy + sg * (cy / 6), oz = sz + sg * (cz / 6);
      fn[k] = (unsigned)(ox + 16) | ((unsigned)(oy + 16) << 5) | ((unsigned)(oz + 16) << 10);
      NewSite[k] = dec_bond(N, fn[k]);
    }
  } else {
    find_nn(M, S, N, NewSite);
    fn[0] = enc_bond_cold(N, NewSite[0]); fn[1] = enc_bond_cold(N, NewSite[1]);
  }
  uint32_t w3[2]; unsigned long long b3[2];              // both cells' words in flight together
#pragma unroll
  for (int k = 0; k < 2; k++) { const uint32_t c3 = ccell(NewSite[k]); w3[k] = D.state[c3]; b3[k] = D.bonds[c3]; }
#pragma unroll
  for (int k = 0; k < 2; k++)
    if (w_occ(w3[k]) && w_coor(w3[k]) <= 3 && w_cov(w3[k]) != 113 && recipr_b(NewSite[k], b3[k], N)) cnew++;
  bN = (unsigned long long)f0 | ((unsigned long long)fNS << 16) | ((unsigned long long)fn[0] << 32) | ((unsigned long long)fn[1] << 48);
  return cnew;
}

struct CoopCounters {                // per-lane partial sums, reduced over the warp at the end of a step
  int nsu, dAtoms, dVoids, dCrystal[3], dCov[3];
};

// ---- one event per TEAM of lanes (8 or 32; a power of two), all teams of the warp in lock step --------------
// EVERY lane of the warp calls this together (full-mask votes, shuffles and __syncwarp() are the convergence
// points): tl = lane index inside its team, go = "this team has an event" (a team without one only takes part in
// the synchronisation).  Loads are issued in code all teams share, the event-kind specific parts are short
// branches between two convergence points, so the memory round trips of the teams overlap instead of queueing.
// Itrans: 1..3 deposit species, 4 evaporate, 5 desorb, 6.. adsorb (main.f:335-357).
// arm_u: uniform used when IF_Arm_Choice.f ties; forced_arm >= 0 overrides the choice (trace replay).
template <int TEAM, int FEAT, class Lists>
__device__ __forceinline__ void coop_event(Lists &L, const int tl, bool go, const P3 S, const uint32_t cS, const int Itrans, const double root,
                                           const double arm_u, const int forced_arm, CoopCounters &K) {
  enum { DEP = 0, EVAP = 1, DES = 2, ABS = 3 };
  constexpr unsigned FULL = 0xFFFFFFFFu;
  constexpr unsigned tm = TEAM == 32 ? 0xFFFFFFFFu : ((1u << (TEAM & 31)) - 1u);
  const int kind = Itrans <= 3 ? DEP : (Itrans == 4 ? EVAP : (Itrans == 5 ? DES : ABS));
  const int tshift = (int)(threadIdx.x & 31u) - tl;         // first lane of the team inside the warp
  if (!(FEAT & F_COV) && go && kind >= DES) { fail_cell(E_BADEV, cS); go = false; }
  if ((FEAT & F_COV) && go && kind == ABS && Itrans - 5 > 3) { fail_cell(E_BADEV, cS); go = false; }
  uint32_t wS = 0u; unsigned long long bS = 0ull;
  if (go) { wS = D.state[cS]; bS = D.bonds[cS]; }
  const int c0 = w_coor(wS);
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
      if (dsp >= 0 && dsp < 3) K.dCrystal[dsp] += sgn;
      if (dcv >= 0 && dcv < 3) K.dCov[dcv] += sgn;
      D.state[cS] = wS;
    }
  }
  __syncwarp();                                            // S's new occupancy is visible before any lane looks back at it

  // deposition.f:49-74: a 1-coordinated site fixes bonds 2..4 (cubic or hexagonal continuation)
  const bool fix = go && kind == DEP && c0 == 1;
  unsigned zig[3] = {0u, 0u, 0u}, arm[3] = {0u, 0u, 0u};
  const unsigned e = fld(bS, 0);
  int cc = 0;
  if (fix) {
    const P3 C = dec_bond(S, e);
    const unsigned long long bC = D.bonds[ccell(C)];
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
      cc = w_coor(D.state[ccell(dec_bond(S, tl < 3 ? fa : fz))]);
    }
  }
  {
    const int CoorArm = __shfl_sync(FULL, cc, 0, TEAM) + __shfl_sync(FULL, cc, 1, TEAM) + __shfl_sync(FULL, cc, 2, TEAM);
    const int CoorZig = __shfl_sync(FULL, cc, 3, TEAM) + __shfl_sync(FULL, cc, 4, TEAM) + __shfl_sync(FULL, cc, 5, TEAM);
    if (fix) {
      bool IF_Arm = CoorZig < CoorArm || (CoorZig == CoorArm && !(arm_u <= D.PZig));
      if (forced_arm >= 0) IF_Arm = forced_arm != 0;
      bS = (unsigned long long)e;
#pragma unroll
      for (int i = 0; i < 3; i++) bS |= (unsigned long long)(IF_Arm ? arm[i] : zig[i]) << (16 * (i + 1));
      if (tl == 4) D.bonds[cS] = bS;
    }
  }

  // ---- phase A: lane i < 4 updates neighbour i ---------------------------------------------------
  int act = -1, qOcc = 0, qSp = 0, qCov = 0, qCoor = 0;
  P3 qN = S; uint32_t qc = cS; unsigned long long qrb = bS;
  int qsl = 0;                                            // the site's list slot, fetched early (off the critical path of the commit)
  if (go && tl == 4) qsl = L.slot_of(cS);
  bool sUp = false, sDown = false;                        // this neighbour raises / lowers S's own coordination
  bool keepable = false;                                  // Evaporation.f:229-258: occupied, under-coordinated, reciprocal neighbour of S
  if (go && tl < 4 && !(kind == DES && c0 == 1 && tl > 0)) {    // desorption.f:54-81: a 1-coordinated site looks at bond 1 only
    const unsigned f = fld(bS, tl);
    const P3 N = dec_bond(S, f);
    const uint32_t cN = ccell(N);
    const uint32_t wN = D.state[cN];
    unsigned long long bN = D.bonds[cN];
    qsl = L.slot_of(cN);
    const int Occ = w_occ(wN), CovN = w_cov(wN), CoorNN = w_coor(wN);
    const bool rec = recip_f(N, bN, S, f);
    const bool wall = Occ && CovN == 113;
    qN = N; qc = cN; qCov = CovN; qOcc = Occ; qSp = Occ ? w_sp(wN) : 0;
    int newCoor = -1; uint32_t wNew = wN;
    if (Occ && !wall && CovN != 0 && kind <= EVAP) fail_cell(E_NNCOV, cN);
    else if (kind == DEP) {
      if (wall) sUp = true;                                // deposition.f:231-232
      else if (Occ) {                                      // :191-229 full neighbour
        if (CoorNN <= 3 && rec) { newCoor = CoorNN + 1; act = A_SETBULK; sUp = (c0 == 1 && tl >= 1); }
      } else if (CoorNN == 0) {                            // :88-99 a new active empty site with one fixed bond
        bN = f ? (unsigned long long)((unsigned)NEGC - f) : (unsigned long long)enc_bond_cold(N, S);
        D.bonds[cN] = bN;
        newCoor = 1; act = A_NEW;
      } else if (CoorNN == 1) {                            // :100-160 second bond: complete the tetrahedron
        const int cnew = second_bond(S, N, f ? (unsigned)NEGC - f : enc_bond_cold(N, S), bN);
        if (cnew) { D.bonds[cN] = bN; newCoor = cnew; act = A_SETVOID; }
      } else if (CoorNN <= 3 && rec) { newCoor = CoorNN + 1; act = A_SETVOID; }   // :161-190
    } else if (kind == EVAP) {
      if (wall) sDown = true;                              // Evaporation.f:211-213
      else if (rec) {
        if (Occ) {                                         // :173-209; a bulk atom becomes active again
          newCoor = CoorNN - 1; act = CoorNN == 4 ? A_BULK2MC : A_SET; keepable = true;
        } else if (CoorNN == 1) {                          // :62-79 orphaned empty site leaves the lattice
          wNew = D.keepT ? (wN & (((1u << LenT) - 1u) << PosT)) : 0u;
          D.bonds[cN] = 0ull;
          act = A_ERASE; qCov = 0;
        } else if (CoorNN >= 2 && CoorNN <= 4) {           // :80-164; 4: a vacancy becomes active again
          newCoor = CoorNN - 1; act = CoorNN == 4 ? A_VOID2MC : A_SET;
          if (CoorNN == 2) D.bonds[cN] = coop_compact(N, bN);          // (rated with the OLD neighbour list, :123)
        }
      }
    } else if ((FEAT & F_COV) && Occ && !wall && CoorNN <= 3 && rec) { act = A_SET; newCoor = CoorNN; }   // absorption.f:43-72 / desorption.f:54-111
    if (newCoor >= 0) { wNew = MVBITS(wN, newCoor, LenCoor, PosCoor); qCoor = newCoor; }
    if (wNew != wN) D.state[cN] = wNew;
    qrb = bN;
  }
  const unsigned upm = (__ballot_sync(FULL, sUp) >> tshift) & tm, downm = (__ballot_sync(FULL, sDown) >> tshift) & tm;
  const int Coor = c0 + __popc(upm) - __popc(downm);
  const unsigned keepm = (__ballot_sync(FULL, keepable) >> tshift) & 15u;
  __syncwarp();                                            // neighbours' words are final and visible to the whole team
  // ---- S's own request (lane 4) -----------------------------------------------------------------------
  uint32_t wS2 = wS;
  if (go && tl == 4) {
    qCoor = Coor; qCov = CovNew; act = A_SET;
    if (kind <= EVAP) wS2 = MVBITS(wS, Coor, LenCoor, PosCoor);
    if (kind == DEP) { act = A_SETBULK; qOcc = 1; qSp = Itrans; }
    else if (kind == EVAP) {
      if (Coor == 0) act = A_ERASE;
      else if (Coor == 1) {                                // Evaporation.f:224-258: keep the last qualifying neighbour only
        if (!keepm) fail(E_STALE_NBUFF, S.x, S.y, S.z);
        D.bonds[cS] = (unsigned long long)fld(bS, 31 - __clz((int)(keepm | 1u)));
      }
    }
  }
  // ---- phase B: lanes 0..4 rate their site -----------------------------------------------------------
  int ev = 0; double p = 0.0;
  if (act >= 0 && act != A_ERASE) { coop_rate<FEAT>(qOcc, qSp, qCov, qCoor, qN, qc, qrb, root, ev, p); K.nsu++; }
  __syncwarp();
  // S's coordination is written after its neighbours were rated (deposition.f:237-250)
  if (go && tl == 4) {
    if (wS2 != wS) D.state[cS] = wS2;
    if (act == A_ERASE) D.bonds[cS] = 0ull;
  }
  // ---- phase C: list bookkeeping (mskc_device.cuh: rq_commit) -------------------------------------------
  if (act >= 0) {
    const bool gone = (ev == 0) && (act == A_SETVOID || act == A_SETBULK);   // the site left the active set
    if (act == A_SETVOID && gone) K.dVoids++;
    if (act == A_SETBULK && gone) K.dAtoms++;
    if (act == A_VOID2MC) K.dVoids--;
    if (act == A_BULK2MC) K.dAtoms--;
    if (act == A_ERASE || ev == 0) L.erase(qN, qc, qsl); else L.put(qN, qc, ev, p, qsl);
  }
  __syncwarp();
}

}  // namespace mskc
