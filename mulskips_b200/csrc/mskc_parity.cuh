// mskc_parity.cuh -- serial-parity mode: the reference's n-fold-way loop replayed on the device.
//
// One persistent single-thread kernel walks main.f:304-371 event by event with the reference's own
// selection structure (complete binary rate tree, AllocateArrays.f / PickTreeEvent.f / UpdateTree.f),
// the reference's list order (Add.f / Erase.f swap-with-last) and its two random streams (ran3 and the
// 1000-number 'T' table), so that the lattice words, the event sequence and Time come out bit-identical.
// Only what the selection needs is kept of the three AoS lists: the MC list is (cell, event) per entry,
// the void list is (cell); bonds live per cell (mskc_device.cuh) and the bulk list is just a counter.
#pragma once
#include "mskc_device.cuh"

namespace mskc {

__device__ const float NUMBERSLIST[1000] = {
#include "numberslist.inc"
};

// loop state that survives between launches (main.f:31-54 locals + random.f SAVEd state)
struct ParityState {
  long long Iter, IterOut, framecounter, counter, events_done, site_updates;
  double Time;
  int NumAdAtom, NumAtoms, NumVoids;
  int CountCrystal[3], CountCov[3];
  int NumAdAtomOcc, NumOcc0;                // LaserAnnealing fork: occupied ad-atoms, stop threshold (la/main.f:714-723)
  int idum, runtype, r_iff, r_inext, r_inextp, r_ma[56];
  int pending, pend_ind, pend_itrans;       // an event picked but not yet applied (frame boundary)
  int last_site[3];
  int stop_reason;
};

struct ParityArgs {
  long long max_iter, out_every, framecounter, event_budget;
  double max_time, out_time;
  int exit_on_time;
};

// device pointers of the parity structures, in constant memory (uploaded with the context)
struct ParityView {
  double *Tree; int Levels, NumMcMax, NumVdMax, leaf0;
  uint32_t *lattind, *mc_cell, *void_cell; int *mc_ev;
  ParityState *ps;
};
__constant__ ParityView PV;

// The single thread that walks the loop keeps what it touches on every event on chip: the top 13 levels of the rate
// tree (nodes 1..8191, 64 KB; the descent of PickTreeEvent.f and the re-sums of UpdateTree.f are chains of dependent
// loads) and the state of ran3.  While the loop runs the mirrored nodes live in shared memory only; all 32 lanes load
// the mirror at the start of a launch and write it back at its end, so downloads and the next launch see the
// complete tree in global memory.
enum { PAR_TOP = 8192 };
struct ParShared { double top[PAR_TOP]; int ma[56]; int inext, inextp, iff, idum; };
#ifndef MSKC_EMU
extern __shared__ __align__(16) unsigned char par_smem[];
__device__ __forceinline__ ParShared &PSH() { return *reinterpret_cast<ParShared *>(par_smem); }
#else
__device__ __forceinline__ ParShared &PSH() { return *reinterpret_cast<ParShared *>(emu::dyn_smem()); }
#endif
__device__ __forceinline__ double tree_ld(int i) { return i < PAR_TOP ? PSH().top[i] : PV.Tree[i]; }
__device__ __forceinline__ void tree_st(int i, double v) { if (i < PAR_TOP) PSH().top[i] = v; else PV.Tree[i] = v; }

struct ParityPol {
  int NumAdAtom, NumAtoms, NumVoids;
  int CountCrystal[3], CountCov[3];
  int NumAdAtomOcc;
  long long n_getprob, n_ran3;

  __device__ __forceinline__ bool predraw() const { return true; }
  __device__ __forceinline__ double tree_root() const { return PSH().top[1]; }

  // mcevent/random.f:22-69
  __device__ __noinline__ double uniform() {
    const int mbig = 1000000000, mseed = 161803398, mz = 0;
    const double fac = (double)(1.0f / 1.0e9f);
    ParShared &s = PSH();
    int *ma = s.ma, mj, mk;
    n_ran3++;
    if (s.idum < 0 || s.iff == 0) {
      s.iff = 1;
      mj = mseed - abs(s.idum); mj = mj % mbig; ma[55] = mj; mk = 1;
      for (int i = 1; i <= 54; i++) { int ii = (21 * i) % 55; ma[ii] = mk; mk = mj - mk; if (mk < mz) mk += mbig; mj = ma[ii]; }
      for (int k = 1; k <= 4; k++) for (int i = 1; i <= 55; i++) { ma[i] = ma[i] - ma[1 + (i + 30) % 55]; if (ma[i] < mz) ma[i] += mbig; }
      s.inext = 0; s.inextp = 31; s.idum = 1;
    }
    if (++s.inext == 56) s.inext = 1;
    if (++s.inextp == 56) s.inextp = 1;
    mj = ma[s.inext] - ma[s.inextp];
    if (mj < mz) mj += mbig;
    ma[s.inext] = mj;
    return mj * fac;
  }

  // mcevent/UpdateTree.f:25-66.  The sibling loads are independent of the running sum, so they are
  // issued together; the 19 adds are the only dependent chain.
  __device__ __noinline__ void update_tree(int Index, double P) {
    if (P == 0 && Index < NumAdAtom) { fail(E_TREE_ZERO, Index); return; }
    if (Index > NumAdAtom) { fail(E_TREE_IDX, Index, NumAdAtom); return; }
    int Pos = Index + PV.leaf0 - 1;
    const int nlev = PV.Levels - 1;
    // levels below the mirror (global memory): ng of them, at most 11 (Levels <= 24); the sibling loads are
    // independent and issued together
    const int ng = nlev > 12 ? nlev - 12 : 0;              // nodes Pos >> l with l < ng have index >= PAR_TOP
    double sib[11];
#pragma unroll
    for (int l = 0; l < 11; l++) if (l < ng) sib[l] = PV.Tree[(Pos >> l) ^ 1];
    tree_st(Pos, P);
    double v = P;
#pragma unroll
    for (int l = 0; l < 11; l++) if (l < ng) { v = v + sib[l]; tree_st(Pos >> (l + 1), v); }
    // mirrored levels: shared memory only
    double *top = PSH().top;
#pragma unroll 1
    for (int l = ng; l < nlev; l++) { v = v + top[(Pos >> l) ^ 1]; top[Pos >> (l + 1)] = v; }
  }
  __device__ __forceinline__ double leaf(int Index) const { return tree_ld(Index + PV.leaf0 - 1); }

  // mcevent/Add.f:24-53
  __device__ __noinline__ void add_mc(P3 s, long long c, const P3 nn[4], int ev, double p) {
    if (p == 0) { fail(E_ADD_ZERO, s.x, s.y, s.z); return; }
    NumAdAtom++;
    if (NumAdAtom >= PV.NumMcMax) { fail(E_MC_FULL, NumAdAtom); NumAdAtom--; return; }
    PV.lattind[c] = (uint32_t)NumAdAtom;
    PV.mc_cell[NumAdAtom] = (uint32_t)c; PV.mc_ev[NumAdAtom] = ev;
    D.bonds[c] = enc_bonds(s, nn);
    update_tree(NumAdAtom, p);
  }
  __device__ __forceinline__ void set_rate(P3 s, long long c, int ev, double p) {
    int ind = (int)PV.lattind[c];
    if (ind < 1 || ind > PV.NumMcMax) { fail(E_ERASE, s.x, s.y, s.z, ind); return; }
    PV.mc_ev[ind] = ev;
    update_tree(ind, p);
  }
  // mcevent/Erase.f:24-78
  __device__ __noinline__ void erase_mc(P3 s, long long c) {
    int IndNN = (int)PV.lattind[c];
    if (IndNN > PV.NumMcMax || IndNN == 0) { fail(E_ERASE, s.x, s.y, s.z, IndNN); return; }
    if (IndNN >= NumAdAtom) {
      PV.mc_cell[IndNN] = 0xFFFFFFFFu; PV.mc_ev[IndNN] = 0;
      PV.lattind[c] = 0; D.bonds[c] = 0ull;
      update_tree(IndNN, 0.0);
      NumAdAtom--;
    } else {
      uint32_t moved = PV.mc_cell[NumAdAtom];
      PV.mc_cell[IndNN] = moved; PV.mc_ev[IndNN] = PV.mc_ev[NumAdAtom];
      update_tree(IndNN, leaf(NumAdAtom));
      update_tree(NumAdAtom, 0.0);
      PV.lattind[moved] = (uint32_t)IndNN;
      PV.mc_cell[NumAdAtom] = 0xFFFFFFFFu; PV.mc_ev[NumAdAtom] = 0;
      PV.lattind[c] = 0; D.bonds[c] = 0ull;
      NumAdAtom--;
    }
  }
  // mcevent/Add.f:82-100, Erase.f:135-179
  __device__ __noinline__ void add_void(P3 s, long long c, const P3 nn[4]) {
    NumVoids++;
    if (NumVoids >= PV.NumVdMax) { fail(E_VOID_FULL, NumVoids); NumVoids--; return; }
    PV.lattind[c] = (uint32_t)NumVoids; PV.void_cell[NumVoids] = (uint32_t)c;
    D.bonds[c] = enc_bonds(s, nn);
  }
  __device__ __noinline__ void erase_void(P3 s, long long c) {
    int IndNN = (int)PV.lattind[c];
    if (IndNN > PV.NumVdMax || IndNN == 0) { fail(E_ERASE, s.x, s.y, s.z, -IndNN); return; }
    if (IndNN != NumVoids) { uint32_t moved = PV.void_cell[NumVoids]; PV.void_cell[IndNN] = moved; PV.lattind[moved] = (uint32_t)IndNN; }
    PV.void_cell[NumVoids] = 0xFFFFFFFFu;
    PV.lattind[c] = 0; D.bonds[c] = 0ull;
    NumVoids--;
  }
  // mcevent/Add.f:58-76, Erase.f:84-128: the bulk list is never enumerated by the loop, a counter suffices
  __device__ __forceinline__ void add_bulk(P3 s, long long c, const P3 nn[4]) { NumAtoms++; PV.lattind[c] = 0; D.bonds[c] = enc_bonds(s, nn); }
  __device__ __forceinline__ void erase_bulk(P3 s, long long c) { NumAtoms--; D.bonds[c] = 0ull; }

  // Get_Prob.f:266-269 looks a coverage neighbour up in ListAdAtom even when it sits on a 4-coordinated
  // empty site, i.e. in ListVoid: it reads the MC record that happens to have the vacancy's index.  The
  // CVD goldens depend on it (26 967 times in CVD-Si-SiH2), so parity mode reproduces the lookup.
  __device__ __noinline__ bool cov_void_recipr(P3 n, long long cn, P3 me) {
    int v = (int)PV.lattind[cn];
    if (v < 1 || v > NumAdAtom) return eq(zero3(), me);          // zeroed record: NextNXYZ = 0
    uint32_t c2 = PV.mc_cell[v];
    if (c2 == 0xFFFFFFFFu) return eq(zero3(), me);
    return recipr_b(cell_xyz((long long)c2), D.bonds[c2], me);
  }
};

__device__ void parity_loop(const ParityArgs &A);
// mcevent/PickTreeEvent.f:25-70 + main.f:304-371
__global__ void __launch_bounds__(32, 1) parity_kernel(ParityArgs A) {
  if (blockIdx.x != 0) return;
  ParityState &S = *PV.ps;
  {                                               // fill the on-chip mirror (all 32 lanes), then lane 0 walks the loop alone
    ParShared &sh = PSH();
    const int ntop = min((int)PAR_TOP, 1 << PV.Levels);
    for (int i = threadIdx.x; i < ntop; i += 32) sh.top[i] = PV.Tree[i];
    for (int i = threadIdx.x; i < 56; i += 32) sh.ma[i] = S.r_ma[i];
    if (threadIdx.x == 0) { sh.inext = S.r_inext; sh.inextp = S.r_inextp; sh.iff = S.r_iff; sh.idum = S.idum; }
    __syncwarp();
  }
  if (threadIdx.x == 0) parity_loop(A);
  __syncwarp();
  {                                               // write the mirror back
    const ParShared &sh = PSH();
    const int ntop = min((int)PAR_TOP, 1 << PV.Levels);
    for (int i = 1 + threadIdx.x; i < ntop; i += 32) PV.Tree[i] = sh.top[i];
  }
}

__device__ __noinline__ void parity_loop(const ParityArgs &A) {
  ParityState &S = *PV.ps;
  ParityPol pol;
  pol.NumAdAtom = S.NumAdAtom; pol.NumAtoms = S.NumAtoms; pol.NumVoids = S.NumVoids;
  for (int i = 0; i < 3; i++) { pol.CountCrystal[i] = S.CountCrystal[i]; pol.CountCov[i] = S.CountCov[i]; }
  pol.n_getprob = 0; pol.n_ran3 = 0;
  pol.NumAdAtomOcc = S.NumAdAtomOcc;
  const bool fork = D.la == 3;
  long long Iter = S.Iter, done = 0;
  double Time = S.Time;
  long long framecounter = A.framecounter;
  int stop = 0;
  const int nlev = PV.Levels - 1;
  const int runtype = S.runtype;
  long long counter = S.counter;
  int nli = (int)(counter % 1000);                 // numberslist cursor: index mod(counter-1,1000)+1 of the NEXT draw (PickTreeEvent.f:44)
  while (true) {
    int Ind, Itrans;
    if (S.pending) { Ind = S.pend_ind; Itrans = S.pend_itrans; S.pending = 0; }
    else {
      if (A.event_budget > 0 && done >= A.event_budget) { stop = 0; break; }
      S.IterOut = Iter;
      if (fork) {                                  // la/main.f:636  CountLiquid = InitCountCrystal - CountCrystal
        int c0 = D.la3_init[0] - pol.CountCrystal[0], c1 = D.la3_init[1] - pol.CountCrystal[1], c2 = D.la3_init[2] - pol.CountCrystal[2];
        volatile int *cl = g_la_CL; cl[0] = c0; cl[1] = c1; cl[2] = c2; cl[3] = c0 + c1 + c2;
      }
      if (pol.NumAdAtom <= 20) { stop = 4; break; }
      int idx = 1;
      const double root = PSH().top[1];
      double node = root;                          // Tree(idx): the child taken at the previous level, already loaded
      for (int lev = 0; lev < nlev; lev++) {
        double r;
        if (runtype == 'T') { counter++; r = (double)NUMBERSLIST[nli]; nli = nli == 999 ? 0 : nli + 1; }
        else r = pol.uniform();
        const double Patom = r * node;
        idx <<= 1;
        double left, right;                        // Tree(2i), Tree(2i+1): one 16-byte load below the mirrored levels
        if (idx < PAR_TOP) { left = PSH().top[idx]; right = PSH().top[idx + 1]; }
        else { const double2 ch = *reinterpret_cast<const double2 *>(&PV.Tree[idx]); left = ch.x; right = ch.y; }
        if (left < Patom) { idx++; node = right; } else node = left;
      }
      Ind = idx - PV.leaf0 + 1;
      if (Ind > pol.NumAdAtom) { fail(E_PICK, Ind, pol.NumAdAtom); stop = -1; break; }
      Time = Time + 1.0 / root;
      Itrans = PV.mc_ev[Ind];
      P3 s = cell_xyz((long long)PV.mc_cell[Ind]);
      S.last_site[0] = s.x; S.last_site[1] = s.y; S.last_site[2] = s.z;
      bool frame = false;
      if (!A.exit_on_time) { if (A.out_every > 0 && (Iter % A.out_every) == 0) frame = true; }
      else if (A.out_time > 0 && Time >= ((double)(framecounter + 1) * A.out_time)) { frame = true; framecounter++; }
      if (frame) { S.pending = 1; S.pend_ind = Ind; S.pend_itrans = Itrans; stop = 3; break; }
    }
    P3 Site = cell_xyz((long long)PV.mc_cell[Ind]);
    // la/main.f:675-681: once a species of the Si/Ge melt is used up, the other one solidifies in its place
    if (fork && D.la3_swap && (Itrans == 1 || Itrans == 2) && pol.CountCrystal[Itrans - 1] >= D.la3_init[Itrans - 1]) Itrans = Itrans % 2 + 1;
    apply_event(pol, Site, Itrans);
    done++;
    Iter++;
    if (D.err[0] != 0) { stop = -1; break; }
    if (!A.exit_on_time) { if (Iter > A.max_iter) { stop = 1; break; } }
    else if (Time > A.max_time) { stop = 2; break; }
    if (fork && pol.NumAtoms + pol.NumAdAtomOcc > S.NumOcc0) { stop = 5; break; }       // la/main.f:714-723
  }
  S.Iter = Iter; S.Time = Time; S.framecounter = framecounter; S.counter = counter;
  { const ParShared &sh = PSH(); for (int i = 0; i < 56; i++) S.r_ma[i] = sh.ma[i]; S.r_inext = sh.inext; S.r_inextp = sh.inextp; S.r_iff = sh.iff; S.idum = sh.idum; }
  S.NumAdAtom = pol.NumAdAtom; S.NumAtoms = pol.NumAtoms; S.NumVoids = pol.NumVoids; S.NumAdAtomOcc = pol.NumAdAtomOcc;
  for (int i = 0; i < 3; i++) { S.CountCrystal[i] = pol.CountCrystal[i]; S.CountCov[i] = pol.CountCov[i]; }
  S.events_done = done; S.site_updates = pol.n_getprob;
  S.stop_reason = stop;
}

}  // namespace mskc
