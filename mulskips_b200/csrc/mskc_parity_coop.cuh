// mskc_parity_coop.cuh -- serial-parity mode, warp-cooperative form.
//
// The loop of main.f:304-371 is strictly sequential (one rate tree, one random stream): what one warp can share is the
// work INSIDE an iteration.  parity_kernel (mskc_parity.cuh) walks everything on one thread, 7.0e3 dependent instructions
// per event; here the 32 lanes of one warp walk the loop together:
//   * selection (PickTreeEvent.f:41-55): the populated prefix of every tree level lives in shared memory (the active
//     list is a prefix of the leaves, so a ~200 KB mirror holds the whole populated tree of every regression case);
//     the descent is 19 dependent shared-memory look-ups;
//   * the transition: the sampler's cooperative event (mskc_coop.cuh, F_PAR) -- four lanes take the four neighbours,
//     one the site itself, their censuses (Get_Prob.f:216-291) run side by side;
//   * random numbers: ran3 (random.f:22-69) is a lagged subtractive generator x[n] = x[n-55] - x[n-24], so 24 numbers
//     at a time are produced by 24 lanes; they are consumed strictly in the reference's order (the tie of
//     IF_Arm_Choice.f first, then one per rated empty site in request order); what was produced but not consumed is
//     un-generated at the end of a launch (the recurrence is reversible), so the saved state is the reference's;
//   * list / tree bookkeeping in the reference's order (Add.f, Erase.f: swap-with-last), but the 19 ancestor re-sums
//     of each UpdateTree.f call are deferred: leaves are written at once, the touched root paths are re-summed level
//     by level (one lane per path) at the end of the event.  Every inner node is the IEEE sum of its two children at
//     the time it was last written, and IEEE addition is commutative, so the tree after the event is bit-identical to
//     the one 6-8 successive UpdateTree calls leave.  The one mid-event reader of the root -- the rate
//     100000*Tree(1) of an isolated atom, Get_Prob.f:305 -- forces the re-sum before it is evaluated.
// Everything observable is the reference's: lattice words, list order, event sequence, Time bit for bit
// (tests/test_parity_gpu.py, tests/test_host_gpu.py run this kernel; MSKC_PARITY_SERIAL=1 selects the one-thread kernel).
#pragma once
#include "mskc_device.cuh"
#include "mskc_parity.cuh"
#include "mskc_coop.cuh"
#include "mskc_parallel.cuh"

namespace mskc {

enum { PC_MAXLEV = 24, PC_UBUF = 64, PC_MIRROR = 24800 /* doubles of tree mirror: 194 KB */ };
struct PcShared {
  int ma[56]; int inext, inextp, iff, idum;      // ran3 state (random.f SAVEd variables)
  double ubuf[PC_UBUF];                          // uniforms produced ahead, ring buffer
  int wl, nlev;                                  // mirrored leaf prefix, level of the leaves
  int NumAdAtom;                                 // for par_cov_void_recipr
  int pend[32];                                  // leaves written in this event (their root paths await the re-sum)
  int quirk[32];                                 // lane consulted the lists while rating (par_cov_void_recipr): its rate is re-evaluated at commit time
  double __align__(16) tree[PC_MIRROR];
};
#ifndef MSKC_EMU
extern __shared__ __align__(16) unsigned char pc_smem[];
__device__ __forceinline__ PcShared &PCS() { return *reinterpret_cast<PcShared *>(pc_smem); }
#else
__device__ __forceinline__ PcShared &PCS() { return *reinterpret_cast<PcShared *>(emu::dyn_smem()); }
#endif

// Node o of level L (level L has 2^L nodes; heap index 2^L + o).  Mirrored: the first pc_W(L) nodes of every level -- the
// ancestors of the first `wl` leaves -- at offset pc_base(L); both are arithmetic in (wl, nlev), no table look-up sits in
// front of the shared-memory access (wl and nlev live in PcShared, uniform).  pc_base(L) = even(wl >> (nlev-L)) + 4L leaves
// room for all lower levels (their widths halve), and keeps every (2o, 2o+1) child pair 16-byte aligned.
__device__ __forceinline__ int pc_W(int L, int wl, int nlev) { return min(1 << L, ((wl >> (nlev - L)) + 3) & ~1); }
__device__ __forceinline__ int pc_base(int L, int wl, int nlev) { return (((wl >> (nlev - L)) + 1) & ~1) + 4 * L; }
__device__ __forceinline__ double pc_ld(int L, int o) {
  const PcShared &s = PCS(); const int wl = s.wl, nlev = s.nlev;
  return o < pc_W(L, wl, nlev) ? s.tree[pc_base(L, wl, nlev) + o] : PV.Tree[(1 << L) + o];
}
__device__ __forceinline__ void pc_st(int L, int o, double v) {
  PcShared &s = PCS(); const int wl = s.wl, nlev = s.nlev;
  if (o < pc_W(L, wl, nlev)) s.tree[pc_base(L, wl, nlev) + o] = v; else PV.Tree[(1 << L) + o] = v;
}
// both children (2o, 2o+1) of node o of level L-1, one 16-byte access
__device__ __forceinline__ double2 pc_ld2(int L, int o2, int wl, int nlev) {
  const PcShared &s = PCS();
  if (o2 < pc_W(L, wl, nlev)) return *reinterpret_cast<const double2 *>(&s.tree[pc_base(L, wl, nlev) + o2]);
  return *reinterpret_cast<const double2 *>(&PV.Tree[(1 << L) + o2]);
}

// Get_Prob.f:266-269 (see ParityPol::cov_void_recipr): the MC record that happens to have the vacancy's index
__device__ __noinline__ bool par_cov_void_recipr(P3 n, uint32_t cn, P3 me) {
  PCS().quirk[threadIdx.x & 31] = 1;
  const int v = (int)PV.lattind[cn];
  if (v < 1 || v > PCS().NumAdAtom) return eq(zero3(), me);
  const uint32_t c2 = PV.mc_cell[v];
  if (c2 == 0xFFFFFFFFu) return eq(zero3(), me);
  return recipr_b(cell_xyz((long long)c2), D.bonds[c2], me);
}

__device__ __forceinline__ int pc_wrap55(int i) { return ((i - 1) % 55) + 1; }
// random.f:22-69, 24 numbers at once (all 32 lanes call; `avail` numbers wait in the ring starting at `head`)
__device__ __noinline__ void pc_generate(int lane, int head, int avail) {
  PcShared &s = PCS();
  const int mbig = 1000000000, mseed = 161803398, mz = 0;
  const double fac = (double)(1.0f / 1.0e9f);
  if (s.idum < 0 || s.iff == 0) {                 // first use: the initialisation of random.f:36-58
    if (lane == 0) {
      int *ma = s.ma, mj, mk;
      s.iff = 1;
      mj = mseed - abs(s.idum); mj = mj % mbig; ma[55] = mj; mk = 1;
      for (int i = 1; i <= 54; i++) { int ii = (21 * i) % 55; ma[ii] = mk; mk = mj - mk; if (mk < mz) mk += mbig; mj = ma[ii]; }
      for (int k = 1; k <= 4; k++) for (int i = 1; i <= 55; i++) { ma[i] = ma[i] - ma[1 + (i + 30) % 55]; if (ma[i] < mz) ma[i] += mbig; }
      s.inext = 0; s.inextp = 31; s.idum = 1;
    }
    __syncwarp();
  }
  int mj = 0, idx = 0;
  if (lane < 24) {
    idx = pc_wrap55(s.inext + 1 + lane);
    mj = s.ma[idx] - s.ma[pc_wrap55(s.inextp + 1 + lane)];
    if (mj < mz) mj += mbig;
  }
  __syncwarp();
  if (lane < 24) { s.ma[idx] = mj; s.ubuf[(head + avail + lane) & (PC_UBUF - 1)] = mj * fac; }
  __syncwarp();
  if (lane == 0) { s.inext = pc_wrap55(s.inext + 24); s.inextp = pc_wrap55(s.inextp + 24); }
  __syncwarp();
}
// un-generate the `left` numbers that were produced but not consumed (lane 0): ma[idx] = ma[idx] + ma[idxp] mod mbig
__device__ __noinline__ void pc_ungenerate(int left) {
  PcShared &s = PCS();
  const int mbig = 1000000000;
  for (int t = 0; t < left; t++) {
    int v = s.ma[s.inext] + s.ma[s.inextp];
    if (v >= mbig) v -= mbig;
    s.ma[s.inext] = v;
    s.inext = s.inext == 1 ? 55 : s.inext - 1;
    s.inextp = s.inextp == 1 ? 55 : s.inextp - 1;
  }
}

// the deferred half of UpdateTree.f:49-65: re-sum the root paths of the leaves written since the last call, level by
// level, one lane per path (duplicate parents are written twice with the same value)
__device__ __forceinline__ void pc_flush(int lane, int nlev, int &npend) {
  PcShared &s = PCS();
  const int n = npend;                             // (uniform over the warp: every lane counts the leaf writes itself)
  npend = 0;
  if (n == 0) return;
  int o = lane < n ? s.pend[lane] : 0;
  __syncwarp();
#pragma unroll 1
  for (int L = nlev - 1; L >= 0; L--) {            // parents at level L of nodes at level L+1
    o >>= 1;
    if (lane < n) { const double2 ch = pc_ld2(L + 1, 2 * o, s.wl, nlev); pc_st(L, o, ch.x + ch.y); }
    __syncwarp();
  }
}

// mcevent/PickTreeEvent.f:25-70 + main.f:304-371, one warp
__global__ void __launch_bounds__(32, 1) parity_coop_kernel(ParityArgs A) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  constexpr int FEAT = F_COV | F_LA | F_PAR;
  const int lane = threadIdx.x;
  PcShared &sh = PCS();
  ParityState &S = *PV.ps;
  const int nlev = PV.Levels - 1;                  // leaves live on level nlev
  int mwl = 0;                                     // mirrored leaf prefix (uniform)
  // ---- mirror: the first W[L] nodes of every level (ancestors of the first PC_LEAVES leaves), W even so that siblings stay together
  {
    const int leaves = 1 << nlev;
    int wl = min(leaves, 12000);                   // largest leaf prefix whose ancestors fit the mirror
    while (wl > 64 && pc_base(nlev, wl, nlev) + pc_W(nlev, wl, nlev) > (int)PC_MIRROR) wl -= 500;
    if (lane == 0) {
      sh.wl = wl; sh.nlev = nlev;
      for (int i = 0; i < 56; i++) sh.ma[i] = S.r_ma[i];
      sh.inext = S.r_inext; sh.inextp = S.r_inextp; sh.iff = S.r_iff; sh.idum = S.idum;
    }
    __syncwarp();
    mwl = sh.wl;
    for (int L = 0; L <= nlev; L++) for (int o = lane; o < pc_W(L, sh.wl, nlev); o += 32) sh.tree[pc_base(L, sh.wl, nlev) + o] = PV.Tree[(1 << L) + o];
    __syncwarp();
  }
  // ---- loop state, uniform over the warp
  int NumAdAtom = S.NumAdAtom, NumAtoms = S.NumAtoms, NumVoids = S.NumVoids, NumAdAtomOcc = S.NumAdAtomOcc;
  int CC[3] = {S.CountCrystal[0], S.CountCrystal[1], S.CountCrystal[2]}, CV[3] = {S.CountCov[0], S.CountCov[1], S.CountCov[2]};
  const bool fork = D.la == 3;
  long long Iter = S.Iter, done = 0, nsu = 0, counter = S.counter, framecounter = A.framecounter;
  double Time = S.Time;
  int stop = 0;
  const int runtype = S.runtype;
  int nli = (int)(counter % 1000);                 // numberslist cursor (PickTreeEvent.f:44)
  int uhead = 0, uavail = 0;                       // ring of ran3 uniforms produced ahead
  int npend = 0;                                   // leaves written since the last re-sum of their root paths
  TileLists Ldummy; Ldummy.own = -1; Ldummy.mark = 0; Ldummy.s_rate = nullptr; Ldummy.s_ce = nullptr; Ldummy.s_n = nullptr;
  CoopCounters K; K.nsu = 0; K.dAtoms = 0; K.dVoids = 0;
  for (int i = 0; i < 3; i++) { K.dCrystal[i] = 0; K.dCov[i] = 0; }
  bool pending = S.pending != 0;
  int Ind = S.pend_ind, Itrans = S.pend_itrans;
  auto next_uniform = [&]() -> double {            // uniform over the warp: every lane advances the same cursor
    if (uavail == 0) { pc_generate(lane, uhead, uavail); uavail += 24; }
    const double u = sh.ubuf[uhead]; uhead = (uhead + 1) & (PC_UBUF - 1); uavail--;
    return u;
  };
  while (true) {
    if (pending) pending = false;
    else {
      if (A.event_budget > 0 && done >= A.event_budget) { stop = 0; break; }
      if (lane == 0) S.IterOut = Iter;
      if (fork) {                                  // la/main.f:636  CountLiquid = InitCountCrystal - CountCrystal
        const int c0 = D.la3_init[0] - CC[0], c1 = D.la3_init[1] - CC[1], c2 = D.la3_init[2] - CC[2];
        if (lane == 0) { volatile int *cl = g_la_CL; cl[0] = c0; cl[1] = c1; cl[2] = c2; cl[3] = c0 + c1 + c2; }
        __syncwarp();
      }
      if (NumAdAtom <= 20) { stop = 4; break; }
      // PickTreeEvent.f:41-55: one fresh uniform per level
      const double root = pc_ld(0, 0);
      double node = root; int o = 0;
#pragma unroll 1
      for (int lev = 0; lev < nlev; lev++) {
        double r;
        if (runtype == 'T') { counter++; r = (double)NUMBERSLIST[nli]; nli = nli == 999 ? 0 : nli + 1; }
        else r = next_uniform();
        const double Patom = r * node;
        const double2 ch = pc_ld2(lev + 1, 2 * o, mwl, nlev); const double left = ch.x, right = ch.y;
        if (left < Patom) { o = 2 * o + 1; node = right; } else { o = 2 * o; node = left; }
      }
      Ind = o + 1;
      if (Ind > NumAdAtom) { fail(E_PICK, Ind, NumAdAtom); stop = -1; break; }
      Time = Time + 1.0 / root;
      Itrans = PV.mc_ev[Ind];
      const P3 s = cell_xyz((long long)PV.mc_cell[Ind]);
      if (lane == 0) { S.last_site[0] = s.x; S.last_site[1] = s.y; S.last_site[2] = s.z; }
      bool frame = false;
      if (!A.exit_on_time) { if (A.out_every > 0 && (Iter % A.out_every) == 0) frame = true; }
      else if (A.out_time > 0 && Time >= ((double)(framecounter + 1) * A.out_time)) { frame = true; framecounter++; }
      if (frame) { pending = true; stop = 3; break; }
    }
    const uint32_t cS = PV.mc_cell[Ind];
    const P3 Site = cell_xyz((long long)cS);
    // la/main.f:675-681: once a species of the Si/Ge melt is used up, the other one solidifies in its place
    if (fork && D.la3_swap && (Itrans == 1 || Itrans == 2) && CC[Itrans - 1] >= D.la3_init[Itrans - 1]) Itrans = Itrans % 2 + 1;
    if (Itrans < 1 || Itrans > 8) { fail(E_BADEV, Site.x, Site.y, Site.z, Itrans); stop = -1; break; }
    // ---- the transition (deposition.f / Evaporation.f / absorption.f / desorption.f), cooperative ------------------
    while (uavail < 8) { pc_generate(lane, uhead, uavail); uavail += 24; }          // the tie of IF_Arm_Choice.f + up to five pre-drawn species
    if (lane == 0) sh.NumAdAtom = NumAdAtom;
    sh.quirk[lane] = 0;
    __syncwarp();
    CoopReq rq; rq.act = -1; rq.ev = 0; rq.p = 0.0; rq.c = 0; rq.w = 0; rq.Coor = 0; rq.self = 0; rq.needs_root = 0; rq.arm_used = 0;
    const int kd0 = K.dCrystal[0], kd1 = K.dCrystal[1], kd2 = K.dCrystal[2], kv0 = K.dCov[0], kv1 = K.dCov[1], kv2 = K.dCov[2];
    coop_event<32, FEAT>(Ldummy, lane, true, Site, cS, Itrans, 0.0, sh.ubuf[uhead], -1, K, 0.5, &rq);
#ifdef MSKC_EMU_DEBUG
    if (lane < 6) printf(" ev Itrans=%d lane %d: act=%d ev=%d p=%g c=%u self=%d nr=%d\n", Itrans, lane, rq.act, rq.ev, rq.p, rq.c, rq.self, rq.needs_root);
#endif
    if (__any_sync(FULL, rq.arm_used != 0)) { uhead = (uhead + 1) & (PC_UBUF - 1); uavail--; }
    // species counters (main.f:341,356; Evaporation.f:41; desorption.f:42-48): lane 4 counted them
    CC[0] += __shfl_sync(FULL, K.dCrystal[0] - kd0, 4); CC[1] += __shfl_sync(FULL, K.dCrystal[1] - kd1, 4); CC[2] += __shfl_sync(FULL, K.dCrystal[2] - kd2, 4);
    CV[0] += __shfl_sync(FULL, K.dCov[0] - kv0, 4); CV[1] += __shfl_sync(FULL, K.dCov[1] - kv1, 4); CV[2] += __shfl_sync(FULL, K.dCov[2] - kv2, 4);
    if (Itrans == 4) NumAdAtomOcc--;               // la/mcevent/Evaporation.f:47
    // Get_Prob.f:160-179 draws the species / adsorbate when it rates an empty site: one uniform per such request, in request order
    const bool rated = rq.act >= 0 && rq.act != A_ERASE;
    const unsigned dm = __ballot_sync(FULL, rated && rq.ev == EV_DRAW);
    if (rated && rq.ev == EV_DRAW) {
      const double u = sh.ubuf[(uhead + __popc(dm & ((1u << lane) - 1u))) & (PC_UBUF - 1)];
      const int T = D.la ? site_T(rq.w, cell_xyz((long long)rq.c)) : 0;
      double ptot = 0.0;
      rq.ev = choose_dep_abs(rq.Coor, T, u, &ptot); rq.p = ptot;
    }
    uhead = (uhead + __popc(dm)) & (PC_UBUF - 1); uavail -= __popc(dm);
    nsu += __popc(__ballot_sync(FULL, rated));
    // list indices of the requests' sites (one round trip for all five)
    int my_ind = rq.act >= 0 ? (int)PV.lattind[rq.c] : 0;
    // ---- commit, in the reference's order: neighbours 1..4, then the site (mskc_device.cuh: rq_commit) -------------
#pragma unroll 1
    for (int k = 0; k < 5; k++) {
      const int act = __shfl_sync(FULL, rq.act, k);
      if (act < 0) continue;
      int ev = __shfl_sync(FULL, rq.ev, k);
      double p = __shfl_sync(FULL, rq.p, k);
      const uint32_t c = __shfl_sync(FULL, rq.c, k);
      const int self = __shfl_sync(FULL, rq.self, k);
      int ind = __shfl_sync(FULL, my_ind, k);
      // Get_Prob.f:266-269 reads the particle lists (the wrong one) for a coverage neighbour on a 4-coordinated empty site: the
      // reference rates request k after it committed requests 1..k-1, so such a rate is evaluated again here, against the
      // lists as they are now
      if (__shfl_sync(FULL, sh.quirk[lane], k)) {
        if (lane == 0) sh.NumAdAtom = NumAdAtom;
        __syncwarp();
        if (lane == k) { int e2 = 0; double p2 = 0.0; coop_rate<FEAT>(rq.Occ, rq.Sp, rq.Cov, rq.Coor, rq.N, rq.c, rq.rb, 0.0, e2, p2); rq.ev = e2; rq.p = p2; rq.needs_root = p2 < 0.0; }
        __syncwarp();
        ev = __shfl_sync(FULL, rq.ev, k); p = __shfl_sync(FULL, rq.p, k);
      }
      if (__shfl_sync(FULL, rq.needs_root, k)) { pc_flush(lane, nlev, npend); p = (double)100000.f * pc_ld(0, 0); }     // Get_Prob.f:305
      // which list operations this request amounts to
#ifdef MSKC_EMU_DEBUG
      if (lane == 0) printf("  commit k=%d act=%d ev=%d p=%g c=%u self=%d ind=%d N=%d NV=%d\n", k, act, ev, p, c, self, ind, NumAdAtom, NumVoids);
#endif
      const bool leaves = ev == 0 && (act == A_SETVOID || act == A_SETBULK);
      const bool do_erase_mc = act == A_ERASE || leaves;
      const bool do_add_mc = (act == A_NEW && ev != 0) || act == A_VOID2MC || act == A_BULK2MC;
      const bool do_set = !do_erase_mc && !do_add_mc && act != A_NEW;
      if (act == A_VOID2MC) {                      // Erase.f:135-179
        if (ind > PV.NumVdMax || ind == 0) { fail(E_ERASE, -1, -1, -1, -ind); stop = -1; break; }
        uint32_t moved = 0xFFFFFFFFu;
        if (ind != NumVoids) moved = PV.void_cell[NumVoids];
        __syncwarp();                              // every lane has read the tail entry before lane 0 overwrites it
        if (lane == 0) {
          if (ind != NumVoids) { PV.void_cell[ind] = moved; PV.lattind[moved] = (uint32_t)ind; }
          PV.void_cell[NumVoids] = 0xFFFFFFFFu; PV.lattind[c] = 0;
        }
        if (ind != NumVoids && rq.c == moved) my_ind = ind;
        NumVoids--;
      }
      if (act == A_BULK2MC) { NumAtoms--; NumAdAtomOcc++; }      // Erase.f:84-128; la/mcevent/Evaporation.f:192
      if (do_erase_mc) {                           // Erase.f:24-78
        if (ind > PV.NumMcMax || ind == 0) { fail(E_ERASE, -1, -1, -1, ind); stop = -1; break; }
        const int leafL = nlev;
        if (ind >= NumAdAtom) {
          if (ind > NumAdAtom) { fail(E_TREE_IDX, ind, NumAdAtom); stop = -1; break; }
          if (lane == 0) { PV.mc_cell[ind] = 0xFFFFFFFFu; PV.mc_ev[ind] = 0; PV.lattind[c] = 0; pc_st(leafL, ind - 1, 0.0); sh.pend[npend] = ind - 1; }
          npend++;
        } else {
          const uint32_t moved = PV.mc_cell[NumAdAtom]; const int mev = PV.mc_ev[NumAdAtom];
          const double lv = pc_ld(leafL, NumAdAtom - 1);
          if (lv == 0) { fail(E_TREE_ZERO, ind); stop = -1; break; }
          __syncwarp();                            // every lane has read the tail entry before lane 0 overwrites it
          if (lane == 0) {
            PV.mc_cell[ind] = moved; PV.mc_ev[ind] = mev;
            pc_st(leafL, ind - 1, lv); pc_st(leafL, NumAdAtom - 1, 0.0); sh.pend[npend] = ind - 1; sh.pend[npend + 1] = NumAdAtom - 1;
            PV.lattind[moved] = (uint32_t)ind;
            PV.mc_cell[NumAdAtom] = 0xFFFFFFFFu; PV.mc_ev[NumAdAtom] = 0; PV.lattind[c] = 0;
          }
          npend += 2;
          if (rq.c == moved) my_ind = ind;
        }
        NumAdAtom--;
        if (act == A_SETVOID) {                    // Add.f:82-100
          NumVoids++;
          if (NumVoids >= PV.NumVdMax) { fail(E_VOID_FULL, NumVoids); stop = -1; break; }
          if (lane == 0) { PV.lattind[c] = (uint32_t)NumVoids; PV.void_cell[NumVoids] = c; }
        } else if (act == A_SETBULK) { NumAtoms++; NumAdAtomOcc--; if (lane == 0) PV.lattind[c] = 0; }      // Add.f:58-76
      }
      if (do_add_mc) {                             // Add.f:24-53
        if (p == 0) { fail(E_ADD_ZERO, -1, -1, -1, (int)c); stop = -1; break; }
        NumAdAtom++;
        if (NumAdAtom >= PV.NumMcMax) { fail(E_MC_FULL, NumAdAtom); stop = -1; break; }
        if (lane == 0) { PV.lattind[c] = (uint32_t)NumAdAtom; PV.mc_cell[NumAdAtom] = c; PV.mc_ev[NumAdAtom] = ev; pc_st(nlev, NumAdAtom - 1, p); sh.pend[npend] = NumAdAtom - 1; }
        npend++;
      }
      if (do_set) {                                // UpdateTree.f:25-48 + the record's event code
        if (ind < 1 || ind > PV.NumMcMax) { fail(E_ERASE, -1, -1, -1, ind); stop = -1; break; }
        if (p == 0 && ind < NumAdAtom) { fail(E_TREE_ZERO, ind); stop = -1; break; }
        if (ind > NumAdAtom) { fail(E_TREE_IDX, ind, NumAdAtom); stop = -1; break; }
        if (lane == 0) { PV.mc_ev[ind] = ev; pc_st(nlev, ind - 1, p); sh.pend[npend] = ind - 1; }
        npend++;
        if (self && act == A_SETBULK) NumAdAtomOcc++;              // la/mcevent/deposition.f:253
      }
      __syncwarp();
    }
    if (stop == -1) break;
    __syncwarp();
    pc_flush(lane, nlev, npend);
    done++;
    Iter++;
    if (__any_sync(FULL, D.err[0] != 0)) { stop = -1; break; }
    if (!A.exit_on_time) { if (Iter > A.max_iter) { stop = 1; break; } }
    else if (Time > A.max_time) { stop = 2; break; }
    if (fork && NumAtoms + NumAdAtomOcc > S.NumOcc0) { stop = 5; break; }       // la/main.f:714-723
  }
  __syncwarp();
  pc_flush(lane, nlev, npend);
  // ---- write the mirror and the loop state back
  for (int L = 0; L <= nlev; L++) for (int o = lane; o < pc_W(L, sh.wl, nlev); o += 32) PV.Tree[(1 << L) + o] = sh.tree[pc_base(L, sh.wl, nlev) + o];
  if (lane == 0) {
    pc_ungenerate(uavail);
    S.Iter = Iter; S.Time = Time; S.framecounter = framecounter; S.counter = counter;
    for (int i = 0; i < 56; i++) S.r_ma[i] = sh.ma[i];
    S.r_inext = sh.inext; S.r_inextp = sh.inextp; S.r_iff = sh.iff; S.idum = sh.idum;
    S.NumAdAtom = NumAdAtom; S.NumAtoms = NumAtoms; S.NumVoids = NumVoids; S.NumAdAtomOcc = NumAdAtomOcc;
    for (int i = 0; i < 3; i++) { S.CountCrystal[i] = CC[i]; S.CountCov[i] = CV[i]; }
    S.events_done = done; S.site_updates = nsu;
    S.stop_reason = stop;
    S.pending = pending ? 1 : 0; S.pend_ind = Ind; S.pend_itrans = Itrans;
  }
}

}  // namespace mskc
