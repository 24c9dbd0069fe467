// mskc_parallel.cuh -- synchronous-sublattice sampler + the dense helper kernels.
//
// The x-y plane is cut into square tiles (edge 24 super-lattice units = 2 cubic cells) coloured 2x2.
// An event at S writes S and its 4 neighbours (<= 1 bond, |d| <= 8 per axis incl. hexagonal bonds) and
// reads at most 2 bonds away (reciprocity of neighbours' neighbours, zig-zag/arm-chair candidates),
// so two tiles of the same colour (>= 24 units apart) never touch the same cell: all tiles of a colour
// run concurrently, one warp per tile, each doing its own n-fold way for a time quantum tau.
//
// HBM layout added to state[]/bonds[] (mskc_device.cuh):
//   rate[list][slot] f64, ce[list][slot] u64 {u32 cell, u32 event}: the per-tile active-site lists, tile-major
//   cnt[tile]           i32  slots in use (tombstones with rate 0 included until the owner compacts)
//   aux[cell]           u32  bits 0-15: 1-based slot of the cell in its tile's list (0 = not listed); bits 16-23: the
//                            cell's census flags (mskc_coop.cuh)
//   tstat[tile]         per-tile counters (events, site updates, per-class counts), summed on demand
//   lag[tile]           f64  time the tile still owes (its last colour step stopped at the event cap)
#pragma once
#include "mskc_device.cuh"
#include "mskc_coop.cuh"
#include "../../include/mskc.h"
#include <algorithm>
#include <dlfcn.h>
#ifndef MSKC_EMU
#include <nccl.h>
#endif
#include <cstdlib>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace mskc {

struct TileStat {            // written only by the tile's own lane
  long long events, site_updates;
  int dep[3], evap, abs_, des;
  int dAtoms, dVoids;        // bulk / vacancy list size changes
  int dCrystal[3]; int dCov[3];
  int maxev, pad;            // most events this tile executed in one colour step
};

struct SweepState {          // one per context, device resident
  double time, tau, total_rate, tree_root;
  long long n_active, sweeps, step;     // step = colour steps done before the current sweep (Philox counter word)
  long long max_cnt;
  // parameters of the current mskc_run_parallel call (written once per call, so that the kernels of a sweep take no
  // per-call arguments and a captured CUDA graph of one sweep can be replayed)
  uint2 key; int event_cap, pad;
  double events_per_tile, tau_fixed, ntiles_global;
  double tau_adaptive;                  // the quantum the rates ask for (tau itself may be the caller's tau_fixed)
};

// Per-tile active lists, tile-major: entry k of list l lives at [l*cap + k]; the warp that owns tile l
// streams its list into shared memory with coalesced loads at the start of its colour step.  Lists of one
// colour are contiguous (l = colour*ntc + index).
struct HaloRec;
struct ParallelCtx {
  int tile = 24, ntx = 0, nty = 0, ntiles = 0, cap = 0;
  double *rate = nullptr; unsigned long long *ce = nullptr; int *cnt = nullptr, *live = nullptr; uint32_t *aux = nullptr; double *tile_rate = nullptr; double *lag = nullptr;
  TileStat *tstat = nullptr; SweepState *ss = nullptr; TileStat *tot = nullptr;
  // slab mode
  int halo = 0, rank = 0, nranks = 1;
  uint32_t *dirty_bits = nullptr, *dirty_cell = nullptr; int *dirty_n = nullptr; int dirty_cap = 0; size_t dirty_words = 0;
  double *red = nullptr;
  unsigned char *sendbuf[2] = {nullptr, nullptr}, *recvbuf[2] = {nullptr, nullptr};
  void *comm = nullptr;                 // ncclComm_t
  unsigned int *pack_done = nullptr;
  // peer-to-peer halo exchange: my receive buffers [2 sides][2 parities] (one allocation, IPC-exported) and the two
  // neighbours' receive buffers mapped into this process
  HaloRec *p2p_recv = nullptr; void *p2p_left = nullptr, *p2p_right = nullptr; bool p2p = false; unsigned int halo_seq = 0;
  long long ncell = 0;
  long long base_atoms = 0, base_voids = 0;
  int baseCrystal[3] = {0, 0, 0}, baseCov[3] = {0, 0, 0};
  TileStat h_tot{}; SweepState h_ss{};
  TileStat extra{};                     // events applied outside the colour steps (trace replay), added to the counters
  double time = 0.0;
  bool ran = false, replayed = false, lists_stale = false, listed_flag_in_state = false, halo_pending = false;
#ifndef MSKC_EMU
  cudaGraphExec_t gexec = nullptr;        // one captured sweep (single GPU)
#endif
  void *g_kern = nullptr; int g_blocks = 0, g_launches = 0, g_sweeps = 0; size_t g_smem = 0;
};

struct TileView {            // what the kernels need (constant memory)
  int tile, ntx, nty, cap, NT, hx, ntc;
  double *rate; unsigned long long *ce; int *cnt, *live /* entries with rate > 0 as of the tile's last own step */; double *tile_rate; TileStat *tstat; SweepState *ss;
  double *lag;                          // [tile] KMC time a tile still owes because its colour step hit the event cap
  // x-decomposed slab (multi-GPU): one halo tile column on each side of the owned block, see "halo exchange" below
  int halo, nranks;
  uint32_t *dirty_bits;                 // bitmap over the two 48-unit shared strips (left: x < 48, right: x >= LX-48)
  uint32_t *dirty_cell; int *dirty_n; int dirty_cap;     // [2][dirty_cap] cells modified this colour step, [2] counts
  double *red;                          // [4] per-sweep sums that cross ranks: total rate, active sites, (spare)
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: no per-tile generator state in memory ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x, hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {       // uniform in (0,1)
  unsigned long long v = (((unsigned long long)a << 32) | b) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

__constant__ TileView V;      // the current context's tile lists (uploaded with the DevCtx)

enum { TILE = 24 };           // tile edge in super-lattice units (2 cubic cells)
__device__ __forceinline__ int list_of(P3 s) {          // list id of the tile that owns site s
  int tx = s.x / TILE, ty = s.y / TILE;
  return (((ty & 1) << 1) | (tx & 1)) * V.ntc + (ty >> 1) * V.hx + (tx >> 1);
}

// list entry k of tile-list l (tile-major: the owning team streams its own list with coalesced loads)
__device__ __forceinline__ long long lidx(int l, int k) { return (long long)l * V.cap + k; }

// One list entry: bits 0-31 cell, 32-39 event class, 40-44 / 45-49 x / y inside the tile, 50-63 z -- the site's
// coordinates come back without integer divisions when the entry fires.
__host__ __device__ __forceinline__ unsigned long long pack_ce(uint32_t cell, int ev, int x, int y, int z) {
  return (unsigned long long)cell | ((unsigned long long)(uint32_t)(ev & 255) << 32) | ((unsigned long long)(uint32_t)(x % TILE) << 40) |
         ((unsigned long long)(uint32_t)(y % TILE) << 45) | ((unsigned long long)(uint32_t)z << 50);
}
__host__ __device__ __forceinline__ int ce_event(unsigned long long e) { return (int)((e >> 32) & 255u); }

// List bookkeeping of the cooperative sampler.  The running tile's own list is mirrored in shared memory
// for the whole colour step (scans and updates never leave the SM) and written back once at the end;
// entries of cells that belong to other (not running) tiles go straight to those tiles' lists in HBM,
// slots claimed with atomics because two running tiles can border the same idle tile.
// Slab mode: remember every cell of the two shared strips whose record changed in this colour step (once
// per cell: bitmap + list), for halo_pack_kernel.
__device__ __noinline__ void mark_dirty(P3 s, uint32_t c) {
  const int side = s.x < 2 * TILE ? 0 : (s.x >= D.LX - 2 * TILE ? 1 : -1);
  if (side < 0) return;
  const int xs = side ? s.x - (D.LX - 2 * TILE) : s.x;
  const uint32_t bit = (uint32_t)(((side * 2 * TILE + xs) * D.LY + s.y) * (D.LZ + 1) + s.z);
  const uint32_t m = 1u << (bit & 31u);
  if (atomicOr(&V.dirty_bits[bit >> 5], m) & m) return;
  const int k = atomicAdd(&V.dirty_n[side], 1);
  if (k >= V.dirty_cap) { fail(E_TILE_FULL, s.x, s.y, s.z, -1 - side); return; }
  V.dirty_cell[(size_t)side * V.dirty_cap + k] = c;
}

struct TileLists {
  int own;                                   // list id of the running tile (-1: none, everything is "foreign")
  int mark;                                  // slab mode: log changed cells of the shared strips
  double *s_rate; unsigned long long *s_ce; int *s_n;
  __device__ __forceinline__ int slot_of(uint32_t c) const { return aux_slot(D.aux[c]); }
  // sl = the cell's current slot (slot_of(c)), read by the caller ahead of time.  The entry lives in the shared-memory mirror
  // when the cell belongs to the running tile, else in that tile's list in HBM (generic pointers: one store path).
  __device__ __forceinline__ void put(P3 s, uint32_t c, int ev, double p, int sl) {
    if (mark) mark_dirty(s, c);
    const int l = list_of(s);
    const bool mine = l == own;
    if (sl == 0) {
      int *cnt = mine ? s_n : &V.cnt[l];
      const int k = atomicAdd(cnt, 1);
      if (k >= V.cap) { atomicSub(cnt, 1); fail(E_TILE_FULL, s.x, s.y, s.z, l); return; }
      st_slot(c, k + 1); sl = k + 1;
    }
    double *rp = mine ? s_rate + (sl - 1) : V.rate + lidx(l, sl - 1);
    unsigned long long *cp = mine ? s_ce + (sl - 1) : V.ce + lidx(l, sl - 1);
    *rp = p; *cp = pack_ce(c, ev, s.x, s.y, s.z);
  }
  __device__ __forceinline__ void erase(P3 s, uint32_t c, int sl) {  // tombstone (rate 0); the owner compacts
    if (mark) mark_dirty(s, c);
    if (!sl) return;
    const int l = list_of(s);
    double *rp = l == own ? s_rate + (sl - 1) : V.rate + lidx(l, sl - 1);
    *rp = 0.0;
  }
};

// tuning knobs (overridable at compile time for A/B builds): resident blocks per SM the register allocation aims at,
// and the list length above which a tile drops its tombstones when it loads its list
#ifndef MSKC_CS_MINB
#define MSKC_CS_MINB 4
#endif
#ifndef MSKC_COMPACT_ABOVE
#define MSKC_COMPACT_ABOVE 64
#endif
enum { CS_THREADS = 128, CS_MINB = MSKC_CS_MINB };
enum { CS_COMPACT_ABOVE = MSKC_COMPACT_ABOVE };
#define CS_BURST 4.0                          /* a state expected to fire more than CS_BURST*cap events per quantum is transient */
#define CS_MAX_LAG 6.0                        /* quanta of debt beyond which a tile's event cap is lifted */               // tombstones are dropped when a list is loaded with more entries than this

// ---- one tile's colour step, as the team that owns it sees it -----------------------------------------------------
// Everything here is team-uniform (every lane of the team holds the same value) except K, the per-lane counters.
struct TileRun {
  int l, x0, y0;                 // list id, tile origin
  int nev, max_ev;               // events executed in this step, cap (see CS_BURST / CS_MAX_LAG)
  bool capped;
  uint32_t draw;                 // Philox draw counter inside the step
  unsigned long long step;       // global colour-step counter (Philox counter words)
  double tloc, budget, R;        // local clock, time budget of the step, current total rate of the tile
  double tau, debt;              // the quantum, and the time the tile owed when the step began
  int ndep0, ndep1, ndep2, nevap, nabs, ndes;
  CoopCounters K;
};
__device__ __forceinline__ void tile_reset(TileRun &t) {
  t.nev = 0; t.capped = false; t.draw = 0u; t.tloc = 0.0; t.R = 0.0;
  t.ndep0 = t.ndep1 = t.ndep2 = t.nevap = t.nabs = t.ndes = 0;
  t.K.nsu = 0; t.K.dAtoms = 0; t.K.dVoids = 0;
  for (int i = 0; i < 3; i++) { t.K.dCrystal[i] = 0; t.K.dCov[i] = 0; }
}
// Time budget of a step.  A step stops after event_cap events (the slowest tile of a colour bounds the whole
// launch, and event counts per step have a long tail: half of the tiles exceed four times the mean at least once
// in 32 sweeps); the unspent time becomes the tile's debt, paid back at no more than one extra quantum per step so
// that catching up rarely hits the cap again.  A tile that owes more than CS_MAX_LAG quanta may run twice the cap,
// beyond four times that it runs uncapped: the debt stays bounded.
__device__ __forceinline__ void tile_budget(TileRun &t, double tau, int event_cap) {
  const double lg = V.lag[t.l];
  t.tau = tau; t.debt = lg; t.budget = tau + fmin(lg, tau);
  t.max_ev = lg > 4.0 * CS_MAX_LAG * tau ? (1 << 14) : (lg > CS_MAX_LAG * tau ? 2 * event_cap : event_cap);
}
// load the tile's list into the team's shared-memory mirror (dropping tombstones when it has grown); team-divergent
template <int TEAM>
__device__ __forceinline__ void tile_load(const int l, const int tl, const int tshift, const unsigned tm, double *s_rate, unsigned long long *s_ce, int *s_n) {
  const unsigned tmask = tm << tshift;
  const int n0 = V.cnt[l];
  const bool compact = n0 > CS_COMPACT_ABOVE;
  int n = 0;
#pragma unroll 1
  for (int base = 0; base < n0; base += TEAM) {
    const int i = base + tl;
    double r = 0.0; unsigned long long e = 0ull;
    if (i < n0) { r = V.rate[lidx(l, i)]; e = V.ce[lidx(l, i)]; }
    if (!compact) { if (i < n0) { s_rate[i] = r; s_ce[i] = e; } }
    else {
      const bool live = (i < n0) && (r > 0.0);
      const unsigned m = (__ballot_sync(tmask, live) >> tshift) & tm;
      const int k = n + __popc(m & ((1u << tl) - 1u));
      if (live) { s_rate[k] = r; s_ce[k] = e; if (k != i) st_slot((uint32_t)e, k + 1); }
      else if (i < n0) st_slot((uint32_t)e, 0);
      n += __popc(m);
    }
  }
  if (!compact) n = n0;
  if (tl == 0) *s_n = n;
}
// One lock-step iteration: every lane of the warp calls this together; a team with active == false only takes part
// in the synchronisation.  Returns whether the team executed an event (false for an active team: its step is over).
// Selection (the n-fold way inside the tile): every lane sums a CONTIGUOUS chunk of the list mirror, a team-wide scan
// of the eight partial sums finds the lane whose chunk holds the target, and that lane walks its chunk.
template <int TEAM, int FEAT>
__device__ __forceinline__ bool tile_iteration(TileRun &t, TileLists &L, const bool active, const int tl, const int tshift, const unsigned tm,
                                               const uint2 key, const double tau, const int event_cap, const double root) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const double *s_rate = L.s_rate;
  const int n = active ? *L.s_n : 0;
  const int chunk = (n + TEAM - 1) / TEAM, i0 = tl * chunk, i1 = min(n, i0 + chunk);
  double part = 0.0;
#pragma unroll 1
  for (int i = i0; i < i1; i++) part += s_rate[i];
  double incl = part;
#pragma unroll
  for (int o = 1; o < TEAM; o <<= 1) { const double v = __shfl_up_sync(FULL, incl, o, TEAM); if (tl >= o) incl += v; }
  const double Rn = __shfl_sync(FULL, incl, TEAM - 1, TEAM);
  if (active) t.R = Rn;
  // the cap never freezes a tile in a transient state that lives much shorter than the quantum (an unstable atom
  // about to re-evaporate): such a state would also poison the total rate the next quantum is derived from
  const bool more = active && (Rn > 0.0) && t.nev < (1 << 14);
  bool go = more && (t.nev < t.max_ev || Rn * tau >= CS_BURST * (double)event_cap);
  if (more && !go) t.capped = true;
  // one counter-based draw per event: x -> waiting time, y + 20 bits of z -> target, 24 bits -> species, 20 bits -> arm-chair tie
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)t.l, (uint32_t)t.step, (uint32_t)(t.step >> 32), t.draw), key);
  if (go) {
    t.draw++;
    t.tloc += (double)__fdividef(-__logf(((float)rnd.x + 0.5f) * 2.3283064365386963e-10f), (float)Rn);
    if (t.tloc > t.budget) go = false;
  }
  const double target = ((double)(((unsigned long long)rnd.y << 20) | (rnd.z >> 12)) + 0.5) * (1.0 / 4503599627370496.0) * Rn;
  // the lane whose chunk contains the target (rounding: the last lane with rate mass), then that lane walks its chunk
  const unsigned hitm = (__ballot_sync(FULL, part > 0.0 && incl > target) >> tshift) & tm;
  const unsigned massm = (__ballot_sync(FULL, part > 0.0) >> tshift) & tm;
  const int sel = hitm ? (__ffs(hitm) - 1) : (massm ? 31 - __clz((int)massm) : 0);
  int pick = -1;
  if (go && tl == sel) {
    double acc = incl - part;
#pragma unroll 1
    for (int i = i0; i < i1; i++) {
      const double r = s_rate[i];
      if (r > 0.0) { pick = i; acc += r; if (acc > target) break; }
    }
  }
  pick = __shfl_sync(FULL, pick, sel, TEAM);
  if (go && pick < 0) go = false;                        // (cannot happen while Rn > 0)
  unsigned long long e = 0ull;
  if (go) e = L.s_ce[pick];
  const uint32_t cell = (uint32_t)e;
  P3 Site; Site.x = t.x0 + (int)((e >> 40) & 31u); Site.y = t.y0 + (int)((e >> 45) & 31u); Site.z = (int)(e >> 50);
  const double us = ((double)(((rnd.z & 0xFFFu) << 12) | (rnd.w >> 20)) + 0.5) * (1.0 / 16777216.0);   // species / adsorbate (Get_Prob.f:164-179)
  const double arm_u = ((double)(rnd.w & 0xFFFFFu) + 0.5) * (1.0 / 1048576.0);
  const int Itrans = coop_event<TEAM, FEAT>(L, tl, go, Site, cell, go ? ce_event(e) : 4, root, arm_u, -1, t.K, us);
  if (go) {
    t.ndep0 += Itrans == 1; t.ndep1 += Itrans == 2; t.ndep2 += Itrans == 3; t.nevap += Itrans == 4; t.ndes += Itrans == 5; t.nabs += Itrans > 5;
    t.nev++;
  }
  return go;
}
// End of a tile's step: every lane of the warp calls this together, the teams with fin == true write their list back
// and add their counters to the tile's statistics.
template <int TEAM>
__device__ __forceinline__ void tile_finish(TileRun &t, const TileLists &L, const bool fin, const int tl) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  __syncwarp();
  int v[9] = {t.K.nsu, t.K.dAtoms, t.K.dVoids, t.K.dCrystal[0], t.K.dCrystal[1], t.K.dCrystal[2], t.K.dCov[0], t.K.dCov[1], t.K.dCov[2]};
#pragma unroll
  for (int k = 0; k < 9; k++)
#pragma unroll
    for (int o = TEAM / 2; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o, TEAM);
  const unsigned fin_mask = __ballot_sync(FULL, fin);
  if (!fin) return;
  const int l = t.l, n = *L.s_n;
  int live = 0;
#pragma unroll 1
  for (int i = tl; i < n; i += TEAM) { const double r = L.s_rate[i]; live += r > 0.0; V.rate[lidx(l, i)] = r; V.ce[lidx(l, i)] = L.s_ce[i]; }
#pragma unroll
  for (int o = TEAM / 2; o > 0; o >>= 1) live += __shfl_xor_sync(fin_mask, live, o, TEAM);
  if (tl == 0) {
    V.cnt[l] = n; V.live[l] = live;
    V.tile_rate[l] = t.R;
    V.lag[l] = t.capped ? t.debt + t.tau - t.tloc : t.debt - fmin(t.debt, t.tau);
    TileStat &st = V.tstat[l];
    st.events += t.nev; st.site_updates += v[0];
    st.dep[0] += t.ndep0; st.dep[1] += t.ndep1; st.dep[2] += t.ndep2; st.evap += t.nevap; st.abs_ += t.nabs; st.des += t.ndes;
    st.dAtoms += v[1]; st.dVoids += v[2];
    for (int i = 0; i < 3; i++) { st.dCrystal[i] += v[3 + i]; st.dCov[i] += v[6 + i]; }
    if (t.nev > st.maxev) st.maxev = t.nev;
  }
}

// One colour step: every tile of colour `colour` advances by tau.  A TEAM of lanes owns one tile
// (mskc_coop.cuh); the 32/TEAM teams of a warp run their event loops in lock step: every lane of the warp
// executes every iteration (full-mask votes and shuffles are the convergence points), a team without an event
// only takes part in the synchronisation.
template <int TEAM, int FEAT>
__global__ void __launch_bounds__(CS_THREADS, CS_MINB) colour_step_kernel(int colour, int pos /* position of this colour step inside its sweep, 0..3 */) {
  constexpr int TPB = CS_THREADS / TEAM;
  constexpr unsigned FULL = 0xFFFFFFFFu;
  MSKC_DYN_SHARED(smem);
  const int lane = threadIdx.x & 31, tl = threadIdx.x % TEAM, tib = threadIdx.x / TEAM;
  const int tshift = lane - tl;
  const unsigned tm = TEAM == 32 ? 0xFFFFFFFFu : ((1u << (TEAM & 31)) - 1u);
  const int wid = blockIdx.x * TPB + tib;
  const int cap = V.cap;
  TileLists L; L.mark = V.halo;
  L.s_rate = reinterpret_cast<double *>(smem) + (size_t)tib * cap;
  L.s_ce = reinterpret_cast<unsigned long long *>(smem + (size_t)TPB * cap * 8) + (size_t)tib * cap;
  L.s_n = reinterpret_cast<int *>(smem + (size_t)TPB * cap * 16) + tib;
  TileRun t; tile_reset(t);
  t.l = colour * V.ntc + wid;
  t.x0 = (2 * (wid % V.hx) + (colour & 1)) * TILE; t.y0 = (2 * (wid / V.hx) + (colour >> 1)) * TILE;   // tile origin (list_of inverted)
  // a team beyond the last tile, or on a halo tile (it belongs to the neighbouring rank), idles through the step
  const bool exists = wid < V.ntc && !(V.halo && (t.x0 == 0 || t.x0 == D.LX - TILE));
  const SweepState ss = *V.ss;
  const uint2 key = ss.key; const int event_cap = ss.event_cap;
  t.step = (unsigned long long)ss.step + (unsigned long long)pos; t.budget = t.tau = ss.tau; t.debt = 0.0; t.max_ev = event_cap;
  L.own = exists ? t.l : -2;
  if (exists) { tile_budget(t, ss.tau, event_cap); tile_load<TEAM>(t.l, tl, tshift, tm, L.s_rate, L.s_ce, L.s_n); }
  __syncwarp();
  bool active = exists;
  while (__any_sync(FULL, active)) active = tile_iteration<TEAM, FEAT>(t, L, active, tl, tshift, tm, key, ss.tau, event_cap, ss.tree_root);
  tile_finish<TEAM>(t, L, exists, tl);
}

// ---- halo exchange (x-decomposed slabs) ------------------------------------------------------
// Every rank holds its block plus one tile column (24 units) of each neighbour.  Tiles of one colour on both
// sides of a rank boundary never touch the same cell (section 5.3 of DESIGN.md), so after a colour step the two
// copies of a shared strip differ only where ONE side wrote: each rank sends the final record of every cell it
// changed in its two 48-unit strips (its border tile column + its halo column) and applies what it receives.
enum { HALO_SPIN_MAX = 1 << 24 };  // polls of a neighbour's header before halo_apply_kernel gives up (each a system-scope load: ~0.5-1 us)
struct HaloRec { uint32_t cell, state; unsigned long long bonds; double rate; uint32_t ev, pad /* census flags */; };     // 32 B
// one launch packs both strips (blockIdx.y = side); the last block to finish clears the dirty counters
// One launch packs both strips (blockIdx.y = side).  out0 / out1 may live in the NEIGHBOUR's memory (peer-mapped over
// NVLink, see p2p_connect): the records are stored straight into its receive buffer, and the last block to finish
// publishes both headers {record count, seq} with one 64-bit store each after a system-wide fence -- the header is the
// flag the neighbour's halo_apply_kernel waits on.  It also clears the dirty counters.
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
#ifndef MSKC_EMU
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
#ifndef MSKC_EMU
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
#else
  return *p;
#endif
}
__global__ void __launch_bounds__(256) halo_pack_kernel(int owned /* width of the owned block */, HaloRec *out0, HaloRec *out1, unsigned int *done, unsigned int seq) {
  const int side = blockIdx.y;
  HaloRec *out = side ? out1 : out0;
  const int shift = side ? -owned : owned;
  const int n = min(V.dirty_n[side], V.dirty_cap);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint32_t c = V.dirty_cell[(size_t)side * V.dirty_cap + k];
    const P3 s = cell_xyz((long long)c);
    const int xs = side ? s.x - (D.LX - 2 * TILE) : s.x;
    const uint32_t bit = (uint32_t)(((side * 2 * TILE + xs) * D.LY + s.y) * (D.LZ + 1) + s.z);
    atomicAnd(&V.dirty_bits[bit >> 5], ~(1u << (bit & 31u)));
    HaloRec r; r.cell = (uint32_t)((long long)c + shift); r.state = D.state[c]; r.bonds = D.bonds[c]; r.rate = 0.0; r.ev = 0; r.pad = 0;
    const uint32_t ax = D.aux[c]; const int sl = aux_slot(ax);
    r.pad = aux_flags(ax);
    if (sl) { const long long a = lidx(list_of(s), sl - 1); r.rate = V.rate[a]; r.ev = (uint32_t)ce_event(V.ce[a]); }
    out[1 + k] = r;
  }
  __threadfence_system();                                  // this thread's records are visible to the peer before the header is
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1u) == gridDim.x * gridDim.y - 1) {
      __threadfence_system();
      const unsigned long long n0 = (unsigned long long)min(V.dirty_n[0], V.dirty_cap), n1 = (unsigned long long)min(V.dirty_n[1], V.dirty_cap);
      st_release_sys(reinterpret_cast<unsigned long long *>(out0), n0 | ((unsigned long long)seq << 32));
      st_release_sys(reinterpret_cast<unsigned long long *>(out1), n1 | ((unsigned long long)seq << 32));
      V.dirty_n[0] = 0; V.dirty_n[1] = 0; *done = 0u;
    }
  }
}
// one launch applies both received buffers (blockIdx.y selects the buffer); every block first waits for the header of
// its buffer to carry this exchange's sequence number (the neighbour's pack kernel has finished writing it)
__global__ void __launch_bounds__(256) halo_apply_kernel(const HaloRec *in0, const HaloRec *in1, unsigned int seq) {
  const HaloRec *in = blockIdx.y ? in1 : in0;
  __shared__ int s_n;
  if (threadIdx.x == 0) {
    // (bounded: a neighbour that stopped with an error, or whose process is gone, never publishes this header -- after
    // HALO_SPIN_MAX polls, ~10 s, this rank stops with E_HALO_TIMEOUT instead of spinning on the GPU for ever; once an error
    // is recorded no further exchange waits at all)
    unsigned long long h = ld_acquire_sys(reinterpret_cast<const unsigned long long *>(in));
    long long spins = 0;
    while ((unsigned int)(h >> 32) != seq && D.err[0] == 0 && ++spins < HALO_SPIN_MAX) h = ld_acquire_sys(reinterpret_cast<const unsigned long long *>(in));
    if ((unsigned int)(h >> 32) != seq) { if (D.err[0] == 0) fail(E_HALO_TIMEOUT, (int)seq, (int)(unsigned int)(h >> 32), (int)blockIdx.y); h = 0ull; }
    s_n = (int)(unsigned int)h;
  }
  __syncthreads();
  const int n = s_n;
  TileLists L; L.own = -1; L.mark = 0; L.s_rate = nullptr; L.s_ce = nullptr; L.s_n = nullptr;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const HaloRec r = in[1 + k];
    D.state[r.cell] = r.state; D.bonds[r.cell] = r.bonds;
    const P3 s = cell_xyz((long long)r.cell);
    const int sl = aux_slot(D.aux[r.cell]);
    st_flags(r.cell, r.pad);
    if (r.rate > 0.0) L.put(s, r.cell, (int)r.ev, r.rate, sl); else L.erase(s, r.cell, sl);
  }
}

// Drop the tombstones of every list (one warp per tile) -- the same compaction the colour-step kernel does
// when it loads a long list; used between replay chunks, where no tile is ever "running".
__global__ void __launch_bounds__(128) compact_lists_kernel(int ntiles) {
  const int lane = threadIdx.x & 31, l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (l >= ntiles) return;
  const int n0 = V.cnt[l];
  int n = 0;
  for (int base = 0; base < n0; base += 32) {       // in place: entry k <= i is written after entry i was read by this warp
    const int i = base + lane;
    double r = 0.0; unsigned long long e = 0ull;
    if (i < n0) { r = V.rate[lidx(l, i)]; e = V.ce[lidx(l, i)]; }
    const bool live = (i < n0) && (r > 0.0);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    const int k = n + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (live) { V.rate[lidx(l, k)] = r; V.ce[lidx(l, k)] = e; if (k != i) st_slot((uint32_t)e, k + 1); }
    else if (i < n0) st_slot((uint32_t)e, 0);
    n += __popc(m);
    __syncwarp();
  }
  if (lane == 0) V.cnt[l] = n;
}

// Slab mode: the two halo tile columns mirror the neighbours' border tiles.  They never run here, so tile_load never
// drops their tombstones; one warp per halo tile does it once per sweep (lists of idle tiles, nothing else touches them
// between two kernels of the stream).
__global__ void __launch_bounds__(128) compact_halo_kernel() {
  const int lane = threadIdx.x & 31, w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= 2 * V.nty) return;
  const int tx = w < V.nty ? 0 : V.ntx - 1, ty = w % V.nty;
  const int l = (((ty & 1) << 1) | (tx & 1)) * V.ntc + (ty >> 1) * V.hx + (tx >> 1);
  const int n0 = V.cnt[l];
  if (n0 <= CS_COMPACT_ABOVE / 2) return;
  int n = 0;
  for (int base = 0; base < n0; base += 32) {
    const int i = base + lane;
    double r = 0.0; unsigned long long e = 0ull;
    if (i < n0) { r = V.rate[lidx(l, i)]; e = V.ce[lidx(l, i)]; }
    const bool live = (i < n0) && (r > 0.0);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    const int k = n + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (live) { V.rate[lidx(l, k)] = r; V.ce[lidx(l, k)] = e; if (k != i) st_slot((uint32_t)e, k + 1); }
    else if (i < n0) st_slot((uint32_t)e, 0);
    n += __popc(m);
    __syncwarp();
  }
  if (lane == 0) V.cnt[l] = n;
}

// Trace replay (test/diagnostic entry point): ONE warp applies a given event sequence -- site, event code
// and the IF_Arm_Choice outcome of every event, as the CPU oracle traced them -- through the same
// cooperative transition code the sampler uses.  Lists are updated in HBM only.
__global__ void __launch_bounds__(32) replay_kernel(const int32_t *ev_site, const int32_t *ev_code, const int32_t *ev_arm, long long n, double root, TileStat *out) {
  const int lane = threadIdx.x;
  TileLists L; L.own = -1; L.mark = 0; L.s_rate = nullptr; L.s_ce = nullptr; L.s_n = nullptr;
  CoopCounters K; K.nsu = 0; K.dAtoms = 0; K.dVoids = 0;
  for (int i = 0; i < 3; i++) { K.dCrystal[i] = 0; K.dCov[i] = 0; }
  long long done = 0;
  const bool fork = D.la == 3;
  int cl0[3] = {0, 0, 0};
  if (fork) { cl0[0] = g_la_CL[0]; cl0[1] = g_la_CL[1]; cl0[2] = g_la_CL[2]; }
  for (long long i = 0; i < n; i++) {
    P3 S; S.x = ev_site[3 * i]; S.y = ev_site[3 * i + 1]; S.z = ev_site[3 * i + 2];
    const int arm = ev_arm[i];
    if (fork) {                                    // la/main.f:636: CountLiquid as of the top of this iteration
      int c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] = cl0[k] - __reduce_add_sync(0xFFFFFFFFu, K.dCrystal[k]);
      if (lane == 0) { volatile int *cl = g_la_CL; cl[0] = c[0]; cl[1] = c[1]; cl[2] = c[2]; cl[3] = c[0] + c[1] + c[2]; }
      __syncwarp();
    }
    coop_event<32, F_COV | F_LA>(L, lane, true, S, (uint32_t)cidx(S), ev_code[i], root, 0.5, arm == 0 ? -1 : arm - 1, K);
    done++;
    if (D.err[0] != 0) break;
  }
  int v[9] = {K.nsu, K.dAtoms, K.dVoids, K.dCrystal[0], K.dCrystal[1], K.dCrystal[2], K.dCov[0], K.dCov[1], K.dCov[2]};
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = __reduce_add_sync(0xFFFFFFFFu, v[k]);
  if (lane == 0) {
    if (fork) { volatile int *cl = g_la_CL; cl[0] = cl0[0] - v[3]; cl[1] = cl0[1] - v[4]; cl[2] = cl0[2] - v[5]; cl[3] = cl[0] + cl[1] + cl[2]; }   // for the next chunk
    out->events += done; out->site_updates += v[0]; out->dAtoms += v[1]; out->dVoids += v[2];
    for (int i = 0; i < 3; i++) { out->dCrystal[i] += v[3 + i]; out->dCov[i] += v[6 + i]; }
  }
}

// Trace replay through the instantiations the sampler itself runs: 32 / TEAM teams of TEAM lanes in lock step (TEAM = 8: the
// throughput configuration; 16: what mskc_run_parallel picks when a colour has few tiles, e.g. the strong-scaling slabs), FEAT as
// mskc_run_parallel picks it.  batch[4*b + t] = index of the event team t applies in lock-step iteration b, or -1 (the
// team idles through the iteration, as a tile without an event does).  Events of one iteration commute (the host puts
// only events more than two tile edges apart into one iteration), so the outcome equals the serial trace.
template <int TEAM, int FEAT>
__global__ void __launch_bounds__(32) replay_teams_kernel(const int32_t *ev_site, const int32_t *ev_code, const int32_t *ev_arm, const int32_t *batch, long long nbatch, double root, TileStat *out) {
  const int lane = threadIdx.x, team = lane / TEAM, tl = lane % TEAM;
  TileLists L; L.own = -1; L.mark = 0; L.s_rate = nullptr; L.s_ce = nullptr; L.s_n = nullptr;
  CoopCounters K; K.nsu = 0; K.dAtoms = 0; K.dVoids = 0;
  for (int i = 0; i < 3; i++) { K.dCrystal[i] = 0; K.dCov[i] = 0; }
  long long done = 0;
  for (long long b = 0; b < nbatch; b++) {
    const int i = batch[4 * b + team];
    const bool go = i >= 0;
    P3 S; S.x = S.y = S.z = 0; int code = 4, arm = 0;
    if (go) { S.x = ev_site[3 * i]; S.y = ev_site[3 * i + 1]; S.z = ev_site[3 * i + 2]; code = ev_code[i]; arm = ev_arm[i]; }
    coop_event<TEAM, FEAT>(L, tl, go, S, go ? (uint32_t)cidx(S) : 0u, code, root, 0.5, arm == 0 ? -1 : arm - 1, K);
    if (go && tl == 0) done++;
    if (__any_sync(0xFFFFFFFFu, D.err[0] != 0)) break;
  }
  int v[9] = {K.nsu, K.dAtoms, K.dVoids, K.dCrystal[0], K.dCrystal[1], K.dCrystal[2], K.dCov[0], K.dCov[1], K.dCov[2]};
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = __reduce_add_sync(0xFFFFFFFFu, v[k]);
  const long long dn = (long long)__reduce_add_sync(0xFFFFFFFFu, (int)done);
  if (lane == 0) {
    out->events += dn; out->site_updates += v[0]; out->dAtoms += v[1]; out->dVoids += v[2];
    for (int i = 0; i < 3; i++) { out->dCrystal[i] += v[3 + i]; out->dCov[i] += v[6 + i]; }
  }
}

__device__ __forceinline__ bool halo_list(int l) {        // is tile-list l one of the two halo tile columns?
  if (!V.halo) return false;
  const int colour = l / V.ntc, idx = l - colour * V.ntc;
  const int tx = 2 * (idx % V.hx) + (colour & 1);
  return tx == 0 || tx == V.ntx - 1;
}
// Per-sweep bookkeeping ("per-class rate sums").  Stage 1: this rank's total rate, active-site count and number of tiles that
// carry rate -> V.red[0..2], the clipped rate the quantum is sized on -> V.red[3].  SUM_BLOCKS blocks each reduce a strided share of the tiles into V.red[8 + 5b ..]; the last
// block to finish adds the partial sums in block order (a fixed order: the totals do not depend on which block came last) and,
// on a single rank, goes straight on to stage 2.  (One block of 1024 threads over the 3.7e4 tiles of the bench box took 24 us
// per sweep -- 4 % of a sweep; this takes ~4.)
enum { SUM_BLOCKS = 64 /* = 2 x 32: the last block's warp 0 takes two partial sums per lane */, RED_DOUBLES = 8 + 5 * SUM_BLOCKS };
// stage 2 (after the all-reduce of V.red over the ranks, if any): clock, next tau, global-coupling terms
__device__ __forceinline__ void sweep_finish(int advance) {
  SweepState &s = *V.ss;
  const double R = V.red[0];
  if (advance) { s.time += (double)advance * s.tau; s.sweeps += advance; s.step += 4 * advance; }      // `advance` sweeps ran on this tau
  s.total_rate = R; s.tree_root = R; s.n_active = (long long)V.red[1];
  // the quantum: events_per_tile events expected per colour step in a tile that carries rate (a nucleus or a ridge
  // occupies few tiles of its box: sizing tau on all tiles would multiply the events per busy tile, i.e. the bias)
  const double nt = V.red[2] > 0.0 ? V.red[2] : s.ntiles_global;
  // ... on the sum of rates with every tile's share clipped at the kernel's own definition of a transient state (see
  // sweep_sum_kernel): one unstable atom about to re-evaporate carries more rate than the whole rest of a 10^9-site lattice,
  // lives for a millionth of a quantum, and would otherwise halve tau for the sweep after every step that happened to end on one
  const double Rq = V.red[3] > 0.0 ? V.red[3] : R;
  s.tau_adaptive = Rq > 0 ? s.events_per_tile * nt / Rq : 0.0;
  s.tau = s.tau_fixed > 0 ? s.tau_fixed : s.tau_adaptive;
}
__global__ void __launch_bounds__(256) sweep_sum_kernel(int ntiles, int finish_advance /* >= 0: single rank, run stage 2 with this `advance`; < 0: leave it to sweep_finish_kernel */) {
  __shared__ double sR[8], sQ[8]; __shared__ long long sN[8]; __shared__ int sM[8], sT[8]; __shared__ bool s_last;
  double r = 0.0, rq = 0.0; long long n = 0; int mx = 0, nt = 0;        // nt: tiles that carry rate (the tiles the quantum is sized for)
  // rq: the rate the next quantum is sized on -- a tile expected to fire more than CS_BURST * event_cap events in the quantum
  // just used is in a transient state (the colour step lifts its event cap for exactly that case) and counts with that bound
  const SweepState &ss = *V.ss;
  const double tq = fmax(ss.tau, ss.tau_adaptive);            // (tau itself may be a caller's short landing quantum)
  const double qcap = tq > 0.0 ? CS_BURST * (double)ss.event_cap / tq : 1.0e300;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) { if (halo_list(t)) continue; const double q = V.tile_rate[t]; r += q; rq += fmin(q, qcap); nt += q > 0.0; n += V.live[t]; mx = max(mx, V.cnt[t]); }
  for (int o = 16; o > 0; o >>= 1) { r += __shfl_down_sync(0xFFFFFFFFu, r, o); rq += __shfl_down_sync(0xFFFFFFFFu, rq, o); n += __shfl_down_sync(0xFFFFFFFFu, n, o); mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, o)); nt += __shfl_down_sync(0xFFFFFFFFu, nt, o); }
  if ((threadIdx.x & 31) == 0) { sR[threadIdx.x >> 5] = r; sQ[threadIdx.x >> 5] = rq; sN[threadIdx.x >> 5] = n; sM[threadIdx.x >> 5] = mx; sT[threadIdx.x >> 5] = nt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double R = 0, Q = 0; long long N = 0; int M = 0, T = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { R += sR[i]; Q += sQ[i]; N += sN[i]; M = max(M, sM[i]); T += sT[i]; }
    double *part = V.red + 8 + 5 * blockIdx.x;
    part[0] = R; part[1] = (double)N; part[2] = (double)T; part[3] = (double)M; part[4] = Q;
    __threadfence();
    unsigned int *done = reinterpret_cast<unsigned int *>(V.red + 4);
    s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    if (s_last) *done = 0u;
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  // the last block: warp 0 adds the SUM_BLOCKS partial sums -- lane l takes blocks l and l + 32, then a shuffle tree: a fixed
  // order whichever block came last (64 dependent L2 reads by one thread took 15 us)
  __threadfence();
  const volatile double *q0 = V.red + 8 + 5 * threadIdx.x, *q1 = V.red + 8 + 5 * (threadIdx.x + 32);
  double tR = q0[0] + q1[0], tN = q0[1] + q1[1], tT = q0[2] + q1[2], tM = fmax(q0[3], q1[3]), tQ = q0[4] + q1[4];
  for (int o = 16; o > 0; o >>= 1) {
    tR += __shfl_down_sync(0xFFFFFFFFu, tR, o); tN += __shfl_down_sync(0xFFFFFFFFu, tN, o); tT += __shfl_down_sync(0xFFFFFFFFu, tT, o);
    tM = fmax(tM, __shfl_down_sync(0xFFFFFFFFu, tM, o)); tQ += __shfl_down_sync(0xFFFFFFFFu, tQ, o);
  }
  if (threadIdx.x == 0) {
    V.red[0] = tR; V.red[1] = tN; V.red[2] = tT; V.red[3] = tQ;
    V.ss->max_cnt = (long long)tM;
    if (finish_advance >= 0) sweep_finish(finish_advance);
  }
}
__global__ void sweep_finish_kernel(int advance) { sweep_finish(advance); }
__global__ void set_params_kernel(uint2 key, int event_cap, double events_per_tile, double tau_fixed, double ntiles_global) {
  SweepState &s = *V.ss;
  s.key = key; s.event_cap = event_cap; s.events_per_tile = events_per_tile; s.tau_fixed = tau_fixed; s.ntiles_global = ntiles_global;
}
__global__ void reset_maxev_kernel(TileStat *ts, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) ts[i].maxev = 0; }

__global__ void __launch_bounds__(256) tile_rate_kernel(int ntiles) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= ntiles) return;
  int n = V.cnt[l], live = 0; double r = 0;
  if (!halo_list(l)) for (int i = 0; i < n; i++) { const double v = V.rate[lidx(l, i)]; r += v; live += v > 0.0; }
  V.tile_rate[l] = r; V.live[l] = live;
}

// per-call counter reduction: 19 integer columns of every tile's statistics, summed into *out (zeroed by the caller; integer
// atomics, so the totals do not depend on the order).  Many blocks: one block over the 3.7e4 tiles of the bench box took
// 150 us per call, 4 % of an 8-sweep call of the host program.
__global__ void __launch_bounds__(256) reduce_tstat_kernel(const TileStat *ts, int ntiles, TileStat *out) {
  __shared__ long long acc[17];
  if (threadIdx.x < 17) acc[threadIdx.x] = 0;
  __syncthreads();
  long long loc[17];
  for (int k = 0; k < 17; k++) loc[k] = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
    const TileStat &s = ts[t];
    loc[0] += s.events; loc[1] += s.site_updates; loc[2] += s.dep[0]; loc[3] += s.dep[1]; loc[4] += s.dep[2]; loc[5] += s.evap; loc[6] += s.abs_; loc[7] += s.des;
    loc[8] += s.dAtoms; loc[9] += s.dVoids; if (s.maxev > loc[16]) loc[16] = s.maxev; loc[10] += s.dCrystal[0]; loc[11] += s.dCrystal[1]; loc[12] += s.dCrystal[2]; loc[13] += s.dCov[0]; loc[14] += s.dCov[1]; loc[15] += s.dCov[2];
  }
  for (int k = 0; k < 17; k++) {
    long long v = loc[k];
    for (int o = 16; o > 0; o >>= 1) { const long long u = __shfl_down_sync(0xFFFFFFFFu, v, o); v = k == 16 ? (u > v ? u : v) : v + u; }
    if ((threadIdx.x & 31) == 0) { if (k == 16) atomicMax((long long *)&acc[16], v); else atomicAdd((unsigned long long *)&acc[k], (unsigned long long)v); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd((unsigned long long *)&out->events, (unsigned long long)acc[0]); atomicAdd((unsigned long long *)&out->site_updates, (unsigned long long)acc[1]);
    atomicAdd(&out->dep[0], (int)acc[2]); atomicAdd(&out->dep[1], (int)acc[3]); atomicAdd(&out->dep[2], (int)acc[4]);
    atomicAdd(&out->evap, (int)acc[5]); atomicAdd(&out->abs_, (int)acc[6]); atomicAdd(&out->des, (int)acc[7]); atomicAdd(&out->dAtoms, (int)acc[8]); atomicAdd(&out->dVoids, (int)acc[9]);
    for (int i = 0; i < 3; i++) { atomicAdd(&out->dCrystal[i], (int)acc[10 + i]); atomicAdd(&out->dCov[i], (int)acc[13 + i]); }
    atomicMax(&out->maxev, (int)acc[16]);
  }
}

// ---- the dense "rate kernel": recompute (event class, rate) of every listed site from the lattice ----
__global__ void __launch_bounds__(256) rerate_kernel(int ntiles, unsigned long long *n_sites) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)ntiles * V.cap;
  const double root = V.ss->tree_root;
  unsigned long long mine = 0;
  for (long long k = gid; k < total; k += (long long)gridDim.x * blockDim.x) {
    int l = (int)(k / V.cap), i = (int)(k - (long long)l * V.cap);
    if (i >= V.cnt[l]) continue;
    double r = V.rate[k];
    if (!(r > 0.0)) continue;
    uint32_t cell = (uint32_t)V.ce[k];
    P3 s = cell_xyz((long long)cell);
    uint32_t w = D.state[cell];
    int ev = 0; double p = 0;
    const unsigned long long b = D.bonds[cell];
    st_flags(cell, census_flags(s, w, b));
    coop_rate<F_COV | F_LA>(w_occ(w), w_sp(w), w_occ(w) ? 0 : w_cov(w), w_coor(w), s, cell, b, root, ev, p);
    if (ev != 0 && p > 0) { V.rate[k] = p; V.ce[k] = pack_ce(cell, ev, s.x, s.y, s.z); mine++; }
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_sites, mine);
}

// ---- the dense census pass: census flags of every cell that carries bonds (after an upload / a device-side build;
// mode 1 = check: count cells whose stored flags differ from the census, and cells whose flag count differs from the
// coordination in their lattice word) -- Get_Prob.f:216-291 as one streaming pass over bonds[] (8 B/cell)
__global__ void __launch_bounds__(256) census_flags_kernel(int check, unsigned long long *out /* [3]: sites, flag mismatches, coordination mismatches */) {
  unsigned long long nsite = 0, nbad = 0, ncoor = 0;
  auto one = [&](unsigned long long b, long long c) {
    if (b == 0ull) return;
    const uint32_t w = D.state[c];
    const P3 q = cell_xyz(c);
    const unsigned fl = census_flags(q, w, b);
    nsite++;
    if (check) { if (fl != aux_flags(D.aux[c])) nbad++; }
    else st_flags((uint32_t)c, fl);
    // consistency of the lattice word with the census (Get_Prob.f:293), where every bond stays inside the box
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 4; i++) { const unsigned f = fld(b, i); if (f && (unsigned)(q.z + cdz(f)) > (unsigned)D.LZ) inside = false; }
    const unsigned nz = (fl | (fl >> 1)) & 0x55u;
    if (inside && w_cov(w) != 113 && __popc(nz) != w_coor(w)) ncoor++;
  };
  const long long n2 = D.ncell >> 1, stride = (long long)gridDim.x * blockDim.x;
  const ulonglong2 *B2 = reinterpret_cast<const ulonglong2 *>(D.bonds);
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n2; q0 += 4 * stride) {      // four 128-bit loads in flight
    ulonglong2 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const long long q = q0 + u * stride; v[u] = q < n2 ? B2[q] : make_ulonglong2(0ull, 0ull); }
#pragma unroll
    for (int u = 0; u < 4; u++) { const long long c = 2 * (q0 + u * stride); one(v[u].x, c); one(v[u].y, c + 1); }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (D.ncell & 1)) one(D.bonds[D.ncell - 1], D.ncell - 1);
  for (int o = 16; o > 0; o >>= 1) { nsite += __shfl_down_sync(0xFFFFFFFFu, nsite, o); nbad += __shfl_down_sync(0xFFFFFFFFu, nbad, o); ncoor += __shfl_down_sync(0xFFFFFFFFu, ncoor, o); }
  if ((threadIdx.x & 31) == 0) { if (nsite) atomicAdd(&out[0], nsite); if (nbad) atomicAdd(&out[1], nbad); if (ncoor) atomicAdd(&out[2], ncoor); }
}

// ---- helper kernels used by mskc_api.cu ----------------------------------------------------------
__global__ void scatter_records(unsigned long long *bonds, uint32_t *lattind, const uint32_t *cell, const unsigned long long *b, const uint32_t *ind, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { bonds[cell[i]] = b[i]; lattind[cell[i]] = ind[i]; }
}
__global__ void gather_bonds(const unsigned long long *bonds, const uint32_t *cell, unsigned long long *out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cell[i] == 0xFFFFFFFFu ? 0ull : bonds[cell[i]];
}
__global__ void collect_bulk(uint32_t *out, long long cap, unsigned long long *cnt) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < D.ncell; c += (long long)gridDim.x * blockDim.x) {
    uint32_t w = D.state[c];
    if (w_occ(w) && w_coor(w) == 4 && w_cov(w) != 113) { unsigned long long k = atomicAdd(cnt, 1ull); if ((long long)k < cap) out[k] = (uint32_t)c; }
  }
}
// initial/SetT.f:36-46: bits 9-21 of every word := T (10 = wall, left alone).  Four cells per thread and iteration
// (128-bit loads of the field and the words); TT = int32 (the reference's array) or uint16 (half the upload)
__device__ __forceinline__ uint32_t with_T(uint32_t w, int t) { return t != 10 ? MVBITS(w, t, LenT, PosT) : w; }
template <class TT>
__global__ void __launch_bounds__(256) set_temperature_kernel(const TT *T, long long n) {
  const long long n4 = n >> 2;
  uint4 *S4 = reinterpret_cast<uint4 *>(D.state);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    int t0, t1, t2, t3;
    if (sizeof(TT) == 4) { const uint4 tv = reinterpret_cast<const uint4 *>(T)[q]; t0 = (int)tv.x; t1 = (int)tv.y; t2 = (int)tv.z; t3 = (int)tv.w; }
    else { const uint2 tv = reinterpret_cast<const uint2 *>(T)[q]; t0 = (int)(tv.x & 0xFFFFu); t1 = (int)(tv.x >> 16); t2 = (int)(tv.y & 0xFFFFu); t3 = (int)(tv.y >> 16); }
    uint4 v = S4[q];
    v.x = with_T(v.x, t0); v.y = with_T(v.y, t1); v.z = with_T(v.z, t2); v.w = with_T(v.w, t3);
    S4[q] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (long long c = n4 << 2; c < n; c++) D.state[c] = with_T(D.state[c], (int)T[c]);
}
// the phase array of the laser-annealing coupling (la/main.f:806-814): coordination of every occupied cell, 10 for a
// wall, 0 elsewhere -- one byte per cell instead of the whole lattice word
__global__ void __launch_bounds__(256) phases_kernel(unsigned char *out, long long n) {
  const long long n4 = n >> 2;
  const uint4 *S4 = reinterpret_cast<const uint4 *>(D.state);
  auto ph = [](uint32_t w) -> uint32_t { return w_occ(w) ? (w_cov(w) == 113 ? 10u : (uint32_t)w_coor(w)) : 0u; };
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const uint4 v = S4[q];
    reinterpret_cast<uint32_t *>(out)[q] = ph(v.x) | (ph(v.y) << 8) | (ph(v.z) << 16) | (ph(v.w) << 24);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (long long c = n4 << 2; c < n; c++) out[c] = (unsigned char)ph(D.state[c]);
}
__device__ __forceinline__ bool is_cubic_site(int x, int y, int z, int &kind) {
  if (x % 3 || y % 3 || z % 3) return false;
  int x1 = x / 3, y1 = y / 3, z1 = z / 3;
  if ((x1 & 1) != (y1 & 1) || (x1 & 1) != (z1 & 1)) return false;
  int m = (x1 + y1 + z1) & 3;
  if (m == 0) { kind = 0; return true; }
  if (m == 3) { kind = 1; return true; }
  return false;
}
// Streaming census (the counting scans of inpout/WriteMolMolXYZFile.f:77-119 fused into one pass,
// 128-bit loads, 4 B/cell of HBM traffic): under-coordinated atoms, coverage, non-epitaxial bulk, ...
__global__ void __launch_bounds__(256) census_kernel(unsigned long long *out, double *hsum) {
  unsigned long long nUnder = 0, nCov = 0, nNoEpi = 0, nBulk = 0, nOcc = 0, nWall = 0; double hz = 0, hz2 = 0;
  const long long n4 = D.ncell >> 2;
  const uint4 *S4 = reinterpret_cast<const uint4 *>(D.state);
  const unsigned plane = (unsigned)D.LX * (unsigned)D.LY, LXu = (unsigned)D.LX, LZm1 = (unsigned)(D.LZ - 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // (cell indices fit 32 bits, DESIGN.md section 2: the coordinates of a non-empty cell cost two 32-bit divisions, and only
  // the classes that need them pay)
  auto one = [&](uint32_t w, long long cl) {
    if (w == 0) return;
    const unsigned c = (unsigned)cl;
    if (V.halo) { const unsigned x = c % LXu; if (x < (unsigned)TILE || x >= LXu - (unsigned)TILE) return; }      // halo columns belong to the neighbours
    const int cov = w_cov(w);
    if (cov == 113) { nWall++; return; }
    // plane ranges as the reference scans them: census 1 (atoms with coordination <= 3, coverage sites) over planes
    // 1..LenZ-1 (WriteMolMolXYZFile.f:77-92), census 2 (non-epitaxial bulk) over planes 0..LenZ-1 (:95-119)
    if (w_occ(w)) {
      nOcc++;
      const int co = w_coor(w);
      if (co > 4) return;
      const unsigned z = c / plane;
      if (co <= 3) { if (z - 1u < LZm1) { nUnder++; hz += (double)z; hz2 += (double)z * (double)z; } }
      else {
        nBulk++;
        const unsigned r = c - z * plane, y = r / LXu, x = r - y * LXu; int kind;
        if (z <= LZm1 && !is_cubic_site((int)x, (int)y, (int)z, kind)) nNoEpi++;
      }
    } else if (cov != 0) { const unsigned z = c / plane; if (z - 1u < LZm1) nCov++; }
  };
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n4; q0 += 4 * stride) {      // four 128-bit loads in flight
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const long long q = q0 + u * stride; v[u] = q < n4 ? S4[q] : make_uint4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
      const long long c = 4 * (q0 + u * stride);
      one(v[u].x, c); one(v[u].y, c + 1); one(v[u].z, c + 2); one(v[u].w, c + 3);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (long long c = n4 << 2; c < D.ncell; c++) one(D.state[c], c);
  unsigned long long vals[6] = {nUnder, nCov, nNoEpi, nBulk, nOcc, nWall};
  for (int k = 0; k < 6; k++) {
    unsigned long long v = vals[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[k], v);
  }
  for (int o = 16; o > 0; o >>= 1) { hz += __shfl_down_sync(0xFFFFFFFFu, hz, o); hz2 += __shfl_down_sync(0xFFFFFFFFu, hz2, o); }
  if ((threadIdx.x & 31) == 0 && (hz != 0 || hz2 != 0)) { atomicAdd(&hsum[0], hz); atomicAdd(&hsum[1], hz2); }
}

// Frame extraction (the output scans of inpout/WriteMolMolXYZFile.f:130-215 fused into one streaming pass, 4 B/cell of
// HBM traffic): every cell the reference's frame writers print -- under-coordinated atoms (.xyz), coverage sites (.xyz),
// non-epitaxial 4-coordinated atoms (_d.xyz) -- is appended as (cell, word) to a compact list; the host sorts by cell to
// get the reference's z,y,x order.  Four 128-bit loads are in flight per thread; a thread keeps a 16-bit hit mask of its
// sixteen words and the warp-aggregated compaction runs only where some lane of the warp has a hit (the printed cells are
// ~1e-3 of the lattice).
__device__ __forceinline__ bool frame_hit(uint32_t w, uint32_t c, unsigned plane) {
  if (w == 0u || w_cov(w) == 113) return false;
  if (!w_occ(w)) return w_cov(w) != 0;
  if (w_coor(w) <= 3) return true;
  if (w_coor(w) != 4) return false;
  const unsigned z = c / plane, r = c - z * plane, y = r / (unsigned)D.LX, x = r - y * (unsigned)D.LX; int kind;
  return !is_cubic_site((int)x, (int)y, (int)z, kind);
}
// MODE 1 ("all"): the union of every test the frame writers apply, vacancies included -- what the host program needs to
// write I*.xyz, _d.xyz and _v.xyz without the lattice (WriteMolMolXYZFile.f:77-260): coverage ids 1 / 17 whatever else the
// word holds (:169-176), atoms of species 1..3 with coordination <= 3 (:177-215), the two counting scans (:77-92: occupied
// with coordination <= 3, empty with a coverage id, walls excluded), non-epitaxial 4-coordinated atoms (:95-159), and the
// empty 4-coordinated sites (the vacancy list, Add.f:82-100).  The plane ranges of the individual scans are applied by the
// caller (it has the coordinates).
__device__ __forceinline__ bool frame_hit_all(uint32_t w, uint32_t c, unsigned plane) {
  if (w == 0u) return false;
  const int cov = w_cov(w), co = w_coor(w), a = w_sp(w);
  const bool occ = w_occ(w) != 0;
  if (cov == 1 || cov == 17) return true;
  if (a >= 1 && a <= 3 && co <= 3) return true;
  if (cov != 113 && (occ ? co <= 3 : (cov != 0 || co == 4))) return true;
  if (!occ || co != 4) return false;
  const unsigned z = c / plane, r = c - z * plane, y = r / (unsigned)D.LX, x = r - y * (unsigned)D.LX; int kind;
  return !is_cubic_site((int)x, (int)y, (int)z, kind);
}
template <int MODE>
__global__ void __launch_bounds__(256) frame_extract_kernel(uint32_t *cells, uint32_t *words, unsigned long long cap, unsigned long long *count) {
  auto hit = [](uint32_t w, uint32_t c, unsigned plane) -> bool { return MODE ? frame_hit_all(w, c, plane) : frame_hit(w, c, plane); };
  const long long n4 = D.ncell >> 2;
  const uint4 *S4 = reinterpret_cast<const uint4 *>(D.state);
  const unsigned plane = (unsigned)D.LX * (unsigned)D.LY;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; q0 < n4; q0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const long long q = q0 + lane + u * stride; v[u] = q < n4 ? S4[q] : make_uint4(0, 0, 0, 0); }
    unsigned hits = 0u;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t c0 = (uint32_t)(4 * (q0 + lane + u * stride));
      if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
      hits |= (hit(v[u].x, c0, plane) ? 1u : 0u) << (4 * u) | (hit(v[u].y, c0 + 1, plane) ? 2u : 0u) << (4 * u) |
              (hit(v[u].z, c0 + 2, plane) ? 4u : 0u) << (4 * u) | (hit(v[u].w, c0 + 3, plane) ? 8u : 0u) << (4 * u);
    }
    if (V.halo && hits) {                                 // halo columns belong to the neighbours
#pragma unroll
      for (int k = 0; k < 16; k++) if ((hits >> k) & 1u) {
        const unsigned x = (unsigned)((4 * (q0 + lane + (k >> 2) * stride) + (k & 3)) % (unsigned)D.LX);
        if (x < (unsigned)TILE || x >= (unsigned)(D.LX - TILE)) hits &= ~(1u << k);
      }
    }
    if (!__any_sync(0xFFFFFFFFu, hits != 0u)) continue;
    const int mine = __popc(hits);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0) + (unsigned long long)(incl - mine);
    for (int k = 0; k < 16; k++) if ((hits >> k) & 1u) {
      const uint4 vv = v[k >> 2]; const int j = k & 3;
      const uint32_t w = j == 0 ? vv.x : (j == 1 ? vv.y : (j == 2 ? vv.z : vv.w));
      if (base < cap) { cells[base] = (uint32_t)(4 * (q0 + lane + (k >> 2) * stride) + j); words[base] = w; }
      base++;
    }
  }
}

// The antisite scan of the SiC frame writer (WriteMolMolXYZFile.f:267-398): every 4-coordinated atom of planes 1..LenZ-1
// (walls excluded) whose number of bonded neighbours of the OTHER species equals `target` is appended as (cell, word).  The
// host calls it with target 0 (atoms the reference flips), patches the flipped words in, calls it with target 3 (the atoms
// it prints into _w.xyz) and patches the words back -- the scan itself stays on the device: 4 B/cell streamed + 8 B bond
// word and four gathered neighbour words per bulk atom (1/216 of the filled cells).
__global__ void __launch_bounds__(256) antisite_kernel(int target, uint32_t *cells, uint32_t *words, unsigned long long cap, unsigned long long *count) {
  const long long n4 = D.ncell >> 2;
  const uint4 *S4 = reinterpret_cast<const uint4 *>(D.state);
  const unsigned plane = (unsigned)D.LX * (unsigned)D.LY, LXu = (unsigned)D.LX, LZm1 = (unsigned)(D.LZ - 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto one = [&](uint32_t w, long long cl) {
    if (w == 0u || !w_occ(w) || w_coor(w) != 4 || w_cov(w) == 113) return;
    const unsigned c = (unsigned)cl, z = c / plane;
    if (z - 1u >= LZm1) return;                                  // planes 1..LenZ-1
    const unsigned r = c - z * plane, y = r / LXu, x = r - y * LXu;
    if (V.halo && (x < (unsigned)TILE || x >= LXu - (unsigned)TILE)) return;
    const int a = w_sp(w);
    int n = 0;
    if (a == 1 || a == 2) {
      const int bit = a == 1 ? PosSiC + 1 : PosSiC;              // a Si atom counts its C neighbours, a C atom its Si neighbours
      const unsigned long long b = D.bonds[c];
      P3 s; s.x = (int)x; s.y = (int)y; s.z = (int)z;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const P3 q = dec_bond(s, (unsigned)((b >> (16 * j)) & 0xFFFFu));
        if (in_box(q)) n += (int)((D.state[cidx_nofail(q)] >> bit) & 1u);
      }
    }
    if (n != target) return;
    const unsigned long long k = atomicAdd(count, 1ull);
    if (k < cap) { cells[k] = c; words[k] = w; }
  };
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n4; q0 += 4 * stride) {      // four 128-bit loads in flight
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const long long q = q0 + u * stride; v[u] = q < n4 ? S4[q] : make_uint4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0u) continue;
      const long long c = 4 * (q0 + u * stride);
      one(v[u].x, c); one(v[u].y, c + 1); one(v[u].z, c + 2); one(v[u].w, c + 3);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (long long c = n4 << 2; c < D.ncell; c++) one(D.state[c], c);
}
__global__ void scatter_state(uint32_t *state, const uint32_t *cell, const uint32_t *word, long long n) {      // cells are distinct
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) state[cell[i]] = word[i];
}
__global__ void gather_state(const uint32_t *state, const uint32_t *cell, uint32_t *out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = state[cell[i]];
}

// Device-side flat (001) builder: geometry of initial/SetSi.f:25-221 (walls, one species) and
// initial/SetSiC3C.f:25-199 (no walls, Si/C).  One thread per (x,y) column of cubic sites.
__global__ void __launch_bounds__(256) init_flat_kernel(int Len, int walls, int two_species, unsigned long long *counters) {
  const long long plane = (long long)D.LX * D.LY;
  const int JD[8][3] = {{3, 3, 3}, {3, -3, -3}, {-3, -3, 3}, {-3, 3, -3}, {3, 3, -3}, {3, -3, 3}, {-3, -3, -3}, {-3, 3, 3}};
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < plane; col += (long long)gridDim.x * blockDim.x) {
    int y = (int)(col / D.LX), x = (int)(col - (long long)y * D.LX);
    if (walls) {
      for (int z = 0; z <= 5; z++) D.state[col + plane * z] = (1u << PosOcc) | 113u;
      for (int z = D.LZ - 6; z <= D.LZ - 1; z++) D.state[col + plane * z] = (1u << PosOcc) | 113u;
    }
    if (x % 3 || y % 3) continue;
    const bool owned = !V.halo || (x >= TILE && x < D.LX - TILE);
    for (int z = (walls ? 6 : 0); z <= Len; z += 1) {
      int kind;
      if (!is_cubic_site(x, y, z, kind)) continue;
      P3 s; s.x = x; s.y = y; s.z = z;
      P3 nn[4];
      for (int i = 0; i < 4; i++) { nn[i].x = x + JD[i + 4 * kind][0]; nn[i].y = y + JD[i + 4 * kind][1]; nn[i].z = z + JD[i + 4 * kind][2]; wrapxy(nn[i]); }
      long long c = col + plane * z;
      int sp = (kind && two_species) ? 2 : 1;
      if (z <= Len - 4) {
        D.state[c] = (1u << (PosSiC + sp - 1)) | (1u << PosOcc) | (4u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        if (owned) atomicAdd(&counters[0], 1ull);
      } else if (z <= Len - 1) {
        D.state[c] = (1u << (PosSiC + sp - 1)) | (1u << PosOcc) | (2u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        double p = D.PE[IDXE(sp, 2, 0, 0, 0, 0, 0)];
        int t = list_of(s);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        V.rate[lidx(t, k)] = p; V.ce[lidx(t, k)] = pack_ce((uint32_t)c, 4, x, y, z); st_slot((uint32_t)c, k + 1);
        atomicAdd(&counters[1], 1ull);
      } else {
        D.state[c] = (2u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        double p = total_dep_abs(2, 0);
        int t = list_of(s);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        V.rate[lidx(t, k)] = p; V.ce[lidx(t, k)] = pack_ce((uint32_t)c, EV_DRAW, x, y, z); st_slot((uint32_t)c, k + 1);
        atomicAdd(&counters[1], 1ull);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (err) *err = std::string("CUDA error ") + cudaGetErrorString(e_) + " in " #call; return -1; } } while (0)

// ---- NCCL, loaded at run time (only multi-rank slab runs need it; torch's bundled libnccl.so.2 is reused when
// the host process has it loaded, else the system one) --------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static inline NcclApi *nccl_api(std::string *err) {
  static NcclApi api; static bool tried = false;
  if (api.lib) return &api;
  if (tried) { if (err) *err = "libnccl.so.2 could not be loaded"; return nullptr; }
  tried = true;
  const char *names[] = {getenv("MSKC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) { if (n && *n) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; } }
  if (!api.lib) { if (err) *err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return nullptr; }
#define MSKC_SYM(f) *(void **)(&api.f) = dlsym(api.lib, "nccl" #f); if (!api.f) { if (err) *err = "libnccl: missing symbol nccl" #f; api.lib = nullptr; return nullptr; }
  MSKC_SYM(GetUniqueId) MSKC_SYM(CommInitRank) MSKC_SYM(CommDestroy) MSKC_SYM(Send) MSKC_SYM(Recv) MSKC_SYM(AllReduce) MSKC_SYM(AllGather) MSKC_SYM(GroupStart) MSKC_SYM(GroupEnd) MSKC_SYM(GetErrorString)
#undef MSKC_SYM
  return &api;
}
#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { if (err) *err = std::string("NCCL error ") + nccl->GetErrorString(r_) + " in " #call; return -1; } } while (0)

enum { HALO_CAP = 32768 };         // records per side per exchange (two colour steps' worth of changed cells)

static inline void parallel_free(ParallelCtx &P) {
  const int halo = P.halo, rank = P.rank, nranks = P.nranks; void *comm = P.comm;
  HaloRec *p2p_recv = P.p2p_recv; void *p2p_left = P.p2p_left, *p2p_right = P.p2p_right; const bool p2p = P.p2p; const unsigned int halo_seq = P.halo_seq;
#ifndef MSKC_EMU
  if (P.gexec) { cudaGraphExecDestroy(P.gexec); P.gexec = nullptr; }
#endif
  cudaFree(P.dirty_bits); cudaFree(P.dirty_cell); cudaFree(P.dirty_n); cudaFree(P.pack_done); cudaFree(P.red);
  for (int i = 0; i < 2; i++) { cudaFree(P.sendbuf[i]); cudaFree(P.recvbuf[i]); }
  cudaFree(P.rate); cudaFree(P.ce); cudaFree(P.cnt); cudaFree(P.live); cudaFree(P.aux); cudaFree(P.tile_rate); cudaFree(P.lag); cudaFree(P.tstat); cudaFree(P.ss); cudaFree(P.tot);
  P = ParallelCtx();
  P.halo = halo; P.rank = rank; P.nranks = nranks; P.comm = comm;
  P.p2p_recv = p2p_recv; P.p2p_left = p2p_left; P.p2p_right = p2p_right; P.p2p = p2p; P.halo_seq = halo_seq;
}
static inline TileView tile_view(const ParallelCtx &P) {
  TileView V; V.tile = P.tile; V.ntx = P.ntx; V.nty = P.nty; V.cap = P.cap; V.NT = P.ntiles; V.hx = P.ntx / 2; V.ntc = (P.ntx / 2) * (P.nty / 2);
  V.rate = P.rate; V.ce = P.ce; V.cnt = P.cnt; V.live = P.live;
  V.tile_rate = P.tile_rate; V.tstat = P.tstat; V.ss = P.ss; V.lag = P.lag;
  V.halo = P.halo; V.nranks = P.nranks; V.dirty_bits = P.dirty_bits; V.dirty_cell = P.dirty_cell; V.dirty_n = P.dirty_n; V.dirty_cap = P.dirty_cap; V.red = P.red;
  return V;
}
static inline int parallel_alloc(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int cap, std::string *err) {
  int tile = 24;
  if (H.LX % (2 * tile) || H.LY % (2 * tile)) { if (err) *err = "parallel sampler needs LenX and LenY to be multiples of 48 (24-unit tiles, 2x2 colouring)"; return -1; }
  if (P.rate && P.cap == cap && P.ncell == H.ncell) {
    PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.live, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.rate, 0, (size_t)P.ntiles * cap * sizeof(double), st));
    PCK(cudaMemsetAsync(P.aux, 0, (size_t)H.ncell * sizeof(uint32_t), st));
    PCK(cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles * sizeof(TileStat), st));
    PCK(cudaMemsetAsync(P.tile_rate, 0, (size_t)P.ntiles * sizeof(double), st));
    PCK(cudaMemsetAsync(P.lag, 0, (size_t)P.ntiles * sizeof(double), st));
    PCK(cudaMemsetAsync(P.ss, 0, sizeof(SweepState), st));
    if (P.halo) { PCK(cudaMemsetAsync(P.dirty_bits, 0, P.dirty_words * 4, st)); PCK(cudaMemsetAsync(P.dirty_n, 0, 2 * sizeof(int), st)); }
    return 0;
  }
  parallel_free(P);
  P.tile = tile; P.ntx = H.LX / tile; P.nty = H.LY / tile; P.ntiles = P.ntx * P.nty; P.cap = cap; P.ncell = H.ncell;
  PCK(cudaMalloc(&P.rate, (size_t)P.ntiles * cap * sizeof(double)));
  PCK(cudaMalloc(&P.ce, (size_t)P.ntiles * cap * sizeof(unsigned long long)));
  PCK(cudaMalloc(&P.cnt, (size_t)P.ntiles * sizeof(int)));
  PCK(cudaMalloc(&P.live, (size_t)P.ntiles * sizeof(int)));
  PCK(cudaMalloc(&P.aux, (size_t)H.ncell * sizeof(uint32_t)));
  PCK(cudaMalloc(&P.tile_rate, (size_t)P.ntiles * sizeof(double)));
  PCK(cudaMalloc(&P.lag, (size_t)P.ntiles * sizeof(double)));
  PCK(cudaMalloc(&P.tstat, (size_t)P.ntiles * sizeof(TileStat)));
  PCK(cudaMalloc(&P.ss, sizeof(SweepState)));
  PCK(cudaMalloc(&P.tot, sizeof(TileStat)));
  PCK(cudaMalloc(&P.red, RED_DOUBLES * sizeof(double)));
  PCK(cudaMemsetAsync(P.red, 0, RED_DOUBLES * sizeof(double), st));
  if (P.halo) {
    if (H.LX < 4 * tile) { if (err) *err = "slab mode needs at least two owned tile columns per rank"; return -1; }
    P.dirty_cap = HALO_CAP;
    P.dirty_words = ((size_t)4 * tile * H.LY * (H.LZ + 1) + 31) / 32;
    PCK(cudaMalloc(&P.dirty_bits, P.dirty_words * 4)); PCK(cudaMemsetAsync(P.dirty_bits, 0, P.dirty_words * 4, st));
    PCK(cudaMalloc(&P.dirty_cell, (size_t)2 * P.dirty_cap * 4));
    PCK(cudaMalloc(&P.dirty_n, 2 * sizeof(int))); PCK(cudaMemsetAsync(P.dirty_n, 0, 2 * sizeof(int), st));
    PCK(cudaMalloc(&P.pack_done, sizeof(unsigned int))); PCK(cudaMemsetAsync(P.pack_done, 0, sizeof(unsigned int), st));
    for (int i = 0; i < 2; i++) {
      PCK(cudaMalloc(&P.sendbuf[i], (size_t)(HALO_CAP + 1) * sizeof(HaloRec))); PCK(cudaMemsetAsync(P.sendbuf[i], 0, sizeof(HaloRec), st));
      PCK(cudaMalloc(&P.recvbuf[i], (size_t)(HALO_CAP + 1) * sizeof(HaloRec))); PCK(cudaMemsetAsync(P.recvbuf[i], 0, sizeof(HaloRec), st));
    }
  }
  PCK(cudaMemsetAsync(P.rate, 0, (size_t)P.ntiles * cap * sizeof(double), st));
  PCK(cudaMemsetAsync(P.ce, 0, (size_t)P.ntiles * cap * sizeof(unsigned long long), st));
  PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.live, 0, (size_t)P.ntiles * sizeof(int), st));
  PCK(cudaMemsetAsync(P.aux, 0, (size_t)H.ncell * sizeof(uint32_t), st));
  PCK(cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles * sizeof(TileStat), st));
  PCK(cudaMemsetAsync(P.tile_rate, 0, (size_t)P.ntiles * sizeof(double), st));
  PCK(cudaMemsetAsync(P.lag, 0, (size_t)P.ntiles * sizeof(double), st));
  PCK(cudaMemsetAsync(P.ss, 0, sizeof(SweepState), st));
  return 0;
}
// census flags of every cell with bonds (check = 0: write them; 1: compare with the stored ones); out[3] as the kernel's
static inline int parallel_census_flags(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int check, unsigned long long out[3], float *ms, cudaEvent_t e0, cudaEvent_t e1, std::string *err) {
  unsigned long long *d = nullptr;
  PCK(cudaMalloc(&d, 24)); PCK(cudaMemsetAsync(d, 0, 24, st));
  if (e0) PCK(cudaEventRecord(e0, st));
  MSKC_LAUNCH(census_flags_kernel, 148 * 8, 256, 0, st, check, d);
  if (e1) PCK(cudaEventRecord(e1, st));
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(out, d, 24, cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  if (ms && e0 && e1) PCK(cudaEventElapsedTime(ms, e0, e1));
  cudaFree(d);
  return 0;
}
#ifndef MSKC_TILE_CAP
#define MSKC_TILE_CAP 192
#endif
static const int TILE_CAP = MSKC_TILE_CAP;

// per-sweep totals: local sums, all-reduce over the ranks of a slab run, then clock / tau / global-coupling terms
static inline int sweep_totals(ParallelCtx &P, cudaStream_t st, int advance, int *launches, std::string *err) {
  MSKC_LAUNCH(sweep_sum_kernel, SUM_BLOCKS, 256, 0, st, P.ntiles, P.nranks > 1 ? -1 : advance);
  *launches += 1;
  if (P.nranks == 1) return 0;
  {
    NcclApi *nccl = nccl_api(err); if (!nccl) return -1;
    if (!P.comm) { if (err) *err = "slab run with several ranks: call mskc_comm_connect first"; return -1; }
    NCK(nccl->AllReduce(P.red, P.red, 4, ncclDouble, ncclSum, (ncclComm_t)P.comm, st));
  }
  MSKC_LAUNCH(sweep_finish_kernel, 1, 1, 0, st, advance);
  *launches += 1;
  return 0;
}
// after a colour step: send the changed records of the two shared strips to the neighbours, apply theirs
// Receive buffers of the halo exchange: [side][parity][1 + HALO_CAP] records in ONE allocation, so that one IPC handle per
// rank is enough.  side 0 receives from the right neighbour (its left strip), side 1 from the left neighbour.
static inline HaloRec *p2p_slot(void *base, int side, int parity) { return reinterpret_cast<HaloRec *>(base) + (size_t)(side * 2 + parity) * (HALO_CAP + 1); }
static inline int p2p_alloc(ParallelCtx &P, cudaStream_t st, std::string *err) {
  if (P.p2p_recv) return 0;
  const size_t bytes = (size_t)4 * (HALO_CAP + 1) * sizeof(HaloRec);
  PCK(cudaMalloc(&P.p2p_recv, bytes));
  PCK(cudaMemsetAsync(P.p2p_recv, 0, bytes, st));
  PCK(cudaStreamSynchronize(st));
  return 0;
}
// Several ranks (one process per GPU): every rank exports its receive buffers through CUDA IPC, the handles travel in one
// ncclAllGather, and each rank maps the buffers of its two neighbours: halo_pack_kernel then stores its records straight
// into the neighbour's memory over NVLink and the neighbour's halo_apply_kernel waits on the header -- no collective call,
// no fixed-size message and no host involvement per colour step.  Falls back to ncclSend/ncclRecv when the mapping fails.
static inline int p2p_connect(ParallelCtx &P, cudaStream_t st, std::string *err) {
  if (p2p_alloc(P, st, err)) return -1;
  P.p2p = false;
#ifndef MSKC_EMU
  if (P.nranks == 1 || getenv("MSKC_HALO_NCCL")) return 0;
  NcclApi *nccl = nccl_api(err); if (!nccl) return -1;
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, P.p2p_recv) != cudaSuccess) { cudaGetLastError(); return 0; }
  unsigned char *d_in = nullptr, *d_all = nullptr;
  const size_t hs = sizeof(cudaIpcMemHandle_t);
  PCK(cudaMalloc(&d_in, hs)); PCK(cudaMalloc(&d_all, hs * P.nranks));
  PCK(cudaMemcpyAsync(d_in, &mine, hs, cudaMemcpyHostToDevice, st));
  NCK(nccl->AllGather(d_in, d_all, hs, ncclUint8, (ncclComm_t)P.comm, st));
  std::vector<cudaIpcMemHandle_t> all(P.nranks);
  PCK(cudaMemcpyAsync(all.data(), d_all, hs * P.nranks, cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  cudaFree(d_in); cudaFree(d_all);
  const int left = (P.rank + P.nranks - 1) % P.nranks, right = (P.rank + 1) % P.nranks;
  bool ok = cudaIpcOpenMemHandle(&P.p2p_left, all[left], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  if (ok) { if (right == left) P.p2p_right = P.p2p_left; else ok = cudaIpcOpenMemHandle(&P.p2p_right, all[right], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess; }
  // every rank must take the same path: agree by all-reduce (sum of successes == nranks)
  double *d_ok = nullptr, h_ok = ok ? 1.0 : 0.0;
  PCK(cudaMalloc(&d_ok, sizeof(double)));
  PCK(cudaMemcpyAsync(d_ok, &h_ok, sizeof(double), cudaMemcpyHostToDevice, st));
  NCK(nccl->AllReduce(d_ok, d_ok, 1, ncclDouble, ncclSum, (ncclComm_t)P.comm, st));
  PCK(cudaMemcpyAsync(&h_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  cudaFree(d_ok);
  if (!ok) cudaGetLastError();
  P.p2p = h_ok > P.nranks - 0.5;
  if (!P.p2p) fprintf(stderr, "mskc: peer-mapped halo buffers unavailable on rank %d, using ncclSend/ncclRecv\n", P.rank);
#endif
  return 0;
}
// after a colour step: the changed records of the two shared strips go to the neighbours, theirs are applied
static inline int halo_exchange(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int *launches, std::string *err) {
  const int owned = H.LX - 2 * P.tile;                         // width of the owned block = shift between the two copies of a strip
  if (p2p_alloc(P, st, err)) return -1;
  const unsigned int seq = ++P.halo_seq;                       // same on every rank (all run the same sequence of colour steps)
  const int par = (int)(seq & 1u);                             // double-buffered: a rank one step ahead never overwrites what its neighbour still reads
  HaloRec *in0 = p2p_slot(P.p2p_recv, 0, par), *in1 = p2p_slot(P.p2p_recv, 1, par);
  // my left strip = the left neighbour's right strip (cells shift by +owned), my right strip = the right neighbour's left strip
  if (P.nranks == 1) {                                         // one rank: it is its own neighbour on both sides (periodic box)
    MSKC_LAUNCH(halo_pack_kernel, dim3(16, 2), 256, 0, st, owned, in0, in1, P.pack_done, seq);
  } else if (P.p2p) {
    MSKC_LAUNCH(halo_pack_kernel, dim3(16, 2), 256, 0, st, owned, p2p_slot(P.p2p_left, 0, par), p2p_slot(P.p2p_right, 1, par), P.pack_done, seq);
  } else {
    NcclApi *nccl = nccl_api(err); if (!nccl) return -1;
    if (!P.comm) { if (err) *err = "slab run with several ranks: call mskc_comm_connect first"; return -1; }
    MSKC_LAUNCH(halo_pack_kernel, dim3(16, 2), 256, 0, st, owned, (HaloRec *)P.sendbuf[0], (HaloRec *)P.sendbuf[1], P.pack_done, seq);
    const size_t bytes = (size_t)(HALO_CAP + 1) * sizeof(HaloRec);
    const int left = (P.rank + P.nranks - 1) % P.nranks, right = (P.rank + 1) % P.nranks;
    NCK(nccl->GroupStart());
    NCK(nccl->Send(P.sendbuf[0], bytes, ncclUint8, left, (ncclComm_t)P.comm, st));
    NCK(nccl->Send(P.sendbuf[1], bytes, ncclUint8, right, (ncclComm_t)P.comm, st));
    NCK(nccl->Recv(in0, bytes, ncclUint8, right, (ncclComm_t)P.comm, st));
    NCK(nccl->Recv(in1, bytes, ncclUint8, left, (ncclComm_t)P.comm, st));
    NCK(nccl->GroupEnd());
  }
  MSKC_LAUNCH(halo_apply_kernel, dim3(16, 2), 256, 0, st, (const HaloRec *)in0, (const HaloRec *)in1, seq);
  *launches += 2;
  return 0;
}

// tile lists from an uploaded MC list (reference order is irrelevant here)
static inline int parallel_build_from_lists(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int n_mc, const uint32_t *mc_cell, const int *mc_ev,
                                            const double *mc_rate, int n_bulk, int n_void, std::string *err) {
  if (H.LX % 48 || H.LY % 48) { P.lists_stale = true; return 0; }      // parity-only box
  if (parallel_alloc(P, H, st, TILE_CAP, err)) return -1;
  std::vector<double> R((size_t)P.ntiles * P.cap, 0.0); std::vector<unsigned long long> CE((size_t)P.ntiles * P.cap, 0ull);
  std::vector<int> cnt(P.ntiles, 0); std::vector<uint32_t> slot((size_t)H.ncell, 0);
  long long plane = (long long)H.LX * H.LY;
  const int hx = P.ntx / 2, ntc = (P.ntx / 2) * (P.nty / 2);
  for (int i = 0; i < n_mc; i++) {
    uint32_t c = mc_cell[i];
    long long r = c % plane; int y = (int)(r / H.LX), x = (int)(r - (long long)y * H.LX);
    int tx = x / P.tile, ty = y / P.tile;
    int t = (((ty & 1) << 1) | (tx & 1)) * ntc + (ty >> 1) * hx + (tx >> 1);
    if (cnt[t] >= P.cap) { if (err) *err = "per-tile active list capacity exceeded at upload"; return -1; }
    // pre-drawn deposition/adsorption codes (1..3, >=6) of empty sites become "draw at fire time",
    // except the forced code-1 entry of Get_Prob.f:182-183 (rate 1e-8) / Get_Prob_Ini.f:104 (1e-6)
    int ev = mc_ev[i];
    bool forced = ev == 1 && (mc_rate[i] == (double)1.e-8f || mc_rate[i] == (double)1.e-6f);
    uint32_t e = (ev == 4 || ev == 5 || forced) ? (uint32_t)ev : (uint32_t)EV_DRAW;
    size_t a = (size_t)t * P.cap + cnt[t];
    R[a] = mc_rate[i]; CE[a] = pack_ce(c, (int)e, x, y, (int)(c / plane));
    slot[c] = (uint32_t)(++cnt[t]);
  }
  PCK(cudaMemcpyAsync(P.rate, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.ce, CE.data(), CE.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.aux, slot.data(), slot.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  PCK(cudaStreamSynchronize(st));
  P.base_atoms = n_bulk; P.base_voids = n_void; P.time = 0; P.ran = false; P.replayed = false; P.lists_stale = false;
  for (int i = 0; i < 3; i++) { P.baseCrystal[i] = 0; P.baseCov[i] = 0; }
  memset(&P.h_tot, 0, sizeof P.h_tot); memset(&P.extra, 0, sizeof P.extra);
  return 0;
}

static inline int parallel_init_flat(ParallelCtx &P, DevCtx &H, cudaStream_t st, int len1, int walls, int two_species, const std::function<int()> &sync_consts, std::string *err) {
  len1 -= len1 % 24;                                                  // main.f:226,240
  if (len1 < 24 || len1 + 12 > H.LZ - (walls ? 6 : 0)) { if (err) *err = "mskc_init_flat: Len1 does not fit the box"; return -1; }
  if (H.LX % 48 || H.LY % 48) { if (err) *err = "mskc_init_flat: LenX, LenY must be multiples of 48"; return -1; }
  if (parallel_alloc(P, H, st, TILE_CAP, err)) return -1;
  if (sync_consts()) { if (err) *err = "constant upload failed"; return -1; }
  PCK(cudaMemsetAsync(H.state, 0, (size_t)H.ncell * sizeof(uint32_t), st));
  PCK(cudaMemsetAsync(H.bonds, 0, (size_t)H.ncell * sizeof(unsigned long long), st));
  unsigned long long *d_cnt = nullptr, h_cnt[2];
  PCK(cudaMalloc(&d_cnt, 16)); PCK(cudaMemsetAsync(d_cnt, 0, 16, st));
  MSKC_LAUNCH(init_flat_kernel, 148 * 8, 256, 0, st, len1, walls, two_species, d_cnt);
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  cudaFree(d_cnt);
  { unsigned long long o[3]; if (parallel_census_flags(P, H, st, 0, o, nullptr, nullptr, nullptr, err)) return -1; }
  P.base_atoms = (long long)h_cnt[0]; P.base_voids = 0; P.time = 0; P.ran = false; P.lists_stale = false;
  memset(&P.h_tot, 0, sizeof P.h_tot);
  return 0;
}

// which team size runs a colour step: every tile of the colour should be resident at once (one wave of 16 warps per SM
// at 128 registers) with as many lanes per tile as that allows -- the step is bound by the serial chain of the busiest
// tile, and a wider team shortens that chain (no other tile's branches in its instruction stream)
static inline int pick_team(int ntc) {
  if (const char *e = getenv("MSKC_TEAM")) { const int t = atoi(e); if (t == 8 || t == 16 || t == 32) return t; }
  const int resident_warps = 148 * 16;
  return ntc <= resident_warps ? 32 : (ntc <= 2 * resident_warps ? 16 : 8);
}

static inline int parallel_run(ParallelCtx &P, DevCtx &H, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, const mskc_parallel_args *a, mskc_parallel_result *r, std::string *err) {
  memset(r, 0, sizeof *r);
  const int ntc = (P.ntx / 2) * (P.nty / 2);
  const int teamsz = pick_team(ntc);
  const int tpb = CS_THREADS / teamsz;
  const int blocks = (ntc + tpb - 1) / tpb;
  size_t smem = (size_t)tpb * P.cap * 16 + tpb * sizeof(int);
  if (const char *e = getenv("MSKC_SMEM_PAD")) smem += (size_t)atoi(e);      // diagnostic: fewer resident blocks per SM
  const int feat = (H.NCov > 0 ? F_COV : 0) | (H.la ? F_LA : 0);
  // the kernel is instantiated per feature set so that a PVD run does not carry coverage / LA code in its hot loop
  typedef void (*cs_kernel_t)(int, int);
  static const cs_kernel_t kern8[4] = {colour_step_kernel<8, 0>, colour_step_kernel<8, 1>, colour_step_kernel<8, 2>, colour_step_kernel<8, 3>};
  static const cs_kernel_t kern16[4] = {colour_step_kernel<16, 0>, colour_step_kernel<16, 1>, colour_step_kernel<16, 2>, colour_step_kernel<16, 3>};
  static const cs_kernel_t kern32[4] = {colour_step_kernel<32, 0>, colour_step_kernel<32, 1>, colour_step_kernel<32, 2>, colour_step_kernel<32, 3>};
  const cs_kernel_t kern = teamsz == 8 ? kern8[feat] : (teamsz == 16 ? kern16[feat] : kern32[feat]);
  PCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 64));
  const double ept = a->events_per_tile > 0 ? a->events_per_tile : 4.0;
  // events one tile may execute in one colour step before the rest of its quantum is carried over (flags & 0xFFFF,
  // 0 = default: three times the target mean; 0xFFFF = no cap)
  int evcap = a->flags & 0xFFFF;
  if (const char *e = getenv("MSKC_EVCAP")) evcap = atoi(e);
  if (evcap == 0) evcap = std::max(4, (int)(3.0 * ept));
  if (evcap == 0xFFFF) evcap = 1 << 14;
  const uint2 key = make_uint2((uint32_t)a->seed, (uint32_t)(a->seed >> 32));
  const double owned = (double)(P.ntx - 2 * P.halo) * P.nty * P.nranks;
  TileStat before = P.h_tot;
  int launches = 0;
  MSKC_LAUNCH(reset_maxev_kernel, (P.ntiles + 255) / 256, 256, 0, st, P.tstat, P.ntiles);
  MSKC_LAUNCH(set_params_kernel, 1, 1, 0, st, key, evcap, ept, a->tau_fixed, owned);
  PCK(cudaEventRecord(e0, st));
  if (!P.ran) {                    // first call: establish total rate and tau
    MSKC_LAUNCH(tile_rate_kernel, (P.ntiles + 255) / 256, 256, 0, st, P.ntiles);
    if (sweep_totals(P, st, 0, &launches, err)) return -1;
    launches += 1;
  } else {                         // later calls: tau of the first sweep from THIS call's parameters (the sums of the last sweep stand)
    MSKC_LAUNCH(sweep_finish_kernel, 1, 1, 0, st, 0);
    launches += 1;
  }
  static const int order[4][2] = {{0, 0}, {1, 1}, {1, 0}, {0, 1}};
  // on request every sampler kernel is timed on its own (CUDA events on the launching stream): this is the per-launch
  // duration the roofline in bench.py is computed from
  const bool timed = (a->flags & MSKC_PAR_TIME_KERNELS) && a->n_sweeps > 0 && a->n_sweeps <= 64;
  std::vector<cudaEvent_t> evs;
  if (timed) { evs.resize((size_t)a->n_sweeps * 8); for (auto &e : evs) PCK(cudaEventCreate(&e)); }
  size_t ke = 0;
  int n_sampler_launches = 0;
  // one sweep = four colour steps (fixed colour order: an alternating ABCD/DCBA order was measured WORSE, a colour then
  // runs twice in a row, doubling its effective quantum -- DESIGN.md section 5.3).  The totals (sum of rates -> next tau,
  // clock, active-site count; one all-reduce over the ranks of a slab run) follow every sweep.  MSKC_TOTALS_EVERY=k takes
  // them once per batch of k sweeps instead (the clock advances by the tau that was used, so this is exact in time): +3 %
  // events/s at k = 4, but tau then lags the total rate by up to k sweeps, which over-sizes the quantum while a surface is
  // still roughening (measured: 1.5x the events per step in the first sweeps after a flat start) -- not the default.
  int totals_every = 1;
  if (const char *e = getenv("MSKC_TOTALS_EVERY")) totals_every = std::max(1, atoi(e));
  auto batch = [&](int k) -> int {
    for (int sw = 0; sw < k; sw++) for (int c = 0; c < 4; c++) {
      if (timed) PCK(cudaEventRecord(evs[ke++], st));
      const int col = order[c][1] * 2 + order[c][0];
      MSKC_LAUNCH(kern, blocks, CS_THREADS, smem, st, col, 4 * sw + c);
      if (timed) PCK(cudaEventRecord(evs[ke++], st));
      launches += 1; n_sampler_launches++;
      // The shared strips cross the seam only when the NEXT colour step runs the other tile-column parity: tiles of the same
      // parity on the two sides of a rank boundary neither write nor read each other's cells (writes <= 8, reads <= 16 units
      // from a tile's edge, 24 between them -- the argument that lets one colour run concurrently), so steps 2 -> 3 of a sweep
      // and 4 -> 1 of the next need no exchange; the changed cells keep accumulating in the dirty list.
      if (P.halo) {
        if (order[c][0] != order[(c + 1) & 3][0]) { if (halo_exchange(P, H, st, &launches, err)) return -1; P.halo_pending = false; }
        else P.halo_pending = true;
      }
    }
    if (P.halo) { MSKC_LAUNCH(compact_halo_kernel, (2 * P.nty * 32 + 127) / 128, 128, 0, st); launches++; }
    return sweep_totals(P, st, k, &launches, err);
  };
  long long left = a->n_sweeps;
#ifndef MSKC_EMU
  // single GPU: the launches of a batch are captured once into a CUDA graph and replayed (launch-bound tail of short
  // colour steps on small boxes; nothing in a sweep depends on host-side arguments any more)
  if (!timed && !P.halo && left >= totals_every && !getenv("MSKC_NO_GRAPH")) {
    if (P.gexec && (P.g_kern != (void *)kern || P.g_blocks != blocks || P.g_smem != smem || P.g_sweeps != totals_every)) { cudaGraphExecDestroy(P.gexec); P.gexec = nullptr; }
    if (!P.gexec) {
      cudaGraph_t g = nullptr; int l0 = launches, n0 = n_sampler_launches;
      PCK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int rc = batch(totals_every);
      PCK(cudaStreamEndCapture(st, &g));
      if (rc) return -1;
      PCK(cudaGraphInstantiate(&P.gexec, g, 0));
      cudaGraphDestroy(g);
      P.g_kern = (void *)kern; P.g_blocks = blocks; P.g_smem = smem; P.g_launches = launches - l0; P.g_sweeps = totals_every;
      launches = l0; n_sampler_launches = n0;
    }
    for (; left >= totals_every; left -= totals_every) { PCK(cudaGraphLaunch(P.gexec, st)); launches += P.g_launches; n_sampler_launches += 4 * totals_every; }
  }
#endif
  for (; left > 0; left -= std::min<long long>(left, totals_every)) if (batch((int)std::min<long long>(left, totals_every))) return -1;
  if (P.halo && P.halo_pending) { if (halo_exchange(P, H, st, &launches, err)) return -1; P.halo_pending = false; }     // seams consistent when the call returns
  PCK(cudaMemsetAsync(P.tot, 0, sizeof(TileStat), st));
  MSKC_LAUNCH(reduce_tstat_kernel, 32, 256, 0, st, P.tstat, P.ntiles, P.tot);
  launches++;
  PCK(cudaEventRecord(e1, st));
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&P.h_tot, P.tot, sizeof(TileStat), cudaMemcpyDeviceToHost, st));
  PCK(cudaMemcpyAsync(&P.h_ss, P.ss, sizeof(SweepState), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  PCK(cudaEventElapsedTime(&r->kernel_ms, e0, e1));
  if (timed) {
    float tot = 0;
    for (size_t i = 0; i + 1 < ke; i += 2) { float ms = 0; PCK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1])); tot += ms; }
    r->sweep_kernel_ms = tot;
    for (auto &e : evs) cudaEventDestroy(e);
  }
  P.ran = true;
  r->events_done = P.h_tot.events - before.events; r->site_updates = P.h_tot.site_updates - before.site_updates;
  r->sweeps_done = a->n_sweeps; r->time = P.h_ss.time; r->total_rate = P.h_ss.total_rate; r->n_active = P.h_ss.n_active;
  for (int i = 0; i < 3; i++) r->count_dep[i] = P.h_tot.dep[i] - before.dep[i];
  r->count_evap = P.h_tot.evap - before.evap; r->count_abs = P.h_tot.abs_ - before.abs_; r->count_des = P.h_tot.des - before.des;
  r->launches = launches; r->max_tile_events = P.h_tot.maxev;
  r->stop_reason = r->n_active <= 20 ? MSKC_STOP_FEWMC : MSKC_STOP_BUDGET;      // main.f:307-310 on the live active-site count
  r->sampler_launches = n_sampler_launches;
  r->tau = P.h_ss.tau_adaptive;
  P.time = P.h_ss.time;
  return 0;
}

static inline int parallel_rerate(ParallelCtx &P, DevCtx &H, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, int64_t *n_sites, double *total_rate, float *ms, std::string *err) {
  unsigned long long *d_n = nullptr, h_n = 0;
  PCK(cudaMalloc(&d_n, 8)); PCK(cudaMemsetAsync(d_n, 0, 8, st));
  PCK(cudaEventRecord(e0, st));
  MSKC_LAUNCH(rerate_kernel, 148 * 8, 256, 0, st, P.ntiles, d_n);
  MSKC_LAUNCH(tile_rate_kernel, (P.ntiles + 255) / 256, 256, 0, st, P.ntiles);
  { int l_ = 0; if (sweep_totals(P, st, 0, &l_, err)) return -1; }
  PCK(cudaEventRecord(e1, st));
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&h_n, d_n, 8, cudaMemcpyDeviceToHost, st));
  PCK(cudaMemcpyAsync(&P.h_ss, P.ss, sizeof(SweepState), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  if (ms) PCK(cudaEventElapsedTime(ms, e0, e1));
  if (n_sites) *n_sites = (int64_t)h_n;
  if (total_rate) *total_rate = P.h_ss.total_rate;
  cudaFree(d_n);
  return 0;
}

static inline void parallel_fill_counters(const ParallelCtx &P, mskc_counters *c) {
  c->num_atoms = (int32_t)(P.base_atoms + P.h_tot.dAtoms + P.extra.dAtoms); c->num_voids = (int32_t)(P.base_voids + P.h_tot.dVoids + P.extra.dVoids);
  c->num_adatom = (int32_t)P.h_ss.n_active;
  for (int i = 0; i < 3; i++) { c->CountCrystal[i] = P.baseCrystal[i] + P.h_tot.dCrystal[i] + P.extra.dCrystal[i]; c->CountCov[i] = P.baseCov[i] + P.h_tot.dCov[i] + P.extra.dCov[i]; }
  c->iter = P.h_tot.events + P.extra.events + 1; c->time = P.h_ss.time;
}

}  // namespace mskc
