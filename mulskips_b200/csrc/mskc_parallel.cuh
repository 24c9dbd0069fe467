// mskc_parallel.cuh -- synchronous-sublattice sampler + the dense helper kernels.
//
// The x-y plane is cut into square tiles (edge 24 super-lattice units = 2 cubic cells) coloured 2x2.
// An event at S writes S and its 4 neighbours (<= 1 bond, |d| <= 8 per axis incl. hexagonal bonds) and
// reads at most 2 bonds away (reciprocity of neighbours' neighbours, zig-zag/arm-chair candidates),
// so two tiles of the same colour (>= 24 units apart) never touch the same cell: all tiles of a colour
// run concurrently, one warp per tile, each doing its own n-fold way for a time quantum tau.
//
// HBM layout added to state[]/bonds[] (mskc_device.cuh):
//   rate[slot][list] f64, ce[slot][list] u64 {u32 cell, u32 event}: the per-tile active-site lists, slot-major
//   cnt[tile]           i32  slots in use (tombstones with rate 0 included until the owner compacts)
//   slot[cell]          u16  1-based slot of the cell in its tile's list, 0 = not listed
//   tstat[tile]         per-tile counters (events, site updates, per-class counts), summed on demand
#pragma once
#include "mskc_device.cuh"
#include "../../include/mskc.h"
#include <algorithm>
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
  long long n_active, sweeps, step;     // step = global colour-step counter (Philox counter word)
  long long max_cnt;
};

// Per-tile active lists, slot-major ("SoA over tiles"): entry k of list l lives at [k*NT + l], and the
// lists of one colour are contiguous (l = colour*ntc + index), so the 32 lanes of a warp -- 32 tiles of
// the running colour -- read 32 consecutive words when they scan slot k together.
struct ParallelCtx {
  int tile = 24, ntx = 0, nty = 0, ntiles = 0, cap = 0;
  double *rate = nullptr; unsigned long long *ce = nullptr; int *cnt = nullptr, *ntomb = nullptr; uint16_t *slot = nullptr; double *tile_rate = nullptr;
  TileStat *tstat = nullptr; SweepState *ss = nullptr; TileStat *tot = nullptr;
  long long ncell = 0;
  long long base_atoms = 0, base_voids = 0;
  int baseCrystal[3] = {0, 0, 0}, baseCov[3] = {0, 0, 0};
  TileStat h_tot{}; SweepState h_ss{};
  double time = 0.0;
  bool ran = false, lists_stale = false, listed_flag_in_state = false;
};

struct TileView {            // what the kernels need (constant memory)
  int tile, ntx, nty, cap, NT, hx, ntc;
  double *rate; unsigned long long *ce; int *cnt, *ntomb; uint16_t *slot; double *tile_rate; TileStat *tstat; SweepState *ss;
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

__device__ __forceinline__ int list_of(P3 s) {          // list id of the tile that owns site s
  int tx = s.x / V.tile, ty = s.y / V.tile;
  return (((ty & 1) << 1) | (tx & 1)) * V.ntc + (ty >> 1) * V.hx + (tx >> 1);
}

struct TilePol {
  uint2 key; uint32_t tile_id; uint32_t step_lo, step_hi; uint32_t draw;
  double root;
  long long n_getprob;
  int CountCrystal[3], CountCov[3];
  int dAtoms, dVoids;

  __device__ __forceinline__ bool predraw() const { return false; }
  __device__ __forceinline__ double tree_root() const { return root; }
  __device__ __forceinline__ uint4 next4() { return philox4x32_10(make_uint4(tile_id, step_lo, step_hi, draw++), key); }
  __device__ __forceinline__ double uniform() { uint4 r = next4(); return u53(r.x, r.y); }

  __device__ __noinline__ void put(P3 s, long long c, int ev, double p) {
    int l = list_of(s);
    int sl = V.slot[c];
    if (sl == 0) {
      int k = atomicAdd(&V.cnt[l], 1);
      if (k >= V.cap) { atomicSub(&V.cnt[l], 1); fail(E_TILE_FULL, s.x, s.y, s.z, l); return; }
      V.slot[c] = (uint16_t)(k + 1);
      sl = k + 1;
    } else if (V.rate[(long long)(sl - 1) * V.NT + l] == 0.0) atomicSub(&V.ntomb[l], 1);       // a tombstone comes back to life
    long long a = (long long)(sl - 1) * V.NT + l;
    V.rate[a] = p;
    V.ce[a] = (unsigned long long)(uint32_t)c | ((unsigned long long)(uint32_t)ev << 32);
  }
  __device__ __forceinline__ void add_mc(P3 s, long long c, const P3 nn[4], int ev, double p) {
    if (p == 0) { fail(E_ADD_ZERO, s.x, s.y, s.z); return; }
    D.bonds[c] = enc_bonds(s, nn);
    put(s, c, ev, p);
  }
  __device__ __forceinline__ void set_rate(P3 s, long long c, int ev, double p) { put(s, c, ev, p); }
  __device__ __noinline__ void erase_mc(P3 s, long long c) {
    int sl = V.slot[c];
    if (sl) {
      int l = list_of(s);
      long long a = (long long)(sl - 1) * V.NT + l;
      if (V.rate[a] != 0.0) { V.rate[a] = 0.0; atomicAdd(&V.ntomb[l], 1); }     // tombstone; the owner compacts
    }
    D.bonds[c] = 0ull;
  }
  __device__ __forceinline__ void add_void(P3 s, long long c, const P3 nn[4]) { dVoids++; D.bonds[c] = enc_bonds(s, nn); }
  __device__ __forceinline__ void erase_void(P3 s, long long c) { dVoids--; D.bonds[c] = 0ull; }
  __device__ __forceinline__ void add_bulk(P3 s, long long c, const P3 nn[4]) { dAtoms++; D.bonds[c] = enc_bonds(s, nn); }
  __device__ __forceinline__ void erase_bulk(P3 s, long long c) { dAtoms--; D.bonds[c] = 0ull; }
  // Coverage on a 4-coordinated empty site: the reference's lookup lands on an unrelated record
  // (see ParityPol::cov_void_recipr), i.e. such a neighbour is effectively never counted.
  __device__ __forceinline__ bool cov_void_recipr(P3, long long, P3) const { return false; }
};

// One colour step: every tile of colour `colour` advances by tau.
// A TEAM of lanes (1..32, power of two) owns one tile, 32/TEAM tiles share a warp.  The team leader runs
// the transition function (a serial dependency chain); the whole team shares (a) the rate sum / pick scan
// over the tile's list and (b) the re-rate requests the event queued (one Get_Prob per lane).  Lists are
// slot-major, so the leaders' / teams' scans are coalesced across the tiles of a warp.
template <int TEAM>
__global__ void __launch_bounds__(32) colour_step_kernel(int colour, uint2 key, int max_events_per_tile, int dbg_lid) {
  constexpr int TPW = 32 / TEAM;
  __shared__ RQ sq[TPW];
  const int lane = threadIdx.x, team = lane / TEAM, tl = lane % TEAM;
  const unsigned tmask = (TEAM == 32) ? 0xFFFFFFFFu : (((1u << TEAM) - 1u) << (team * TEAM));
  const int leader = team * TEAM;
  const int wid = blockIdx.x * TPW + team;
  if (wid >= V.ntc) return;
  const int l = colour * V.ntc + wid;
  const long long NT = V.NT;
  const SweepState ss = *V.ss;
  const double tau = ss.tau;
  double *rate = V.rate + l; unsigned long long *ce = V.ce + l;
  RQ &q = sq[team];
  if (tl == 0) {
    int n = V.cnt[l];
    if (V.ntomb[l] > 0) {               // compaction: drop tombstones (swap with last), keep slot[] consistent
      int i = 0;
      while (i < n) {
        if (rate[i * NT] == 0.0) {
          V.slot[(uint32_t)ce[i * NT]] = 0;
          n--;
          if (i < n) { double r = rate[n * NT]; unsigned long long e = ce[n * NT]; rate[i * NT] = r; ce[i * NT] = e; if (r != 0.0) V.slot[(uint32_t)e] = (uint16_t)(i + 1); }
        } else i++;
      }
      V.cnt[l] = n; V.ntomb[l] = 0;
    }
    q.n = 0;
  }
  __syncwarp(tmask);
  TilePol pol;
  pol.key = key; pol.tile_id = (uint32_t)l; pol.step_lo = (uint32_t)ss.step; pol.step_hi = (uint32_t)(ss.step >> 32); pol.draw = 0;
  pol.root = ss.tree_root; pol.n_getprob = 0; pol.dAtoms = 0; pol.dVoids = 0;
  for (int i = 0; i < 3; i++) { pol.CountCrystal[i] = 0; pol.CountCov[i] = 0; }
  int nev = 0, ndep[3] = {0, 0, 0}, nevap = 0, nabs = 0, ndes = 0;
  double tloc = 0.0, R = 0.0;
  while (true) {
    const int n = V.cnt[l];
    // team-parallel rate sum (strided over the slots), inclusive prefix over the team
    double part = 0.0;
    for (int i = tl; i < n; i += TEAM) part += rate[i * NT];
    double incl = part;
#pragma unroll
    for (int o = 1; o < TEAM; o <<= 1) { double v = __shfl_up_sync(tmask, incl, o, TEAM); if (tl >= o) incl += v; }
    R = __shfl_sync(tmask, incl, leader + TEAM - 1);
    int stop = 0;
    if (tl == 0) stop = (nev >= max_events_per_tile) || (D.err[0] != 0);
    stop = __shfl_sync(tmask, stop, leader);
    if (stop || !(R > 0.0)) break;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (tl == 0) rnd = pol.next4();
    rnd.x = __shfl_sync(tmask, rnd.x, leader); rnd.y = __shfl_sync(tmask, rnd.y, leader);
    rnd.z = __shfl_sync(tmask, rnd.z, leader); rnd.w = __shfl_sync(tmask, rnd.w, leader);
    tloc += -log(u53(rnd.x, rnd.y)) / R;
    if (tloc > tau) break;
    const double target = u53(rnd.z, rnd.w) * R;
    unsigned m = __ballot_sync(tmask, incl > target) & tmask;
    int sel = m ? (__ffs(m) - 1) : (leader + TEAM - 1);
    double before = __shfl_sync(tmask, incl - part, sel);
    int pick = -1;
    if (lane == sel) {
      double acc = before; int last = -1;
      for (int i = tl; i < n; i += TEAM) { double r = rate[i * NT]; if (r > 0.0) { last = i; acc += r; if (acc > target) { pick = i; break; } } }
      if (pick < 0) pick = last;
    }
    pick = __shfl_sync(tmask, pick, sel);
    if (pick < 0) {            // rounding left the chosen lane without a live entry: take any live one
      int mine = -1;
      for (int i = tl; i < n; i += TEAM) if (rate[i * NT] > 0.0) { mine = i; break; }
      unsigned mm = __ballot_sync(tmask, mine >= 0) & tmask;
      if (!mm) break;
      pick = __shfl_sync(tmask, mine, __ffs(mm) - 1);
    }
    if (tl == 0) {
      unsigned long long e = ce[pick * NT];
      uint32_t cell = (uint32_t)e; int Itrans = (int)(e >> 32);
      P3 Site = cell_xyz((long long)cell);
      if (Itrans == EV_DRAW) {                       // species / adsorbate drawn now (Get_Prob.f:164-179)
        uint32_t w = D.state[cell];
        int T = D.la ? site_T(w, Site) : 0;
        Itrans = choose_dep_abs(w_coor(w), T, pol.uniform(), nullptr);
      }
      if (Itrans >= 1 && Itrans <= 3) ndep[Itrans - 1]++; else if (Itrans == 4) nevap++; else if (Itrans == 5) ndes++; else nabs++;
      if (l == dbg_lid) printf("DBG step %lld ev#%d site %d %d %d Itrans %d w=%08x rate=%g R=%g tloc=%g tau=%g n=%d\n", ss.step, nev, Site.x, Site.y, Site.z, Itrans, D.state[cell], rate[pick * NT], R, tloc, tau, n);
      apply_event_q(pol, q, Site, Itrans);
      nev++;
    }
    __syncwarp(tmask);
    // the queued re-rate requests: one Get_Prob per lane, then the list updates (any order: slots are
    // claimed with atomics and every request concerns a different cell)
    const int nq = q.n;
    for (int k0 = 0; k0 < nq; k0 += TEAM) {
      const int k = k0 + tl;
      int ev = 0; double p = 0.0;
      if (k < nq && q.it[k].act != A_ERASE) { RQItem &it = q.it[k]; get_prob(pol, it.Occ, it.NSiC, it.Cov, it.Coor, it.N, it.nn, ev, p); }
      __syncwarp(tmask);
      if (k < nq) rq_commit(pol, q.it[k], ev, p);
      __syncwarp(tmask);
    }
    if (tl == 0) q.n = 0;
    __syncwarp(tmask);
  }
  // team totals -> leader
  long long ngp = pol.n_getprob; int dA = pol.dAtoms, dV = pol.dVoids;
#pragma unroll
  for (int o = TEAM / 2; o > 0; o >>= 1) {
    ngp += __shfl_down_sync(tmask, ngp, o, TEAM); dA += __shfl_down_sync(tmask, dA, o, TEAM); dV += __shfl_down_sync(tmask, dV, o, TEAM);
  }
  if (tl == 0) {
    V.tile_rate[l] = R;
    TileStat &st = V.tstat[l];
    st.events += nev; st.site_updates += ngp;
    for (int i = 0; i < 3; i++) { st.dep[i] += ndep[i]; st.dCrystal[i] += pol.CountCrystal[i]; st.dCov[i] += pol.CountCov[i]; }
    st.evap += nevap; st.abs_ += nabs; st.des += ndes; st.dAtoms += dA; st.dVoids += dV;
    if (nev > st.maxev) st.maxev = nev;
  }
}

// Per-sweep bookkeeping ("per-class rate sums"): total rate, active-site count, next tau, clock.
__global__ void __launch_bounds__(1024) sweep_stats_kernel(int ntiles, double events_per_tile, double tau_fixed, int advance) {
  __shared__ double sR[32]; __shared__ long long sN[32]; __shared__ int sM[32];
  double r = 0.0; long long n = 0; int mx = 0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) { r += V.tile_rate[t]; int c = V.cnt[t]; n += c; mx = max(mx, c); }
  for (int o = 16; o > 0; o >>= 1) { r += __shfl_down_sync(0xFFFFFFFFu, r, o); n += __shfl_down_sync(0xFFFFFFFFu, n, o); mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { sR[threadIdx.x >> 5] = r; sN[threadIdx.x >> 5] = n; sM[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double R = 0; long long N = 0; int M = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { R += sR[i]; N += sN[i]; M = max(M, sM[i]); }
    SweepState &s = *V.ss;
    if (advance) { s.time += s.tau; s.sweeps++; }
    s.total_rate = R; s.tree_root = R; s.n_active = N; s.max_cnt = M;
    s.tau = tau_fixed > 0 ? tau_fixed : (R > 0 ? events_per_tile * (double)ntiles / R : 0.0);
  }
}
__global__ void bump_step_kernel(SweepState *ss) { ss->step++; }

__global__ void __launch_bounds__(256) tile_rate_kernel(int ntiles) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= ntiles) return;
  int n = V.cnt[l]; double r = 0;
  for (int i = 0; i < n; i++) r += V.rate[(long long)i * V.NT + l];
  V.tile_rate[l] = r;
}

__global__ void __launch_bounds__(256) reduce_tstat_kernel(const TileStat *ts, int ntiles, TileStat *out) {
  // one block; 19 integer columns
  __shared__ long long acc[19];
  if (threadIdx.x < 19) acc[threadIdx.x] = 0;
  __syncthreads();
  long long loc[19];
  for (int k = 0; k < 19; k++) loc[k] = 0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    const TileStat &s = ts[t];
    loc[0] += s.events; loc[1] += s.site_updates; loc[2] += s.dep[0]; loc[3] += s.dep[1]; loc[4] += s.dep[2]; loc[5] += s.evap; loc[6] += s.abs_; loc[7] += s.des;
    loc[8] += s.dAtoms; loc[9] += s.dVoids; if (s.maxev > loc[16]) loc[16] = s.maxev; loc[10] += s.dCrystal[0]; loc[11] += s.dCrystal[1]; loc[12] += s.dCrystal[2]; loc[13] += s.dCov[0]; loc[14] += s.dCov[1]; loc[15] += s.dCov[2];
  }
  for (int k = 0; k < 16; k++) atomicAdd((unsigned long long *)&acc[k], (unsigned long long)loc[k]);
  atomicMax((long long *)&acc[16], loc[16]);
  __syncthreads();
  if (threadIdx.x == 0) {
    out->events = acc[0]; out->site_updates = acc[1]; out->dep[0] = (int)acc[2]; out->dep[1] = (int)acc[3]; out->dep[2] = (int)acc[4];
    out->evap = (int)acc[5]; out->abs_ = (int)acc[6]; out->des = (int)acc[7]; out->dAtoms = (int)acc[8]; out->dVoids = (int)acc[9];
    for (int i = 0; i < 3; i++) { out->dCrystal[i] = (int)acc[10 + i]; out->dCov[i] = (int)acc[13 + i]; }
    out->maxev = (int)acc[16];
  }
}

// ---- the dense "rate kernel": recompute (event class, rate) of every listed site from the lattice ----
struct RatePol {             // Get_Prob without any bookkeeping side effects
  double root; long long n_getprob; int CountCrystal[3], CountCov[3];
  __device__ __forceinline__ bool predraw() const { return false; }
  __device__ __forceinline__ double tree_root() const { return root; }
  __device__ __forceinline__ double uniform() { return 0.5; }
  __device__ __forceinline__ bool cov_void_recipr(P3, long long, P3) const { return false; }
};
__global__ void __launch_bounds__(256) rerate_kernel(int ntiles, unsigned long long *n_sites) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)ntiles * V.cap;
  RatePol pol; pol.root = V.ss->tree_root; pol.n_getprob = 0;
  unsigned long long mine = 0;
  for (long long k = gid; k < total; k += (long long)gridDim.x * blockDim.x) {
    int i = (int)(k / ntiles), l = (int)(k - (long long)i * ntiles);
    if (i >= V.cnt[l]) continue;
    double r = V.rate[k];
    if (!(r > 0.0)) continue;
    uint32_t cell = (uint32_t)V.ce[k];
    P3 s = cell_xyz((long long)cell);
    uint32_t w = D.state[cell];
    P3 nn[4]; dec_bonds(s, D.bonds[cell], nn);
    int ev = 0; double p = 0;
    get_prob(pol, w_occ(w), w_sp(w), w_occ(w) ? 0 : w_cov(w), w_coor(w), s, nn, ev, p);
    if (ev != 0 && p > 0) { V.rate[k] = p; V.ce[k] = (unsigned long long)cell | ((unsigned long long)(uint32_t)ev << 32); mine++; }
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_sites, mine);
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
// initial/SetT.f:36-46
__global__ void set_temperature_kernel(const int32_t *T, long long n) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    int t = T[c];
    if (t != 10) D.state[c] = MVBITS(D.state[c], t, LenT, PosT);
  }
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
  const long long plane = (long long)D.LX * D.LY;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    uint4 v = S4[q];
    uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t w = ww[k];
      if (w == 0) continue;
      int cov = w_cov(w);
      if (cov == 113) { nWall++; continue; }
      if (w_occ(w)) {
        nOcc++;
        long long c = 4 * q + k; int z = (int)(c / plane);
        if (w_coor(w) <= 3) { nUnder++; hz += z; hz2 += (double)z * z; }
        else if (w_coor(w) == 4) {
          nBulk++;
          int r = (int)(c - (long long)z * plane), y = r / D.LX, x = r - y * D.LX, kind;
          if (!is_cubic_site(x, y, z, kind)) nNoEpi++;
        }
      } else if (cov != 0) nCov++;
    }
  }
  unsigned long long vals[6] = {nUnder, nCov, nNoEpi, nBulk, nOcc, nWall};
  for (int k = 0; k < 6; k++) {
    unsigned long long v = vals[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[k], v);
  }
  for (int o = 16; o > 0; o >>= 1) { hz += __shfl_down_sync(0xFFFFFFFFu, hz, o); hz2 += __shfl_down_sync(0xFFFFFFFFu, hz2, o); }
  if ((threadIdx.x & 31) == 0 && (hz != 0 || hz2 != 0)) { atomicAdd(&hsum[0], hz); atomicAdd(&hsum[1], hz2); }
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
        atomicAdd(&counters[0], 1ull);
      } else if (z <= Len - 1) {
        D.state[c] = (1u << (PosSiC + sp - 1)) | (1u << PosOcc) | (2u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        double p = D.PE[IDXE(sp, 2, 0, 0, 0, 0, 0)];
        int t = list_of(s);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        V.rate[(long long)k * V.NT + t] = p; V.ce[(long long)k * V.NT + t] = (unsigned long long)(uint32_t)c | (4ull << 32); V.slot[c] = (uint16_t)(k + 1);
        atomicAdd(&counters[1], 1ull);
      } else {
        D.state[c] = (2u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        double p = total_dep_abs(2, 0);
        int t = list_of(s);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        V.rate[(long long)k * V.NT + t] = p; V.ce[(long long)k * V.NT + t] = (unsigned long long)(uint32_t)c | ((unsigned long long)EV_DRAW << 32); V.slot[c] = (uint16_t)(k + 1);
        atomicAdd(&counters[1], 1ull);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (err) *err = std::string("CUDA error ") + cudaGetErrorString(e_) + " in " #call; return -1; } } while (0)

static inline void parallel_free(ParallelCtx &P) {
  cudaFree(P.rate); cudaFree(P.ce); cudaFree(P.cnt); cudaFree(P.ntomb); cudaFree(P.slot); cudaFree(P.tile_rate); cudaFree(P.tstat); cudaFree(P.ss); cudaFree(P.tot);
  P = ParallelCtx();
}
static inline TileView tile_view(const ParallelCtx &P) {
  TileView V; V.tile = P.tile; V.ntx = P.ntx; V.nty = P.nty; V.cap = P.cap; V.NT = P.ntiles; V.hx = P.ntx / 2; V.ntc = (P.ntx / 2) * (P.nty / 2);
  V.rate = P.rate; V.ce = P.ce; V.cnt = P.cnt; V.ntomb = P.ntomb; V.slot = P.slot;
  V.tile_rate = P.tile_rate; V.tstat = P.tstat; V.ss = P.ss; return V;
}
static inline int parallel_alloc(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int cap, std::string *err) {
  int tile = 24;
  if (H.LX % (2 * tile) || H.LY % (2 * tile)) { if (err) *err = "parallel sampler needs LenX and LenY to be multiples of 48 (24-unit tiles, 2x2 colouring)"; return -1; }
  if (P.rate && P.cap == cap && P.ncell == H.ncell) {
    PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.ntomb, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.rate, 0, (size_t)P.ntiles * cap * sizeof(double), st));
    PCK(cudaMemsetAsync(P.slot, 0, (size_t)H.ncell * sizeof(uint16_t), st));
    PCK(cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles * sizeof(TileStat), st));
    PCK(cudaMemsetAsync(P.tile_rate, 0, (size_t)P.ntiles * sizeof(double), st));
    PCK(cudaMemsetAsync(P.ss, 0, sizeof(SweepState), st));
    return 0;
  }
  parallel_free(P);
  P.tile = tile; P.ntx = H.LX / tile; P.nty = H.LY / tile; P.ntiles = P.ntx * P.nty; P.cap = cap; P.ncell = H.ncell;
  PCK(cudaMalloc(&P.rate, (size_t)P.ntiles * cap * sizeof(double)));
  PCK(cudaMalloc(&P.ce, (size_t)P.ntiles * cap * sizeof(unsigned long long)));
  PCK(cudaMalloc(&P.cnt, (size_t)P.ntiles * sizeof(int)));
  PCK(cudaMalloc(&P.ntomb, (size_t)P.ntiles * sizeof(int)));
  PCK(cudaMalloc(&P.slot, (size_t)H.ncell * sizeof(uint16_t)));
  PCK(cudaMalloc(&P.tile_rate, (size_t)P.ntiles * sizeof(double)));
  PCK(cudaMalloc(&P.tstat, (size_t)P.ntiles * sizeof(TileStat)));
  PCK(cudaMalloc(&P.ss, sizeof(SweepState)));
  PCK(cudaMalloc(&P.tot, sizeof(TileStat)));
  PCK(cudaMemsetAsync(P.rate, 0, (size_t)P.ntiles * cap * sizeof(double), st));
  PCK(cudaMemsetAsync(P.ce, 0, (size_t)P.ntiles * cap * sizeof(unsigned long long), st));
  PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
  PCK(cudaMemsetAsync(P.ntomb, 0, (size_t)P.ntiles * sizeof(int), st));
  PCK(cudaMemsetAsync(P.slot, 0, (size_t)H.ncell * sizeof(uint16_t), st));
  PCK(cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles * sizeof(TileStat), st));
  PCK(cudaMemsetAsync(P.tile_rate, 0, (size_t)P.ntiles * sizeof(double), st));
  PCK(cudaMemsetAsync(P.ss, 0, sizeof(SweepState), st));
  return 0;
}
static const int TILE_CAP = 192;

// tile lists from an uploaded MC list (reference order is irrelevant here)
static inline int parallel_build_from_lists(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int n_mc, const uint32_t *mc_cell, const int *mc_ev,
                                            const double *mc_rate, int n_bulk, int n_void, std::string *err) {
  if (H.LX % 48 || H.LY % 48) { P.lists_stale = true; return 0; }      // parity-only box
  if (parallel_alloc(P, H, st, TILE_CAP, err)) return -1;
  std::vector<double> R((size_t)P.ntiles * P.cap, 0.0); std::vector<unsigned long long> CE((size_t)P.ntiles * P.cap, 0ull);
  std::vector<int> cnt(P.ntiles, 0); std::vector<uint16_t> slot((size_t)H.ncell, 0);
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
    size_t a = (size_t)cnt[t] * P.ntiles + t;
    R[a] = mc_rate[i]; CE[a] = (unsigned long long)c | ((unsigned long long)e << 32);
    slot[c] = (uint16_t)(++cnt[t]);
  }
  PCK(cudaMemcpyAsync(P.rate, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.ce, CE.data(), CE.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  PCK(cudaMemcpyAsync(P.slot, slot.data(), slot.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  PCK(cudaStreamSynchronize(st));
  P.base_atoms = n_bulk; P.base_voids = n_void; P.time = 0; P.ran = false; P.lists_stale = false;
  for (int i = 0; i < 3; i++) { P.baseCrystal[i] = 0; P.baseCov[i] = 0; }
  memset(&P.h_tot, 0, sizeof P.h_tot);
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
  init_flat_kernel<<<148 * 8, 256, 0, st>>>(len1, walls, two_species, d_cnt);
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  cudaFree(d_cnt);
  P.base_atoms = (long long)h_cnt[0]; P.base_voids = 0; P.time = 0; P.ran = false; P.lists_stale = false;
  memset(&P.h_tot, 0, sizeof P.h_tot);
  return 0;
}

static inline int parallel_run(ParallelCtx &P, DevCtx &H, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, const mskc_parallel_args *a, mskc_parallel_result *r, std::string *err) {
  memset(r, 0, sizeof *r);
  const int ntc = (P.ntx / 2) * (P.nty / 2);
  int teamsz = 8; if (const char *e = getenv("MSKC_TEAM")) teamsz = atoi(e);
  if (teamsz != 1 && teamsz != 2 && teamsz != 4 && teamsz != 8 && teamsz != 16 && teamsz != 32) teamsz = 8;
  const int tpw = 32 / teamsz;
  const int blocks = (ntc + tpw - 1) / tpw;
  int dbg_lid = -1; if (const char *e = getenv("MSKC_DBG_LID")) dbg_lid = atoi(e);
  const double ept = a->events_per_tile > 0 ? a->events_per_tile : 4.0;
  uint2 key = make_uint2((uint32_t)a->seed, (uint32_t)(a->seed >> 32));
  TileStat before = P.h_tot;
  int launches = 0;
  PCK(cudaEventRecord(e0, st));
  if (!P.ran) {                    // first call: establish total rate and tau
    tile_rate_kernel<<<(P.ntiles + 255) / 256, 256, 0, st>>>(P.ntiles);
    sweep_stats_kernel<<<1, 1024, 0, st>>>(P.ntiles, ept, a->tau_fixed, 0);
    launches += 2;
  }
  static const int order[4][2] = {{0, 0}, {1, 1}, {1, 0}, {0, 1}};
  // short calls time every colour-step kernel on its own (CUDA events on the launching stream): this is
  // the per-launch duration the roofline in bench.py is computed from
  const bool timed = a->n_sweeps > 0 && a->n_sweeps <= 64;
  std::vector<cudaEvent_t> evs;
  if (timed) { evs.resize((size_t)a->n_sweeps * 8); for (auto &e : evs) PCK(cudaEventCreate(&e)); }
  size_t ke = 0;
  for (long long s = 0; s < a->n_sweeps; s++) {
    for (int c = 0; c < 4; c++) {
      if (timed) PCK(cudaEventRecord(evs[ke++], st));
      const int col = order[c][1] * 2 + order[c][0];
      switch (teamsz) {
        case 1: colour_step_kernel<1><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
        case 2: colour_step_kernel<2><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
        case 4: colour_step_kernel<4><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
        case 8: colour_step_kernel<8><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
        case 16: colour_step_kernel<16><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
        default: colour_step_kernel<32><<<blocks, 32, 0, st>>>(col, key, 1 << 20, dbg_lid); break;
      }
      if (timed) PCK(cudaEventRecord(evs[ke++], st));
      bump_step_kernel<<<1, 1, 0, st>>>(P.ss);
      launches += 2;
    }
    sweep_stats_kernel<<<1, 1024, 0, st>>>(P.ntiles, ept, a->tau_fixed, 1);
    launches++;
  }
  reduce_tstat_kernel<<<1, 256, 0, st>>>(P.tstat, P.ntiles, P.tot);
  launches++;
  PCK(cudaEventRecord(e1, st));
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(&P.h_tot, P.tot, sizeof(TileStat), cudaMemcpyDeviceToHost, st));
  PCK(cudaMemcpyAsync(&P.h_ss, P.ss, sizeof(SweepState), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
  PCK(cudaEventElapsedTime(&r->kernel_ms, e0, e1));
  if (timed) {
    float tot = 0;
    for (size_t i = 0; i + 1 < evs.size(); i += 2) { float ms = 0; PCK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1])); tot += ms; }
    r->sweep_kernel_ms = tot;
    for (auto &e : evs) cudaEventDestroy(e);
  }
  P.ran = true;
  r->events_done = P.h_tot.events - before.events; r->site_updates = P.h_tot.site_updates - before.site_updates;
  r->sweeps_done = a->n_sweeps; r->time = P.h_ss.time; r->total_rate = P.h_ss.total_rate; r->n_active = P.h_ss.n_active;
  for (int i = 0; i < 3; i++) r->count_dep[i] = P.h_tot.dep[i] - before.dep[i];
  r->count_evap = P.h_tot.evap - before.evap; r->count_abs = P.h_tot.abs_ - before.abs_; r->count_des = P.h_tot.des - before.des;
  r->launches = launches; r->stop_reason = P.h_tot.maxev;   /* diagnostic: max events of one tile in one colour step */
  (void)ke;
  P.time = P.h_ss.time;
  return 0;
}

static inline int parallel_rerate(ParallelCtx &P, DevCtx &H, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, int64_t *n_sites, double *total_rate, float *ms, std::string *err) {
  unsigned long long *d_n = nullptr, h_n = 0;
  PCK(cudaMalloc(&d_n, 8)); PCK(cudaMemsetAsync(d_n, 0, 8, st));
  PCK(cudaEventRecord(e0, st));
  rerate_kernel<<<148 * 8, 256, 0, st>>>(P.ntiles, d_n);
  tile_rate_kernel<<<(P.ntiles + 255) / 256, 256, 0, st>>>(P.ntiles);
  sweep_stats_kernel<<<1, 1024, 0, st>>>(P.ntiles, 4.0, 0.0, 0);
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
  c->num_atoms = (int32_t)(P.base_atoms + P.h_tot.dAtoms); c->num_voids = (int32_t)(P.base_voids + P.h_tot.dVoids);
  c->num_adatom = (int32_t)P.h_ss.n_active;
  for (int i = 0; i < 3; i++) { c->CountCrystal[i] = P.baseCrystal[i] + P.h_tot.dCrystal[i]; c->CountCov[i] = P.baseCov[i] + P.h_tot.dCov[i]; }
  c->iter = P.h_tot.events + 1; c->time = P.h_ss.time;
}

}  // namespace mskc
