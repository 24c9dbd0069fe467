// mskc_parallel.cuh -- synchronous-sublattice sampler + the dense helper kernels.
//
// The x-y plane is cut into square tiles (edge 24 super-lattice units = 2 cubic cells) coloured 2x2.
// An event at S writes S and its 4 neighbours (<= 1 bond, |d| <= 8 per axis incl. hexagonal bonds) and
// reads at most 2 bonds away (reciprocity of neighbours' neighbours, zig-zag/arm-chair candidates),
// so two tiles of the same colour (>= 24 units apart) never touch the same cell: all tiles of a colour
// run concurrently, one warp per tile, each doing its own n-fold way for a time quantum tau.
//
// HBM layout added to state[]/bonds[] (mskc_device.cuh):
//   entries[tile][cap]  16 B {f64 rate, u32 cell, u32 event}   the tile's active-site list
//   cnt[tile]           i32  slots in use (tombstones with rate 0 included until the owner compacts)
//   slot[cell]          u16  1-based slot of the cell in its tile's list, 0 = not listed
//   tstat[tile]         per-tile counters (events, site updates, per-class counts), summed on demand
#pragma once
#include "mskc_device.cuh"
#include "../../include/mskc.h"
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

namespace mskc {

struct __align__(16) TileEntry { double rate; uint32_t cell; uint32_t ev; };

struct TileStat {            // 64 B, written only by the tile's own warp
  long long events, site_updates;
  int dep[3], evap, abs_, des;
  int dAtoms, dVoids;        // bulk / vacancy list size changes
  int dCrystal[3]; int dCov[3];
};

struct SweepState {          // one per context, device resident
  double time, tau, total_rate, tree_root;
  long long n_active, sweeps, step;     // step = global colour-step counter (Philox counter word)
  long long max_cnt;
};

struct ParallelCtx {
  int tile = 24, ntx = 0, nty = 0, ntiles = 0, cap = 0;
  TileEntry *entries = nullptr; int *cnt = nullptr; uint16_t *slot = nullptr; double *tile_rate = nullptr;
  TileStat *tstat = nullptr; SweepState *ss = nullptr; TileStat *tot = nullptr;
  long long ncell = 0;
  long long base_atoms = 0, base_voids = 0;
  int baseCrystal[3] = {0, 0, 0}, baseCov[3] = {0, 0, 0};
  TileStat h_tot{}; SweepState h_ss{};
  double time = 0.0;
  bool ran = false, lists_stale = false, listed_flag_in_state = false;
};

struct TileView {            // what the kernels need, by value
  int tile, ntx, nty, cap;
  TileEntry *entries; int *cnt; uint16_t *slot; double *tile_rate; TileStat *tstat; SweepState *ss;
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
  __device__ __forceinline__ int tile_of(P3 s) const { return (s.y / V.tile) * V.ntx + (s.x / V.tile); }

  __device__ __noinline__ void put(P3 s, long long c, int ev, double p) {
    int t = tile_of(s);
    int sl = V.slot[c];
    TileEntry e; e.rate = p; e.cell = (uint32_t)c; e.ev = (uint32_t)ev;
    if (sl == 0) {
      int k = atomicAdd(&V.cnt[t], 1);
      if (k >= V.cap) { atomicSub(&V.cnt[t], 1); fail(E_TILE_FULL, s.x, s.y, s.z, t); return; }
      V.slot[c] = (uint16_t)(k + 1);
      sl = k + 1;
    }
    V.entries[(long long)t * V.cap + sl - 1] = e;
  }
  __device__ __forceinline__ void add_mc(P3 s, long long c, const P3 nn[4], int ev, double p) {
    if (p == 0) { fail(E_ADD_ZERO, s.x, s.y, s.z); return; }
    D.bonds[c] = enc_bonds(s, nn);
    put(s, c, ev, p);
  }
  __device__ __forceinline__ void set_rate(P3 s, long long c, int ev, double p) { put(s, c, ev, p); }
  __device__ __forceinline__ void erase_mc(P3 s, long long c) {
    int sl = V.slot[c];
    if (sl) V.entries[(long long)tile_of(s) * V.cap + sl - 1].rate = 0.0;      // tombstone; the owner compacts
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

// One colour step: every tile of colour (cx,cy) advances by tau.  One warp per tile; the lanes share
// the selection scan, lane 0 applies the event (the transition rules are a serial dependency chain).
__global__ void __launch_bounds__(128) colour_step_kernel(int cx, int cy, uint2 key, int max_events_per_tile) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int hx = V.ntx >> 1, hy = V.nty >> 1;
  if (wid >= hx * hy) return;
  const int tx = 2 * (wid % hx) + cx, ty = 2 * (wid / hx) + cy;
  const int t = ty * V.ntx + tx;
  TileEntry *E = V.entries + (long long)t * V.cap;
  const SweepState ss = *V.ss;
  const double tau = ss.tau;
  int n = V.cnt[t];

  // -- compaction: drop tombstones (rate 0), keep slot[] consistent
  {
    int w = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      TileEntry e; e.rate = 0; e.cell = 0; e.ev = 0;
      bool live = false;
      if (i < n) { e = E[i]; live = e.rate > 0.0; }
      unsigned m = __ballot_sync(0xFFFFFFFFu, live);
      int pos = w + __popc(m & ((1u << lane) - 1u));
      __syncwarp();
      if (i < n) {
        if (live) { if (pos != i) { E[pos] = e; V.slot[e.cell] = (uint16_t)(pos + 1); } }
        else V.slot[e.cell] = 0;
      }
      w += __popc(m);
      __syncwarp();
    }
    if (w != n && lane == 0) V.cnt[t] = w;
    n = w;
  }
  __syncwarp();

  TilePol pol;
  pol.key = key; pol.tile_id = (uint32_t)t; pol.step_lo = (uint32_t)ss.step; pol.step_hi = (uint32_t)(ss.step >> 32); pol.draw = 0;
  pol.root = ss.tree_root; pol.n_getprob = 0; pol.dAtoms = 0; pol.dVoids = 0;
  for (int i = 0; i < 3; i++) { pol.CountCrystal[i] = 0; pol.CountCov[i] = 0; }
  int nev = 0, ndep[3] = {0, 0, 0}, nevap = 0, nabs = 0, ndes = 0;
  double tloc = 0.0, Rlast = 0.0;

  while (true) {
    n = V.cnt[t];                                   // own-tile appends happen through lane 0 of this warp
    // per-lane partial sums over entries lane, lane+32, ...
    double part = 0.0;
    for (int i = lane; i < n; i += 32) part += E[i].rate;
    double incl = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
    double R = __shfl_sync(0xFFFFFFFFu, incl, 31);
    Rlast = R;
    if (!(R > 0.0) || nev >= max_events_per_tile || D.err[0] != 0) break;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (lane == 0) rnd = pol.next4();
    rnd.x = __shfl_sync(0xFFFFFFFFu, rnd.x, 0); rnd.y = __shfl_sync(0xFFFFFFFFu, rnd.y, 0);
    rnd.z = __shfl_sync(0xFFFFFFFFu, rnd.z, 0); rnd.w = __shfl_sync(0xFFFFFFFFu, rnd.w, 0);
    tloc += -log(u53(rnd.x, rnd.y)) / R;
    if (tloc > tau) break;
    double target = u53(rnd.z, rnd.w) * R;
    // the lane whose inclusive prefix first exceeds target
    unsigned m = __ballot_sync(0xFFFFFFFFu, incl > target);
    int sel_lane = m ? (__ffs(m) - 1) : 31;
    double before = __shfl_sync(0xFFFFFFFFu, incl - part, sel_lane);
    int pick = -1;
    if (lane == sel_lane) {
      double acc = before; int last = -1;
      for (int i = lane; i < n; i += 32) { double r = E[i].rate; if (r > 0.0) { last = i; acc += r; if (acc > target) { pick = i; break; } } }
      if (pick < 0) pick = last;
    }
    pick = __shfl_sync(0xFFFFFFFFu, pick, sel_lane);
    if (pick < 0) break;
    {
      // Cooperative gather: the transition rules are one serial dependency chain (lane 0), but every
      // address they will touch is known up front -- the site, its 4 neighbours and their 16 neighbours.
      // Lanes 0-3 / 0-15 pull those words into this SM's L1 in two round trips so that lane 0 then runs
      // on L1 hits instead of ~20 dependent L2/HBM round trips.
      uint32_t cS = E[pick].cell;
      unsigned long long bS = D.bonds[cS];
      touch32(D.state + cS);
      P3 S = cell_xyz((long long)cS);
      unsigned long long bN = 0ull; P3 N = zero3(); bool okN = false;
      if (lane < 4) {
        N = dec_bond(S, (unsigned)((bS >> (16 * lane)) & 0xFFFFu));
        okN = in_box(N);
        if (okN) { long long cN = cidx_nofail(N); bN = D.bonds[cN]; touch32(D.state + cN); touch16(V.slot + cN); }
      }
      int src = lane >> 2;
      unsigned long long bq = __shfl_sync(0xFFFFFFFFu, bN, src & 3);
      P3 Nq; Nq.x = __shfl_sync(0xFFFFFFFFu, N.x, src & 3); Nq.y = __shfl_sync(0xFFFFFFFFu, N.y, src & 3); Nq.z = __shfl_sync(0xFFFFFFFFu, N.z, src & 3);
      bool okq = __shfl_sync(0xFFFFFFFFu, (int)okN, src & 3) != 0;
      if (lane < 16 && okq) {
        P3 M = dec_bond(Nq, (unsigned)((bq >> (16 * (lane & 3))) & 0xFFFFu));
        if (in_box(M)) { long long cM = cidx_nofail(M); touch32(D.state + cM); touch64(D.bonds + cM); }
      }
    }
    if (lane == 0) {
      TileEntry e = E[pick];
      P3 Site = cell_xyz((long long)e.cell);
      int Itrans = (int)e.ev;
      if (Itrans == EV_DRAW) {                       // species / adsorbate drawn now (Get_Prob.f:164-179)
        uint32_t w = D.state[e.cell];
        int T = D.la ? site_T(w, Site) : 0;
        Itrans = choose_dep_abs(w_coor(w), T, pol.uniform(), nullptr);
      }
      if (Itrans >= 1 && Itrans <= 3) ndep[Itrans - 1]++; else if (Itrans == 4) nevap++; else if (Itrans == 5) ndes++; else nabs++;
      apply_event(pol, Site, Itrans);
      nev++;
    }
    __syncwarp();
  }
  if (lane == 0) {
    V.tile_rate[t] = Rlast;
    TileStat &st = V.tstat[t];
    st.events += nev; st.site_updates += pol.n_getprob;
    for (int i = 0; i < 3; i++) { st.dep[i] += ndep[i]; st.dCrystal[i] += pol.CountCrystal[i]; st.dCov[i] += pol.CountCov[i]; }
    st.evap += nevap; st.abs_ += nabs; st.des += ndes; st.dAtoms += pol.dAtoms; st.dVoids += pol.dVoids;
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
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= ntiles) return;
  const TileEntry *E = V.entries + (long long)t * V.cap;
  int n = V.cnt[t]; double part = 0;
  for (int i = lane; i < n; i += 32) part += E[i].rate;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, o);
  if (lane == 0) V.tile_rate[t] = part;
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
    loc[8] += s.dAtoms; loc[9] += s.dVoids; loc[10] += s.dCrystal[0]; loc[11] += s.dCrystal[1]; loc[12] += s.dCrystal[2]; loc[13] += s.dCov[0]; loc[14] += s.dCov[1]; loc[15] += s.dCov[2];
  }
  for (int k = 0; k < 16; k++) atomicAdd((unsigned long long *)&acc[k], (unsigned long long)loc[k]);
  __syncthreads();
  if (threadIdx.x == 0) {
    out->events = acc[0]; out->site_updates = acc[1]; out->dep[0] = (int)acc[2]; out->dep[1] = (int)acc[3]; out->dep[2] = (int)acc[4];
    out->evap = (int)acc[5]; out->abs_ = (int)acc[6]; out->des = (int)acc[7]; out->dAtoms = (int)acc[8]; out->dVoids = (int)acc[9];
    for (int i = 0; i < 3; i++) { out->dCrystal[i] = (int)acc[10 + i]; out->dCov[i] = (int)acc[13 + i]; }
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
    int t = (int)(k / V.cap), i = (int)(k - (long long)t * V.cap);
    if (i >= V.cnt[t]) continue;
    TileEntry e = V.entries[k];
    if (!(e.rate > 0.0)) continue;
    P3 s = cell_xyz((long long)e.cell);
    uint32_t w = D.state[e.cell];
    P3 nn[4]; dec_bonds(s, D.bonds[e.cell], nn);
    int ev = 0; double p = 0;
    get_prob(pol, w_occ(w), w_sp(w), w_occ(w) ? 0 : w_cov(w), w_coor(w), s, nn, ev, p);
    if (ev != 0 && p > 0) { e.rate = p; e.ev = (uint32_t)ev; V.entries[k] = e; mine++; }
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
        int t = (y / V.tile) * V.ntx + (x / V.tile);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        TileEntry e; e.rate = p; e.cell = (uint32_t)c; e.ev = 4; V.entries[(long long)t * V.cap + k] = e; V.slot[c] = (uint16_t)(k + 1);
        atomicAdd(&counters[1], 1ull);
      } else {
        D.state[c] = (2u << PosCoor);
        D.bonds[c] = enc_bonds(s, nn);
        double p = total_dep_abs(2, 0);
        int t = (y / V.tile) * V.ntx + (x / V.tile);
        int k = atomicAdd(&V.cnt[t], 1);
        if (k >= V.cap) { fail(E_TILE_FULL, x, y, z, t); continue; }
        TileEntry e; e.rate = p; e.cell = (uint32_t)c; e.ev = EV_DRAW; V.entries[(long long)t * V.cap + k] = e; V.slot[c] = (uint16_t)(k + 1);
        atomicAdd(&counters[1], 1ull);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (err) *err = std::string("CUDA error ") + cudaGetErrorString(e_) + " in " #call; return -1; } } while (0)

static inline void parallel_free(ParallelCtx &P) {
  cudaFree(P.entries); cudaFree(P.cnt); cudaFree(P.slot); cudaFree(P.tile_rate); cudaFree(P.tstat); cudaFree(P.ss); cudaFree(P.tot);
  P = ParallelCtx();
}
static inline TileView tile_view(const ParallelCtx &P) {
  TileView V; V.tile = P.tile; V.ntx = P.ntx; V.nty = P.nty; V.cap = P.cap; V.entries = P.entries; V.cnt = P.cnt; V.slot = P.slot;
  V.tile_rate = P.tile_rate; V.tstat = P.tstat; V.ss = P.ss; return V;
}
static inline int parallel_alloc(ParallelCtx &P, const DevCtx &H, cudaStream_t st, int cap, std::string *err) {
  int tile = 24;
  if (H.LX % (2 * tile) || H.LY % (2 * tile)) { if (err) *err = "parallel sampler needs LenX and LenY to be multiples of 48 (24-unit tiles, 2x2 colouring)"; return -1; }
  if (P.entries && P.cap == cap && P.ncell == H.ncell) {
    PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
    PCK(cudaMemsetAsync(P.slot, 0, (size_t)H.ncell * sizeof(uint16_t), st));
    PCK(cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles * sizeof(TileStat), st));
    PCK(cudaMemsetAsync(P.tile_rate, 0, (size_t)P.ntiles * sizeof(double), st));
    PCK(cudaMemsetAsync(P.ss, 0, sizeof(SweepState), st));
    return 0;
  }
  parallel_free(P);
  P.tile = tile; P.ntx = H.LX / tile; P.nty = H.LY / tile; P.ntiles = P.ntx * P.nty; P.cap = cap; P.ncell = H.ncell;
  PCK(cudaMalloc(&P.entries, (size_t)P.ntiles * cap * sizeof(TileEntry)));
  PCK(cudaMalloc(&P.cnt, (size_t)P.ntiles * sizeof(int)));
  PCK(cudaMalloc(&P.slot, (size_t)H.ncell * sizeof(uint16_t)));
  PCK(cudaMalloc(&P.tile_rate, (size_t)P.ntiles * sizeof(double)));
  PCK(cudaMalloc(&P.tstat, (size_t)P.ntiles * sizeof(TileStat)));
  PCK(cudaMalloc(&P.ss, sizeof(SweepState)));
  PCK(cudaMalloc(&P.tot, sizeof(TileStat)));
  PCK(cudaMemsetAsync(P.entries, 0, (size_t)P.ntiles * cap * sizeof(TileEntry), st));
  PCK(cudaMemsetAsync(P.cnt, 0, (size_t)P.ntiles * sizeof(int), st));
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
  std::vector<TileEntry> E((size_t)P.ntiles * P.cap); std::vector<int> cnt(P.ntiles, 0); std::vector<uint16_t> slot((size_t)H.ncell, 0);
  memset(E.data(), 0, E.size() * sizeof(TileEntry));
  long long plane = (long long)H.LX * H.LY;
  for (int i = 0; i < n_mc; i++) {
    uint32_t c = mc_cell[i];
    long long r = c % plane; int y = (int)(r / H.LX), x = (int)(r - (long long)y * H.LX);
    int t = (y / P.tile) * P.ntx + (x / P.tile);
    if (cnt[t] >= P.cap) { if (err) *err = "per-tile active list capacity exceeded at upload"; return -1; }
    TileEntry e; e.rate = mc_rate[i]; e.cell = c;
    // pre-drawn deposition/adsorption codes (1..3, >=6) of empty sites become "draw at fire time",
    // except the forced code-1 entry of Get_Prob.f:182-183 (rate 1e-8)
    int ev = mc_ev[i];
    bool forced = (ev == 1 && mc_rate[i] == (double)1.e-8f) || (ev == 1 && mc_rate[i] == (double)1.e-6f);
    e.ev = (ev == 4 || ev == 5 || forced) ? (uint32_t)ev : (uint32_t)EV_DRAW;
    E[(size_t)t * P.cap + cnt[t]] = e; slot[c] = (uint16_t)(++cnt[t]);
  }
  PCK(cudaMemcpyAsync(P.entries, E.data(), E.size() * sizeof(TileEntry), cudaMemcpyHostToDevice, st));
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
  const int wpb = 4, blocks = (ntc + wpb - 1) / wpb;
  const double ept = a->events_per_tile > 0 ? a->events_per_tile : 4.0;
  uint2 key = make_uint2((uint32_t)a->seed, (uint32_t)(a->seed >> 32));
  TileStat before = P.h_tot;
  int launches = 0;
  PCK(cudaEventRecord(e0, st));
  if (!P.ran) {                    // first call: establish total rate and tau
    tile_rate_kernel<<<(P.ntiles * 32 + 255) / 256, 256, 0, st>>>(P.ntiles);
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
      colour_step_kernel<<<blocks, wpb * 32, 0, st>>>(order[c][0], order[c][1], key, 1 << 20);
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
  r->launches = launches; r->stop_reason = MSKC_STOP_BUDGET;
  (void)ke;
  P.time = P.h_ss.time;
  return 0;
}

static inline int parallel_rerate(ParallelCtx &P, DevCtx &H, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, int64_t *n_sites, double *total_rate, float *ms, std::string *err) {
  unsigned long long *d_n = nullptr, h_n = 0;
  PCK(cudaMalloc(&d_n, 8)); PCK(cudaMemsetAsync(d_n, 0, 8, st));
  PCK(cudaEventRecord(e0, st));
  rerate_kernel<<<148 * 8, 256, 0, st>>>(P.ntiles, d_n);
  tile_rate_kernel<<<(P.ntiles * 32 + 255) / 256, 256, 0, st>>>(P.ntiles);
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
