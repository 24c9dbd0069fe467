// mskc_api.cu -- the C ABI of include/mskc.h: context, state upload/download, kernel launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 (see __graft_entry__.build()).
#include "../../include/mskc.h"
#include "mskc_device.cuh"
#include "mskc_parity.cuh"
#include "mskc_parallel.cuh"
#include "mskc_parity_coop.cuh"
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <string>
#include <algorithm>
#include <vector>

using namespace mskc;

static thread_local std::string g_last_error;

struct mskc_ctx {
  DevCtx D{};                 // device pointers inside
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  std::string err;
  // tables
  double *d_PE = nullptr, *d_PD = nullptr, *d_PAbs = nullptr, *d_PDes = nullptr;
  std::vector<double> h_PE, h_PD, h_PAbs, h_PDes;
  short *d_la_slot = nullptr; double *d_la_E = nullptr, *d_la_D = nullptr, *d_la_A = nullptr, *d_la_S = nullptr;
  int *d_err = nullptr;
  int32_t *d_stage = nullptr;
  uint32_t *d_frame_cell = nullptr, *d_frame_word = nullptr; unsigned long long *d_frame_n = nullptr; int64_t frame_cap = 0;
  unsigned short *d_arm_tab = nullptr;
  uint32_t *d_pa_tab = nullptr;
  // parity mode
  int NumMcMax = 500000, NumVdMax = 100000, Levels = 20;
  double *d_Tree = nullptr; uint32_t *d_lattind = nullptr, *d_mc_cell = nullptr, *d_void_cell = nullptr; int *d_mc_ev = nullptr;
  ParityState *d_ps = nullptr;
  ParityState h_ps{};
  bool parity_ready = false;
  // parallel mode
  ParallelCtx PC{};
  bool parallel_ready = false;
  bool have_rates = false, have_state = false;
  int idum = 0, runtype = 'R';
  // LaserAnnealing fork (include/mskc.h: mskc_set_tree_variant / mskc_set_la_fork)
  int tree_variant = 0;
  double *d_la3_F = nullptr;
  int fork_occ = 0, fork_occ0 = 0;
  float last_dense_ms = 0.0f;
};

// lattice words and event codes of the LaserAnnealing fork <-> the internal (mulskips-source) ones
static inline uint32_t word_to_dev(uint32_t w) { return (w & ~(7u << 27)) | (((w >> 28) & 3u) << 27); }     // la/modules/defsystem.f:99-100 -> src/modules/defsystem.f:90-91
static inline uint32_t word_from_dev(uint32_t w) { return (w & ~(7u << 27)) | (((w >> 27) & 3u) << 28); }
static inline int ev_to_dev(int e) { return e <= 2 ? e : e + 1; }                                           // la/main.f:671-699 -> src/main.f:335-357
static inline int ev_from_dev(int e) { return (e <= 2 || e == 100) ? e : e - 1; }
static int push_fork_CL(mskc_ctx *h);

static int upload_consts(mskc_ctx *h);
static int set_err(mskc_ctx *h, const char *fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf;
  g_last_error = buf;
  return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(h, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); } while (0)

// Make h the context the kernels see: DevCtx, tile lists and parity structures go to constant memory.
// Every API call that launches kernels does this first; calls return only after their stream is idle, so
// no kernel of another context is in flight when the constants change (one host thread per process).
static int upload_consts(mskc_ctx *h) {
  CK(cudaSetDevice(h->device));
  h->D.aux = h->PC.aux;
  TileView tv = tile_view(h->PC);
  ParityView pv{};
  pv.Tree = h->d_Tree; pv.Levels = h->Levels; pv.NumMcMax = h->NumMcMax; pv.NumVdMax = h->NumVdMax; pv.leaf0 = 1 << (h->Levels - 1);
  pv.lattind = h->d_lattind; pv.mc_cell = h->d_mc_cell; pv.void_cell = h->d_void_cell; pv.mc_ev = h->d_mc_ev; pv.ps = h->d_ps;
  CK(cudaMemcpyToSymbolAsync(mskc::D, &h->D, sizeof(DevCtx), 0, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyToSymbolAsync(mskc::V, &tv, sizeof(TileView), 0, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyToSymbolAsync(mskc::PV, &pv, sizeof(ParityView), 0, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static const char *dev_err_text(int code) {
  switch (code) {
    case E_OOB: return "lattice access outside 0..LenZ (the reference would read foreign memory)";
    case E_FARBOND: return "bond vector outside the +-15 range of the packed bond code";
    case E_TREE_ZERO: return "error stabilizzato  da dentro Update!";
    case E_TREE_IDX: return "Index > NumAdAtom";
    case E_PICK: return "write Index> NumAdAtom";
    case E_ADD_ZERO: return "provo ad aggiungere prob nulla";
    case E_MC_FULL: return "troppe particelle cinetiche";
    case E_VOID_FULL: return "troppe vacanze";
    case E_ERASE: return "err: EraseMC/EraseVoid";
    case E_COVOCC: return "ERROR Covind!=0 with Occ!=0 in get_prob";
    case E_COVLIST: return "ERROR: could not find CovInd inside ListCov";
    case E_CENSUS: return "Get Prob Error N_Si+N_C+N_B+N_wall.NE.Coor";
    case E_TEMP0: return "Temperature is zero in get_prob().";
    case E_NNCOV: return "ERROR in deposition: Site can only have NN with CovInd=0 or 113 when its Occ=1";
    case E_TILE_FULL: return "per-tile active list capacity exceeded";
    case E_STALE_NBUFF: return "Evaporation: no reciprocal occupied neighbour left (reference uses an uninitialised buffer here)";
    case E_TRANGE: return "site temperature outside the table range given to mskc_set_la";
    case E_BADEV: return "invalid event code";
    case E_FLAGS: return "incremental census flags disagree with the full neighbour census";
    case E_HALO_TIMEOUT: return "slab mode: the neighbouring rank did not deliver its halo records (it stopped, or its process is gone)";
  }
  return "unknown device error";
}
static int check_dev_err(mskc_ctx *h) {
  int e[8];
  CK(cudaMemcpyAsync(e, h->d_err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (e[0] != 0) return set_err(h, "%s [code %d; %d %d %d %d]", dev_err_text(e[0]), e[0], e[1], e[2], e[3], e[4]);
  return 0;
}

extern "C" const char *mskc_last_error(const mskc_ctx *h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int mskc_create(int lenx, int leny, int lenz, int ncrystal, int ncov, const int32_t *listcrystal, const int32_t *listcov,
                           int device_id, int nummcmax, mskc_ctx **out) {
  mskc_ctx *h = nullptr;
  if (!out) return set_err(nullptr, "mskc_create: out is NULL");
  *out = nullptr;
  if (lenx < 12 || leny < 12 || lenz < 12) return set_err(nullptr, "mskc_create: box too small");
  if (ncrystal < 1 || ncrystal > 3) return set_err(nullptr, "ERR NOT IMPLEMENTED: NCrystal must be <= 3");
  if (ncov < 0 || ncov > 3) return set_err(nullptr, "ERR NOT IMPLEMENTED: NCov must be <= 3");
  if ((long long)lenx * leny * ((long long)lenz + 1) >= 0xFFFFFFF0ll) return set_err(nullptr, "mskc_create: more than 2^32 cells per GPU");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return set_err(nullptr, "mskc_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  if (device_id < 0 || device_id >= ndev) return set_err(nullptr, "mskc_create: device %d of %d", device_id, ndev);
  h = new mskc_ctx();
  h->device = device_id;
  CK(cudaSetDevice(device_id));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1)); CK(cudaEventCreate(&h->ev2)); CK(cudaEventCreate(&h->ev3));
  DevCtx &D = h->D;
  D.LX = lenx; D.LY = leny; D.LZ = lenz; D.ncell = (long long)lenx * leny * ((long long)lenz + 1);
  D.NCrystal = ncrystal; D.NCov = ncov;
  for (int i = 0; i < 3; i++) { D.ListCrystal[i] = (i < ncrystal && listcrystal) ? listcrystal[i] : 0; D.ListCov[i] = (i < ncov && listcov) ? listcov[i] : 0; }
  CK(cudaMalloc(&D.state, D.ncell * sizeof(uint32_t)));
  CK(cudaMalloc(&D.bonds, D.ncell * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(D.state, 0, D.ncell * sizeof(uint32_t), h->stream));
  CK(cudaMemsetAsync(D.bonds, 0, D.ncell * sizeof(unsigned long long), h->stream));
  CK(cudaMalloc(&h->d_PE, 12288 * sizeof(double))); CK(cudaMalloc(&h->d_PD, 9 * sizeof(double)));
  CK(cudaMalloc(&h->d_PAbs, 12 * sizeof(double))); CK(cudaMalloc(&h->d_PDes, 12 * sizeof(double)));
  CK(cudaMalloc(&h->d_err, 8 * sizeof(int))); CK(cudaMemsetAsync(h->d_err, 0, 8 * sizeof(int), h->stream));
  D.PE = h->d_PE; D.PD = h->d_PD; D.PAbs = h->d_PAbs; D.PDes = h->d_PDes; D.err = h->d_err;
  // FIND_NN.f:27,51-52: Pi is a REAL(4) literal; cos/sin from the host libm, as in the CPU reference
  const double Pi = (double)3.14159265358979323f;
  for (int irot = 1; irot <= 2; irot++) { D.rot_c[irot - 1] = cos((double)(2 * irot) * Pi / 3.0); D.rot_s[irot - 1] = sin((double)(2 * irot) * Pi / 3.0); }
  // arm-chair offsets of FIND_SN.f:66-70 for every bond code: NINT(8.6602254037f * unit(S - C)), S - C = -v(code),
  // by the reference's own IEEE sequence (divide by the norm, multiply, NINT); entries within 1e-6 of a
  // rounding tie (where NINT(sn + x) could differ from sn + NINT(x)) and the zero vector stay 0 = "use find_sn()"
  {
    std::vector<unsigned short> tab(32768, 0);
    const double ARM = (double)8.6602254037f;
    for (int code = 1; code < 32768; code++) {
      int vx = (code & 31) - 16, vy = ((code >> 5) & 31) - 16, vz = ((code >> 10) & 31) - 16;
      if (vx == -16 || vy == -16 || vz == -16 || (vx == 0 && vy == 0 && vz == 0)) continue;
      double ax[3] = {(double)-vx, (double)-vy, (double)-vz};
      double snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      int o[3]; bool ok = true;
      for (int k = 0; k < 3; k++) {
        ax[k] = ax[k] / snorm;
        double x = ARM * ax[k];
        double fr = fabs(x) - floor(fabs(x));
        if (fabs(fr - 0.5) < 1e-6) ok = false;
        o[k] = (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5));
      }
      if (ok) tab[code] = (unsigned short)((o[0] + 16) | ((o[1] + 16) << 5) | ((o[2] + 16) << 10));
    }
    CK(cudaMalloc(&h->d_arm_tab, tab.size() * sizeof(unsigned short)));
    CK(cudaMemcpyAsync(h->d_arm_tab, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    D.arm_tab = h->d_arm_tab;
  }
  // the neighbour-update rules of the four events as a table (mskc_coop.cuh: pa_rule)
  {
    std::vector<uint32_t> tab(PA_ENTRIES, 0u);
    for (int kind = 0; kind < 4; kind++) for (int occ = 0; occ < 2; occ++) for (int wall = 0; wall < 2; wall++) for (int rec = 0; rec < 2; rec++) for (int c = 0; c <= 5; c++)
      tab[pa_index(kind, occ, wall, rec, c)] = pa_rule(kind, occ, wall && occ, rec, c);
    CK(cudaMalloc(&h->d_pa_tab, tab.size() * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(h->d_pa_tab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    D.pa_tab = h->d_pa_tab;
  }
  // FIND_NN.f on the ideal bond shell: integer form -(b1+b2)/2 -+ (b1 x b2)/6 against the reference's double-precision
  // rotation (same IEEE sequence as the device routine) for every tetrahedral pair of the 32 |b|^2 = 27 vectors
  {
    auto nint = [](double x) { return (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); };
    std::vector<std::array<int, 3>> shell;
    for (int x = -5; x <= 5; x++) for (int y = -5; y <= 5; y++) for (int z = -5; z <= 5; z++) if (x * x + y * y + z * z == 27) shell.push_back({x, y, z});
    int ok = 1, npairs = 0;
    for (auto &b1 : shell) for (auto &b2 : shell) {
      int d = (b1[0] - b2[0]) * (b1[0] - b2[0]) + (b1[1] - b2[1]) * (b1[1] - b2[1]) + (b1[2] - b2[2]) * (b1[2] - b2[2]);
      if (d != 72) continue;
      npairs++;
      double ax[3] = {-(double)b1[0], -(double)b1[1], -(double)b1[2]};
      double snorm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      for (int k = 0; k < 3; k++) ax[k] = ax[k] / snorm;
      double v[3] = {(double)b2[0], (double)b2[1], (double)b2[2]};
      int c[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
      for (int irot = 0; irot < 2; irot++) {
        double Cth = D.rot_c[irot], Sth = D.rot_s[irot], Rot[3][3], R[3];
        for (int i = 0; i < 3; i++) Rot[i][i] = ax[i] * ax[i] * (1 - Cth) + Cth;
        Rot[1][0] = ax[0] * ax[1] * (1 - Cth) - ax[2] * Sth; Rot[0][1] = ax[0] * ax[1] * (1 - Cth) + ax[2] * Sth;
        Rot[2][0] = ax[0] * ax[2] * (1 - Cth) + ax[1] * Sth; Rot[0][2] = ax[0] * ax[2] * (1 - Cth) - ax[1] * Sth;
        Rot[2][1] = ax[1] * ax[2] * (1 - Cth) - ax[0] * Sth; Rot[1][2] = ax[1] * ax[2] * (1 - Cth) + ax[0] * Sth;
        for (int k = 0; k < 3; k++) { R[k] = 0.0; for (int l = 0; l < 3; l++) R[k] = R[k] + Rot[l][k] * v[l]; }
        const int sg = irot == 0 ? -1 : 1;
        for (int k = 0; k < 3; k++) {
          if ((b1[k] + b2[k]) % 2 || c[k] % 6) ok = 0;
          if (nint(R[k]) != -(b1[k] + b2[k]) / 2 + sg * (c[k] / 6)) ok = 0;
          if (fabs(fabs(R[k] - floor(R[k])) - 0.5) < 1e-6) ok = 0;
        }
      }
    }
    D.nn_exact = (ok && npairs == 120) ? 1 : 0;
  }
  // AllocateArrays.f:30-43
  h->NumMcMax = nummcmax > 0 ? nummcmax : 500000;
  h->Levels = 0;
  for (int ind = 0; ind <= 31; ind++) if ((h->NumMcMax >> ind) & 1) h->Levels = ind;
  h->Levels += (h->NumMcMax > (1 << h->Levels)) ? 2 : 1;
  CK(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

extern "C" int mskc_destroy(mskc_ctx *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->D.state); cudaFree(h->D.bonds);
  cudaFree(h->d_PE); cudaFree(h->d_PD); cudaFree(h->d_PAbs); cudaFree(h->d_PDes);
  cudaFree(h->d_stage); cudaFree(h->d_arm_tab); cudaFree(h->d_pa_tab); cudaFree(h->d_frame_cell); cudaFree(h->d_frame_word); cudaFree(h->d_frame_n);
  cudaFree(h->d_la_slot); cudaFree(h->d_la_E); cudaFree(h->d_la_D); cudaFree(h->d_la_A); cudaFree(h->d_la_S); cudaFree(h->d_la3_F); cudaFree(h->d_err);
  cudaFree(h->d_Tree); cudaFree(h->d_lattind); cudaFree(h->d_mc_cell); cudaFree(h->d_void_cell); cudaFree(h->d_mc_ev); cudaFree(h->d_ps);
#ifndef MSKC_EMU
  if (h->PC.p2p_left) cudaIpcCloseMemHandle(h->PC.p2p_left);
  if (h->PC.p2p_right && h->PC.p2p_right != h->PC.p2p_left) cudaIpcCloseMemHandle(h->PC.p2p_right);
#endif
  cudaFree(h->PC.p2p_recv); h->PC.p2p_recv = nullptr; h->PC.p2p_left = h->PC.p2p_right = nullptr;
  if (h->PC.comm) { std::string e; if (NcclApi *nccl = nccl_api(&e)) nccl->CommDestroy((ncclComm_t)h->PC.comm); h->PC.comm = nullptr; }
  parallel_free(h->PC);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" int mskc_set_rates(mskc_ctx *h, const double *PtransE, const double *PtransD, const double *PtransAbs, const double *PtransDes, double PtransZig) {
  if (!h || !PtransE || !PtransD) return set_err(h, "mskc_set_rates: NULL argument");
  CK(cudaSetDevice(h->device));
  h->h_PE.assign(PtransE, PtransE + 12288); h->h_PD.assign(PtransD, PtransD + 9);
  h->h_PAbs.assign(12, 0.0); h->h_PDes.assign(12, 0.0);
  if (PtransAbs) h->h_PAbs.assign(PtransAbs, PtransAbs + 12);
  if (PtransDes) h->h_PDes.assign(PtransDes, PtransDes + 12);
  CK(cudaMemcpyAsync(h->d_PE, h->h_PE.data(), 12288 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_PD, h->h_PD.data(), 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_PAbs, h->h_PAbs.data(), 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_PDes, h->h_PDes.data(), 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->D.PZig = PtransZig;
  for (int coor = 1; coor <= 3; coor++) {            // same association as total_dep_abs(): (PD1+PD2+PD3) + (PAbs1+PAbs2+PAbs3)
    const double *PD = h->h_PD.data(), *PA = h->h_PAbs.data();
    double PDep = PD[0 + 3 * (coor - 1)] + PD[1 + 3 * (coor - 1)] + PD[2 + 3 * (coor - 1)];
    double PAbs = 0; for (int k = 1; k <= 3; k++) PAbs += PA[(k - 1) + 3 * coor];
    h->D.PDepAbs[coor] = PDep + PAbs;
  }
  h->D.PDepAbs[0] = 0.0;
  for (int i = 0; i < 9; i++) h->D.cPD[i] = h->h_PD[i];
  h->have_rates = true;
  return 0;
}

extern "C" int mskc_set_la(mskc_ctx *h, int enable, double Tm, double P0, double sigma, double Tflex, int tmin, int tmax) {
  if (!h) return set_err(h, "mskc_set_la: NULL handle");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  if (!enable) { D.la = 0; D.keepT = 0; return 0; }
  if (!h->have_rates) return set_err(h, "mskc_set_la: call mskc_set_rates first");
  if (tmin < 1 || tmax > 8191 || tmax < tmin) return set_err(h, "mskc_set_la: temperature range must lie in 1..8191 K (13-bit field)");
  const double kb = 0.00008617330350;               // Get_Prob.f:70
  std::vector<short> slot(12288, -1);
  int nnz = 0;
  for (int i = 0; i < 12288; i++) if (h->h_PE[i] != 0.0) slot[i] = (short)nnz++;
  if (nnz == 0) nnz = 1;
  int nt = tmax - tmin + 1;
  std::vector<double> E((size_t)nt * nnz, 0.0), Dt((size_t)nt * 9, 0.0);
  for (int t = tmin; t <= tmax; t++) {
    int temperature = t;
    for (int i = 0; i < 12288; i++) if (slot[i] >= 0)   // Get_Prob.f:143-144
      E[(size_t)(t - tmin) * nnz + slot[i]] = P0 * exp(-h->h_PE[i] / (kb * temperature)) * 1.0;
    for (int i = 0; i < 9; i++) {                       // Get_Prob.f:146-148
      double d = h->h_PD[i];
      Dt[(size_t)(t - tmin) * 9 + i] = 0.5 * P0 * exp(-d / (kb * Tm)) * (1.0 + erf((temperature - Tflex) / sigma)) * (d != 0.0 ? 1.0 : 0.0);
    }
  }
  cudaFree(h->d_la_slot); cudaFree(h->d_la_E); cudaFree(h->d_la_D);
  h->d_la_slot = nullptr; h->d_la_E = h->d_la_D = nullptr;
  CK(cudaMalloc(&h->d_la_slot, 12288 * sizeof(short)));
  CK(cudaMalloc(&h->d_la_E, E.size() * sizeof(double)));
  CK(cudaMalloc(&h->d_la_D, Dt.size() * sizeof(double)));
  CK(cudaMemcpy(h->d_la_slot, slot.data(), 12288 * sizeof(short), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_E, E.data(), E.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_D, Dt.data(), Dt.size() * sizeof(double), cudaMemcpyHostToDevice));
  D.la = 1; D.keepT = 1; D.la_tmin = tmin; D.la_tmax = tmax; D.la_nnz = nnz;
  D.la_slot = h->d_la_slot; D.la_E = h->d_la_E; D.la_D = h->d_la_D;
  return 0;
}

extern "C" int mskc_set_tree_variant(mskc_ctx *h, int variant) {
  if (!h) return set_err(h, "mskc_set_tree_variant: NULL handle");
  if (variant != MSKC_TREE_SRC && variant != MSKC_TREE_LA) return set_err(h, "mskc_set_tree_variant: unknown variant %d", variant);
  if (variant == MSKC_TREE_LA && h->D.NCrystal > 2) return set_err(h, "mskc_set_tree_variant: the LaserAnnealing fork has two species bits (NCrystal <= 2)");
  h->tree_variant = variant;
  return 0;
}

extern "C" int mskc_set_la_fork(mskc_ctx *h, int enable, const double *Tm, const double *P0, const double *sigma, const double *Tflex,
                                int tmin, int tmax, const int32_t *InitCountCrystal, int num_occ0, int num_adatom_occ) {
  if (!h) return set_err(h, "mskc_set_la_fork: NULL handle");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  if (!enable) { D.la = 0; D.keepT = 0; return 0; }
  if (!Tm || !P0 || !sigma || !Tflex || !InitCountCrystal) return set_err(h, "mskc_set_la_fork: NULL argument");
  if (!h->have_rates) return set_err(h, "mskc_set_la_fork: call mskc_set_rates first");
  if (D.NCrystal > 2) return set_err(h, "mskc_set_la_fork: the LaserAnnealing fork knows two crystal species");
  if (tmin < 1 || tmax > 8191 || tmax < tmin) return set_err(h, "mskc_set_la_fork: temperature range must lie in 1..8191 K (13-bit field)");
  const double kb = 0.00008617330350;               // la/mcevent/Get_Prob.f:70
  std::vector<short> slot(12288, -1);
  int nnz = 0;
  for (int i = 0; i < 12288; i++) if (h->h_PE[i] != 0.0) slot[i] = (short)nnz++;
  if (nnz == 0) nnz = 1;
  const int nt = tmax - tmin + 1;
  std::vector<double> E((size_t)nt * nnz, 0.0), F((size_t)nt * 3, 0.0);
  for (int t = tmin; t <= tmax; t++) {
    const int temperature = t;
    for (int i = 0; i < 12288; i++) if (slot[i] >= 0) {   // :144-146  PE(k,...) = P0(k) * EXP(-PtransE(k,...)/(kb*temperature)) * merge
      const int k = i % 3;
      E[(size_t)(t - tmin) * nnz + slot[i]] = k < D.NCrystal ? P0[k] * exp(-h->h_PE[i] / (kb * temperature)) * 1.0 : 0.0;
    }
    for (int k = 0; k < D.NCrystal; k++) F[(size_t)(t - tmin) * 3 + k] = 1.0 + erf((temperature - Tflex[k]) / sigma[k]);   // :150
  }
  for (int k = 0; k < 3; k++) {
    D.la3_halfP0[k] = k < D.NCrystal ? 0.5 * P0[k] : 0.0;
    D.la3_init[k] = InitCountCrystal[k];
    for (int c = 0; c < 3; c++) {
      const double d = h->h_PD[k + 3 * c];
      D.la3_expD[k + 3 * c] = k < D.NCrystal ? exp(-d / (kb * Tm[k])) : 0.0;     // :149
      D.la3_merge[k + 3 * c] = d != 0.0 ? 1.0 : 0.0;                              // :151
    }
  }
  D.la3_swap = (D.NCrystal == 2 && D.ListCrystal[0] == 14 && D.ListCrystal[1] == 32) ? 1 : 0;    // la/main.f:675-676
  cudaFree(h->d_la_slot); cudaFree(h->d_la_E); cudaFree(h->d_la3_F);
  h->d_la_slot = nullptr; h->d_la_E = nullptr; h->d_la3_F = nullptr;
  CK(cudaMalloc(&h->d_la_slot, 12288 * sizeof(short)));
  CK(cudaMalloc(&h->d_la_E, E.size() * sizeof(double)));
  CK(cudaMalloc(&h->d_la3_F, F.size() * sizeof(double)));
  CK(cudaMemcpy(h->d_la_slot, slot.data(), 12288 * sizeof(short), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_E, E.data(), E.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la3_F, F.data(), F.size() * sizeof(double), cudaMemcpyHostToDevice));
  D.la = 3; D.keepT = 1; D.la_tmin = tmin; D.la_tmax = tmax; D.la_nnz = nnz;
  D.la_slot = h->d_la_slot; D.la_E = h->d_la_E; D.la3_F = h->d_la3_F;
  h->fork_occ = num_adatom_occ; h->fork_occ0 = num_occ0;
  if (h->parity_ready) {                               // state already uploaded: the two counters go into the loop state
    h->h_ps.NumAdAtomOcc = num_adatom_occ; h->h_ps.NumOcc0 = num_occ0;
    static_assert(offsetof(ParityState, NumOcc0) == offsetof(ParityState, NumAdAtomOcc) + sizeof(int), "the two counters are copied together");
    CK(cudaMemcpy(reinterpret_cast<char *>(h->d_ps) + offsetof(ParityState, NumAdAtomOcc), &h->h_ps.NumAdAtomOcc, 2 * sizeof(int), cudaMemcpyHostToDevice));
  }
  return 0;
}

// la = 3: CountLiquid for the kernels that do not walk the serial loop (sampler, re-rate, trace replay): as of now
static int push_fork_CL(mskc_ctx *h) {
  if (h->D.la != 3) return 0;
  mskc_counters c;
  if (mskc_get_counters(h, &c)) return -1;
  int cl[4];
  for (int k = 0; k < 3; k++) cl[k] = h->D.la3_init[k] - c.CountCrystal[k];
  cl[3] = cl[0] + cl[1] + cl[2];
  CK(cudaMemcpyToSymbol(g_la_CL, cl, sizeof cl));
  return 0;
}

// mcevent/InterpProbFromField.f:22-57 for one table entry: the two tabulated blocks that bracket `temperature`,
// mixed with INTEGER distance weights.  Returns false where the reference would index its arrays out of bounds
// (temperature outside the tabulated range).
static bool st_bracket(int ntemp, const int32_t *reftemp, int temperature, int *ilower, int *iupper, int *wlower, int *wupper) {
  for (int i = 0; i < ntemp; i++) if (reftemp[i] - temperature == 0) { *ilower = *iupper = i; *wlower = 1; *wupper = 0; return true; }   // FINDLOC(diff, 0)
  int up = -1, best = 0;
  for (int i = 0; i < ntemp; i++) { int d = reftemp[i] - temperature; if (d > 0 && (up < 0 || d < best)) { up = i; best = d; } }   // MINLOC(diff, MASK=diff>0)
  if (up <= 0) return false;
  *iupper = up; *ilower = up - 1;
  *wlower = abs(reftemp[up] - temperature); *wupper = abs(reftemp[up - 1] - temperature);
  return true;
}
static inline double st_mix(int il, int iu, int wl, int wu, double lo, double up) {
  if (il == iu) return lo;                                                 // exact hit: the block is copied
  return ((double)wl * lo + (double)wu * up) / (double)(wl + wu);
}

extern "C" int mskc_set_rate_field(mskc_ctx *h, int ntemp, const int32_t *reftemp, const double *refPE, const double *refPD,
                                   const double *refPAbs, const double *refPDes, double PtransZig, int tmin, int tmax, int fixed_temperature) {
  if (!h || !reftemp || !refPE || !refPD) return set_err(h, "mskc_set_rate_field: NULL argument");
  if (ntemp < 1 || ntemp > 300) return set_err(h, "mskc_set_rate_field: ntemp must be 1..300 (ntempMax, defsystem.f:72)");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  std::vector<double> zero12((size_t)ntemp * 12, 0.0);
  if (!refPAbs) refPAbs = zero12.data();
  if (!refPDes) refPDes = zero12.data();
  int il, iu, wl, wu;
  if (fixed_temperature > 0) {
    // SetTrenchSi.f:65-67: Ttop == Tbottom -> interpolate once, plain tables from then on; Evaporation.f:70 still
    // keeps the temperature bits of a cell it clears
    if (!st_bracket(ntemp, reftemp, fixed_temperature, &il, &iu, &wl, &wu))
      return set_err(h, "mskc_set_rate_field: temperature %d K outside the tabulated range %d..%d K", fixed_temperature, reftemp[0], reftemp[ntemp - 1]);
    std::vector<double> PE(12288), PD(9), PA(12), PS(12);
    for (int i = 0; i < 12288; i++) PE[i] = st_mix(il, iu, wl, wu, refPE[(size_t)il * 12288 + i], refPE[(size_t)iu * 12288 + i]);
    for (int i = 0; i < 9; i++) PD[i] = st_mix(il, iu, wl, wu, refPD[(size_t)il * 9 + i], refPD[(size_t)iu * 9 + i]);
    for (int i = 0; i < 12; i++) PA[i] = st_mix(il, iu, wl, wu, refPAbs[(size_t)il * 12 + i], refPAbs[(size_t)iu * 12 + i]);
    for (int i = 0; i < 12; i++) PS[i] = st_mix(il, iu, wl, wu, refPDes[(size_t)il * 12 + i], refPDes[(size_t)iu * 12 + i]);
    int rc = mskc_set_rates(h, PE.data(), PD.data(), PA.data(), PS.data(), PtransZig);
    if (rc) return rc;
    D.la = 0; D.keepT = 1;
    return 0;
  }
  if (tmin < 1 || tmax > 8191 || tmax < tmin) return set_err(h, "mskc_set_rate_field: temperature range must lie in 1..8191 K (13-bit field)");
  if (!st_bracket(ntemp, reftemp, tmin, &il, &iu, &wl, &wu) || !st_bracket(ntemp, reftemp, tmax, &il, &iu, &wl, &wu))
    return set_err(h, "mskc_set_rate_field: temperature range %d..%d K outside the tabulated range %d..%d K", tmin, tmax, reftemp[0], reftemp[ntemp - 1]);
  // the plain tables stay defined (block 1) so that mskc_set_la / diagnostics have something to read; the kernels use the per-kelvin ones
  { int rc = mskc_set_rates(h, refPE, refPD, refPAbs, refPDes, PtransZig); if (rc) return rc; }
  std::vector<short> slot(12288, -1);
  int nnz = 0;
  for (int i = 0; i < 12288; i++) for (int t = 0; t < ntemp; t++) if (refPE[(size_t)t * 12288 + i] != 0.0) { slot[i] = (short)nnz++; break; }
  if (nnz == 0) nnz = 1;
  const int nt = tmax - tmin + 1;
  std::vector<double> E((size_t)nt * nnz, 0.0), Dt((size_t)nt * 9, 0.0), A((size_t)nt * 12, 0.0), S((size_t)nt * 12, 0.0);
  for (int t = tmin; t <= tmax; t++) {
    if (!st_bracket(ntemp, reftemp, t, &il, &iu, &wl, &wu)) return set_err(h, "mskc_set_rate_field: reftemp must increase monotonically");
    const size_t r = (size_t)(t - tmin);
    for (int i = 0; i < 12288; i++) if (slot[i] >= 0) E[r * nnz + slot[i]] = st_mix(il, iu, wl, wu, refPE[(size_t)il * 12288 + i], refPE[(size_t)iu * 12288 + i]);
    for (int i = 0; i < 9; i++) Dt[r * 9 + i] = st_mix(il, iu, wl, wu, refPD[(size_t)il * 9 + i], refPD[(size_t)iu * 9 + i]);
    for (int i = 0; i < 12; i++) A[r * 12 + i] = st_mix(il, iu, wl, wu, refPAbs[(size_t)il * 12 + i], refPAbs[(size_t)iu * 12 + i]);
    for (int i = 0; i < 12; i++) S[r * 12 + i] = st_mix(il, iu, wl, wu, refPDes[(size_t)il * 12 + i], refPDes[(size_t)iu * 12 + i]);
  }
  cudaFree(h->d_la_slot); cudaFree(h->d_la_E); cudaFree(h->d_la_D); cudaFree(h->d_la_A); cudaFree(h->d_la_S);
  h->d_la_slot = nullptr; h->d_la_E = h->d_la_D = h->d_la_A = h->d_la_S = nullptr;
  CK(cudaMalloc(&h->d_la_slot, 12288 * sizeof(short)));
  CK(cudaMalloc(&h->d_la_E, E.size() * sizeof(double)));
  CK(cudaMalloc(&h->d_la_D, Dt.size() * sizeof(double)));
  CK(cudaMalloc(&h->d_la_A, A.size() * sizeof(double)));
  CK(cudaMalloc(&h->d_la_S, S.size() * sizeof(double)));
  CK(cudaMemcpy(h->d_la_slot, slot.data(), 12288 * sizeof(short), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_E, E.data(), E.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_D, Dt.data(), Dt.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_A, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_la_S, S.data(), S.size() * sizeof(double), cudaMemcpyHostToDevice));
  D.la = 2; D.keepT = 1; D.la_tmin = tmin; D.la_tmax = tmax; D.la_nnz = nnz;
  D.la_slot = h->d_la_slot; D.la_E = h->d_la_E; D.la_D = h->d_la_D; D.la_A = h->d_la_A; D.la_S = h->d_la_S;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// state upload
static inline uint64_t h_enc_bonds(const DevCtx &D, const int32_t s[3], const int32_t nn[12], bool *far) {
  uint64_t r = 0;
  for (int i = 0; i < 4; i++) {
    const int32_t *n = nn + 3 * i;
    unsigned raw = 0;
    if (!(n[0] == 0 && n[1] == 0 && n[2] == 0)) {
      int dx = n[0] - s[0], dy = n[1] - s[1], dz = n[2] - s[2];
      if (dx > D.LX / 2) dx -= D.LX; else if (dx < -(D.LX / 2)) dx += D.LX;
      if (dy > D.LY / 2) dy -= D.LY; else if (dy < -(D.LY / 2)) dy += D.LY;
      if (dx < -15 || dx > 15 || dy < -15 || dy > 15 || dz < -15 || dz > 15) { *far = true; dx = dy = dz = 0; }
      raw = (unsigned)(dx + 16) | ((unsigned)(dy + 16) << 5) | ((unsigned)(dz + 16) << 10);
    }
    r |= (uint64_t)raw << (16 * i);
  }
  return r;
}
static inline void h_dec_bonds(const DevCtx &D, const int32_t s[3], uint64_t b, int32_t nn[12]) {
  for (int i = 0; i < 4; i++) {
    unsigned raw = (unsigned)((b >> (16 * i)) & 0xFFFFu);
    int32_t *n = nn + 3 * i;
    if (raw == 0) { n[0] = n[1] = n[2] = 0; continue; }
    n[0] = s[0] + (int)(raw & 31u) - 16; n[1] = s[1] + (int)((raw >> 5) & 31u) - 16; n[2] = s[2] + (int)((raw >> 10) & 31u) - 16;
    if (n[0] >= D.LX) n[0] -= D.LX; else if (n[0] < 0) n[0] += D.LX;
    if (n[1] >= D.LY) n[1] -= D.LY; else if (n[1] < 0) n[1] += D.LY;
  }
}

static int ensure_parity_buffers(mskc_ctx *h) {
  if (h->d_Tree) return 0;
  size_t tsz = (size_t)1 << h->Levels;
  CK(cudaMalloc(&h->d_Tree, tsz * sizeof(double)));
  CK(cudaMalloc(&h->d_lattind, h->D.ncell * sizeof(uint32_t)));
  CK(cudaMalloc(&h->d_mc_cell, ((size_t)h->NumMcMax + 2) * sizeof(uint32_t)));
  CK(cudaMalloc(&h->d_mc_ev, ((size_t)h->NumMcMax + 2) * sizeof(int)));
  CK(cudaMalloc(&h->d_void_cell, ((size_t)h->NumVdMax + 2) * sizeof(uint32_t)));
  CK(cudaMalloc(&h->d_ps, sizeof(ParityState)));
  return 0;
}

extern "C" int mskc_upload_state(mskc_ctx *h, const int32_t *LattCoo,
                                 int n_mc, const int32_t *mc_xyz, const int32_t *mc_nn, const int32_t *mc_event, const double *mc_rate,
                                 int n_bulk, const int32_t *bulk_xyz, const int32_t *bulk_nn,
                                 int n_void, const int32_t *void_xyz, const int32_t *void_nn) {
  if (!h || !LattCoo) return set_err(h, "mskc_upload_state: NULL argument");
  if (n_mc >= h->NumMcMax) return set_err(h, "troppe particelle cinetiche (%d >= %d)", n_mc, h->NumMcMax);
  if (n_void >= h->NumVdMax) return set_err(h, "troppe vacanze");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  if (ensure_parity_buffers(h)) return -1;
  std::vector<uint32_t> conv; std::vector<int32_t> conv_ev;
  if (h->tree_variant == MSKC_TREE_LA) {               // the caller speaks the LaserAnnealing fork's layout
    conv.resize((size_t)D.ncell);
    for (long long i = 0; i < D.ncell; i++) conv[(size_t)i] = word_to_dev((uint32_t)LattCoo[i]);
    LattCoo = (const int32_t *)conv.data();
    conv_ev.resize((size_t)n_mc + 1);
    for (int i = 0; i < n_mc; i++) conv_ev[i] = ev_to_dev(mc_event[i]);
    mc_event = conv_ev.data();
  }
  CK(cudaMemcpyAsync(D.state, LattCoo, D.ncell * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  // bonds + lattind are assembled on the host (sparse) and scattered on the device
  size_t nrec = (size_t)n_mc + n_bulk + n_void;
  std::vector<uint32_t> rc(nrec ? nrec : 1), rind(nrec ? nrec : 1); std::vector<uint64_t> rb(nrec ? nrec : 1);
  bool far = false; size_t k = 0;
  auto cell_of = [&](const int32_t *p) { return (uint32_t)((long long)p[0] + (long long)D.LX * ((long long)p[1] + (long long)D.LY * (long long)p[2])); };
  for (int i = 0; i < n_bulk; i++, k++) { rc[k] = cell_of(bulk_xyz + 3 * i); rb[k] = h_enc_bonds(D, bulk_xyz + 3 * i, bulk_nn + 12 * i, &far); rind[k] = 0; }
  for (int i = 0; i < n_void; i++, k++) { rc[k] = cell_of(void_xyz + 3 * i); rb[k] = h_enc_bonds(D, void_xyz + 3 * i, void_nn + 12 * i, &far); rind[k] = (uint32_t)(i + 1); }
  for (int i = 0; i < n_mc; i++, k++) { rc[k] = cell_of(mc_xyz + 3 * i); rb[k] = h_enc_bonds(D, mc_xyz + 3 * i, mc_nn + 12 * i, &far); rind[k] = (uint32_t)(i + 1); }
  if (far) return set_err(h, "mskc_upload_state: %s", dev_err_text(E_FARBOND));
  uint32_t *d_rc = nullptr, *d_rind = nullptr; unsigned long long *d_rb = nullptr;
  CK(cudaMalloc(&d_rc, rc.size() * 4)); CK(cudaMalloc(&d_rind, rind.size() * 4)); CK(cudaMalloc(&d_rb, rb.size() * 8));
  CK(cudaMemcpyAsync(d_rc, rc.data(), rc.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_rind, rind.data(), rind.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_rb, rb.data(), rb.size() * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(D.bonds, 0, D.ncell * sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->d_lattind, 0, D.ncell * sizeof(uint32_t), h->stream));
  if (nrec) MSKC_LAUNCH(scatter_records, (unsigned)((nrec + 255) / 256), 256, 0, h->stream, D.bonds, h->d_lattind, d_rc, d_rb, d_rind, (long long)nrec);
  // MC list, void list, rate tree
  std::vector<uint32_t> mcc((size_t)h->NumMcMax + 2, 0xFFFFFFFFu), vdc((size_t)h->NumVdMax + 2, 0xFFFFFFFFu);
  std::vector<int> mce((size_t)h->NumMcMax + 2, 0);
  size_t tsz = (size_t)1 << h->Levels, leaf0 = tsz >> 1;
  std::vector<double> tree(tsz, 0.0);
  for (int i = 0; i < n_mc; i++) { mcc[i + 1] = cell_of(mc_xyz + 3 * i); mce[i + 1] = mc_event[i]; tree[leaf0 + i] = mc_rate[i]; }
  for (int i = 0; i < n_void; i++) vdc[i + 1] = cell_of(void_xyz + 3 * i);
  // The tree as n successive UpdateTree calls leave it: every inner node is left+right of its final
  // children (each node is last written with both children final), so one bottom-up pass is identical.
  for (size_t p = leaf0 - 1; p >= 1; p--) tree[p] = tree[2 * p] + tree[2 * p + 1];
  CK(cudaMemcpyAsync(h->d_Tree, tree.data(), tsz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_mc_cell, mcc.data(), mcc.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_mc_ev, mce.data(), mce.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_void_cell, vdc.data(), vdc.size() * 4, cudaMemcpyHostToDevice, h->stream));
  ParityState &S = h->h_ps;
  int idum = S.idum, runtype = S.runtype, iff = S.r_iff, inext = S.r_inext, inextp = S.r_inextp; long long counter = S.counter;
  int ma[56]; memcpy(ma, S.r_ma, sizeof ma);
  memset(&S, 0, sizeof S);
  S.idum = idum; S.runtype = runtype ? runtype : 'R'; S.r_iff = iff; S.r_inext = inext; S.r_inextp = inextp; S.counter = counter; memcpy(S.r_ma, ma, sizeof ma);
  S.Iter = 1; S.Time = 0.0; S.NumAdAtom = n_mc; S.NumAtoms = n_bulk; S.NumVoids = n_void;
  S.NumAdAtomOcc = h->fork_occ; S.NumOcc0 = h->fork_occ0;
  CK(cudaMemcpyAsync(h->d_ps, &S, sizeof S, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->d_err, 0, 8 * sizeof(int), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_rc); cudaFree(d_rind); cudaFree(d_rb);
  h->parity_ready = true; h->have_state = true;
  // the parallel sampler's tile lists are (re)built lazily from the uploaded MC list
  if (parallel_build_from_lists(h->PC, D, h->stream, n_mc, mcc.data() + 1, mce.data() + 1, mc_rate, n_bulk, n_void, &h->err)) { g_last_error = h->err; return -1; }
  if (!h->PC.lists_stale) {                            // census flags of every site (dense pass; mskc_coop.cuh)
    if (upload_consts(h)) return -1;
    unsigned long long o[3];
    if (parallel_census_flags(h->PC, h->D, h->stream, 0, o, nullptr, nullptr, nullptr, &h->err)) { g_last_error = h->err; return -1; }
  }
  h->parallel_ready = true;
  return check_dev_err(h);
}

extern "C" int mskc_seed(mskc_ctx *h, int idum, int runtype, int64_t counter0, const int32_t *ran3_state) {
  if (!h) return set_err(h, "mskc_seed: NULL handle");
  CK(cudaSetDevice(h->device));
  ParityState &S = h->h_ps;
  if (h->parity_ready) { CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost)); }
  S.idum = idum; S.runtype = runtype; S.counter = counter0;
  if (ran3_state) { S.r_iff = ran3_state[0]; S.r_inext = ran3_state[1]; S.r_inextp = ran3_state[2]; for (int i = 1; i <= 55; i++) S.r_ma[i] = ran3_state[2 + i]; }
  else { S.r_iff = 0; S.r_inext = 0; S.r_inextp = 0; memset(S.r_ma, 0, sizeof S.r_ma); }
  if (h->parity_ready) { CK(cudaMemcpy(h->d_ps, &S, sizeof S, cudaMemcpyHostToDevice)); }
  return 0;
}
extern "C" int mskc_get_rng(mskc_ctx *h, int32_t *ran3_state, int64_t *counter, int32_t *idum) {
  if (!h || !h->parity_ready) return set_err(h, "mskc_get_rng: no state");
  CK(cudaSetDevice(h->device));
  ParityState &S = h->h_ps;
  CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost));
  if (ran3_state) { ran3_state[0] = S.r_iff; ran3_state[1] = S.r_inext; ran3_state[2] = S.r_inextp; for (int i = 1; i <= 55; i++) ran3_state[2 + i] = S.r_ma[i]; }
  if (counter) *counter = S.counter;
  if (idum) *idum = S.idum;
  return 0;
}
extern "C" int mskc_set_clock(mskc_ctx *h, int64_t iter, double time, const int32_t *CountCrystal, const int32_t *CountCov) {
  if (!h || !h->parity_ready) return set_err(h, "mskc_set_clock: upload a state first");
  CK(cudaSetDevice(h->device));
  ParityState &S = h->h_ps;
  CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost));
  S.Iter = iter; S.Time = time;
  for (int i = 0; i < 3; i++) { if (CountCrystal) S.CountCrystal[i] = CountCrystal[i]; if (CountCov) S.CountCov[i] = CountCov[i]; }
  CK(cudaMemcpy(h->d_ps, &S, sizeof S, cudaMemcpyHostToDevice));
  h->PC.time = time;
  for (int i = 0; i < 3; i++) { if (CountCrystal) h->PC.baseCrystal[i] = CountCrystal[i]; if (CountCov) h->PC.baseCov[i] = CountCov[i]; }
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" int mskc_run_parity(mskc_ctx *h, const mskc_parity_args *a, mskc_parity_result *r) {
  if (!h || !a || !r) return set_err(h, "mskc_run_parity: NULL argument");
  if (!h->parity_ready || !h->have_rates) return set_err(h, "mskc_run_parity: rates and state must be uploaded first");
  if (h->PC.halo) return set_err(h, "mskc_run_parity: serial parity mode is single-GPU (not available on a slab context)");
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  ParityArgs A{};
  A.max_iter = a->max_iter; A.out_every = a->out_every; A.framecounter = a->framecounter; A.max_time = a->max_time; A.out_time = a->out_time;
  A.exit_on_time = a->exit_on_time;
  // the warp-cooperative loop (mskc_parity_coop.cuh) is the default; MSKC_PARITY_SERIAL=1 selects the one-thread kernel
  const bool serial = getenv("MSKC_PARITY_SERIAL") != nullptr;
  CK(cudaFuncSetAttribute(parity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ParShared)));
  CK(cudaFuncSetAttribute(parity_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PcShared)));
  // The kernel is launched in slices so that a single launch stays well below any watchdog.
  const long long slice = 2000000;
  long long remaining = a->event_budget > 0 ? a->event_budget : -1;
  memset(r, 0, sizeof *r);
  float total_ms = 0; int launches = 0;
  while (true) {
    A.event_budget = (remaining < 0 || remaining > slice) ? slice : remaining;
    CK(cudaEventRecord(h->ev0, h->stream));
    if (serial) MSKC_LAUNCH(parity_kernel, 1, 32, sizeof(ParShared), h->stream, A);
    else MSKC_LAUNCH(parity_coop_kernel, 1, 32, sizeof(PcShared), h->stream, A);
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h->h_ps, h->d_ps, sizeof(ParityState), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); total_ms += ms; launches++;
    ParityState &S = h->h_ps;
    r->events_done += S.events_done; r->site_updates += S.site_updates;
    A.framecounter = S.framecounter;
    if (remaining > 0) remaining -= S.events_done;
    if (S.stop_reason != MSKC_STOP_BUDGET || remaining == 0) break;
  }
  ParityState &S = h->h_ps;
  r->iter = S.Iter; r->iter_out = S.IterOut; r->time = S.Time;
  for (int i = 0; i < 3; i++) r->last_site[i] = S.last_site[i];
  r->stop_reason = S.stop_reason; r->num_atoms = S.NumAtoms; r->num_voids = S.NumVoids; r->num_adatom = S.NumAdAtom;
  r->kernel_ms = total_ms; r->launches = launches;
  CK(cudaMemcpy(&r->tree_root, h->d_Tree + 1, sizeof(double), cudaMemcpyDeviceToHost));
  h->PC.lists_stale = true;
  if (S.stop_reason == MSKC_STOP_ERROR) { check_dev_err(h); return -1; }
  return check_dev_err(h);
}

// ------------------------------------------------------------------------------------------------
extern "C" int mskc_download_lattice(mskc_ctx *h, int32_t *LattCoo) {
  if (!h || !LattCoo) return set_err(h, "mskc_download_lattice: NULL argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(LattCoo, h->D.state, h->D.ncell * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->PC.listed_flag_in_state) { for (long long i = 0; i < h->D.ncell; i++) LattCoo[i] &= ~(int32_t)(1u << 30); }
  if (h->tree_variant == MSKC_TREE_LA) for (long long i = 0; i < h->D.ncell; i++) LattCoo[i] = (int32_t)word_from_dev((uint32_t)LattCoo[i]);
  return 0;
}
extern "C" int mskc_patch_lattice(mskc_ctx *h, int n, const int64_t *cell, const int32_t *word) {
  if (!h) return set_err(h, "mskc_patch_lattice: NULL handle");
  if (n <= 0) return 0;
  if (!cell || !word) return set_err(h, "mskc_patch_lattice: NULL argument");
  CK(cudaSetDevice(h->device));
  // one upload + one scatter kernel (the frame writer's antisite pass patches a list of cells twice per frame); a cell that
  // occurs twice keeps its LAST word, as a loop of single stores would
  std::vector<uint32_t> c((size_t)n), w((size_t)n);
  for (int i = 0; i < n; i++) {
    if (cell[i] < 0 || cell[i] >= h->D.ncell) return set_err(h, "mskc_patch_lattice: cell out of range");
    c[(size_t)i] = (uint32_t)cell[i];
    w[(size_t)i] = h->tree_variant == MSKC_TREE_LA ? word_to_dev((uint32_t)word[i]) : (uint32_t)word[i];
  }
  {                                                           // duplicates: keep the last occurrence only
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int p, int q) { return c[(size_t)p] < c[(size_t)q]; });
    std::vector<uint32_t> c2, w2;
    for (int k = 0; k < n; k++) if (k + 1 == n || c[(size_t)order[(size_t)k]] != c[(size_t)order[(size_t)k + 1]]) { c2.push_back(c[(size_t)order[(size_t)k]]); w2.push_back(w[(size_t)order[(size_t)k]]); }
    c.swap(c2); w.swap(w2);
  }
  const size_t m = c.size();
  uint32_t *d_c = nullptr, *d_w = nullptr;
  CK(cudaMalloc(&d_c, m * 4)); CK(cudaMalloc(&d_w, m * 4));
  CK(cudaMemcpyAsync(d_c, c.data(), m * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_w, w.data(), m * 4, cudaMemcpyHostToDevice, h->stream));
  MSKC_LAUNCH(scatter_state, (unsigned)((m + 255) / 256), 256, 0, h->stream, h->D.state, (const uint32_t *)d_c, (const uint32_t *)d_w, (long long)m);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_c); cudaFree(d_w);
  return 0;
}
extern "C" int mskc_get_counters(mskc_ctx *h, mskc_counters *c) {
  if (!h || !c) return set_err(h, "mskc_get_counters: NULL argument");
  if (!h->parity_ready && !h->parallel_ready) return set_err(h, "mskc_get_counters: no state");
  CK(cudaSetDevice(h->device));
  memset(c, 0, sizeof *c);
  if (!h->parity_ready) { parallel_fill_counters(h->PC, c); return 0; }
  ParityState &S = h->h_ps;
  CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost));
  c->num_atoms = S.NumAtoms; c->num_voids = S.NumVoids; c->num_adatom = S.NumAdAtom; c->num_adatom_occ = S.NumAdAtomOcc;
  for (int i = 0; i < 3; i++) { c->CountCrystal[i] = S.CountCrystal[i]; c->CountCov[i] = S.CountCov[i]; }
  c->iter = S.Iter; c->time = S.Time;
  if (h->PC.ran || h->PC.replayed) {
    parallel_fill_counters(h->PC, c);
    if (h->D.la == 3) {       // the sampler does not carry NumAdAtomOcc: occupied ad-atoms = occupied, under-coordinated, non-wall cells
      mskc_census cs; if (mskc_census_lattice(h, &cs, nullptr)) return -1;
      c->num_adatom_occ = (int32_t)cs.undercoordinated;
    }
  }
  return 0;
}

extern "C" int mskc_download_lists(mskc_ctx *h, int32_t *mc_xyz, int32_t *mc_nn, int32_t *mc_event, double *mc_rate,
                                   int32_t *void_xyz, int32_t *void_nn, int32_t *bulk_xyz, int32_t *bulk_nn, int64_t bulk_cap, int64_t *n_bulk_found) {
  if (!h || !h->parity_ready) return set_err(h, "mskc_download_lists: no state");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  ParityState &S = h->h_ps;
  CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost));
  long long plane = (long long)D.LX * D.LY;
  auto xyz_of = [&](uint32_t c, int32_t *p) { p[2] = (int32_t)(c / plane); long long r = c - (long long)p[2] * plane; p[1] = (int32_t)(r / D.LX); p[0] = (int32_t)(r - (long long)p[1] * D.LX); };
  auto fetch_bonds = [&](const std::vector<uint32_t> &cells, std::vector<uint64_t> &out) -> int {
    out.assign(cells.size(), 0);
    if (cells.empty()) return 0;
    uint32_t *d_c = nullptr; unsigned long long *d_b = nullptr;
    CK(cudaMalloc(&d_c, cells.size() * 4)); CK(cudaMalloc(&d_b, cells.size() * 8));
    CK(cudaMemcpy(d_c, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice));
    MSKC_LAUNCH(gather_bonds, (unsigned)((cells.size() + 255) / 256), 256, 0, h->stream, D.bonds, d_c, d_b, (long long)cells.size());
    CK(cudaMemcpyAsync(out.data(), d_b, cells.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_c); cudaFree(d_b);
    return 0;
  };
  if (mc_xyz || mc_nn || mc_event || mc_rate) {
    int n = S.NumAdAtom;
    std::vector<uint32_t> cells(n); std::vector<int> evs(n); std::vector<double> rates(n);
    if (n) {
      CK(cudaMemcpy(cells.data(), h->d_mc_cell + 1, (size_t)n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(evs.data(), h->d_mc_ev + 1, (size_t)n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(rates.data(), h->d_Tree + ((size_t)1 << (h->Levels - 1)), (size_t)n * 8, cudaMemcpyDeviceToHost));
    }
    std::vector<uint64_t> b; if (fetch_bonds(cells, b)) return -1;
    for (int i = 0; i < n; i++) {
      int32_t p[3]; xyz_of(cells[i], p);
      if (mc_xyz) memcpy(mc_xyz + 3 * i, p, 12);
      if (mc_nn) h_dec_bonds(D, p, b[i], mc_nn + 12 * i);
      if (mc_event) mc_event[i] = h->tree_variant == MSKC_TREE_LA ? ev_from_dev(evs[i]) : evs[i];
      if (mc_rate) mc_rate[i] = rates[i];
    }
  }
  if (void_xyz || void_nn) {
    int n = S.NumVoids;
    std::vector<uint32_t> cells(n);
    if (n) CK(cudaMemcpy(cells.data(), h->d_void_cell + 1, (size_t)n * 4, cudaMemcpyDeviceToHost));
    std::vector<uint64_t> b; if (fetch_bonds(cells, b)) return -1;
    for (int i = 0; i < n; i++) {
      int32_t p[3]; xyz_of(cells[i], p);
      if (void_xyz) memcpy(void_xyz + 3 * i, p, 12);
      if (void_nn) h_dec_bonds(D, p, b[i], void_nn + 12 * i);
    }
  }
  if (bulk_xyz || bulk_nn || n_bulk_found) {
    // bulk atoms = occupied, non-wall, coordination 4 cells in z,y,x order (compacted on the device)
    long long cap = bulk_cap > 0 ? bulk_cap : 1;
    uint32_t *d_c = nullptr; unsigned long long *d_cnt = nullptr;
    CK(cudaMalloc(&d_c, (size_t)cap * 4)); CK(cudaMalloc(&d_cnt, 8)); CK(cudaMemset(d_cnt, 0, 8));
    if (upload_consts(h)) return -1;
    MSKC_LAUNCH(collect_bulk, 1184, 256, 0, h->stream, d_c, cap, d_cnt);
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_bulk_found) *n_bulk_found = (int64_t)cnt;
    long long n = (long long)cnt < cap ? (long long)cnt : cap;
    if ((bulk_xyz || bulk_nn) && n > 0) {
      std::vector<uint32_t> cells((size_t)n);
      CK(cudaMemcpy(cells.data(), d_c, (size_t)n * 4, cudaMemcpyDeviceToHost));
      std::sort(cells.begin(), cells.end());
      std::vector<uint64_t> b; if (fetch_bonds(cells, b)) return -1;
      for (long long i = 0; i < n; i++) {
        int32_t p[3]; xyz_of(cells[i], p);
        if (bulk_xyz) memcpy(bulk_xyz + 3 * i, p, 12);
        if (bulk_nn) h_dec_bonds(D, p, b[i], bulk_nn + 12 * i);
      }
    }
    cudaFree(d_c); cudaFree(d_cnt);
  }
  return 0;
}

extern "C" int mskc_state_hash(mskc_ctx *h, uint64_t *hash) {
  if (!h || !hash || !h->parity_ready) return set_err(h, "mskc_state_hash: no state");
  CK(cudaSetDevice(h->device));
  // FNV-1a is sequential by definition: hash on the host from a downloaded copy (test/diagnostic path)
  std::vector<int32_t> L((size_t)h->D.ncell);
  if (mskc_download_lattice(h, L.data())) return -1;
  ParityState &S = h->h_ps;
  CK(cudaMemcpy(&S, h->d_ps, sizeof S, cudaMemcpyDeviceToHost));
  uint64_t x = 1469598103934665603ull;
  for (size_t i = 0; i < L.size(); i++) { x ^= (uint32_t)L[i]; x *= 1099511628211ull; }
  uint64_t t; memcpy(&t, &S.Time, 8);
  uint64_t extra[5] = {(uint64_t)S.NumAdAtom, (uint64_t)S.NumAtoms, (uint64_t)S.NumVoids, t, (uint64_t)S.Iter};
  for (int i = 0; i < 5; i++) { x ^= extra[i]; x *= 1099511628211ull; }
  *hash = x;
  return 0;
}

extern "C" int mskc_set_temperature(mskc_ctx *h, const int32_t *T) {
  if (!h || !T) return set_err(h, "mskc_set_temperature: NULL argument");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  size_t n = (size_t)D.LX * D.LY * D.LZ;
  if (!h->d_stage) CK(cudaMalloc(&h->d_stage, n * 4));        // staging plane stack, kept for the next upload
  if (upload_consts(h)) return -1;
  CK(cudaMemcpyAsync(h->d_stage, T, n * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->ev2, h->stream));
  MSKC_LAUNCH(set_temperature_kernel<int32_t>, 148 * 8, 256, 0, h->stream, (const int32_t *)h->d_stage, (long long)n);
  CK(cudaEventRecord(h->ev3, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_dense_ms, h->ev2, h->ev3));
  return 0;
}
extern "C" int mskc_set_temperature_u16(mskc_ctx *h, const uint16_t *T) {
  if (!h || !T) return set_err(h, "mskc_set_temperature_u16: NULL argument");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  size_t n = (size_t)D.LX * D.LY * D.LZ;
  if (!h->d_stage) CK(cudaMalloc(&h->d_stage, n * 4));
  if (upload_consts(h)) return -1;
  CK(cudaMemcpyAsync(h->d_stage, T, n * 2, cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->ev2, h->stream));
  MSKC_LAUNCH(set_temperature_kernel<uint16_t>, 148 * 8, 256, 0, h->stream, (const uint16_t *)h->d_stage, (long long)n);
  CK(cudaEventRecord(h->ev3, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_dense_ms, h->ev2, h->ev3));
  return 0;
}
extern "C" int mskc_download_phases(mskc_ctx *h, uint8_t *phases) {
  if (!h || !phases) return set_err(h, "mskc_download_phases: NULL argument");
  if (!h->have_state) return set_err(h, "mskc_download_phases: no state");
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  size_t n = (size_t)D.LX * D.LY * D.LZ;
  if (!h->d_stage) CK(cudaMalloc(&h->d_stage, n * 4));
  if (upload_consts(h)) return -1;
  CK(cudaEventRecord(h->ev2, h->stream));
  MSKC_LAUNCH(phases_kernel, 148 * 8, 256, 0, h->stream, (unsigned char *)h->d_stage, (long long)n);
  CK(cudaEventRecord(h->ev3, h->stream));
  CK(cudaMemcpyAsync(phases, h->d_stage, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_dense_ms, h->ev2, h->ev3));
  return 0;
}
// device time of the last dense helper kernel (mskc_set_temperature*, mskc_download_phases), for the roofline in bench.py
extern "C" int mskc_last_dense_ms(mskc_ctx *h, float *ms) {
  if (!h || !ms) return set_err(h, "mskc_last_dense_ms: NULL argument");
  *ms = h->last_dense_ms;
  return 0;
}

extern "C" int mskc_census_lattice(mskc_ctx *h, mskc_census *c, float *kernel_ms) {
  if (!h || !c) return set_err(h, "mskc_census_lattice: NULL argument");
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  unsigned long long *d_out = nullptr; double *d_h = nullptr;
  CK(cudaMalloc(&d_out, 8 * 8)); CK(cudaMalloc(&d_h, 2 * 8));
  CK(cudaMemsetAsync(d_out, 0, 64, h->stream)); CK(cudaMemsetAsync(d_h, 0, 16, h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  MSKC_LAUNCH(census_kernel, 148 * 8, 256, 0, h->stream, d_out, d_h);
  CK(cudaEventRecord(h->ev1, h->stream));
  unsigned long long o[8]; double hh[2];
  CK(cudaMemcpyAsync(o, d_out, 64, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(hh, d_h, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, h->ev0, h->ev1));
  c->undercoordinated = (int64_t)o[0]; c->coverage = (int64_t)o[1]; c->noepi_bulk = (int64_t)o[2]; c->bulk = (int64_t)o[3]; c->occupied = (int64_t)o[4]; c->walls = (int64_t)o[5];
  c->mean_height = o[0] ? hh[0] / (double)o[0] : 0.0; c->sq_height = o[0] ? hh[1] / (double)o[0] : 0.0;
  cudaFree(d_out); cudaFree(d_h);
  return 0;
}

// ---- parallel sampler entry points (kernels in mskc_parallel.cuh) -----------------------------------
extern "C" int mskc_init_flat(mskc_ctx *h, int len1, int walls, int two_species) {
  if (!h) return set_err(h, "mskc_init_flat: NULL handle");
  if (!h->have_rates) return set_err(h, "mskc_init_flat: call mskc_set_rates first");
  CK(cudaSetDevice(h->device));
  if (parallel_init_flat(h->PC, h->D, h->stream, len1, walls, two_species, [h]() { return upload_consts(h); }, &h->err)) { g_last_error = h->err; return -1; }
  h->parallel_ready = true; h->have_state = true;
  return check_dev_err(h);
}
extern "C" int mskc_run_parallel(mskc_ctx *h, const mskc_parallel_args *a, mskc_parallel_result *r) {
  if (!h || !a || !r) return set_err(h, "mskc_run_parallel: NULL argument");
  if (!h->parallel_ready || !h->have_rates) return set_err(h, "mskc_run_parallel: rates and state must be set first");
  CK(cudaSetDevice(h->device));
  if (h->PC.lists_stale) return set_err(h, "mskc_run_parallel: the state was advanced in parity mode; upload it again to rebuild the tile lists");
  if (upload_consts(h)) return -1;
  if (push_fork_CL(h)) return -1;
  if (parallel_run(h->PC, h->D, h->stream, h->ev0, h->ev1, a, r, &h->err)) { g_last_error = h->err; return -1; }
  return check_dev_err(h);
}
extern "C" int mskc_rerate_all(mskc_ctx *h, int64_t *n_sites, double *total_rate, float *kernel_ms) {
  if (!h || !h->parallel_ready) return set_err(h, "mskc_rerate_all: no parallel state");
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  if (push_fork_CL(h)) return -1;
  if (parallel_rerate(h->PC, h->D, h->stream, h->ev0, h->ev1, n_sites, total_rate, kernel_ms, &h->err)) { g_last_error = h->err; return -1; }
  return check_dev_err(h);
}

// diagnostic / test entry point: recompute the census flags of every cell with bonds and compare them with the stored
// (incrementally maintained) ones; also counts cells whose coordination differs from their census
extern "C" int mskc_check_flags(mskc_ctx *h, int64_t *n_sites, int64_t *n_flag_mismatch, int64_t *n_coor_mismatch, float *kernel_ms) {
  if (!h || !h->parallel_ready || h->PC.lists_stale) return set_err(h, "mskc_check_flags: no parallel state");
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  unsigned long long o[3] = {0, 0, 0};
  if (parallel_census_flags(h->PC, h->D, h->stream, 1, o, kernel_ms, h->ev0, h->ev1, &h->err)) { g_last_error = h->err; return -1; }
  if (n_sites) *n_sites = (int64_t)o[0];
  if (n_flag_mismatch) *n_flag_mismatch = (int64_t)o[1];
  if (n_coor_mismatch) *n_coor_mismatch = (int64_t)o[2];
  return check_dev_err(h);
}

// diagnostics: raw per-tile counters (TileStat, 80 B each: i64 events, i64 site_updates, i32 dep[3], evap, abs, des, dAtoms, dVoids,
// dCrystal[3], dCov[3], maxev, pad), list id order (colour-major)
extern "C" int mskc_debug_tile_stats(mskc_ctx *h, void *out, int64_t max_tiles, int64_t *ntiles) {
  if (!h || !h->parallel_ready) return set_err(h, "mskc_debug_tile_stats: no parallel state");
  CK(cudaSetDevice(h->device));
  if (ntiles) *ntiles = h->PC.ntiles;
  int64_t n = h->PC.ntiles < max_tiles ? h->PC.ntiles : max_tiles;
  if (out && n > 0) CK(cudaMemcpy(out, h->PC.tstat, (size_t)n * sizeof(TileStat), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- test/diagnostic entry points of the cooperative transition code --------------------------------------
extern "C" int mskc_replay_events(mskc_ctx *h, int64_t n, const int32_t *site_xyz, const int32_t *event, const int32_t *arm, double tree_root) {
  if (!h || !site_xyz || !event || !arm) return set_err(h, "mskc_replay_events: NULL argument");
  if (!h->parallel_ready || !h->have_rates || h->PC.lists_stale) return set_err(h, "mskc_replay_events: rates and a parallel state must be set first");
  if (n <= 0) return 0;
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  int32_t *d_s = nullptr, *d_e = nullptr, *d_a = nullptr; TileStat *d_out = nullptr; TileStat out{};
  CK(cudaMalloc(&d_s, (size_t)n * 12)); CK(cudaMalloc(&d_e, (size_t)n * 4)); CK(cudaMalloc(&d_a, (size_t)n * 4)); CK(cudaMalloc(&d_out, sizeof(TileStat)));
  CK(cudaMemcpyAsync(d_s, site_xyz, (size_t)n * 12, cudaMemcpyHostToDevice, h->stream));
  std::vector<int32_t> conv_ev;
  if (h->tree_variant == MSKC_TREE_LA) { conv_ev.resize((size_t)n); for (int64_t i = 0; i < n; i++) conv_ev[(size_t)i] = ev_to_dev(event[i]); event = conv_ev.data(); }
  if (push_fork_CL(h)) return -1;
  CK(cudaMemcpyAsync(d_e, event, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_a, arm, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(d_out, 0, sizeof(TileStat), h->stream));
  const int64_t chunk = 1000;                    // tombstones are dropped between chunks (no tile is ever "running" here)
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t m = n - i0 < chunk ? n - i0 : chunk;
    MSKC_LAUNCH(compact_lists_kernel, (h->PC.ntiles * 32 + 127) / 128, 128, 0, h->stream, h->PC.ntiles);
    MSKC_LAUNCH(replay_kernel, 1, 32, 0, h->stream, d_s + 3 * i0, d_e + i0, d_a + i0, (long long)m, tree_root, d_out);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&out, d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));               // (also keeps conv_ev alive until the copy is done)
  cudaFree(d_s); cudaFree(d_e); cudaFree(d_a); cudaFree(d_out);
  ParallelCtx &P = h->PC;
  P.replayed = true;
  P.extra.events += out.events; P.extra.site_updates += out.site_updates; P.extra.dAtoms += out.dAtoms; P.extra.dVoids += out.dVoids;
  for (int i = 0; i < 3; i++) { P.extra.dCrystal[i] += out.dCrystal[i]; P.extra.dCov[i] += out.dCov[i]; }
  if (check_dev_err(h)) return -1;
  if (out.events != n) return set_err(h, "mskc_replay_events: stopped after %lld of %lld events", (long long)out.events, (long long)n);
  return 0;
}

// The same trace through coop_event<8, FEAT>, four teams in lock step (the instantiation mskc_run_parallel launches).
// Consecutive events whose sites are more than 48 units apart in x or y (minimum image) touch disjoint cells
// (mskc_parallel.cuh: writes reach 8, reads 16 units), so up to four of them share a lock-step iteration.
extern "C" int mskc_replay_teams(mskc_ctx *h, int64_t n, const int32_t *site_xyz, const int32_t *event, const int32_t *arm, double tree_root, int64_t *n_iterations, int64_t *n_shared) {
  if (!h || !site_xyz || !event || !arm) return set_err(h, "mskc_replay_teams: NULL argument");
  if (!h->parallel_ready || !h->have_rates || h->PC.lists_stale) return set_err(h, "mskc_replay_teams: rates and a parallel state must be set first");
  if (h->D.la == 3) return set_err(h, "mskc_replay_teams: the LaserAnnealing fork's per-event liquid fraction needs the serial replay (mskc_replay_events)");
  if (n <= 0) return 0;
  CK(cudaSetDevice(h->device));
  if (upload_consts(h)) return -1;
  const DevCtx &D = h->D;
  std::vector<int32_t> conv_ev;
  if (h->tree_variant == MSKC_TREE_LA) { conv_ev.resize((size_t)n); for (int64_t i = 0; i < n; i++) conv_ev[(size_t)i] = ev_to_dev(event[i]); event = conv_ev.data(); }
  // greedy batching of consecutive commuting events; the first event of iteration b goes to team b % 4
  auto far = [&](int64_t a, int64_t b) {
    int dx = abs(site_xyz[3 * a] - site_xyz[3 * b]), dy = abs(site_xyz[3 * a + 1] - site_xyz[3 * b + 1]);
    dx = std::min(dx, D.LX - dx); dy = std::min(dy, D.LY - dy);
    return dx > 48 || dy > 48;
  };
  // lanes per team: 8 (four teams, the throughput configuration) or, MSKC_REPLAY_TEAM=16, two 16-lane teams (the instantiation
  // mskc_run_parallel picks when a colour has few tiles)
  int team = 8;
  if (const char *e = getenv("MSKC_REPLAY_TEAM")) { if (atoi(e) == 16) team = 16; }
  const int nteams = 32 / team;
  std::vector<int32_t> batch; int64_t nb = 0, shared = 0;
  for (int64_t i = 0; i < n;) {
    int64_t j = i + 1;
    while (j < n && j - i < nteams) { bool ok = true; for (int64_t k = i; k < j; k++) ok = ok && far(k, j); if (!ok) break; j++; }
    int32_t row[4] = {-1, -1, -1, -1};
    for (int64_t k = i; k < j; k++) row[(nb + (k - i)) % nteams] = (int32_t)k;
    batch.insert(batch.end(), row, row + 4);
    if (j - i > 1) shared++;
    nb++; i = j;
  }
  int32_t *d_s = nullptr, *d_e = nullptr, *d_a = nullptr, *d_b = nullptr; TileStat *d_out = nullptr; TileStat out{};
  CK(cudaMalloc(&d_s, (size_t)n * 12)); CK(cudaMalloc(&d_e, (size_t)n * 4)); CK(cudaMalloc(&d_a, (size_t)n * 4)); CK(cudaMalloc(&d_b, batch.size() * 4)); CK(cudaMalloc(&d_out, sizeof(TileStat)));
  CK(cudaMemcpyAsync(d_s, site_xyz, (size_t)n * 12, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_e, event, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_a, arm, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_b, batch.data(), batch.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(d_out, 0, sizeof(TileStat), h->stream));
  const int feat = (D.NCov > 0 ? F_COV : 0) | (D.la ? F_LA : 0);
  typedef void (*rp_kernel_t)(const int32_t *, const int32_t *, const int32_t *, const int32_t *, long long, double, TileStat *);
  static const rp_kernel_t kern[4] = {replay_teams_kernel<8, 0>, replay_teams_kernel<8, 1>, replay_teams_kernel<8, 2>, replay_teams_kernel<8, 3>};
  static const rp_kernel_t kern16[4] = {replay_teams_kernel<16, 0>, replay_teams_kernel<16, 1>, replay_teams_kernel<16, 2>, replay_teams_kernel<16, 3>};
  const rp_kernel_t k = team == 16 ? kern16[feat] : kern[feat];
  const int64_t chunk = 1000;                    // iterations per launch; tombstones are dropped between launches
  for (int64_t b0 = 0; b0 < nb; b0 += chunk) {
    const int64_t m = nb - b0 < chunk ? nb - b0 : chunk;
    MSKC_LAUNCH(compact_lists_kernel, (h->PC.ntiles * 32 + 127) / 128, 128, 0, h->stream, h->PC.ntiles);
    MSKC_LAUNCH(k, 1, 32, 0, h->stream, d_s, d_e, d_a, d_b + 4 * b0, (long long)m, tree_root, d_out);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&out, d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_s); cudaFree(d_e); cudaFree(d_a); cudaFree(d_b); cudaFree(d_out);
  ParallelCtx &P = h->PC;
  P.replayed = true;
  P.extra.events += out.events; P.extra.site_updates += out.site_updates; P.extra.dAtoms += out.dAtoms; P.extra.dVoids += out.dVoids;
  for (int i = 0; i < 3; i++) { P.extra.dCrystal[i] += out.dCrystal[i]; P.extra.dCov[i] += out.dCov[i]; }
  if (n_iterations) *n_iterations = nb;
  if (n_shared) *n_shared = shared;
  if (check_dev_err(h)) return -1;
  if (out.events != n) return set_err(h, "mskc_replay_teams: stopped after %lld of %lld events", (long long)out.events, (long long)n);
  return 0;
}

extern "C" int mskc_download_active(mskc_ctx *h, int64_t cap, int32_t *xyz, int32_t *event, double *rate, int64_t *n) {
  if (!h || !n) return set_err(h, "mskc_download_active: NULL argument");
  if (!h->parallel_ready) return set_err(h, "mskc_download_active: no parallel state");
  CK(cudaSetDevice(h->device));
  ParallelCtx &P = h->PC; DevCtx &D = h->D;
  std::vector<double> R((size_t)P.ntiles * P.cap); std::vector<unsigned long long> CE(R.size()); std::vector<int> cnt(P.ntiles);
  CK(cudaMemcpy(R.data(), P.rate, R.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(CE.data(), P.ce, CE.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cnt.data(), P.cnt, cnt.size() * 4, cudaMemcpyDeviceToHost));
  long long plane = (long long)D.LX * D.LY; int64_t k = 0;
  for (int l = 0; l < P.ntiles; l++)
    for (int i = 0; i < cnt[l]; i++) {
      size_t a = (size_t)l * P.cap + i;
      if (!(R[a] > 0.0)) continue;
      if (k < cap) {
        uint32_t c = (uint32_t)CE[a];
        if (xyz) { int32_t z = (int32_t)(c / plane); long long r = c - (long long)z * plane; xyz[3 * k + 2] = z; xyz[3 * k + 1] = (int32_t)(r / D.LX); xyz[3 * k] = (int32_t)(r % D.LX); }
        if (event) event[k] = h->tree_variant == MSKC_TREE_LA ? ev_from_dev((int)ce_event(CE[a])) : (int32_t)ce_event(CE[a]);
        if (rate) rate[k] = R[a];
      }
      k++;
    }
  *n = k;
  return 0;
}

extern "C" int mskc_download_bonds(mskc_ctx *h, int64_t n, const int32_t *xyz, int32_t *nn) {
  if (!h || !xyz || !nn) return set_err(h, "mskc_download_bonds: NULL argument");
  if (n <= 0) return 0;
  CK(cudaSetDevice(h->device));
  DevCtx &D = h->D;
  std::vector<uint32_t> cells((size_t)n); std::vector<uint64_t> b((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const int32_t *p = xyz + 3 * i;
    if (p[0] < 0 || p[0] >= D.LX || p[1] < 0 || p[1] >= D.LY || p[2] < 0 || p[2] > D.LZ) return set_err(h, "mskc_download_bonds: site outside the box");
    cells[i] = (uint32_t)((long long)p[0] + (long long)D.LX * ((long long)p[1] + (long long)D.LY * (long long)p[2]));
  }
  uint32_t *d_c = nullptr; unsigned long long *d_b = nullptr;
  CK(cudaMalloc(&d_c, (size_t)n * 4)); CK(cudaMalloc(&d_b, (size_t)n * 8));
  CK(cudaMemcpyAsync(d_c, cells.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  MSKC_LAUNCH(gather_bonds, (unsigned)((n + 255) / 256), 256, 0, h->stream, D.bonds, d_c, d_b, (long long)n);
  CK(cudaMemcpyAsync(b.data(), d_b, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_c); cudaFree(d_b);
  for (int64_t i = 0; i < n; i++) h_dec_bonds(D, xyz + 3 * i, b[i], nn + 12 * i);
  return 0;
}

// ---- x-decomposed slabs (several GPUs) -----------------------------------------------------------------------
extern "C" int mskc_create_slab(int global_lenx, int leny, int lenz, int ncrystal, int ncov, const int32_t *listcrystal, const int32_t *listcov,
                                int rank, int nranks, int device_id, mskc_ctx **out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_err(nullptr, "mskc_create_slab: rank %d of %d", rank, nranks);
  if (global_lenx % (48 * nranks) || leny % 48) return set_err(nullptr, "mskc_create_slab: global LenX must be a multiple of 48*nranks and LenY of 48");
  const int owned = global_lenx / nranks;
  if (owned < 48) return set_err(nullptr, "mskc_create_slab: fewer than two tile columns per rank");
  if (mskc_create(owned + 48, leny, lenz, ncrystal, ncov, listcrystal, listcov, device_id, 0, out)) return -1;
  (*out)->PC.halo = 1; (*out)->PC.rank = rank; (*out)->PC.nranks = nranks;
  return 0;
}
extern "C" int mskc_comm_unique_id(void *id128) {
  if (!id128) return set_err(nullptr, "mskc_comm_unique_id: NULL argument");
  std::string e; NcclApi *nccl = nccl_api(&e);
  if (!nccl) return set_err(nullptr, "%s", e.c_str());
  ncclUniqueId id;
  ncclResult_t r = nccl->GetUniqueId(&id);
  if (r != ncclSuccess) return set_err(nullptr, "ncclGetUniqueId: %s", nccl->GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}
extern "C" int mskc_comm_connect(mskc_ctx *h, const void *id128) {
  if (!h || !id128) return set_err(h, "mskc_comm_connect: NULL argument");
  if (!h->PC.halo) return set_err(h, "mskc_comm_connect: not a slab context (use mskc_create_slab)");
  if (h->PC.nranks == 1) return 0;
  CK(cudaSetDevice(h->device));
  std::string e; NcclApi *nccl = nccl_api(&e);
  if (!nccl) return set_err(h, "%s", e.c_str());
  ncclUniqueId id; memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = nccl->CommInitRank(&comm, h->PC.nranks, id, h->PC.rank);
  if (r != ncclSuccess) return set_err(h, "ncclCommInitRank: %s", nccl->GetErrorString(r));
  h->PC.comm = (void *)comm;
  // halo buffers of the neighbours mapped into this process (CUDA IPC over NVLink); ncclSend/ncclRecv is the fallback
  if (p2p_connect(h->PC, h->stream, &h->err)) { g_last_error = h->err; return -1; }
  return 0;
}

// which = 0: the cells of mskc_extract_frame; 1: mskc_extract_frame_all; 2 / 3: mskc_extract_antisites with target 0 / 3
static int extract_cells(mskc_ctx *h, const char *who, int which, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms) {
  if (!h || !n) return set_err(h, "%s: NULL argument", who);
  if (!h->have_state) return set_err(h, "%s: no state", who);
  CK(cudaSetDevice(h->device));
  if (cap < 0) cap = 0;
  if (cap > h->frame_cap || !h->d_frame_n) {
    cudaFree(h->d_frame_cell); cudaFree(h->d_frame_word); h->d_frame_cell = h->d_frame_word = nullptr;
    h->frame_cap = cap > 0 ? cap : 1;
    CK(cudaMalloc(&h->d_frame_cell, (size_t)h->frame_cap * 4)); CK(cudaMalloc(&h->d_frame_word, (size_t)h->frame_cap * 4));
    if (!h->d_frame_n) CK(cudaMalloc(&h->d_frame_n, 8));
  }
  if (upload_consts(h)) return -1;
  CK(cudaMemsetAsync(h->d_frame_n, 0, 8, h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  if (which == 0) MSKC_LAUNCH(frame_extract_kernel<0>, 148 * 8, 256, 0, h->stream, h->d_frame_cell, h->d_frame_word, (unsigned long long)cap, h->d_frame_n);
  else if (which == 1) MSKC_LAUNCH(frame_extract_kernel<1>, 148 * 8, 256, 0, h->stream, h->d_frame_cell, h->d_frame_word, (unsigned long long)cap, h->d_frame_n);
  else MSKC_LAUNCH(antisite_kernel, 148 * 8, 256, 0, h->stream, which == 2 ? 0 : 3, h->d_frame_cell, h->d_frame_word, (unsigned long long)cap, h->d_frame_n);
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaGetLastError());
  unsigned long long cnt = 0;
  CK(cudaMemcpyAsync(&cnt, h->d_frame_n, 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int64_t m = (int64_t)cnt < cap ? (int64_t)cnt : cap;
  if (m > 0 && cell) CK(cudaMemcpyAsync(cell, h->d_frame_cell, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
  if (m > 0 && word) CK(cudaMemcpyAsync(word, h->d_frame_word, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->tree_variant == MSKC_TREE_LA && word) for (int64_t i = 0; i < m; i++) word[i] = word_from_dev(word[i]);
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, h->ev0, h->ev1));
  *n = (int64_t)cnt;
  return 0;
}
extern "C" int mskc_extract_frame(mskc_ctx *h, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms) {
  return extract_cells(h, "mskc_extract_frame", 0, cap, cell, word, n, kernel_ms);
}
extern "C" int mskc_extract_frame_all(mskc_ctx *h, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms) {
  return extract_cells(h, "mskc_extract_frame_all", 1, cap, cell, word, n, kernel_ms);
}
extern "C" int mskc_extract_antisites(mskc_ctx *h, int target, int64_t cap, uint32_t *cell, uint32_t *word, int64_t *n, float *kernel_ms) {
  if (target != 0 && target != 3) return set_err(h, "mskc_extract_antisites: target must be 0 or 3 (WriteMolMolXYZFile.f:330,352)");
  return extract_cells(h, "mskc_extract_antisites", target == 0 ? 2 : 3, cap, cell, word, n, kernel_ms);
}
extern "C" int mskc_gather_lattice(mskc_ctx *h, int64_t n, const int64_t *cell, int32_t *word) {
  if (!h || (n > 0 && (!cell || !word))) return set_err(h, "mskc_gather_lattice: NULL argument");
  if (!h->have_state) return set_err(h, "mskc_gather_lattice: no state");
  if (n <= 0) return 0;
  CK(cudaSetDevice(h->device));
  std::vector<uint32_t> c((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    if (cell[i] < 0 || cell[i] >= h->D.ncell) return set_err(h, "mskc_gather_lattice: cell out of range");
    c[(size_t)i] = (uint32_t)cell[i];
  }
  uint32_t *d_c = nullptr, *d_w = nullptr;
  CK(cudaMalloc(&d_c, (size_t)n * 4)); CK(cudaMalloc(&d_w, (size_t)n * 4));
  CK(cudaMemcpyAsync(d_c, c.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  MSKC_LAUNCH(gather_state, (unsigned)((n + 255) / 256), 256, 0, h->stream, (const uint32_t *)h->D.state, (const uint32_t *)d_c, d_w, (long long)n);
  CK(cudaMemcpyAsync(word, d_w, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_c); cudaFree(d_w);
  if (h->tree_variant == MSKC_TREE_LA) for (int64_t i = 0; i < n; i++) word[i] = (int32_t)word_from_dev((uint32_t)word[i]);
  return 0;
}
