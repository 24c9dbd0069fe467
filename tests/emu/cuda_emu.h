// cuda_emu.h -- functional SIMT emulator for running the library's CUDA sources on a CPU.
//
// TEST INFRASTRUCTURE ONLY.  The build container has no GPU; this header lets g++ compile
// mulskips_b200/csrc/mskc_api.cu (-DMSKC_EMU -include cuda_emu.h) into tests/emu/_build/libmskc.so, an
// EMULATED copy of the C ABI in which every kernel launch runs its thread blocks one after another, each
// thread of a block as a fiber (ucontext), and the warp collectives (__shfl*_sync, __ballot_sync, __any_sync,
// __syncwarp, __syncthreads) are rendez-vous points between the fibers.  It is a debugging aid for kernel
// logic (lock-step teams, shuffles, list bookkeeping) before GPU time is spent: nothing in the product
// loads it, `-m gpu` tests and bench.py never see it, and it is orders of magnitude slower than a CPU port.
// The only consumers are tests that set MSKC_LIB to the emulated library explicitly (tests/test_emu_*.py).
#pragma once
#include <sys/mman.h>
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>
#include <dlfcn.h>

// ---- language extensions -> nothing ---------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __constant__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

using std::min; using std::max; using std::abs;
static inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
static inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
static inline long long max(long long a, int b) { return a > b ? a : (long long)b; }

// ---- the fiber scheduler -----------------------------------------------------------------------------------
// user-level context switch (x86-64 SysV: callee-saved registers + stack pointer); swapcontext() would make a
// sigprocmask system call per switch, and a kernel with warp collectives switches ~10^8 times in a test
extern "C" void emu_switch(void **save_sp, void *load_sp);
__asm__(
    ".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch,.-emu_switch\n");
namespace emu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Coll { unsigned long long gen = 0; unsigned arrived = 0, fin[2] = {0, 0}; unsigned long long buf[2][32]; };
struct Fiber { void *sp = nullptr; char *stack = nullptr; bool done = true; unsigned tid = 0; };
struct State {
  Idx threadIdx, blockIdx; dim3 blockDim, gridDim;
  std::vector<Fiber> fibers; void *sched_sp = nullptr; int cur = -1; unsigned nthreads = 0;
  std::vector<unsigned> alive;                                // per warp: lanes that have not returned yet
  std::map<std::pair<unsigned, unsigned>, Coll> colls;        // (warp, mask) -> rendez-vous state
  unsigned long long bar_gen = 0; unsigned bar_arrived = 0, n_alive = 0;
  std::vector<unsigned char> smem; std::function<void()> body;
  unsigned long long progress = 0; bool poison = true;
};
inline State &S() { static State s; return s; }
enum { STACK = 256 * 1024 };
inline void yield() { State &s = S(); emu_switch(&s.fibers[s.cur].sp, s.sched_sp); }
inline void trampoline() {
  State &s = S(); Fiber &f = s.fibers[s.cur];
  s.body();
  f.done = true; s.alive[f.tid >> 5] &= ~(1u << (f.tid & 31)); s.n_alive--; s.progress++;
  emu_switch(&f.sp, s.sched_sp);
  abort();                                                   // a finished fiber is never resumed
}
inline void run_block() {
  State &s = S();
  const unsigned n = s.blockDim.x * s.blockDim.y * s.blockDim.z;
  if (s.fibers.size() < n) {
    size_t old = s.fibers.size(); s.fibers.resize(n);
    for (size_t i = old; i < n; i++) {
      s.fibers[i].stack = (char *)mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (s.fibers[i].stack == (char *)MAP_FAILED) { fprintf(stderr, "cuda_emu: cannot allocate fiber stacks\n"); abort(); }
    }
  }
  s.nthreads = n; s.alive.assign((n + 31) / 32, 0u); s.colls.clear(); s.bar_gen = 0; s.bar_arrived = 0; s.n_alive = n;
  for (unsigned i = 0; i < n; i++) {
    Fiber &f = s.fibers[i]; f.done = false; f.tid = i; s.alive[i >> 5] |= 1u << (i & 31);
    // initial frame: six callee-saved registers, the entry point as emu_switch's return address, a fake return slot
    void **top = reinterpret_cast<void **>((reinterpret_cast<uintptr_t>(f.stack) + STACK) & ~(uintptr_t)15);
    *--top = nullptr; *--top = reinterpret_cast<void *>(&trampoline);
    for (int r = 0; r < 6; r++) *--top = nullptr;
    f.sp = top;
  }
  unsigned long long last = s.progress; int idle_rounds = 0;
  while (s.n_alive) {
    for (unsigned i = 0; i < n; i++) {
      Fiber &f = s.fibers[i];
      if (f.done) continue;
      s.cur = (int)i; s.threadIdx.x = i % s.blockDim.x; s.threadIdx.y = (i / s.blockDim.x) % s.blockDim.y; s.threadIdx.z = i / (s.blockDim.x * s.blockDim.y);
      emu_switch(&s.sched_sp, f.sp);
    }
    if (s.progress == last) { if (++idle_rounds > 1000) { fprintf(stderr, "cuda_emu: deadlock (a collective waits for lanes that never arrive)\n"); abort(); } }
    else { idle_rounds = 0; last = s.progress; }
  }
  s.cur = -1;
}
template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F &&f) {
  State &s = S();
  s.gridDim = grid; s.blockDim = block; s.body = std::function<void()>(f);
  s.smem.assign(smem + 64, 0);
  for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
    s.blockIdx.x = bx; s.blockIdx.y = by; s.blockIdx.z = bz;
    run_block();
  }
}
inline unsigned char *dyn_smem() { State &s = S(); return reinterpret_cast<unsigned char *>(((uintptr_t)s.smem.data() + 15) & ~(uintptr_t)15); }
inline unsigned lane() { return S().fibers[S().cur].tid & 31u; }
inline unsigned warp() { return S().fibers[S().cur].tid >> 5; }
// rendez-vous of the lanes in `mask` (of the calling lane's warp); returns the slots all of them wrote
inline const unsigned long long *exchange(unsigned mask, unsigned long long v, unsigned *arrived_out) {
  State &s = S(); const unsigned l = lane(), w = warp();
  Coll &c = s.colls[std::make_pair(w, mask)];
  const unsigned long long g = c.gen; const int par = (int)(g & 1);
  c.buf[par][l] = v; c.arrived |= 1u << l;
  while (c.gen == g) {
    const unsigned need = mask & s.alive[w];
    if ((c.arrived & need) == need) { c.fin[par] = c.arrived; c.gen++; c.arrived = 0; s.progress++; break; }
    yield();
  }
  if (arrived_out) *arrived_out = c.fin[par];
  return c.buf[par];
}
inline void syncthreads() {
  State &s = S(); const unsigned long long g = s.bar_gen;
  s.bar_arrived++;
  while (s.bar_gen == g) {
    if (s.bar_arrived >= s.n_alive) { s.bar_gen++; s.bar_arrived = 0; s.progress++; break; }
    yield();
  }
}
template <class T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu

#define threadIdx (emu::S().threadIdx)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)

// ---- warp collectives ----------------------------------------------------------------------------------------
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned arr; const unsigned l = emu::lane(); const unsigned long long *b = emu::exchange(mask, emu::to_bits(v), &arr);
  const unsigned s = (l & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
  return ((arr >> s) & 1u) ? emu::from_bits<T>(b[s]) : v;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned arr; const unsigned l = emu::lane(); const unsigned long long *b = emu::exchange(mask, emu::to_bits(v), &arr);
  if ((l & (unsigned)(width - 1)) < delta) return v;
  const unsigned s = l - delta;
  return ((arr >> s) & 1u) ? emu::from_bits<T>(b[s]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned arr; const unsigned l = emu::lane(); const unsigned long long *b = emu::exchange(mask, emu::to_bits(v), &arr);
  if ((l & (unsigned)(width - 1)) + delta >= (unsigned)width) return v;
  const unsigned s = l + delta;
  return ((arr >> s) & 1u) ? emu::from_bits<T>(b[s]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  unsigned arr; const unsigned l = emu::lane(); const unsigned long long *b = emu::exchange(mask, emu::to_bits(v), &arr);
  const unsigned s = l ^ (unsigned)lanemask;
  if ((s & ~(unsigned)(width - 1)) != (l & ~(unsigned)(width - 1))) return v;
  return ((arr >> s) & 1u) ? emu::from_bits<T>(b[s]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned arr; const unsigned long long *b = emu::exchange(mask, pred ? 1ull : 0ull, &arr);
  unsigned r = 0; for (int i = 0; i < 32; i++) if (((arr >> i) & 1u) && b[i]) r |= 1u << i;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { unsigned arr; const unsigned long long *b = emu::exchange(mask, pred ? 1ull : 0ull, &arr); for (int i = 0; i < 32; i++) if (((arr >> i) & 1u) && !b[i]) return 0; return 1; }
static inline int __reduce_add_sync(unsigned mask, int v) {
  unsigned arr; const unsigned long long *b = emu::exchange(mask, emu::to_bits(v), &arr);
  int r = 0; for (int i = 0; i < 32; i++) if ((arr >> i) & 1u) r += emu::from_bits<int>(b[i]);
  return r;
}
static inline unsigned __activemask() { return 0xFFFFFFFFu; }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::exchange(mask, 0ull, nullptr); }
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

// ---- scalar intrinsics ------------------------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#define __logf logf      /* glibc declares __logf itself */
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldca(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (s & 31)); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  unsigned long long t = ((unsigned long long)b << 32) | a; unsigned r = 0;
  for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 7u; r |= (unsigned)((t >> (8 * sel)) & 0xFFu) << (8 * i); }
  return r;
}

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }

// ---- runtime API --------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
typedef struct emu_stream_ *cudaStream_t;
struct emu_event_ { std::chrono::steady_clock::time_point t; };
typedef emu_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = getenv("MSKC_EMU_NO_DEVICE") ? 0 : 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
  void *q = malloc(n ? n : 1); if (!q) return 2;
  if (emu::S().poison) memset(q, 0xCD, n);               // uninitialised device memory is not zero
  *p = (T *)q; return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) { memcpy((char *)&sym + off, s, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T &sym, const void *s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy((char *)&sym + off, s, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyFromSymbol(void *d, const T &sym, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost) { memcpy(d, (const char *)&sym + off, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event_(); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- NCCL types (the library loads libnccl with dlopen, and only for multi-rank slab runs) -------------------------
typedef int ncclResult_t; enum { ncclSuccess = 0 };
typedef struct emu_nccl_comm_ *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int ncclDataType_t; enum { ncclUint8 = 1, ncclDouble = 8 };
typedef int ncclRedOp_t; enum { ncclSum = 0 };

// ---- kernel launch ---------------------------------------------------------------------------------------------------
#define MSKC_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define MSKC_DYN_SHARED(name) unsigned char *name = emu::dyn_smem()
