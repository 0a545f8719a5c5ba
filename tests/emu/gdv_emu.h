// gdv_emu.h — TEST INFRASTRUCTURE.  The CUDA C++ surface the generated kernels and the device
// function library use, restated for a host compiler, so the *unmodified* generated source can be
// executed on a box without a GPU by the functional SIMT simulator in tests/emu/ (fake libcuda +
// fake libnvrtc, see tests/emu/README.md).  Nothing under gandiva_b200/ includes, links or loads
// this: the simulator is substituted for the CUDA driver by LD_LIBRARY_PATH in a test subprocess.
// It checks logic (null propagation, tails, scans, look-back, staging protocols), not performance
// and not the memory model.
//
// Execution model: one CTA at a time, its threads are cooperative fibers on one OS thread; a
// thread runs until it reaches a warp / CTA collective or a spin-wait, where it yields.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline __attribute__((noinline))
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct gdv_emu_uint3 {
  unsigned x, y, z;
};
struct gdv_emu_thread {
  gdv_emu_uint3 tid, bid, bdim, gdim;
  void* dyn_smem;
  unsigned lane, warp;
};
extern "C" {
extern gdv_emu_thread* gdv_emu_cur;
void gdv_emu_yield();
void gdv_emu_syncthreads();
// Every lane named in `mask` of the calling warp deposits v; returns once all have, with the
// 32 deposited values in out[] (lanes outside the mask: 0).
void gdv_emu_warp_gather(unsigned mask, unsigned long long v, unsigned long long* out);
}
#define threadIdx (gdv_emu_cur->tid)
#define blockIdx (gdv_emu_cur->bid)
#define blockDim (gdv_emu_cur->bdim)
#define gridDim (gdv_emu_cur->gdim)
#define warpSize 32

struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct alignas(16) int4 {
  int x, y, z, w;
};
struct alignas(8) uint2 {
  unsigned x, y;
};
struct alignas(8) int2 {
  int x, y;
};
struct alignas(16) longlong2 {
  long long x, y;
};
struct alignas(16) ulonglong2 {
  unsigned long long x, y;
};
struct alignas(16) double2 {
  double x, y;
};

// ---- warp / CTA collectives --------------------------------------------------------------
template <typename T>
inline unsigned long long gdv_emu_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned long long b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
inline T gdv_emu_unbits(unsigned long long b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
inline void __syncthreads() { gdv_emu_syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, 0ull, o);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, pred ? 1ull : 0ull, o);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && o[l]) r |= 1u << l;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, gdv_emu_bits(v), o);
  const int lane = (int)gdv_emu_cur->lane;
  const int base = lane & ~(width - 1);
  return gdv_emu_unbits<T>(o[base + (src & (width - 1))]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, gdv_emu_bits(v), o);
  const int lane = (int)gdv_emu_cur->lane;
  const int base = lane & ~(width - 1);
  const int src = lane - (int)delta;
  return src < base ? v : gdv_emu_unbits<T>(o[src]);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, gdv_emu_bits(v), o);
  const int lane = (int)gdv_emu_cur->lane;
  const int base = lane & ~(width - 1);
  const int src = lane + (int)delta;
  return src >= base + width ? v : gdv_emu_unbits<T>(o[src]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  unsigned long long o[32];
  gdv_emu_warp_gather(mask, gdv_emu_bits(v), o);
  const int lane = (int)gdv_emu_cur->lane;
  const int src = lane ^ lanemask;
  return (src & ~(width - 1)) != (lane & ~(width - 1)) ? v : gdv_emu_unbits<T>(o[src]);
}

// ---- scalar intrinsics -----------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(v >> (sh & 31u));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)((v << (sh & 31u)) >> 32);
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
inline long long __double_as_longlong(double d) { return gdv_emu_unbits<long long>(gdv_emu_bits(d)); }
inline double __longlong_as_double(long long v) { return gdv_emu_unbits<double>(gdv_emu_bits(v)); }
inline unsigned __float_as_uint(float f) { return gdv_emu_unbits<unsigned>(gdv_emu_bits(f)); }
inline float __uint_as_float(unsigned v) { return gdv_emu_unbits<float>(gdv_emu_bits(v)); }
inline int __float_as_int(float f) { return gdv_emu_unbits<int>(gdv_emu_bits(f)); }
inline float __int_as_float(int v) { return gdv_emu_unbits<float>(gdv_emu_bits(v)); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
inline void __stcg(T* p, T v) { *p = v; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { gdv_emu_yield(); }

// atomics: fibers never preempt each other, plain read-modify-write is atomic here
template <typename T, typename U>
inline T atomicAdd(T* p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U>
inline T atomicMax(T* p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U>
inline T atomicMin(T* p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U>
inline T atomicOr(T* p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U>
inline T atomicAnd(T* p, U v) { const T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename U>
inline T atomicExch(T* p, U v) { const T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V>
inline T atomicCAS(T* p, U cmp, V v) { const T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- the primitives gdv_device_lib.cuh implements with inline PTX (#ifndef GDV_HOST_EMU there) ----
// mbarrier object (one u64 in shared memory): [63] phase, [62:48] expected arrivals,
// [47:32] pending arrivals, [31:0] pending transaction bytes.
inline void gdv_emu_mbar_settle(unsigned long long* bar) {
  unsigned long long s = *bar;
  const unsigned pending = (unsigned)((s >> 32) & 0xffffu), tx = (unsigned)s;
  if (pending == 0u && tx == 0u) {
    const unsigned long long init = (s >> 48) & 0x7fffull;
    s = ((s ^ (1ull << 63)) & (1ull << 63)) | (init << 48) | (init << 32);
    *bar = s;
  }
}
inline void gdv_mbar_init(unsigned long long* bar, unsigned count) {
  *bar = ((unsigned long long)count << 48) | ((unsigned long long)count << 32);
}
inline void gdv_fence_mbar_init() {}
inline void gdv_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  unsigned long long s = *bar;
  const unsigned pending = (unsigned)((s >> 32) & 0xffffu) - 1u;
  const unsigned tx = (unsigned)s + bytes;
  *bar = (s & 0xffff000000000000ull) | ((unsigned long long)(pending & 0xffffu) << 32) | tx;
  gdv_emu_mbar_settle(bar);
}
inline unsigned long long gdv_policy_evict_first() { return 0ull; }
inline void gdv_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                         unsigned long long) {
  // cp.async.bulk contract: 16-byte aligned on both sides, size a multiple of 16
  if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) != 0u) __builtin_trap();
  std::memcpy(dst, src, bytes);
  *bar = (*bar & 0xffffffff00000000ull) | (unsigned)((unsigned)*bar - bytes);
  gdv_emu_mbar_settle(bar);
}
inline bool gdv_mbar_try_wait(unsigned long long* bar, unsigned parity) {
  if ((unsigned)(*bar >> 63) != (parity & 1u)) return true;
  gdv_emu_yield();
  return false;
}
inline void gdv_mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!gdv_mbar_try_wait(bar, parity)) {
  }
}
inline void gdv_cp_async16(void* smem_dst, const void* gsrc) {
  if ((((uintptr_t)smem_dst | (uintptr_t)gsrc) & 15u) != 0u) __builtin_trap();
  std::memcpy(smem_dst, gsrc, 16);
}
inline void gdv_cp_async_commit() {}
template <int N>
inline void gdv_cp_async_wait() {}
inline unsigned gdv_lanemask_lt() { return (1u << gdv_emu_cur->lane) - 1u; }
inline unsigned long long gdv_ld_relaxed(const unsigned long long* p) {
  const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
  gdv_emu_yield();  // look-back polls: let the producers run
  return v;
}
inline void gdv_st_relaxed(unsigned long long* p, unsigned long long v) { *p = v; }
inline unsigned long long gdv_ld_acquire_sys(const unsigned long long* p) {
  const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
  gdv_emu_yield();
  return v;
}
inline void gdv_st_release_sys(unsigned long long* p, unsigned long long v) { *p = v; }
