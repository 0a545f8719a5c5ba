// fake_cuda.cc — TEST INFRASTRUCTURE: a functional stand-in for libcuda.so.1 so that the whole,
// unmodified product stack (C-ABI -> runtime -> driver calls -> generated kernels) can be
// exercised on a box without a GPU.  Built to tests/emu/lib/libcuda.so.1 and put in front of the
// real driver with LD_LIBRARY_PATH by tests/test_emu.py only; see tests/emu/README.md.
//
// "Device memory" is host memory, streams execute synchronously at enqueue time, and
// cuLaunchKernel runs the grid one CTA at a time with every CUDA thread as a ucontext fiber
// (tests/emu/gdv_emu.h has the device-side half).  A module is either the blob fake_nvrtc.cc
// returns ("GDVEMU1\0<path of a host-compiled .so>") or, for any other image (the embedded
// static-kernel cubin), the host build of device/static_kernels.cu (libgdv_emu_static.so).
#include <cuda.h>
#include <dlfcn.h>
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

struct gdv_emu_uint3 {
  unsigned x, y, z;
};
struct gdv_emu_thread {
  gdv_emu_uint3 tid, bid, bdim, gdim;
  void* dyn_smem;
  unsigned lane, warp;
};

namespace {

constexpr size_t kStackBytes = 256 * 1024;

struct WarpSync {
  unsigned arrived = 0, gen = 0;
  unsigned long long vals[32];
  unsigned long long snap[32];
};
struct Fiber {
  ucontext_t ctx;
  gdv_emu_thread th;
  bool done = false;
};
struct CtaRun {
  std::vector<Fiber> fibers;
  std::vector<WarpSync> warps;
  unsigned live = 0, bar_arrived = 0, bar_gen = 0;
  ucontext_t sched;
  void (*entry)(void**) = nullptr;
  void** params = nullptr;
};
CtaRun* g_cta = nullptr;
std::vector<char*> g_stacks;
std::recursive_mutex g_mu;

struct Module {
  void* handle = nullptr;
};
struct Function {
  void (*entry)(void**) = nullptr;
  std::string name;
  int max_dyn_smem = 48 * 1024;
};

}  // namespace

extern "C" {
gdv_emu_thread* gdv_emu_cur = nullptr;

void gdv_emu_yield() {
  CtaRun* c = g_cta;
  Fiber* f = reinterpret_cast<Fiber*>(reinterpret_cast<char*>(gdv_emu_cur) - offsetof(Fiber, th));
  swapcontext(&f->ctx, &c->sched);
}

void gdv_emu_syncthreads() {
  CtaRun* c = g_cta;
  const unsigned my = c->bar_gen;
  if (++c->bar_arrived >= c->live) {
    c->bar_arrived = 0;
    ++c->bar_gen;
    return;
  }
  while (c->bar_gen == my) gdv_emu_yield();
}

void gdv_emu_warp_gather(unsigned mask, unsigned long long v, unsigned long long* out) {
  CtaRun* c = g_cta;
  gdv_emu_thread* t = gdv_emu_cur;
  WarpSync& w = c->warps[t->warp];
  if (((mask >> t->lane) & 1u) == 0u) {
    std::fprintf(stderr, "gdv_emu: lane %u calls a warp collective with mask %08x\n", t->lane, mask);
    std::abort();
  }
  w.vals[t->lane] = v;
  w.arrived |= 1u << t->lane;
  const unsigned my = w.gen;
  if (w.arrived == mask) {
    for (int l = 0; l < 32; ++l) w.snap[l] = ((mask >> l) & 1u) ? w.vals[l] : 0ull;
    w.arrived = 0;
    ++w.gen;
  } else {
    if ((w.arrived & ~mask) != 0u) {
      std::fprintf(stderr, "gdv_emu: warp collective with diverging masks\n");
      std::abort();
    }
    while (w.gen == my) gdv_emu_yield();
  }
  std::memcpy(out, w.snap, sizeof(w.snap));
}
}

namespace {

void FiberEntry() {
  CtaRun* c = g_cta;
  Fiber* f = reinterpret_cast<Fiber*>(reinterpret_cast<char*>(gdv_emu_cur) - offsetof(Fiber, th));
  c->entry(c->params);
  f->done = true;
  --c->live;
  // a thread that exits no longer takes part in CTA barriers: release one that is now complete
  if (c->live > 0 && c->bar_arrived >= c->live) {
    c->bar_arrived = 0;
    ++c->bar_gen;
  }
  swapcontext(&f->ctx, &c->sched);
}

bool RunCta(Function* fn, const gdv_emu_uint3& bid, const gdv_emu_uint3& bdim,
            const gdv_emu_uint3& gdim, void* dyn_smem, void** params) {
  const unsigned nthreads = bdim.x * bdim.y * bdim.z;
  CtaRun cta;
  cta.fibers.resize(nthreads);
  cta.warps.resize((nthreads + 31) / 32);
  cta.live = nthreads;
  cta.entry = fn->entry;
  cta.params = params;
  while (g_stacks.size() < nthreads) g_stacks.push_back(static_cast<char*>(std::malloc(kStackBytes)));
  g_cta = &cta;
  for (unsigned i = 0; i < nthreads; ++i) {
    Fiber& f = cta.fibers[i];
    f.th.tid = {i % bdim.x, (i / bdim.x) % bdim.y, i / (bdim.x * bdim.y)};
    f.th.bid = bid;
    f.th.bdim = bdim;
    f.th.gdim = gdim;
    f.th.dyn_smem = dyn_smem;
    f.th.lane = i & 31u;
    f.th.warp = i >> 5;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = g_stacks[i];
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &cta.sched;
    makecontext(&f.ctx, FiberEntry, 0);
  }
  unsigned long long idle_rounds = 0;
  while (cta.live > 0) {
    const unsigned live_before = cta.live;
    const unsigned gen_before = cta.bar_gen;
    for (unsigned i = 0; i < nthreads; ++i) {
      Fiber& f = cta.fibers[i];
      if (f.done) continue;
      gdv_emu_cur = &f.th;
      swapcontext(&cta.sched, &f.ctx);
    }
    // crude hang detector: nothing exits and no barrier completes for a very long time
    if (cta.live == live_before && cta.bar_gen == gen_before) {
      if (++idle_rounds > 100'000ull) {
        std::fprintf(stderr, "gdv_emu: kernel %s appears to hang (CTA %u)\n", fn->name.c_str(), bid.x);
        g_cta = nullptr;
        return false;
      }
    } else {
      idle_rounds = 0;
    }
  }
  g_cta = nullptr;
  gdv_emu_cur = nullptr;
  return true;
}

int g_device_count = 1;
int g_sm_count = 4;

}  // namespace

#define EMU_OK return CUDA_SUCCESS

extern "C" {

CUresult cuInit(unsigned int) {
  if (const char* e = std::getenv("GDV_EMU_DEVICES")) g_device_count = std::atoi(e);
  if (const char* e = std::getenv("GDV_EMU_SMS")) g_sm_count = std::atoi(e);
  EMU_OK;
}
CUresult cuDeviceGetCount(int* n) { *n = g_device_count; EMU_OK; }
CUresult cuDeviceGet(CUdevice* d, int ordinal) {
  if (ordinal < 0 || ordinal >= g_device_count) return CUDA_ERROR_INVALID_DEVICE;
  *d = ordinal;
  EMU_OK;
}
CUresult cuDeviceGetAttribute(int* v, CUdevice_attribute a, CUdevice) {
  switch (a) {
    case CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT: *v = g_sm_count; break;
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR: *v = 10; break;
    case CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR: *v = 0; break;
    case CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN: *v = 227 * 1024; break;
    case CU_DEVICE_ATTRIBUTE_WARP_SIZE: *v = 32; break;
    default: *v = 0; break;
  }
  EMU_OK;
}
CUresult cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice d) {
  *c = reinterpret_cast<CUcontext>(static_cast<uintptr_t>(0x1000 + d));
  EMU_OK;
}
CUresult cuCtxSetCurrent(CUcontext) { EMU_OK; }
CUresult cuCtxGetCurrent(CUcontext* c) { *c = reinterpret_cast<CUcontext>(static_cast<uintptr_t>(0x1000)); EMU_OK; }

// allocations are padded on both sides: the staged loaders legitimately read a few bytes around a
// column (16-byte aligned windows), which is inside the allocation granule on a real device
constexpr size_t kPad = 512;
static std::map<uintptr_t, size_t> g_allocs;
CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t n) {
  std::lock_guard<std::recursive_mutex> lock(g_mu);
  void* raw = nullptr;
  if (posix_memalign(&raw, 512, n + 2 * kPad) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
  std::memset(raw, 0xCD, n + 2 * kPad);  // garbage, like a fresh device allocation
  const uintptr_t user = reinterpret_cast<uintptr_t>(raw) + kPad;
  g_allocs[user] = n;
  *p = static_cast<CUdeviceptr>(user);
  EMU_OK;
}
// Guard bands: the kPad bytes on either side of every allocation keep their 0xCD fill unless a
// kernel (or a copy) wrote out of bounds.  Checked after every launch and at free.
static bool GuardsIntact(const char* when) {
  bool ok = true;
  for (const auto& kv : g_allocs) {
    const unsigned char* lo = reinterpret_cast<const unsigned char*>(kv.first - kPad);
    const unsigned char* hi = reinterpret_cast<const unsigned char*>(kv.first + kv.second);
    for (size_t i = 0; i < kPad; ++i) {
      if (lo[i] != 0xCD || hi[i] != 0xCD) {
        std::fprintf(stderr, "gdv_emu: out-of-bounds WRITE detected %s: allocation of %zu bytes at %p, guard byte %s%zu\n",
                     when, kv.second, reinterpret_cast<void*>(kv.first), lo[i] != 0xCD ? "-" : "+",
                     lo[i] != 0xCD ? kPad - i : i);
        ok = false;
        break;
      }
    }
  }
  return ok;
}
CUresult cuMemFree_v2(CUdeviceptr p) {
  std::lock_guard<std::recursive_mutex> lock(g_mu);
  if (p == 0) EMU_OK;
  auto it = g_allocs.find(static_cast<uintptr_t>(p));
  if (it == g_allocs.end()) return CUDA_ERROR_INVALID_VALUE;
  if (!GuardsIntact("at cuMemFree")) return CUDA_ERROR_ILLEGAL_ADDRESS;
  g_allocs.erase(it);
  std::free(reinterpret_cast<void*>(static_cast<uintptr_t>(p) - kPad));
  EMU_OK;
}
CUresult cuMemGetAddressRange_v2(CUdeviceptr* base, size_t* size, CUdeviceptr p) {
  std::lock_guard<std::recursive_mutex> lock(g_mu);
  auto it = g_allocs.upper_bound(static_cast<uintptr_t>(p));
  if (it == g_allocs.begin()) return CUDA_ERROR_INVALID_VALUE;
  --it;
  if (static_cast<uintptr_t>(p) >= it->first + it->second) return CUDA_ERROR_INVALID_VALUE;
  if (base) *base = static_cast<CUdeviceptr>(it->first);
  if (size) *size = it->second;
  EMU_OK;
}
CUresult cuMemHostAlloc(void** p, size_t n, unsigned int) {
  *p = std::malloc(n ? n : 1);
  return *p ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY;
}
CUresult cuMemFreeHost(void* p) { std::free(p); EMU_OK; }
CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr d, const void* s, size_t n, CUstream) {
  std::memcpy(reinterpret_cast<void*>(d), s, n);
  EMU_OK;
}
CUresult cuMemcpyDtoHAsync_v2(void* d, CUdeviceptr s, size_t n, CUstream) {
  std::memcpy(d, reinterpret_cast<const void*>(s), n);
  EMU_OK;
}
CUresult cuMemcpyDtoDAsync_v2(CUdeviceptr d, CUdeviceptr s, size_t n, CUstream) {
  std::memmove(reinterpret_cast<void*>(d), reinterpret_cast<const void*>(s), n);
  EMU_OK;
}
CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream) {
  std::memset(reinterpret_cast<void*>(d), v, n);
  EMU_OK;
}
CUresult cuStreamCreate(CUstream* s, unsigned int) {
  static uintptr_t next = 0x2000;
  *s = reinterpret_cast<CUstream>(next += 16);
  EMU_OK;
}
CUresult cuStreamSynchronize(CUstream) { EMU_OK; }
CUresult cuStreamDestroy_v2(CUstream) { EMU_OK; }
CUresult cuStreamWaitEvent(CUstream, CUevent, unsigned int) { EMU_OK; }
CUresult cuEventCreate(CUevent* e, unsigned int) {
  static uintptr_t next = 0x3000;
  *e = reinterpret_cast<CUevent>(next += 16);
  EMU_OK;
}
CUresult cuEventRecord(CUevent, CUstream) { EMU_OK; }
CUresult cuEventSynchronize(CUevent) { EMU_OK; }
CUresult cuEventDestroy_v2(CUevent) { EMU_OK; }

CUresult cuModuleLoadData(CUmodule* m, const void* image) {
  std::lock_guard<std::recursive_mutex> lock(g_mu);
  std::string path;
  if (std::memcmp(image, "GDVEMU1", 8) == 0) {
    path = static_cast<const char*>(image) + 8;
  } else {
    const char* s = std::getenv("GDV_EMU_STATIC_LIB");
    if (s == nullptr) return CUDA_ERROR_INVALID_IMAGE;
    path = s;
  }
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (h == nullptr) {
    std::fprintf(stderr, "gdv_emu: dlopen(%s): %s\n", path.c_str(), dlerror());
    return CUDA_ERROR_INVALID_IMAGE;
  }
  Module* mod = new Module();
  mod->handle = h;
  *m = reinterpret_cast<CUmodule>(mod);
  EMU_OK;
}
CUresult cuModuleUnload(CUmodule m) {
  delete reinterpret_cast<Module*>(m);  // the .so stays mapped: kernels keep function-local statics
  EMU_OK;
}
CUresult cuModuleGetFunction(CUfunction* f, CUmodule m, const char* name) {
  Module* mod = reinterpret_cast<Module*>(m);
  const std::string tramp = std::string(name) + "__emu";
  void* p = dlsym(mod->handle, tramp.c_str());
  if (p == nullptr) return CUDA_ERROR_NOT_FOUND;
  Function* fn = new Function();
  fn->entry = reinterpret_cast<void (*)(void**)>(p);
  fn->name = name;
  *f = reinterpret_cast<CUfunction>(fn);
  EMU_OK;
}
CUresult cuFuncGetAttribute(int* v, CUfunction_attribute a, CUfunction) {
  switch (a) {
    case CU_FUNC_ATTRIBUTE_NUM_REGS: *v = 32; break;
    case CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK: *v = 1024; break;
    default: *v = 0; break;
  }
  EMU_OK;
}
CUresult cuFuncSetAttribute(CUfunction f, CUfunction_attribute a, int v) {
  if (a == CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES) {
    if (v > 227 * 1024) return CUDA_ERROR_INVALID_VALUE;
    reinterpret_cast<Function*>(f)->max_dyn_smem = v;
  }
  EMU_OK;
}
CUresult cuOccupancyMaxActiveBlocksPerMultiprocessor(int* n, CUfunction, int block, size_t smem) {
  int by_threads = 2048 / (block > 0 ? block : 1);
  int by_smem = smem > 0 ? static_cast<int>((227 * 1024) / smem) : 32;
  *n = by_threads < by_smem ? by_threads : by_smem;
  if (*n > 32) *n = 32;
  EMU_OK;
}
CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                        unsigned bz, unsigned smem, CUstream, void** params, void**) {
  std::lock_guard<std::recursive_mutex> lock(g_mu);
  Function* fn = reinterpret_cast<Function*>(f);
  if (bx * by * bz == 0 || bx * by * bz > 1024 || gx * gy * gz == 0) return CUDA_ERROR_INVALID_VALUE;
  if (static_cast<int>(smem) > fn->max_dyn_smem) return CUDA_ERROR_INVALID_VALUE;
  void* dyn = nullptr;
  if (posix_memalign(&dyn, 1024, smem + 1024) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
  const gdv_emu_uint3 bdim{bx, by, bz}, gdim{gx, gy, gz};
  bool ok = true;
  for (unsigned z = 0; z < gz && ok; ++z)
    for (unsigned y = 0; y < gy && ok; ++y)
      for (unsigned x = 0; x < gx && ok; ++x) {
        std::memset(dyn, 0xA5, smem + 1024);  // shared memory starts undefined
        ok = RunCta(fn, gdv_emu_uint3{x, y, z}, bdim, gdim, dyn, params);
      }
  std::free(dyn);
  if (ok && !GuardsIntact(("after kernel " + fn->name).c_str())) return CUDA_ERROR_ILLEGAL_ADDRESS;
  return ok ? CUDA_SUCCESS : CUDA_ERROR_LAUNCH_TIMEOUT;
}
CUresult cuGetErrorString(CUresult r, const char** s) {
  *s = r == CUDA_SUCCESS ? "no error" : "emulated driver error";
  EMU_OK;
}
CUresult cuPointerGetAttribute(void*, CUpointer_attribute, CUdeviceptr) { return CUDA_ERROR_INVALID_VALUE; }
CUresult cuCtxEnablePeerAccess(CUcontext, unsigned int) { EMU_OK; }
CUresult cuDeviceCanAccessPeer(int* can, CUdevice, CUdevice) { *can = 0; EMU_OK; }
CUresult cuIpcGetMemHandle(CUipcMemHandle*, CUdeviceptr) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuIpcOpenMemHandle_v2(CUdeviceptr*, CUipcMemHandle, unsigned int) { return CUDA_ERROR_NOT_SUPPORTED; }
CUresult cuIpcCloseMemHandle(CUdeviceptr) { return CUDA_ERROR_NOT_SUPPORTED; }
}
