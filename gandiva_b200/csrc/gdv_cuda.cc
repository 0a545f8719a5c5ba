#include "gdv_cuda.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace gdv {

namespace {

template <typename F>
bool Sym(void* h, const char* name, F* out, std::string* err) {
  void* p = dlsym(h, name);
  if (p == nullptr) {
    *err = std::string("missing symbol ") + name;
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}

DriverApi g_drv;
NvrtcApi g_nvrtc;
std::once_flag g_drv_once, g_nvrtc_once;

void LoadDriver() {
  DriverApi& d = g_drv;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    d.load_error = std::string("cannot load libcuda.so.1: ") + dlerror();
    return;
  }
  std::string e;
  bool ok = Sym(h, "cuInit", &d.Init, &e) && Sym(h, "cuDeviceGetCount", &d.DeviceGetCount, &e) &&
            Sym(h, "cuDeviceGet", &d.DeviceGet, &e) &&
            Sym(h, "cuDeviceGetAttribute", &d.DeviceGetAttribute, &e) &&
            Sym(h, "cuDevicePrimaryCtxRetain", &d.DevicePrimaryCtxRetain, &e) &&
            Sym(h, "cuCtxSetCurrent", &d.CtxSetCurrent, &e) &&
            Sym(h, "cuCtxGetCurrent", &d.CtxGetCurrent, &e) &&
            Sym(h, "cuMemAlloc_v2", &d.MemAlloc, &e) && Sym(h, "cuMemFree_v2", &d.MemFree, &e) &&
            Sym(h, "cuMemHostAlloc", &d.MemHostAlloc, &e) &&
            Sym(h, "cuMemFreeHost", &d.MemFreeHost, &e) &&
            Sym(h, "cuMemcpyHtoDAsync_v2", &d.MemcpyHtoDAsync, &e) &&
            Sym(h, "cuMemcpyDtoHAsync_v2", &d.MemcpyDtoHAsync, &e) &&
            Sym(h, "cuMemcpyDtoDAsync_v2", &d.MemcpyDtoDAsync, &e) &&
            Sym(h, "cuMemsetD8Async", &d.MemsetD8Async, &e) &&
            Sym(h, "cuStreamCreate", &d.StreamCreate, &e) &&
            Sym(h, "cuStreamSynchronize", &d.StreamSynchronize, &e) &&
            Sym(h, "cuStreamDestroy_v2", &d.StreamDestroy, &e) &&
            Sym(h, "cuStreamWaitEvent", &d.StreamWaitEvent, &e) &&
            Sym(h, "cuEventCreate", &d.EventCreate, &e) &&
            Sym(h, "cuEventRecord", &d.EventRecord, &e) &&
            Sym(h, "cuEventSynchronize", &d.EventSynchronize, &e) &&
            Sym(h, "cuEventDestroy_v2", &d.EventDestroy, &e) &&
            Sym(h, "cuModuleLoadData", &d.ModuleLoadData, &e) &&
            Sym(h, "cuModuleUnload", &d.ModuleUnload, &e) &&
            Sym(h, "cuModuleGetFunction", &d.ModuleGetFunction, &e) &&
            Sym(h, "cuFuncGetAttribute", &d.FuncGetAttribute, &e) &&
            Sym(h, "cuFuncSetAttribute", &d.FuncSetAttribute, &e) &&
            Sym(h, "cuOccupancyMaxActiveBlocksPerMultiprocessor",
                &d.OccupancyMaxActiveBlocksPerMultiprocessor, &e) &&
            Sym(h, "cuLaunchKernel", &d.LaunchKernel, &e) &&
            Sym(h, "cuGetErrorString", &d.GetErrorString, &e) &&
            Sym(h, "cuPointerGetAttribute", &d.PointerGetAttribute, &e) &&
            Sym(h, "cuCtxEnablePeerAccess", &d.CtxEnablePeerAccess, &e) &&
            Sym(h, "cuDeviceCanAccessPeer", &d.DeviceCanAccessPeer, &e) &&
            Sym(h, "cuIpcGetMemHandle", &d.IpcGetMemHandle, &e) &&
            Sym(h, "cuIpcOpenMemHandle_v2", &d.IpcOpenMemHandle, &e) &&
            Sym(h, "cuIpcCloseMemHandle", &d.IpcCloseMemHandle, &e) &&
            Sym(h, "cuMemGetAddressRange_v2", &d.MemGetAddressRange, &e);
  if (!ok) {
    d.load_error = "libcuda.so.1: " + e;
    return;
  }
  CUresult r = d.Init(0);
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    d.GetErrorString(r, &s);
    d.load_error = std::string("cuInit failed: ") + (s ? s : "unknown");
    return;
  }
  d.loaded = true;
}

void LoadNvrtc() {
  NvrtcApi& n = g_nvrtc;
  void* h = nullptr;
  std::string tried;
  const char* env = std::getenv("GDV_NVRTC_PATH");
  const char* candidates[] = {env, "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12",
                              "libnvrtc.so"};
  for (const char* c : candidates) {
    if (c == nullptr) continue;
    h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
    if (h != nullptr) break;
    tried += std::string(c) + ": " + dlerror() + "; ";
  }
  if (h == nullptr) {
    n.load_error = "cannot load NVRTC (" + tried + ")";
    return;
  }
  std::string e;
  bool ok = Sym(h, "nvrtcCreateProgram", &n.CreateProgram, &e) &&
            Sym(h, "nvrtcCompileProgram", &n.CompileProgram, &e) &&
            Sym(h, "nvrtcGetProgramLogSize", &n.GetProgramLogSize, &e) &&
            Sym(h, "nvrtcGetProgramLog", &n.GetProgramLog, &e) &&
            Sym(h, "nvrtcGetCUBINSize", &n.GetCUBINSize, &e) &&
            Sym(h, "nvrtcGetCUBIN", &n.GetCUBIN, &e) &&
            Sym(h, "nvrtcGetPTXSize", &n.GetPTXSize, &e) && Sym(h, "nvrtcGetPTX", &n.GetPTX, &e) &&
            Sym(h, "nvrtcDestroyProgram", &n.DestroyProgram, &e) &&
            Sym(h, "nvrtcGetErrorString", &n.GetErrorString, &e);
  if (!ok) {
    n.load_error = "NVRTC: " + e;
    return;
  }
  n.loaded = true;
}

}  // namespace

const DriverApi& Driver() {
  std::call_once(g_drv_once, LoadDriver);
  return g_drv;
}

const NvrtcApi& Nvrtc() {
  std::call_once(g_nvrtc_once, LoadNvrtc);
  return g_nvrtc;
}

Status CuCheck(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return Status::OK();
  const char* s = nullptr;
  if (g_drv.loaded) g_drv.GetErrorString(r, &s);
  return Status::Make(GDV_CUDA_ERROR, std::string(what) + ": CUDA error " +
                                          std::to_string(static_cast<int>(r)) + " (" +
                                          (s ? s : "?") + ")");
}

Status CompileToCubin(const std::string& source, const std::string& arch, bool optimize,
                      bool want_ptx, std::vector<char>* cubin, std::string* ptx,
                      std::string* log) {
  const NvrtcApi& n = Nvrtc();
  if (!n.loaded) return Status::Make(GDV_CUDA_ERROR, n.load_error);
  nvrtcProgram prog;
  const char* hdr_src[] = {gdv_device_lib_text};
  const char* hdr_name[] = {"gdv_device_lib.cuh"};
  nvrtcResult r = n.CreateProgram(&prog, source.c_str(), "gdv_fused.cu", 1, hdr_src, hdr_name);
  if (r != NVRTC_SUCCESS)
    return Status::Make(GDV_CODEGEN_ERROR,
                        std::string("nvrtcCreateProgram: ") + n.GetErrorString(r));
  const std::string arch_opt = "--gpu-architecture=" + arch;
  std::vector<const char*> opts = {arch_opt.c_str(), "--std=c++17", "--fmad=false",
                                   "--device-int128", "-lineinfo"};
  if (!optimize) opts.push_back("--ptxas-options=-O0");
  r = n.CompileProgram(prog, static_cast<int>(opts.size()), opts.data());
  size_t log_size = 0;
  n.GetProgramLogSize(prog, &log_size);
  std::string lg(log_size, '\0');
  if (log_size > 0) n.GetProgramLog(prog, &lg[0]);
  if (log != nullptr) *log = lg;
  if (r != NVRTC_SUCCESS) {
    n.DestroyProgram(&prog);
    return Status::Make(GDV_CODEGEN_ERROR, std::string("NVRTC compilation failed (") +
                                               n.GetErrorString(r) + "):\n" + lg);
  }
  size_t sz = 0;
  n.GetCUBINSize(prog, &sz);
  cubin->resize(sz);
  n.GetCUBIN(prog, cubin->data());
  if (want_ptx && ptx != nullptr) {
    size_t psz = 0;
    if (n.GetPTXSize(prog, &psz) == NVRTC_SUCCESS && psz > 0) {
      ptx->resize(psz);
      n.GetPTX(prog, &(*ptx)[0]);
    }
  }
  n.DestroyProgram(&prog);
  if (sz == 0) return Status::Make(GDV_CODEGEN_ERROR, "NVRTC produced an empty cubin");
  return Status::OK();
}

}  // namespace gdv
