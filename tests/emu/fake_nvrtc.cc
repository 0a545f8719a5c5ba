// fake_nvrtc.cc — TEST INFRASTRUCTURE: stands in for libnvrtc.so.12 under the functional SIMT
// simulator (tests/emu/README.md).  "Compiling" a program means compiling the SAME generated
// CUDA source for the host with g++ against tests/emu/gdv_emu.h (-DGDV_HOST_EMU swaps the inline
// PTX primitives of the device library for C++ restatements); the "cubin" handed back is a blob
// that names the resulting shared object, which tests/emu/fake_cuda.cc dlopens.
//
// Source-level rewrites (the only ones): `extern __shared__ uint4 gdv_smem[];` becomes a pointer
// to the CTA's dynamic shared memory, and every `extern "C" __global__ void NAME(params)` gets a
// trampoline `NAME__emu(void** params)` because a host caller cannot pass a by-value struct
// through a type-erased launch.
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Program {
  std::string source;
  std::vector<std::pair<std::string, std::string>> headers;
  std::string log;
  std::vector<char> blob;
};

unsigned long long Fnv(const std::string& s, unsigned long long h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

std::string EmuDir() {
  const char* e = std::getenv("GDV_EMU_DIR");  // tests/emu
  return e ? e : ".";
}
std::string CacheDir() {
  const char* e = std::getenv("GDV_EMU_CACHE");
  std::string d = e ? e : "/tmp/gdv_emu_cache";
  mkdir(d.c_str(), 0755);
  return d;
}

std::string ReadFile(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// `const __grid_constant__ gdv_args A` -> `gdv_args`; `u64* tiles` -> `u64*`
std::string ParamType(std::string p) {
  p = std::regex_replace(p, std::regex("__grid_constant__"), "");
  p = std::regex_replace(p, std::regex("__restrict__"), "");
  while (!p.empty() && std::isspace(static_cast<unsigned char>(p.back()))) p.pop_back();
  size_t e = p.size();
  while (e > 0 && (std::isalnum(static_cast<unsigned char>(p[e - 1])) || p[e - 1] == '_')) --e;
  return p.substr(0, e);
}

std::string Rewrite(const std::string& src, std::vector<std::string>* kernels) {
  std::string s = std::regex_replace(src, std::regex(R"(extern\s+__shared__\s+uint4\s+gdv_smem\[\];)"),
                                     "uint4* const gdv_smem = reinterpret_cast<uint4*>(gdv_emu_cur->dyn_smem);");
  std::string tramp;
  std::regex sig(R"re(extern\s+"C"\s+__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(([^)]*)\))re");
  for (std::sregex_iterator it(s.begin(), s.end(), sig), end; it != end; ++it) {
    const std::string name = (*it)[1], params = (*it)[2];
    kernels->push_back(name);
    std::vector<std::string> types;
    std::stringstream ps(params);
    std::string one;
    while (std::getline(ps, one, ',')) types.push_back(ParamType(one));
    tramp += "extern \"C\" void " + name + "__emu(void** p) {\n  " + name + "(";
    for (size_t i = 0; i < types.size(); ++i) {
      if (i) tramp += ", ";
      // strip a top-level const of a by-value parameter so the cast names an object type
      std::string t = std::regex_replace(types[i], std::regex(R"(^\s*const\s+(\w+)\s*$)"), "$1");
      tramp += "*reinterpret_cast<" + t + "*>(p[" + std::to_string(i) + "])";
    }
    tramp += ");\n}\n";
  }
  return s + "\n" + tramp;
}

}  // namespace

extern "C" {

nvrtcResult nvrtcCreateProgram(nvrtcProgram* prog, const char* src, const char*, int nh,
                               const char* const* headers, const char* const* names) {
  Program* p = new Program();
  p->source = src;
  for (int i = 0; i < nh; ++i) p->headers.emplace_back(names[i], headers[i]);
  *prog = reinterpret_cast<nvrtcProgram>(p);
  return NVRTC_SUCCESS;
}

nvrtcResult nvrtcCompileProgram(nvrtcProgram prog, int, const char* const*) {
  Program* p = reinterpret_cast<Program*>(prog);
  const std::string emu_h = ReadFile(EmuDir() + "/gdv_emu.h");
  if (emu_h.empty()) {
    p->log = "fake nvrtc: cannot read " + EmuDir() + "/gdv_emu.h (GDV_EMU_DIR)";
    return NVRTC_ERROR_COMPILATION;
  }
  std::vector<std::string> kernels;
  const std::string body = Rewrite(p->source, &kernels);
  unsigned long long h = Fnv(emu_h, Fnv(body));
  for (auto& hd : p->headers) h = Fnv(hd.second, h);
  char hex[32];
  std::snprintf(hex, sizeof(hex), "%016llx", h);
  const std::string dir = CacheDir();
  const std::string so = dir + "/k_" + hex + ".so";
  if (access(so.c_str(), R_OK) != 0) {
    const std::string inc = dir + "/inc_" + hex;
    mkdir(inc.c_str(), 0755);
    for (auto& hd : p->headers) std::ofstream(inc + "/" + hd.first, std::ios::binary) << hd.second;
    const std::string cc = dir + "/k_" + hex + ".cc";
    std::ofstream(cc, std::ios::binary) << body;
    const std::string tmp = so + ".tmp" + std::to_string(getpid());
    const std::string logf = dir + "/k_" + hex + ".log";
    const char* opt = std::getenv("GDV_EMU_CXXFLAGS");
    const std::string cmd = std::string("g++ -std=c++17 -x c++ ") + (opt ? opt : "-O1") +
                            " -fPIC -shared -ffp-contract=off -fno-strict-aliasing -w -DGDV_HOST_EMU -include " +
                            EmuDir() + "/gdv_emu.h -I" + inc + " " + cc + " -o " + tmp + " > " + logf + " 2>&1";
    const int rc = std::system(cmd.c_str());
    p->log = ReadFile(logf);
    if (rc != 0) {
      p->log = "fake nvrtc: " + cmd + "\n" + p->log;
      return NVRTC_ERROR_COMPILATION;
    }
    std::rename(tmp.c_str(), so.c_str());
  }
  p->blob.assign(8, '\0');
  std::memcpy(p->blob.data(), "GDVEMU1", 8);
  p->blob.insert(p->blob.end(), so.begin(), so.end());
  p->blob.push_back('\0');
  for (auto& k : kernels) {
    p->blob.insert(p->blob.end(), k.begin(), k.end());
    p->blob.push_back('\0');
  }
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram prog, size_t* n) {
  *n = reinterpret_cast<Program*>(prog)->log.size() + 1;
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetProgramLog(nvrtcProgram prog, char* out) {
  Program* p = reinterpret_cast<Program*>(prog);
  std::memcpy(out, p->log.c_str(), p->log.size() + 1);
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram prog, size_t* n) {
  *n = reinterpret_cast<Program*>(prog)->blob.size();
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetCUBIN(nvrtcProgram prog, char* out) {
  Program* p = reinterpret_cast<Program*>(prog);
  std::memcpy(out, p->blob.data(), p->blob.size());
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetPTXSize(nvrtcProgram, size_t* n) {
  *n = 0;
  return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetPTX(nvrtcProgram, char*) { return NVRTC_SUCCESS; }
nvrtcResult nvrtcDestroyProgram(nvrtcProgram* prog) {
  delete reinterpret_cast<Program*>(*prog);
  *prog = nullptr;
  return NVRTC_SUCCESS;
}
const char* nvrtcGetErrorString(nvrtcResult r) { return r == NVRTC_SUCCESS ? "NVRTC_SUCCESS" : "fake nvrtc error"; }
}
