"""TEST INFRASTRUCTURE: functional SIMT simulator for the generated kernels (see README.md here).

`build()` compiles the two stand-in libraries (libcuda.so.1, libnvrtc.so.12) into tests/emu/lib and
the host build of device/static_kernels.cu; `env()` returns the environment a SUBPROCESS needs so
that the unmodified libgandiva_b200.so dlopens them instead of the CUDA driver / NVRTC.  Nothing
in gandiva_b200/ refers to this package."""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIBDIR = os.path.join(HERE, "lib")              # libcuda.so.1 stand-in: goes on LD_LIBRARY_PATH
NVRTC_DIR = os.path.join(HERE, "lib_nvrtc")     # libnvrtc.so.12 stand-in: named by GDV_NVRTC_PATH only (a
#                                                 real NVRTC must never find it on the loader path)
# compiled host builds of the generated kernels (hundreds of MB over a session): kept OUTSIDE the repo so that
# they never travel with a gpurun snapshot (the snapshot has a 512 MiB limit); GDV_EMU_CACHE overrides
CACHE = os.environ.get("GDV_EMU_CACHE_DIR", os.path.join("/tmp", "gdv_emu_cache_%d" % os.getuid()))
CUDA_INC = "/usr/local/cuda/include"
DEVICE_DIR = os.path.join(ROOT, "gandiva_b200", "csrc", "device")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _cxx(out: str, src: str, extra: list[str]) -> None:
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-I" + CUDA_INC, src, "-o", out] + extra
    subprocess.run(cmd, check=True)


def build() -> dict:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(NVRTC_DIR, exist_ok=True)
    os.makedirs(CACHE, exist_ok=True)
    cuda = os.path.join(LIBDIR, "libcuda.so.1")
    nvrtc = os.path.join(NVRTC_DIR, "libnvrtc.so.12")
    stale = os.path.join(LIBDIR, "libnvrtc.so.12")
    if os.path.exists(stale):
        os.remove(stale)
    if not _newer(cuda, os.path.join(HERE, "fake_cuda.cc")):
        _cxx(cuda, os.path.join(HERE, "fake_cuda.cc"), ["-ldl"])
    if not _newer(nvrtc, os.path.join(HERE, "fake_nvrtc.cc")):
        _cxx(nvrtc, os.path.join(HERE, "fake_nvrtc.cc"), [])
    return {"cuda": cuda, "nvrtc": nvrtc, "static": _build_static(nvrtc)}


def _build_static(nvrtc_path: str) -> str:
    """Host build of device/static_kernels.cu through the same path generated kernels take
    (fake nvrtcCompileProgram); returns the shared object the fake driver maps the embedded
    static-kernel cubin to."""
    os.environ["GDV_EMU_DIR"] = HERE
    os.environ["GDV_EMU_CACHE"] = CACHE
    lib = ctypes.CDLL(nvrtc_path)
    src = open(os.path.join(DEVICE_DIR, "static_kernels.cu"), "rb").read()
    hdr = open(os.path.join(DEVICE_DIR, "gdv_device_lib.cuh"), "rb").read()
    prog = ctypes.c_void_p()
    hdrs = (ctypes.c_char_p * 1)(hdr)
    names = (ctypes.c_char_p * 1)(b"gdv_device_lib.cuh")
    lib.nvrtcCreateProgram(ctypes.byref(prog), src, b"static_kernels.cu", 1, hdrs, names)
    rc = lib.nvrtcCompileProgram(prog, 0, None)
    if rc != 0:
        n = ctypes.c_size_t()
        lib.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        lib.nvrtcGetProgramLog(prog, buf)
        raise RuntimeError("host build of static_kernels.cu failed:\n" + buf.value.decode(errors="replace"))
    n = ctypes.c_size_t()
    lib.nvrtcGetCUBINSize(prog, ctypes.byref(n))
    blob = ctypes.create_string_buffer(n.value)
    lib.nvrtcGetCUBIN(prog, blob)
    return blob.raw[8:].split(b"\0", 1)[0].decode()


def env(sms: int = 4) -> dict:
    """Environment for a subprocess that should run the product against the simulator."""
    libs = build()
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = LIBDIR + os.pathsep + e.get("LD_LIBRARY_PATH", "")
    e["GDV_NVRTC_PATH"] = libs["nvrtc"]
    e["GDV_EMU_DIR"] = HERE
    e["GDV_EMU_CACHE"] = CACHE
    e["GDV_EMU_STATIC_LIB"] = libs["static"]
    e["GDV_EMU_SMS"] = str(sms)
    e["GDV_EMU"] = "1"  # tests/ read this to pick simulator-sized inputs
    # three CTAs resident at once (each on its own copy of the kernel's shared object), scheduled in
    # a pseudo-random order: look-back waits and ticket order are exercised, not just sequential tiles
    e.setdefault("GDV_EMU_CTAS", "3")
    e.setdefault("GDV_EMU_SCHED", "2")
    e.pop("GDV_CUBIN_CACHE_DIR", None)  # the simulator's "cubins" must never reach a real cache
    return e
