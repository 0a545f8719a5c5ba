// static_kernels.cu — kernels that do not depend on the expression: the synthetic TPC-H
// lineitem column generator used by bench.py / tests (SURVEY.md §8d concretisation) and an
// L2 flush.  Compiled by nvcc to an sm_100a cubin at build time and embedded in
// libgandiva_b200.so; loaded through the driver API.  Including the device library here
// also makes nvcc type-check every function in it (the build-time gate for code that
// NVRTC otherwise only sees at Make()).
#include "gdv_device_lib.cuh"

// Counter-based generator: value = mix(seed, rng column, row).  oracle/lineitem.h restates
// the same arithmetic on the CPU so host- and device-generated columns are identical.
__device__ __forceinline__ u64 gdv_mix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ u64 gdv_rng(u64 seed, u32 col, i64 row) {
  return gdv_mix64(gdv_mix64(seed ^ ((u64)(col + 1u) * 0xD6E8FEB86659FD93ull)) +
                   (u64)row * 0x9E3779B97F4A7C15ull);
}

// column kinds (rng column in brackets):
//  0 l_shipdate  date32  uniform 1992-01-01 .. 1998-12-01            [0]
//  1 l_discount  f64     {0.00 .. 0.10} step 0.01                    [5]
//  2 l_quantity  f64     {1 .. 50}                                   [2]
//  3 l_quantity  int64   {1 .. 50}                                   [2]
//  4 l_extendedprice decimal128(15,2)  cents in [90000, 10500000)     [4]
//  5 l_discount  decimal128(15,2)  cents {0 .. 10}                    [5]
//  6 l_tax       decimal128(15,2)  cents {0 .. 8}                     [6]
//  7 l_extendedprice f64 = cents / 100.0                             [4]
//  8 l_tax       f64 = cents / 100.0                                 [6]
//  9 int32 uniform in [-2^30, 2^30)                                   [9]
// 10 int32 uniform in [-2^30, 2^30)                                   [10]
// validity: null iff mix(rng ^ NULLSALT) % 1000 < null_permille
extern "C" __global__ void __launch_bounds__(256)
gdv_gen_lineitem(int kind, u64 seed, i64 first_row, i64 num_rows, void* values, u32* validity,
                 int null_permille) {
  const u32 lane = threadIdx.x & 31u;
  const i64 warp = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const i64 nwarps = (i64)gridDim.x * (blockDim.x >> 5);
  const i64 n_w = (num_rows + 31) / 32;
  for (i64 w = warp; w < n_w; w += nwarps) {
    const i64 i = w * 32 + lane;
    const bool in = i < num_rows;
    const i64 row = first_row + i;
    u32 col;
    switch (kind) {
      case 0: col = 0; break;
      case 1: case 5: col = 5; break;
      case 2: case 3: col = 2; break;
      case 4: case 7: col = 4; break;
      case 6: case 8: col = 6; break;
      default: col = (u32)kind; break;
    }
    const u64 r = gdv_rng(seed, col, row);
    if (in) {
      switch (kind) {
        case 0: reinterpret_cast<i32*>(values)[i] = (i32)(8035 + (i64)(r % 2527ull)); break;
        case 1: reinterpret_cast<f64*>(values)[i] = (f64)(r % 11ull) / 100.0; break;
        case 2: reinterpret_cast<f64*>(values)[i] = (f64)(1ull + r % 50ull); break;
        case 3: reinterpret_cast<i64*>(values)[i] = (i64)(1ull + r % 50ull); break;
        case 4: gdv_st<i128>(values, i, (i128)(90000ull + r % 10410000ull)); break;
        case 5: gdv_st<i128>(values, i, (i128)(r % 11ull)); break;
        case 6: gdv_st<i128>(values, i, (i128)(r % 9ull)); break;
        case 7: reinterpret_cast<f64*>(values)[i] = (f64)(90000ull + r % 10410000ull) / 100.0; break;
        case 8: reinterpret_cast<f64*>(values)[i] = (f64)(r % 9ull) / 100.0; break;
        default:
          reinterpret_cast<i32*>(values)[i] = (i32)((i64)(r % 2147483648ull) - 1073741824ll);
          break;
      }
    }
    if (validity != nullptr) {
      const bool ok = in && (gdv_mix64(r ^ 0xA5A5A5A55A5A5A5Aull) % 1000ull) >= (u64)null_permille;
      const u32 m = __ballot_sync(GDV_FULL, ok);
      if (lane == 0u && w * 32 < num_rows) validity[w] = m;
    }
  }
}

// Writes `words` 32-bit words: used to flush L2 between timed iterations.
extern "C" __global__ void __launch_bounds__(256) gdv_fill_u32(u32* p, i64 words, u32 v) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) p[i] = v;
}
