// lineitem.h — CPU restatement of the synthetic TPC-H lineitem column generator
// (gandiva_b200/csrc/device/static_kernels.cu: gdv_gen_lineitem).  TEST INFRASTRUCTURE:
// used by tests (generator parity) and by bench.py's cpu_baseline / --impl reference legs to
// build host-resident inputs.  Column kinds and value ranges: SURVEY.md §8(d), DESIGN.md.
#pragma once
#include <cstdint>
#include <cstring>

static inline uint64_t gdv_li_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline uint64_t gdv_li_rng(uint64_t seed, uint32_t col, int64_t row) {
  return gdv_li_mix64(gdv_li_mix64(seed ^ ((uint64_t)(col + 1u) * 0xD6E8FEB86659FD93ull)) +
                      (uint64_t)row * 0x9E3779B97F4A7C15ull);
}
static inline uint32_t gdv_li_rng_col(int kind) {
  switch (kind) {
    case 0: return 0;
    case 1: case 5: return 5;
    case 2: case 3: return 2;
    case 4: case 7: return 4;
    case 6: case 8: return 6;
    default: return (uint32_t)kind;
  }
}
// Fills rows [b, e) of the output (indexed from 0; the table row is first_row + i).
// validity (may be null) is an LSB-first bitmap; [b, e) must not share bytes across threads.
static inline void gdv_lineitem_fill(int kind, uint64_t seed, int64_t first_row, int64_t b,
                                     int64_t e, void* values, uint8_t* validity,
                                     int null_permille) {
  const uint32_t col = gdv_li_rng_col(kind);
  for (int64_t i = b; i < e; ++i) {
    const uint64_t r = gdv_li_rng(seed, col, first_row + i);
    switch (kind) {
      case 0: ((int32_t*)values)[i] = (int32_t)(8035 + (int64_t)(r % 2527ull)); break;
      case 1: ((double*)values)[i] = (double)(r % 11ull) / 100.0; break;
      case 2: ((double*)values)[i] = (double)(1ull + r % 50ull); break;
      case 3: ((int64_t*)values)[i] = (int64_t)(1ull + r % 50ull); break;
      case 4: { unsigned __int128 v = 90000ull + r % 10410000ull; std::memcpy((uint8_t*)values + 16 * i, &v, 16); break; }
      case 5: { unsigned __int128 v = r % 11ull; std::memcpy((uint8_t*)values + 16 * i, &v, 16); break; }
      case 6: { unsigned __int128 v = r % 9ull; std::memcpy((uint8_t*)values + 16 * i, &v, 16); break; }
      case 7: ((double*)values)[i] = (double)(90000ull + r % 10410000ull) / 100.0; break;
      case 8: ((double*)values)[i] = (double)(r % 9ull) / 100.0; break;
      default: ((int32_t*)values)[i] = (int32_t)((int64_t)(r % 2147483648ull) - 1073741824ll); break;
    }
    if (validity != nullptr) {
      const bool ok = (gdv_li_mix64(r ^ 0xA5A5A5A55A5A5A5Aull) % 1000ull) >= (uint64_t)null_permille;
      if (ok) validity[i >> 3] |= (uint8_t)(1u << (i & 7));
      else validity[i >> 3] &= (uint8_t)~(1u << (i & 7));
    }
  }
}
