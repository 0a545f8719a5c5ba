// cpu_proxy.cc — "fused-cxx-proxy": hand-fused, compiled, multi-threaded CPU row loops for the five
// BASELINE.json workloads.  BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference
// legs and tests/test_cpu_proxy.py); nothing under gandiva_b200/ may load it.
//
// Why it exists: the reference's CPU path (Gandiva's LLVM JIT, /root/reference/README.md:19 ->
// Arrow cpp/src/gandiva) cannot be built here (no source, no LLVM).  What that JIT emits for one
// expression is a fused, compiled row loop over the Arrow buffers plus a bitmap pass
// (BitmapAccumulator; Filter: bitmap -> index list).  This file is
// that shape written by hand and compiled with g++ -O3 -march=native: AVX-512 compares straight
// into mask registers where the CPU has them, one pass over the inputs, validity ANDed 64 rows at
// a time, bitmap -> ascending indices in a second pass over the (160x smaller) bitmap.  It is a
// GENEROUS stand-in: the real JIT evaluates one row loop per output expression (an eight-output
// projector re-reads shared inputs) and calls out-of-line functions for decimals and strings.
// Every function is checked bit-for-bit against the scalar oracle (tests/test_cpu_proxy.py).
//
// Threading: a persistent pool of pinned threads (one per hardware thread); every call splits the
// batch into contiguous 64-row-aligned row ranges, one per thread — the way an executor runs one
// Gandiva evaluator per core over different RecordBatches.  proxy_generate() fills the columns
// with the same partition, so pages are first-touched by the thread that later reads them.
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__AVX512F__) && defined(__AVX512BW__) && defined(__AVX512VL__)
#include <immintrin.h>
#define PROXY_AVX512 1
#else
#define PROXY_AVX512 0
#endif

#include "lineitem.h"

namespace {

class Pool {
 public:
  Pool(int n, const cpu_set_t& allowed) : allowed_(allowed), n_(n) {
    for (int t = 0; t < n_; ++t) threads_.emplace_back([this, t] { Loop(t); });
  }
  int size() const { return n_; }
  // Runs fn(t) for t in [0, n) on the pinned workers and returns when all are done.  The calling
  // thread only waits: it is never pinned (threads it creates later would inherit the mask).
  void Run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> g(mu_);
    fn_ = &fn;
    pending_ = n_;
    ++epoch_;
    cv_.notify_all();
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  void Loop(int t) {
    int k = 0;
    for (int c = 0; c < CPU_SETSIZE; ++c) {
      if (!CPU_ISSET(c, &allowed_)) continue;
      if (k++ == t) {
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(c, &one);
        pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
        break;
      }
    }
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return epoch_ != seen; });
        seen = epoch_;
        fn = fn_;
      }
      (*fn)(t);
      {
        std::lock_guard<std::mutex> g(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  cpu_set_t allowed_;
  int n_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  uint64_t epoch_ = 0;
  int pending_ = 0;
};

Pool* g_pool = nullptr;
std::mutex g_pool_mu;

// One pool per process, sized by the first call (want <= 0: every CPU the process may run on).
Pool& ThePool(int want) {
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (g_pool == nullptr) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int n = sched_getaffinity(0, sizeof(allowed), &allowed) == 0 ? CPU_COUNT(&allowed) : 0;
    if (n <= 0) {
      n = (int)std::thread::hardware_concurrency();
      for (int c = 0; c < n; ++c) CPU_SET(c, &allowed);
    }
    if (want > 0 && want < n) n = want;
    g_pool = new Pool(n < 1 ? 1 : n, allowed);
  }
  return *g_pool;
}

// Row range [b, e) of thread t out of T over n rows, aligned to 64 rows.
inline void Range(int64_t n, int t, int T, int64_t* b, int64_t* e) {
  const int64_t words = (n + 63) / 64;
  const int64_t wb = words * t / T, we = words * (t + 1) / T;
  *b = wb * 64 < n ? wb * 64 : n;
  *e = we * 64 < n ? we * 64 : n;
}

// 64 validity bits of rows [r, r+64) of an LSB-first bitmap; r % 64 == 0; null bitmap = all valid.
inline uint64_t Valid64(const uint8_t* v, int64_t r, int64_t n) {
  if (v == nullptr) return ~0ull;
  uint64_t w = 0;
  const int64_t bytes = (n + 7) / 8 - r / 8;
  memcpy(&w, v + r / 8, bytes >= 8 ? 8 : (size_t)bytes);
  return w;
}
inline uint64_t TailMask(int64_t r, int64_t n) { return n - r >= 64 ? ~0ull : ((1ull << (n - r)) - 1ull); }

// ---- Q6: shipdate in [lo, hi) AND disc in [0.05, 0.07] AND qty < 24 -> 64-row truth word ------
inline uint64_t Q6Word(const int32_t* ship, const double* disc, const double* qty, int64_t r, int64_t n) {
  uint64_t m = 0;
#if PROXY_AVX512
  if (n - r >= 64) {
    const __m512i lo = _mm512_set1_epi32(8766), hi = _mm512_set1_epi32(9131);
    const __m512d d0 = _mm512_set1_pd(0.05), d1 = _mm512_set1_pd(0.07), q = _mm512_set1_pd(24.0);
    for (int k = 0; k < 4; ++k) {
      const __m512i s = _mm512_loadu_si512((const void*)(ship + r + 16 * k));
      const uint32_t ms = _mm512_cmp_epi32_mask(s, lo, _MM_CMPINT_NLT) & _mm512_cmp_epi32_mask(s, hi, _MM_CMPINT_LT);
      uint32_t md = 0;
      for (int h = 0; h < 2; ++h) {
        const __m512d d = _mm512_loadu_pd(disc + r + 16 * k + 8 * h);
        const __m512d qq = _mm512_loadu_pd(qty + r + 16 * k + 8 * h);
        const uint32_t mm = (uint32_t)(_mm512_cmp_pd_mask(d, d0, _CMP_GE_OQ) & _mm512_cmp_pd_mask(d, d1, _CMP_LE_OQ) &
                                       _mm512_cmp_pd_mask(qq, q, _CMP_LT_OQ));
        md |= mm << (8 * h);
      }
      m |= (uint64_t)(ms & md) << (16 * k);
    }
    return m;
  }
#endif
  const int64_t e = n - r < 64 ? n - r : 64;
  for (int64_t j = 0; j < e; ++j) {
    const bool t = ship[r + j] >= 8766 && ship[r + j] < 9131 && disc[r + j] >= 0.05 && disc[r + j] <= 0.07 &&
                   qty[r + j] < 24.0;
    m |= (uint64_t)t << j;
  }
  return m;
}

// bitmap words [wb, we) -> ascending row indices at out[pos...]
template <typename I>
inline int64_t BitsToIndices(const uint64_t* bits, int64_t wb, int64_t we, I* out, int64_t pos, uint64_t base) {
  for (int64_t w = wb; w < we; ++w) {
    uint64_t m = bits[w];
    while (m) {
      out[pos++] = (I)(base + (uint64_t)w * 64 + (uint64_t)__builtin_ctzll(m));
      m &= m - 1;
    }
  }
  return pos;
}

// Two passes: (1) every thread evaluates its row range into the shared truth bitmap and counts,
// (2) after a prefix sum of the counts every thread expands its words at its final offset.
template <typename WordFn>
int64_t FilterDriver(int64_t n, uint32_t* out, uint64_t* bits, int threads, WordFn word) {
  Pool& pool = ThePool(threads);
  const int T = pool.size();
  std::vector<int64_t> counts((size_t)T + 1, 0);
  pool.Run([&](int t) {
    int64_t b, e;
    Range(n, t, T, &b, &e);
    int64_t c = 0;
    for (int64_t r = b; r < e; r += 64) {
      const uint64_t m = word(r) & TailMask(r, n);
      bits[r / 64] = m;
      c += __builtin_popcountll(m);
    }
    counts[(size_t)t + 1] = c;
  });
  for (int t = 0; t < T; ++t) counts[(size_t)t + 1] += counts[(size_t)t];
  pool.Run([&](int t) {
    int64_t b, e;
    Range(n, t, T, &b, &e);
    BitsToIndices<uint32_t>(bits, b / 64, (e + 63) / 64, out, counts[(size_t)t], 0);
  });
  return counts[(size_t)T];
}

// ---- string filter: like(upper(substr(c, 1, 32)), '%SPECIAL%REQUESTS%') -------------------------
inline const uint8_t* FindFold(const uint8_t* p, const uint8_t* end, const char* key, int klen) {
  // case-insensitive (ASCII) search for an upper-case key
  for (; p + klen <= end; ++p) {
    if ((uint8_t)(*p & 0xDF) != (uint8_t)key[0] && *p != (uint8_t)key[0]) continue;
    int j = 1;
    for (; j < klen; ++j) {
      uint8_t c = p[j];
      if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
      if (c != (uint8_t)key[j]) break;
    }
    if (j == klen) {
      uint8_t c0 = p[0];
      if (c0 >= 'a' && c0 <= 'z') c0 = (uint8_t)(c0 - 32);
      if (c0 == (uint8_t)key[0]) return p;
    }
  }
  return nullptr;
}
inline bool CommentRow(const uint8_t* s, int64_t len) {
  // substr(c, 1, 32): the first 32 glyphs (a byte that is not a UTF-8 continuation starts a glyph)
  int64_t cut = len;
  if (len > 32) {
    int64_t glyphs = 0, i = 0;
    for (; i < len; ++i) {
      if ((s[i] & 0xC0) != 0x80) {
        if (glyphs == 32) break;
        ++glyphs;
      }
    }
    cut = i;
  }
  const uint8_t* end = s + cut;
  const uint8_t* a = FindFold(s, end, "SPECIAL", 7);
  if (a == nullptr) return false;
  return FindFold(a + 7, end, "REQUESTS", 8) != nullptr;
}

// ---- decimal128 helpers for Q1 -----------------------------------------------------------------
typedef __int128 i128;
inline i128 Pow10_38() {
  i128 p = 1;
  for (int i = 0; i < 38; ++i) p *= 10;
  return p;
}
inline i128 MulChecked(i128 a, i128 b, i128 lim) {
  i128 r;
  if (__builtin_mul_overflow(a, b, &r)) return 0;
  if (r >= lim || r <= -lim) return 0;   // needs more than 38 digits -> 0
  return r;
}

}  // namespace

extern "C" {

int proxy_threads(int want) { return ThePool(want).size(); }
int proxy_simd() { return PROXY_AVX512 ? 512 : 0; }

void* proxy_alloc(size_t bytes) {
  void* p = nullptr;
  if (posix_memalign(&p, 4096, bytes + 64) != 0) return nullptr;
  return p;
}
void proxy_free(void* p) { free(p); }

// Fills a lineitem column (oracle/lineitem.h generator) with the pool's row partition (first touch).
void proxy_generate(int kind, uint64_t seed, int64_t first_row, int64_t n, void* values, uint8_t* validity,
                    int null_permille, int threads) {
  Pool& pool = ThePool(threads);
  const int T = pool.size();
  pool.Run([&](int t) {
    int64_t b, e;
    Range(n, t, T, &b, &e);
    if (validity != nullptr && e > b) memset(validity + b / 8, 0, (size_t)((e + 7) / 8 - b / 8));
    gdv_lineitem_fill(kind, seed, first_row, b, e, values, validity, null_permille);
  });
}

// Copies src -> dst with the pool's row partition (first touch of dst by its reader).
void proxy_copy_rows(void* dst, const void* src, int64_t n, int width, int threads) {
  Pool& pool = ThePool(threads);
  const int T = pool.size();
  pool.Run([&](int t) {
    int64_t b, e;
    Range(n, t, T, &b, &e);
    memcpy((uint8_t*)dst + b * width, (const uint8_t*)src + b * width, (size_t)((e - b) * width));
  });
}

// Q6 Filter.  bits: scratch of ceil(n/64) words.  Returns the number of selected rows.
int64_t proxy_q6_filter(const int32_t* ship, const double* disc, const double* qty, const uint8_t* v_ship,
                        const uint8_t* v_disc, const uint8_t* v_qty, int64_t n, uint32_t* out, uint64_t* bits,
                        int threads) {
  return FilterDriver(n, out, bits, threads, [=](int64_t r) {
    return Q6Word(ship, disc, qty, r, n) & Valid64(v_ship, r, n) & Valid64(v_disc, r, n) & Valid64(v_qty, r, n);
  });
}

// The same on the calling thread alone: Gandiva evaluates one RecordBatch on one thread.
void proxy_add_i32_inline(const int32_t* a, const int32_t* b, const uint8_t* va, const uint8_t* vb, int64_t n,
                          int32_t* out, uint8_t* vout) {
  for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)((uint32_t)a[i] + (uint32_t)b[i]);
  if (vout != nullptr) {
    for (int64_t r = 0; r < n; r += 64) {
      const uint64_t w = Valid64(va, r, n) & Valid64(vb, r, n) & TailMask(r, n);
      const int64_t bytes = (n + 7) / 8 - r / 8;
      memcpy(vout + r / 8, &w, bytes >= 8 ? 8 : (size_t)bytes);
    }
  }
}

// Projector add(int32, int32): values + validity (AND of the inputs') in one pass.
void proxy_add_i32(const int32_t* a, const int32_t* b, const uint8_t* va, const uint8_t* vb, int64_t n, int32_t* out,
                   uint8_t* vout, int threads) {
  Pool& pool = ThePool(threads);
  const int T = pool.size();
  pool.Run([&](int t) {
    int64_t rb, re;
    Range(n, t, T, &rb, &re);
    for (int64_t i = rb; i < re; ++i) out[i] = (int32_t)((uint32_t)a[i] + (uint32_t)b[i]);
    if (vout != nullptr) {
      for (int64_t r = rb; r < re; r += 64) {
        const uint64_t w = Valid64(va, r, n) & Valid64(vb, r, n) & TailMask(r, n);
        const int64_t bytes = (n + 7) / 8 - r / 8;
        memcpy(vout + r / 8, &w, bytes >= 8 ? 8 : (size_t)bytes);
      }
    }
  });
}

// String Filter like(upper(substr(c,1,32)), '%SPECIAL%REQUESTS%') over an Arrow utf8 column.
int64_t proxy_comment_filter(const int32_t* offsets, const uint8_t* data, const uint8_t* validity, int64_t n,
                             uint32_t* out, uint64_t* bits, int threads) {
  return FilterDriver(n, out, bits, threads, [=](int64_t r) {
    const uint64_t v = Valid64(validity, r, n);
    const int64_t e = n - r < 64 ? n - r : 64;
    uint64_t m = 0;
    for (int64_t j = 0; j < e; ++j) {
      if (!((v >> j) & 1)) continue;
      const int32_t o0 = offsets[r + j], o1 = offsets[r + j + 1];
      m |= (uint64_t)CommentRow(data + o0, o1 - o0) << j;
    }
    return m;
  });
}

// Q1 eight-output Projector (tests/cases.py q1_outputs).  cols: 8 inputs in schema order
// (qty i64, ext/disc/tax decimal128(15,2), ext_f/disc_f/tax_f f64, shipdate date32), each with an
// optional validity bitmap; outs: 8 value buffers + 8 validity bitmaps.
void proxy_q1_project(const void* const* in, const uint8_t* const* vin, int64_t n, void* const* outv,
                      uint8_t* const* outb, int threads) {
  Pool& pool = ThePool(threads);
  const int T = pool.size();
  const i128 lim = Pow10_38();
  const int64_t* qty = (const int64_t*)in[0];
  const i128* ext = (const i128*)in[1];
  const i128* disc = (const i128*)in[2];
  const i128* tax = (const i128*)in[3];
  const double* extf = (const double*)in[4];
  const double* discf = (const double*)in[5];
  const double* taxf = (const double*)in[6];
  const int32_t* ship = (const int32_t*)in[7];
  pool.Run([&](int t) {
    int64_t rb, re;
    Range(n, t, T, &rb, &re);
    i128* d1 = (i128*)outv[0];
    i128* d2 = (i128*)outv[1];
    double* f1 = (double*)outv[2];
    double* f2 = (double*)outv[3];
    int64_t* q2 = (int64_t*)outv[4];
    double* c1 = (double*)outv[5];
    int64_t* c2 = (int64_t*)outv[6];
    int64_t* c3 = (int64_t*)outv[7];
    for (int64_t r = rb; r < re; r += 64) {
      const int64_t e = n - r < 64 ? n - r : 64;
      const uint64_t tm = TailMask(r, n);
      const uint64_t vq = Valid64(vin[0], r, n), ve = Valid64(vin[1], r, n), vd = Valid64(vin[2], r, n),
                     vt = Valid64(vin[3], r, n), vef = Valid64(vin[4], r, n), vdf = Valid64(vin[5], r, n),
                     vtf = Valid64(vin[6], r, n), vs = Valid64(vin[7], r, n);
      uint64_t taken = 0, m_ship = 0;   // CASE conditions that are valid and true
      for (int64_t j = 0; j < e; ++j) {
        const int64_t i = r + j;
        i128 xe, xd, xt;   // unaligned-safe loads (Arrow guarantees 8-byte alignment only)
        memcpy(&xe, ext + i, 16);
        memcpy(&xd, disc + i, 16);
        memcpy(&xt, tax + i, 16);
        const i128 p1 = MulChecked(xe, (i128)100 - xd, lim);
        const i128 p2 = MulChecked(p1, (i128)100 + xt, lim);
        memcpy(d1 + i, &p1, 16);
        memcpy(d2 + i, &p2, 16);
        const double g1 = extf[i] * (1.0 - discf[i]);
        f1[i] = g1;
        f2[i] = g1 * (1.0 + taxf[i]);
        q2[i] = (int64_t)((uint64_t)qty[i] + (uint64_t)qty[i]);
        const bool cd = ((vdf >> j) & 1) && discf[i] > 0.05;      // a null condition takes ELSE
        const bool cq = ((vq >> j) & 1) && qty[i] < 24;
        c1[i] = cd ? extf[i] : 0.0;
        c2[i] = cq ? 1 : 0;
        c3[i] = qty[i];
        taken |= (uint64_t)cd << j;
        m_ship |= (uint64_t)(ship[i] <= 10471) << j;
      }
      const uint64_t w[8] = {ve & vd, ve & vd & vt, vef & vdf, vef & vdf & vtf, vq,
                             (taken & vef) | ~taken, ~0ull, vs & m_ship & vq};
      const int64_t bytes = (n + 7) / 8 - r / 8;
      for (int k = 0; k < 8; ++k) {
        if (outb[k] == nullptr) continue;
        const uint64_t x = w[k] & tm;
        memcpy(outb[k] + r / 8, &x, bytes >= 8 ? 8 : (size_t)bytes);
      }
    }
  });
}

}  // extern "C"
