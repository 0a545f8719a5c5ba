// gdv_oracle.cc — CPU oracle: a scalar, row-at-a-time interpreter of the expression tree.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (gandiva_b200/) never links, imports or calls it.
//
// PARITY UNPINNED: /root/reference holds no source (README.md:19 only points
// at Apache Arrow cpp/src/gandiva), no libgandiva / LLVM exists in this image, so this is a
// restatement of the reference's *documented behaviour*: the API and known-answer vectors
// in the descendant's binding tests (site-packages/pyarrow/tests/test_gandiva.py:25-393,
// replayed in tests/test_golden.py) plus the semantics table in DESIGN.md, each row of
// which is cross-checked against pyarrow.compute where the two coincide
// (tests/test_oracle_vs_arrow.py).  Structure mirrors the reference's CPU path
// (SURVEY.md §3 B/C): for each expression one pass over the rows calling scalar functions
// (the "precompiled function library"), validity computed per row, then for filters a pass
// that turns the boolean result into an ascending index list.
//
// Pinned beyond those vectors by published standards: MurmurHash3 (reference answers), MD5 / SHA-1 /
// SHA-256 (RFC 1321, FIPS 180-4 known answers + hashlib), CRC-32 (zlib).  exp / log / log10 / cbrt
// are, on purpose, the same IEEE operation sequences as the device library (no libm on either
// side), measured against the host libm in tests/.
//
// Deliberately naive and independent of the GPU implementation: strings are materialised
// as std::string (the GPU uses lazy views), decimals use 32-bit-limb schoolbook arithmetic
// (the GPU uses 64-bit limbs), the tree is parsed from an s-expression emitted by the test
// harness rather than shared with the product's node classes.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "lineitem.h"

namespace {

typedef __int128 i128;
typedef unsigned __int128 u128;

enum TypeId {
  T_BOOL = 1, T_UINT8 = 2, T_INT8 = 3, T_UINT16 = 4, T_INT16 = 5, T_UINT32 = 6, T_INT32 = 7,
  T_UINT64 = 8, T_INT64 = 9, T_FLOAT = 11, T_DOUBLE = 12, T_STRING = 13, T_BINARY = 14,
  T_DATE32 = 16, T_DATE64 = 17, T_TIMESTAMP = 18, T_TIME32 = 19, T_TIME64 = 20, T_DECIMAL = 23
};

struct Type {
  int id = 0, precision = 0, scale = 0;
  bool is_signed_int() const {
    return id == T_INT8 || id == T_INT16 || id == T_INT32 || id == T_INT64 || id == T_DATE32 ||
           id == T_DATE64 || id == T_TIMESTAMP || id == T_TIME32 || id == T_TIME64;
  }
  bool is_unsigned_int() const {
    return id == T_UINT8 || id == T_UINT16 || id == T_UINT32 || id == T_UINT64;
  }
  bool is_string() const { return id == T_STRING || id == T_BINARY; }
  int width() const {
    switch (id) {
      case T_UINT8: case T_INT8: return 1;
      case T_UINT16: case T_INT16: return 2;
      case T_UINT32: case T_INT32: case T_FLOAT: case T_DATE32: case T_TIME32: return 4;
      case T_UINT64: case T_INT64: case T_DOUBLE: case T_DATE64: case T_TIMESTAMP: case T_TIME64:
        return 8;
      case T_DECIMAL: return 16;
      default: return 0;
    }
  }
  int bits() const { return width() * 8; }
};

// One value.  The member that matches the node's type is meaningful.
struct Val {
  bool ok = false;
  bool b = false;
  int64_t i = 0;
  uint64_t u = 0;
  float f = 0;
  double d = 0;
  i128 dec = 0;
  std::string s;
};

struct Column {  // same layout as gdv_column_t
  const void* validity;
  const void* values;
  const void* var_data;
  int64_t offset;
  int64_t var_data_size;
};

struct EvalCtx {
  const Column* cols;
  int error = 0;  // 1 = divide by zero, 4 = not an integer of the target type, 5 = not a date / timestamp
};

enum Kind { K_FIELD, K_LIT, K_FN, K_IF, K_AND, K_OR, K_IN };

struct Node {
  Kind kind;
  Type type;
  int col = -1;            // field
  Val lit;                 // literal
  std::string name;        // function
  std::vector<std::unique_ptr<Node>> kids;
  std::vector<int64_t> in_ints;
  std::vector<std::string> in_strs;
  std::vector<double> in_dbls;   // IN over float32 / float64
  // LIKE pattern tokens: 0..255 literal byte, 256 = '_', 257 = '%'
  std::vector<int> like;
  std::shared_ptr<struct OrcRe> regex;  // regexp_matches: parsed pattern
};

// ------------------------------------------------------------------------------------------
// s-expression parser
// ------------------------------------------------------------------------------------------
struct Parser {
  const char* p;
  std::string err;
  void ws() { while (*p == ' ' || *p == '\n' || *p == '\t') ++p; }
  std::string tok() {
    ws();
    std::string t;
    while (*p && *p != ' ' && *p != '(' && *p != ')' && *p != '\n') t.push_back(*p++);
    return t;
  }
  bool type(Type* t) {
    std::string s = tok();
    static const struct { const char* n; int id; } names[] = {
        {"bool", T_BOOL}, {"uint8", T_UINT8}, {"int8", T_INT8}, {"uint16", T_UINT16},
        {"int16", T_INT16}, {"uint32", T_UINT32}, {"int32", T_INT32}, {"uint64", T_UINT64},
        {"int64", T_INT64}, {"float32", T_FLOAT}, {"float64", T_DOUBLE}, {"utf8", T_STRING},
        {"binary", T_BINARY}, {"date32", T_DATE32}, {"date64", T_DATE64},
        {"timestamp", T_TIMESTAMP}, {"time32", T_TIME32}, {"time64", T_TIME64}};
    if (s.rfind("decimal128:", 0) == 0) {
      t->id = T_DECIMAL;
      if (std::sscanf(s.c_str(), "decimal128:%d:%d", &t->precision, &t->scale) != 2) {
        err = "bad decimal type " + s;
        return false;
      }
      return true;
    }
    for (const auto& n : names)
      if (s == n.n) {
        t->id = n.id;
        return true;
      }
    err = "unknown type '" + s + "'";
    return false;
  }
  static int hexv(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
  }
  // literal value token: "null", decimal integer, "x<hex bits>" for floats, "h<hex bytes>" strings
  bool value(const Type& t, Val* v) {
    std::string s = tok();
    if (s == "null") {
      v->ok = false;
      return true;
    }
    v->ok = true;
    if (t.is_string()) {
      if (s.empty() || s[0] != 'h') { err = "string literal must be h<hex>"; return false; }
      for (size_t k = 1; k + 1 < s.size(); k += 2)
        v->s.push_back(static_cast<char>(hexv(s[k]) * 16 + hexv(s[k + 1])));
      return true;
    }
    if (t.id == T_FLOAT || t.id == T_DOUBLE) {
      if (s.empty() || s[0] != 'x') { err = "float literal must be x<hexbits>"; return false; }
      uint64_t bits = std::strtoull(s.c_str() + 1, nullptr, 16);
      if (t.id == T_FLOAT) {
        uint32_t b32 = static_cast<uint32_t>(bits);
        std::memcpy(&v->f, &b32, 4);
      } else {
        std::memcpy(&v->d, &bits, 8);
      }
      return true;
    }
    // integers (possibly 128-bit) in decimal
    bool neg = false;
    size_t k = 0;
    if (!s.empty() && s[0] == '-') { neg = true; k = 1; }
    u128 m = 0;
    for (; k < s.size(); ++k) {
      if (s[k] < '0' || s[k] > '9') { err = "bad integer literal " + s; return false; }
      m = m * 10 + static_cast<unsigned>(s[k] - '0');
    }
    i128 val = neg ? static_cast<i128>(~m + 1) : static_cast<i128>(m);
    if (t.id == T_BOOL) v->b = val != 0;
    else if (t.id == T_DECIMAL) v->dec = val;
    else if (t.is_unsigned_int()) v->u = static_cast<uint64_t>(static_cast<u128>(val));
    else v->i = static_cast<int64_t>(val);
    return true;
  }
  std::unique_ptr<Node> node() {
    ws();
    if (*p != '(') { err = "expected '('"; return nullptr; }
    ++p;
    std::string head = tok();
    std::unique_ptr<Node> n(new Node());
    if (head == "field") {
      n->kind = K_FIELD;
      n->col = std::atoi(tok().c_str());
      if (!type(&n->type)) return nullptr;
    } else if (head == "lit") {
      n->kind = K_LIT;
      if (!type(&n->type) || !value(n->type, &n->lit)) return nullptr;
    } else if (head == "fn") {
      n->kind = K_FN;
      n->name = tok();
      if (!type(&n->type)) return nullptr;
      if (!kids(n.get())) return nullptr;
    } else if (head == "if") {
      n->kind = K_IF;
      if (!type(&n->type)) return nullptr;
      if (!kids(n.get())) return nullptr;
      if (n->kids.size() != 3) { err = "if needs 3 children"; return nullptr; }
    } else if (head == "and" || head == "or") {
      n->kind = head == "and" ? K_AND : K_OR;
      n->type.id = T_BOOL;
      if (!kids(n.get())) return nullptr;
    } else if (head == "in") {
      n->kind = K_IN;
      Type vt;
      if (!type(&vt)) return nullptr;
      n->type.id = T_BOOL;
      ws();
      std::unique_ptr<Node> c = node();
      if (!c) return nullptr;
      n->kids.push_back(std::move(c));
      ws();
      while (*p && *p != ')') {
        Val v;
        if (!value(vt, &v)) return nullptr;
        if (vt.is_string()) n->in_strs.push_back(v.s);
        else if (vt.id == T_FLOAT) n->in_dbls.push_back(static_cast<double>(v.f));
        else if (vt.id == T_DOUBLE) n->in_dbls.push_back(v.d);
        else n->in_ints.push_back(v.i);
        ws();
      }
    } else {
      err = "unknown head '" + head + "'";
      return nullptr;
    }
    ws();
    if (*p != ')') { err = "expected ')' after " + head; return nullptr; }
    ++p;
    return n;
  }
  bool kids(Node* n) {
    ws();
    while (*p == '(') {
      std::unique_ptr<Node> c = node();
      if (!c) return false;
      n->kids.push_back(std::move(c));
      ws();
    }
    return true;
  }
};

// ------------------------------------------------------------------------------------------
// helpers: Arrow buffers
// ------------------------------------------------------------------------------------------
inline bool GetBit(const void* bits, int64_t i) {
  return (static_cast<const uint8_t*>(bits)[i >> 3] >> (i & 7)) & 1;
}

void LoadField(const Node& n, const EvalCtx& cx, int64_t row, Val* out) {
  const Column& c = cx.cols[n.col];
  const int64_t r = row + c.offset;
  out->ok = c.validity == nullptr ? true : GetBit(c.validity, r);
  const uint8_t* v = static_cast<const uint8_t*>(c.values);
  switch (n.type.id) {
    case T_BOOL: out->b = GetBit(v, r); break;
    case T_INT8: out->i = reinterpret_cast<const int8_t*>(v)[r]; break;
    case T_INT16: out->i = reinterpret_cast<const int16_t*>(v)[r]; break;
    case T_INT32: case T_DATE32: case T_TIME32: out->i = reinterpret_cast<const int32_t*>(v)[r]; break;
    case T_INT64: case T_DATE64: case T_TIMESTAMP: case T_TIME64:
      out->i = reinterpret_cast<const int64_t*>(v)[r];
      break;
    case T_UINT8: out->u = v[r]; break;
    case T_UINT16: out->u = reinterpret_cast<const uint16_t*>(v)[r]; break;
    case T_UINT32: out->u = reinterpret_cast<const uint32_t*>(v)[r]; break;
    case T_UINT64: out->u = reinterpret_cast<const uint64_t*>(v)[r]; break;
    case T_FLOAT: out->f = reinterpret_cast<const float*>(v)[r]; break;
    case T_DOUBLE: out->d = reinterpret_cast<const double*>(v)[r]; break;
    case T_DECIMAL: std::memcpy(&out->dec, v + 16 * r, 16); break;
    case T_STRING: case T_BINARY: {
      const int32_t* off = reinterpret_cast<const int32_t*>(v);
      const char* data = static_cast<const char*>(c.var_data);
      out->s.assign(data + off[r], static_cast<size_t>(off[r + 1] - off[r]));
      break;
    }
    default: break;
  }
}

// ------------------------------------------------------------------------------------------
// scalar function library
// ------------------------------------------------------------------------------------------
// Truncate a wide integer to the node's integer type (two's-complement wrap).
int64_t WrapSigned(int64_t v, int bits) {
  if (bits >= 64) return v;
  const uint64_t mask = (uint64_t(1) << bits) - 1;
  uint64_t u = static_cast<uint64_t>(v) & mask;
  if (u >> (bits - 1)) u |= ~mask;
  return static_cast<int64_t>(u);
}
uint64_t WrapUnsigned(uint64_t v, int bits) {
  return bits >= 64 ? v : (v & ((uint64_t(1) << bits) - 1));
}

int64_t FloorDiv(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
// days -> ms with two's-complement wrap: month / year starts of dates at the very ends of the
// int64 range fall outside it; the result is defined (and equal to the kernel's), not UB
int64_t DaysToMs(int64_t days) { return static_cast<int64_t>(static_cast<uint64_t>(days) * 86400000ull); }

// Gregorian calendar from days since 1970-01-01 (independent formulation: walk via
// 400/100/4/1-year cycles instead of the era/doe closed form used on the GPU).
struct Ymd { int64_t y; int m, d, doy; };
bool IsLeap(int64_t y) { return (y % 4 == 0) && (y % 100 != 0 || y % 400 == 0); }
Ymd CivilFromDays(int64_t days) {
  // shift to 0001-01-01 based day number (proleptic): 1970-01-01 is day 719162 (0-based)
  int64_t n = days + 719162;
  int64_t cycles400 = FloorDiv(n, 146097);
  n -= cycles400 * 146097;
  int64_t c100 = std::min<int64_t>(n / 36524, 3);
  n -= c100 * 36524;
  int64_t c4 = n / 1461;
  n -= c4 * 1461;
  int64_t c1 = std::min<int64_t>(n / 365, 3);
  n -= c1 * 365;
  Ymd r;
  r.y = cycles400 * 400 + c100 * 100 + c4 * 4 + c1 + 1;
  r.doy = static_cast<int>(n) + 1;
  static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int m = 0;
  int64_t rem = n;
  while (true) {
    int len = mdays[m] + ((m == 1 && IsLeap(r.y)) ? 1 : 0);
    if (rem < len) break;
    rem -= len;
    ++m;
  }
  r.m = m + 1;
  r.d = static_cast<int>(rem) + 1;
  return r;
}

// ---- decimals: sign + magnitude with eight 32-bit limbs (256 bits) ----------------------
struct Big {
  uint32_t w[8];
  Big() { std::memset(w, 0, sizeof(w)); }
  static Big From(u128 v) {
    Big b;
    for (int k = 0; k < 4; ++k) b.w[k] = static_cast<uint32_t>(v >> (32 * k));
    return b;
  }
  bool IsZero() const {
    for (int k = 0; k < 8; ++k) if (w[k]) return false;
    return true;
  }
  void MulSmall(uint32_t m) {
    uint64_t carry = 0;
    for (int k = 0; k < 8; ++k) {
      uint64_t cur = static_cast<uint64_t>(w[k]) * m + carry;
      w[k] = static_cast<uint32_t>(cur);
      carry = cur >> 32;
    }
  }
  uint32_t DivSmall(uint32_t d) {
    uint64_t rem = 0;
    for (int k = 7; k >= 0; --k) {
      uint64_t cur = (rem << 32) | w[k];
      w[k] = static_cast<uint32_t>(cur / d);
      rem = cur % d;
    }
    return static_cast<uint32_t>(rem);
  }
  void Add(const Big& o) {
    uint64_t carry = 0;
    for (int k = 0; k < 8; ++k) {
      uint64_t cur = static_cast<uint64_t>(w[k]) + o.w[k] + carry;
      w[k] = static_cast<uint32_t>(cur);
      carry = cur >> 32;
    }
  }
  void Sub(const Big& o) {  // requires *this >= o
    int64_t borrow = 0;
    for (int k = 0; k < 8; ++k) {
      int64_t cur = static_cast<int64_t>(w[k]) - o.w[k] - borrow;
      borrow = cur < 0 ? 1 : 0;
      if (cur < 0) cur += (int64_t(1) << 32);
      w[k] = static_cast<uint32_t>(cur);
    }
  }
  int Cmp(const Big& o) const {
    for (int k = 7; k >= 0; --k)
      if (w[k] != o.w[k]) return w[k] < o.w[k] ? -1 : 1;
    return 0;
  }
  static Big Mul(const Big& a, const Big& b) {  // low 256 bits of the product
    Big r;
    for (int i = 0; i < 8; ++i) {
      uint64_t carry = 0;
      for (int j = 0; i + j < 8; ++j) {
        uint64_t cur = static_cast<uint64_t>(a.w[i]) * b.w[j] + r.w[i + j] + carry;
        r.w[i + j] = static_cast<uint32_t>(cur);
        carry = cur >> 32;
      }
    }
    return r;
  }
  void MulPow10(int e) { for (int k = 0; k < e; ++k) MulSmall(10); }
  // *this *= 10^e; false when a carry left the 256 bits
  bool MulPow10Checked(int e) {
    for (int k = 0; k < e; ++k) {
      uint64_t carry = 0;
      for (int j = 0; j < 8; ++j) {
        uint64_t cur = static_cast<uint64_t>(w[j]) * 10u + carry;
        w[j] = static_cast<uint32_t>(cur);
        carry = cur >> 32;
      }
      if (carry) return false;
    }
    return true;
  }
  bool Bit(int i) const { return (w[i >> 5] >> (i & 31)) & 1u; }
  // schoolbook binary long division: n = q * d + r  (d != 0, d < 2^255)
  static void DivMod(const Big& n, const Big& d, Big* q, Big* r) {
    Big quo, rem;
    for (int i = 255; i >= 0; --i) {
      uint32_t carry = n.Bit(i) ? 1u : 0u;
      for (int j = 0; j < 8; ++j) {
        const uint32_t top = rem.w[j] >> 31;
        rem.w[j] = (rem.w[j] << 1) | carry;
        carry = top;
      }
      if (rem.Cmp(d) >= 0) {
        rem.Sub(d);
        quo.w[i >> 5] |= 1u << (i & 31);
      }
    }
    *q = quo;
    *r = rem;
  }
  bool FitsDigits(int digits) const {
    Big lim = Big::From(1);
    lim.MulPow10(digits);
    return Cmp(lim) < 0;
  }
  u128 Low128() const {
    u128 v = 0;
    for (int k = 3; k >= 0; --k) v = (v << 32) | w[k];
    return v;
  }
};

// magnitude / 10^e rounded half away from zero: add 5*10^(e-1) then truncate.
Big DivPow10HalfUp(Big mag, int e) {
  if (e <= 0) return mag;
  Big half = Big::From(5);
  half.MulPow10(e - 1);
  mag.Add(half);
  for (int k = 0; k < e; ++k) mag.DivSmall(10);
  return mag;
}

struct SignedBig {
  bool neg;
  Big mag;
};
SignedBig ToSigned(i128 v) {
  SignedBig s;
  s.neg = v < 0;
  u128 m = s.neg ? (~static_cast<u128>(v) + 1) : static_cast<u128>(v);
  s.mag = Big::From(m);
  return s;
}
// result -> i128, 0 when it needs more than `digits` decimal digits
i128 FromSigned(const SignedBig& s, int digits) {
  if (!s.mag.FitsDigits(digits)) return 0;
  u128 m = s.mag.Low128();
  if (m == 0) return 0;
  return s.neg ? static_cast<i128>(~m + 1) : static_cast<i128>(m);
}

i128 DecimalAddSub(i128 x, int xs, i128 y, int ys, int os, bool subtract) {
  const int ms = std::max(xs, ys);
  SignedBig a = ToSigned(x), b = ToSigned(y);
  if (subtract) b.neg = !b.neg;
  a.mag.MulPow10(ms - xs);
  b.mag.MulPow10(ms - ys);
  SignedBig r;
  if (a.neg == b.neg) {
    r.neg = a.neg;
    r.mag = a.mag;
    r.mag.Add(b.mag);
  } else if (a.mag.Cmp(b.mag) >= 0) {
    r.neg = a.neg;
    r.mag = a.mag;
    r.mag.Sub(b.mag);
  } else {
    r.neg = b.neg;
    r.mag = b.mag;
    r.mag.Sub(a.mag);
  }
  if (os < ms) r.mag = DivPow10HalfUp(r.mag, ms - os);
  if (os > ms) {
    if (!r.mag.FitsDigits(38)) return 0;
    r.mag.MulPow10(os - ms);
  }
  return FromSigned(r, 38);
}

i128 DecimalMultiply(i128 x, int xs, i128 y, int ys, int os) {
  SignedBig a = ToSigned(x), b = ToSigned(y);
  SignedBig r;
  r.neg = a.neg != b.neg;
  r.mag = Big::Mul(a.mag, b.mag);
  const int delta = xs + ys - os;
  if (delta > 0) r.mag = DivPow10HalfUp(r.mag, std::min(delta, 38));
  if (delta < 0) {
    if (!r.mag.FitsDigits(38)) return 0;
    r.mag.MulPow10(-delta);
  }
  return FromSigned(r, 38);
}

int DecimalCompare(i128 x, int xs, i128 y, int ys) {
  const int ms = std::max(xs, ys);
  SignedBig a = ToSigned(x), b = ToSigned(y);
  a.mag.MulPow10(ms - xs);
  b.mag.MulPow10(ms - ys);
  if (a.mag.IsZero()) a.neg = false;
  if (b.mag.IsZero()) b.neg = false;
  if (a.neg != b.neg) return a.neg ? -1 : 1;
  int c = a.mag.Cmp(b.mag);
  return a.neg ? -c : c;
}

// Drops xs - rs digits under `mode` (0 half away from zero, 1 toward zero, 2 toward +inf, 3 toward
// -inf), then expresses the result at (op, os): digit by digit, with a sticky remainder.
i128 DecimalRoundTo(i128 x, int xs, int64_t rs64, int mode, int op, int os) {
  SignedBig a = ToSigned(x);
  const int rs = static_cast<int>(std::max<int64_t>(rs64, -39));
  int cur = xs;
  if (rs < xs) {
    const int d = xs - rs;
    bool sticky = false;
    uint32_t last = 0;
    for (int k = 0; k < d; ++k) {
      sticky = sticky || last != 0;
      last = a.mag.DivSmall(10);
    }
    bool up;
    if (mode == 0) up = last >= 5;
    else if (mode == 1) up = false;
    else if (mode == 2) up = !a.neg && (last != 0 || sticky);
    else up = a.neg && (last != 0 || sticky);
    if (up) a.mag.Add(Big::From(1));
    cur = rs;
  }
  if (os > cur) {
    if (!a.mag.MulPow10Checked(os - cur)) return 0;
  } else if (os < cur) {
    a.mag = DivPow10HalfUp(a.mag, std::min(cur - os, 38));
  }
  return FromSigned(a, op);
}

i128 DecimalRescale(i128 x, int xs, int op, int os) {
  SignedBig a = ToSigned(x);
  if (os > xs) a.mag.MulPow10(os - xs);
  if (os < xs) a.mag = DivPow10HalfUp(a.mag, xs - os);
  return FromSigned(a, op);
}

// x / y at scale os: |x| * 10^(os - xs + ys) / |y| rounded half away from zero; 0 on overflow.
// *div_zero is set when y == 0.
i128 DecimalDivide(i128 x, int xs, i128 y, int ys, int os, bool* div_zero) {
  if (y == 0) { *div_zero = true; return 0; }
  SignedBig a = ToSigned(x), b = ToSigned(y);
  const int delta = os - xs + ys;
  if (delta > 0 && !a.mag.MulPow10Checked(delta)) return 0;
  if (delta < 0 && !b.mag.MulPow10Checked(-delta)) return 0;
  Big q, r;
  Big::DivMod(a.mag, b.mag, &q, &r);
  Big r2 = r;
  r2.Add(r);
  if (r2.Cmp(b.mag) >= 0) q.Add(Big::From(1));
  SignedBig res;
  res.neg = a.neg != b.neg;
  res.mag = q;
  return FromSigned(res, 38);
}

i128 DecimalMod(i128 x, int xs, i128 y, int ys, int os, bool* div_zero) {
  if (y == 0) { *div_zero = true; return 0; }
  const int ms = std::max(xs, ys);
  SignedBig a = ToSigned(x), b = ToSigned(y);
  a.mag.MulPow10(ms - xs);
  b.mag.MulPow10(ms - ys);
  Big q, r;
  Big::DivMod(a.mag, b.mag, &q, &r);
  if (os < ms) r = DivPow10HalfUp(r, ms - os);
  if (os > ms) {
    if (!r.FitsDigits(38)) return 0;
    r.MulPow10(os - ms);
  }
  SignedBig res;
  res.neg = a.neg;
  res.mag = r;
  return FromSigned(res, 38);
}

// double -> decimal(op, os); the IEEE operation sequence is the one DESIGN.md states.
i128 DecimalFromDouble(double v, int op, int os) {
  double p = 1.0;
  for (int k = 0; k < os; ++k) p = p * 10.0;
  const double s = v * p;
  const double a = std::fabs(s);
  if (!(a < 1.0e38)) return 0;
  double t = std::floor(a);
  if (a - t >= 0.5) t = t + 1.0;
  // integer-valued double -> 128-bit integer, by peeling 32-bit digits from the top
  u128 m = 0;
  {
    int e = 0;
    const double fr = std::frexp(t, &e);  // t = fr * 2^e, fr in [0.5, 1)
    double f = fr;
    int bits_left = e;
    while (bits_left > 0) {
      const int take = bits_left >= 32 ? 32 : bits_left;
      f = std::ldexp(f, take);
      const double ip = std::floor(f);
      m = (m << take) | static_cast<u128>(static_cast<uint64_t>(ip));
      f -= ip;
      bits_left -= take;
    }
  }
  Big lim = Big::From(1);
  lim.MulPow10(op);
  if (Big::From(m).Cmp(lim) >= 0) return 0;
  if (m == 0) return 0;
  return s < 0.0 ? static_cast<i128>(~m + 1) : static_cast<i128>(m);
}

// ---- MurmurHash3 over byte buffers (Austin Appleby's published algorithm) ---------------------
uint32_t Rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint64_t Rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
uint32_t Murmur3_x86_32(const unsigned char* data, int len, uint32_t seed) {
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  uint32_t h1 = seed;
  const int nblocks = len / 4;
  for (int i = 0; i < nblocks; ++i) {
    uint32_t k1;
    std::memcpy(&k1, data + 4 * i, 4);
    k1 *= c1; k1 = Rotl32(k1, 15); k1 *= c2;
    h1 ^= k1; h1 = Rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
  }
  const unsigned char* tail = data + nblocks * 4;
  uint32_t k1 = 0;
  switch (len & 3) {
    case 3: k1 ^= static_cast<uint32_t>(tail[2]) << 16;  // fall through
    case 2: k1 ^= static_cast<uint32_t>(tail[1]) << 8;   // fall through
    case 1: k1 ^= tail[0];
            k1 *= c1; k1 = Rotl32(k1, 15); k1 *= c2; h1 ^= k1;
  }
  h1 ^= static_cast<uint32_t>(len);
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}
uint64_t Fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
// low 64 bits (h1) of MurmurHash3_x64_128; the seed initialises both lanes
uint64_t Murmur3_x64_128_lo(const unsigned char* data, int len, uint64_t seed) {
  const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
  uint64_t h1 = seed, h2 = seed;
  const int nblocks = len / 16;
  for (int i = 0; i < nblocks; ++i) {
    uint64_t k1, k2;
    std::memcpy(&k1, data + 16 * i, 8);
    std::memcpy(&k2, data + 16 * i + 8, 8);
    k1 *= c1; k1 = Rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = Rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = Rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = Rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const unsigned char* tail = data + nblocks * 16;
  uint64_t k1 = 0, k2 = 0;
  const int rest = len & 15;
  for (int i = rest - 1; i >= 8; --i) k2 ^= static_cast<uint64_t>(tail[i]) << (8 * (i - 8));
  if (rest > 8) { k2 *= c2; k2 = Rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  for (int i = std::min(rest, 8) - 1; i >= 0; --i) k1 ^= static_cast<uint64_t>(tail[i]) << (8 * i);
  if (rest > 0) { k1 *= c1; k1 = Rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= static_cast<uint64_t>(len); h2 ^= static_cast<uint64_t>(len);
  h1 += h2; h2 += h1;
  h1 = Fmix64(h1); h2 = Fmix64(h2);
  h1 += h2;
  return h1;
}

double DecimalToDouble(i128 x, int xs) {
  const bool neg = x < 0;
  u128 m = neg ? (~static_cast<u128>(x) + 1) : static_cast<u128>(x);
  const double hi = static_cast<double>(static_cast<uint64_t>(m >> 64));
  const double lo = static_cast<double>(static_cast<uint64_t>(m));
  double v = hi * 18446744073709551616.0 + lo;
  double p = 1.0;
  for (int k = 0; k < xs; ++k) p = p * 10.0;
  v = v / p;
  return neg ? -v : v;
}

// ---- strings ------------------------------------------------------------------------------
int GlyphLen(unsigned char c) {
  if (c < 0x80) return 1;
  if ((c & 0xE0) == 0xC0) return 2;
  if ((c & 0xF0) == 0xE0) return 3;
  if ((c & 0xF8) == 0xF0) return 4;
  return 1;
}
std::vector<size_t> GlyphStarts(const std::string& s) {
  std::vector<size_t> st;
  size_t pos = 0;
  while (pos < s.size()) {
    st.push_back(pos);
    pos += static_cast<size_t>(GlyphLen(static_cast<unsigned char>(s[pos])));
  }
  return st;
}
std::string Substr(const std::string& s, int64_t offset, int64_t length) {
  if (length <= 0 || s.empty()) return "";
  std::vector<size_t> st = GlyphStarts(s);
  const int64_t glyphs = static_cast<int64_t>(st.size());
  int64_t from;
  if (offset > 0) from = offset - 1;
  else if (offset < 0) from = glyphs + offset;
  else from = 0;
  if (from < 0 || from >= glyphs) return "";
  // length may be huge (INT64_MAX-like); clamp before adding
  int64_t avail = glyphs - from;
  int64_t take = length < avail ? length : avail;
  const size_t b = st[static_cast<size_t>(from)];
  const size_t e = (from + take >= glyphs) ? s.size() : st[static_cast<size_t>(from + take)];
  return s.substr(b, e - b);
}
// Recursive LIKE matcher over tokens (memoisation unnecessary at test sizes).
bool LikeRec(const std::string& s, size_t i, const std::vector<int>& pat, size_t j) {
  while (j < pat.size()) {
    const int t = pat[j];
    if (t == 257) {
      // collapse runs of %
      while (j + 1 < pat.size() && pat[j + 1] == 257) ++j;
      if (j + 1 == pat.size()) return true;
      for (size_t k = i;; ) {
        if (LikeRec(s, k, pat, j + 1)) return true;
        if (k >= s.size()) return false;
        k += static_cast<size_t>(GlyphLen(static_cast<unsigned char>(s[k])));
        if (k > s.size()) k = s.size();
      }
    }
    if (i >= s.size()) return false;
    if (t == 256) {
      i += static_cast<size_t>(GlyphLen(static_cast<unsigned char>(s[i])));
      if (i > s.size()) i = s.size();
    } else {
      if (static_cast<unsigned char>(s[i]) != static_cast<unsigned>(t)) return false;
      ++i;
    }
    ++j;
  }
  return i == s.size();
}
std::vector<int> CompileLike(const std::string& pat, bool has_esc, char esc) {
  std::vector<int> out;
  for (size_t k = 0; k < pat.size(); ++k) {
    if (has_esc && pat[k] == esc && k + 1 < pat.size()) out.push_back(static_cast<unsigned char>(pat[++k]));
    else if (pat[k] == '%') out.push_back(257);
    else if (pat[k] == '_') out.push_back(256);
    else out.push_back(static_cast<unsigned char>(pat[k]));
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
void Eval(const Node& n, EvalCtx& cx, int64_t row, Val* out);

template <typename F>
bool Relop(const std::string& name, F cmp3) {  // cmp3() returns -1/0/1; NaN handled by callers
  const int c = cmp3();
  if (name == "equal" || name == "eq" || name == "same") return c == 0;
  if (name == "not_equal") return c != 0;
  if (name == "less_than") return c < 0;
  if (name == "less_than_or_equal_to") return c <= 0;
  if (name == "greater_than") return c > 0;
  return c >= 0;  // greater_than_or_equal_to
}
bool IsRelop(const std::string& n) {
  return n == "equal" || n == "eq" || n == "same" || n == "not_equal" || n == "less_than" ||
         n == "less_than_or_equal_to" || n == "greater_than" || n == "greater_than_or_equal_to";
}
template <typename T>
bool RelopNum(const std::string& name, T a, T b) {
  if (name == "equal" || name == "eq" || name == "same") return a == b;
  if (name == "not_equal") return a != b;
  if (name == "less_than") return a < b;
  if (name == "less_than_or_equal_to") return a <= b;
  if (name == "greater_than") return a > b;
  return a >= b;
}

// ---- message digests (RFC 1321 MD5, FIPS 180-4 SHA-1 / SHA-256): whole padded message in a
// vector, full message schedules -- written for clarity, unlike the kernel's rolling windows.
std::vector<uint8_t> PadMessage(const std::string& m, bool big_endian_len) {
  std::vector<uint8_t> p(m.begin(), m.end());
  p.push_back(0x80);
  while (p.size() % 64 != 56) p.push_back(0);
  const uint64_t bits = static_cast<uint64_t>(m.size()) * 8;
  for (int k = 0; k < 8; ++k)
    p.push_back(static_cast<uint8_t>(bits >> (big_endian_len ? 8 * (7 - k) : 8 * k)));
  return p;
}
uint32_t Rol(uint32_t v, int d) { return (v << d) | (v >> (32 - d)); }
uint32_t Ror(uint32_t v, int d) { return (v >> d) | (v << (32 - d)); }
std::string HexOfWords(const uint32_t* w, int n, bool big_endian) {
  static const char* digits = "0123456789abcdef";
  std::string out;
  for (int k = 0; k < n; ++k)
    for (int b = 0; b < 4; ++b) {
      const uint32_t byte = big_endian ? (w[k] >> (24 - 8 * b)) & 0xff : (w[k] >> (8 * b)) & 0xff;
      out.push_back(digits[byte >> 4]);
      out.push_back(digits[byte & 15]);
    }
  return out;
}
std::string Sha256Hex(const std::string& m) {
  uint32_t K[64];
  {  // K[i] = first 32 bits of the fractional part of the cube root of the i-th prime
    int count = 0;
    for (int p = 2; count < 64; ++p) {
      bool prime = true;
      for (int q = 2; q * q <= p; ++q) prime = prime && p % q != 0;
      if (!prime) continue;
      const long double r = cbrtl(static_cast<long double>(p));
      K[count++] = static_cast<uint32_t>((r - floorl(r)) * 4294967296.0L);
    }
  }
  uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  const std::vector<uint8_t> p = PadMessage(m, true);
  for (size_t off = 0; off < p.size(); off += 64) {
    uint32_t W[64];
    for (int t = 0; t < 16; ++t)
      W[t] = (uint32_t(p[off + 4 * t]) << 24) | (uint32_t(p[off + 4 * t + 1]) << 16) |
             (uint32_t(p[off + 4 * t + 2]) << 8) | uint32_t(p[off + 4 * t + 3]);
    for (int t = 16; t < 64; ++t) {
      const uint32_t s0 = Ror(W[t - 15], 7) ^ Ror(W[t - 15], 18) ^ (W[t - 15] >> 3);
      const uint32_t s1 = Ror(W[t - 2], 17) ^ Ror(W[t - 2], 19) ^ (W[t - 2] >> 10);
      W[t] = W[t - 16] + s0 + W[t - 7] + s1;
    }
    uint32_t v[8];
    std::memcpy(v, H, sizeof(v));
    for (int t = 0; t < 64; ++t) {
      const uint32_t t1 = v[7] + (Ror(v[4], 6) ^ Ror(v[4], 11) ^ Ror(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + K[t] + W[t];
      const uint32_t t2 = (Ror(v[0], 2) ^ Ror(v[0], 13) ^ Ror(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
      for (int k = 7; k > 0; --k) v[k] = v[k - 1];
      v[4] += t1;
      v[0] = t1 + t2;
    }
    for (int k = 0; k < 8; ++k) H[k] += v[k];
  }
  return HexOfWords(H, 8, true);
}
std::string Sha1Hex(const std::string& m) {
  uint32_t H[5] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476, 0xc3d2e1f0};
  const std::vector<uint8_t> p = PadMessage(m, true);
  for (size_t off = 0; off < p.size(); off += 64) {
    uint32_t W[80];
    for (int t = 0; t < 16; ++t)
      W[t] = (uint32_t(p[off + 4 * t]) << 24) | (uint32_t(p[off + 4 * t + 1]) << 16) |
             (uint32_t(p[off + 4 * t + 2]) << 8) | uint32_t(p[off + 4 * t + 3]);
    for (int t = 16; t < 80; ++t) W[t] = Rol(W[t - 3] ^ W[t - 8] ^ W[t - 14] ^ W[t - 16], 1);
    uint32_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4];
    for (int t = 0; t < 80; ++t) {
      uint32_t f, k;
      if (t < 20) { f = (b & c) | (~b & d); k = 0x5a827999; }
      else if (t < 40) { f = b ^ c ^ d; k = 0x6ed9eba1; }
      else if (t < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8f1bbcdc; }
      else { f = b ^ c ^ d; k = 0xca62c1d6; }
      const uint32_t tmp = Rol(a, 5) + f + e + k + W[t];
      e = d; d = c; c = Rol(b, 30); b = a; a = tmp;
    }
    H[0] += a; H[1] += b; H[2] += c; H[3] += d; H[4] += e;
  }
  return HexOfWords(H, 5, true);
}
std::string Md5Hex(const std::string& m) {
  uint32_t K[64];
  for (int i = 0; i < 64; ++i)  // K[i] = floor(2^32 * |sin(i + 1)|)
    K[i] = static_cast<uint32_t>(floorl(fabsl(sinl(static_cast<long double>(i + 1))) * 4294967296.0L));
  static const int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
  uint32_t H[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
  const std::vector<uint8_t> p = PadMessage(m, false);
  for (size_t off = 0; off < p.size(); off += 64) {
    uint32_t M[16];
    for (int t = 0; t < 16; ++t)
      M[t] = uint32_t(p[off + 4 * t]) | (uint32_t(p[off + 4 * t + 1]) << 8) | (uint32_t(p[off + 4 * t + 2]) << 16) |
             (uint32_t(p[off + 4 * t + 3]) << 24);
    uint32_t a = H[0], b = H[1], c = H[2], d = H[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      switch (i / 16) {
        case 0: f = (b & c) | (~b & d); g = i; break;
        case 1: f = (d & b) | (~d & c); g = (5 * i + 1) % 16; break;
        case 2: f = b ^ c ^ d; g = (3 * i + 5) % 16; break;
        default: f = c ^ (b | ~d); g = (7 * i) % 16; break;
      }
      const uint32_t x = a + f + K[i] + M[g];
      a = d; d = c; c = b;
      b = b + Rol(x, S[i / 16][i % 4]);
    }
    H[0] += a; H[1] += b; H[2] += c; H[3] += d;
  }
  return HexOfWords(H, 4, false);
}

// ---- exp / log / log10 / cbrt: the same IEEE operation sequences as the device library (fdlibm's
// reductions and polynomials, Sun Microsystems 1993); the two must agree bit for bit, and both are
// measured against the host libm in tests/test_oracle_vs_arrow.py.
double F64FromBits(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }
uint64_t F64Bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }
// ---- sin / cos / tan / cot: the kernel's sequence (device/gdv_device_lib.cuh gdv_trig), operation for
// operation: integer Payne-Hanek reduction against 1280 bits of 2/pi, Taylor kernels, double-double
// quotient for tan / cot.  Constants from tools/derive_trig_constants.py.  Checked against the host
// libm (<= 1 ULP) and against exact rational evaluation in tests/test_oracle_vs_arrow.py.
const uint64_t kTwoOverPi[20] = {
    0xa2f9836e4e441529ull, 0xfc2757d1f534ddc0ull, 0xdb6295993c439041ull, 0xfe5163abdebbc561ull,
    0xb7246e3a424dd2e0ull, 0x06492eea09d1921cull, 0xfe1deb1cb129a73eull, 0xe88235f52ebb4484ull,
    0xe99c7026b45f7e41ull, 0x3991d639835339f4ull, 0x9c845f8bbdf9283bull, 0x1ff897ffde05980full,
    0xef2f118b5a0a6d1full, 0x6d367ecf27cb09b7ull, 0x4f463f669e5fea2dull, 0x7527bac7ebe5f17bull,
    0x3d0739f78a5292eaull, 0x6bfb5fb11f8d5d08ull, 0x56033046fc7b6babull, 0xf0cfbc209af4361dull};
// ax finite, > pi/4: ax = quad * pi/2 + (y0 + y1), |y0 + y1| <= pi/4
void OrcRemPio2(double ax, double* y0, double* y1, int32_t* quad) {
  const uint64_t bits = F64Bits(ax);
  const int32_t e = (int32_t)(bits >> 52) - 1075;  // ax = M * 2^e, M in [2^52, 2^53)
  const uint64_t M = (bits & 0x000fffffffffffffull) | 0x0010000000000000ull;
  const int32_t i0 = e >= 2 ? e - 1 : 1;       // first bit of 2/pi (1 = the 2^-1 bit) that matters mod 4
  const int32_t s = e >= 2 ? 190 : 192 - e;    // the product below is (ax * 2/pi mod 4) * 2^s
  const int32_t j = (i0 - 1) >> 6, sh = (i0 - 1) & 63;
  uint64_t w[3];
  for (int32_t k = 0; k < 3; ++k)
    w[k] = sh ? (kTwoOverPi[j + k] << sh) | (kTwoOverPi[j + k + 1] >> (64 - sh)) : kTwoOverPi[j + k];
  const unsigned __int128 p2 = (unsigned __int128)M * w[2], p1 = (unsigned __int128)M * w[1], p0 = (unsigned __int128)M * w[0];
  uint64_t P[4];
  P[0] = (uint64_t)p2;
  unsigned __int128 acc = (p2 >> 64) + (unsigned __int128)(uint64_t)p1;
  P[1] = (uint64_t)acc;
  acc = (acc >> 64) + (p1 >> 64) + (unsigned __int128)(uint64_t)p0;
  P[2] = (uint64_t)acc;
  acc = (acc >> 64) + (p0 >> 64);
  P[3] = (uint64_t)acc;
  int32_t q = (int32_t)((P[s >> 6] >> (s & 63)) & 1ull) | ((int32_t)((P[(s + 1) >> 6] >> ((s + 1) & 63)) & 1ull) << 1);
  const bool half = ((P[(s - 1) >> 6] >> ((s - 1) & 63)) & 1ull) != 0ull;
  // keep the s fraction bits; past one half, go to the next quadrant and negate the fraction
  const uint64_t top_mask = (1ull << (s & 63)) - 1ull;
  for (int32_t k = 0; k < 4; ++k) {
    if (half) P[k] = ~P[k];
    if (k == (s >> 6)) P[k] &= top_mask;
    if (k > (s >> 6)) P[k] = 0ull;
  }
  if (half) {  // two's complement: + 1 (cannot carry out of s bits: the fraction was not zero)
    ++q;
    for (int32_t k = 0; k < 4; ++k) {
      P[k] += 1ull;
      if (P[k] != 0ull) break;
    }
  }
  *quad = q & 3;
  int32_t p = -1;
  for (int32_t k = 3; k >= 0 && p < 0; --k)
    if (P[k] != 0ull) p = k * 64 + 63 - __builtin_clzll(P[k]);
  if (p < 0) {
    *y0 = 0.0;
    *y1 = 0.0;
    return;
  }
  // the top 106 bits of the fraction as two 53-bit integers
  const int32_t up = 255 - p, uw = up >> 6, ub = up & 63;
  uint64_t G[4];
  for (int32_t k = 3; k >= 0; --k) {
    const uint64_t hi = k - uw >= 0 ? P[k - uw] : 0ull;
    const uint64_t lo = k - uw - 1 >= 0 ? P[k - uw - 1] : 0ull;
    G[k] = ub ? (hi << ub) | (lo >> (64 - ub)) : hi;
  }
  const uint64_t H = G[3] >> 11, L = ((G[3] & 0x7ffull) << 42) | (G[2] >> 22);
  const double sc = F64FromBits((uint64_t)(int64_t)(1023 + p - 52 - s) << 52);
  const double fh = (double)(int64_t)H * sc, fl = ((double)(int64_t)L * sc) * 1.1102230246251565e-16;  // * 2^-53
  const double ph = 1.5707963267948966, pl = 6.123233995736766e-17;
  const double t = fh * ph;
  const double err = std::fma(fh, ph, -t);
  const double lo = err + (fh * pl + fl * ph);
  double r0 = t + lo;
  double r1 = (t - r0) + lo;
  if (half) {
    r0 = -r0;
    r1 = -r1;
  }
  *y0 = r0;
  *y1 = r1;
}
double OrcKSinPoly(double z) {  // (sin(x) - x + x^3/6) / x^5, z = x^2
  const double S2 = 0.008333333333333333, S3 = -0.0001984126984126984, S4 = 2.7557319223985893e-06,
            S5 = -2.505210838544172e-08, S6 = 1.6059043836821613e-10, S7 = -7.647163731819816e-13,
            S8 = 2.8114572543455206e-15;
  return S2 + z * (S3 + z * (S4 + z * (S5 + z * (S6 + z * (S7 + z * S8)))));
}
double OrcKCosPoly(double z) {
  const double C1 = 0.041666666666666664, C2 = -0.001388888888888889, C3 = 2.48015873015873e-05,
            C4 = -2.755731922398589e-07, C5 = 2.08767569878681e-09, C6 = -1.1470745597729725e-11,
            C7 = 4.779477332387385e-14, C8 = -1.5619206968586225e-16;
  return z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * (C6 + z * (C7 + z * C8)))))));
}
// fn: 0 sin, 1 cos, 2 tan, 3 cot
double OrcTrig(double x, int32_t fn) {
  const uint64_t bits = F64Bits(x);
  const uint64_t ab = bits & 0x7fffffffffffffffull;
  if (ab >= 0x7ff0000000000000ull) return F64FromBits(0x7ff8000000000000ull);  // inf, nan -> the canonical nan
  const bool neg = (bits >> 63) != 0ull;
  const double ax = F64FromBits(ab);
  if (ab < 0x3e40000000000000ull) {  // |x| < 2^-27
    if (fn == 1) return 1.0;
    if (fn == 3) return 1.0 / x;
    return x;
  }
  double y0 = ax, y1 = 0.0;
  int32_t q = 0;
  if (ax > 0.7853981633974483) OrcRemPio2(ax, &y0, &y1, &q);
  // sin = y0 - (y0 + y1)^3 / 6 + ..., cos = 1 - (y0 + y1)^2 / 2 + ...: the squares, the cube and 1/6
  // carry their rounding errors along, so that both come out as double-doubles good to ~2^-58
  const double S1 = -0.16666666666666666, S1L = -9.25185853854297e-18;
  const double z = y0 * y0;
  const double zl = std::fma(y0, y0, -z) + (2.0 * y0) * y1;
  const double v = z * y0;
  const double vl = std::fma(z, y0, -v) + (zl * y0 + z * y1);
  const double t3 = v * S1;
  const double t3l = (std::fma(v, S1, -t3) + v * S1L) + vl * S1;
  const double s_rest = (t3l + y1) + ((z * z) * y0) * OrcKSinPoly(z);
  const double s_a = y0 + t3, s_b = ((y0 - s_a) + t3) + s_rest;
  const double s_hi = s_a + s_b, s_lo = (s_a - s_hi) + s_b;
  const double ch = -0.5 * z;
  const double c_rest = z * OrcKCosPoly(z) - 0.5 * zl;
  const double c_a = 1.0 + ch, c_b = ((1.0 - c_a) + ch) + c_rest;
  const double c_hi = c_a + c_b, c_lo = (c_a - c_hi) + c_b;
  double res;
  if (fn <= 1) {
    const int32_t k = (q + fn) & 3;  // cos(t) = sin(t + pi/2)
    res = (k & 1) ? c_hi : s_hi;
    if (k & 2) res = -res;
    if (fn == 0 && neg) res = -res;
    return res;
  }
  // tan / cot: one double-double divided by the other
  const bool sin_over_cos = ((q & 1) != 0) == (fn == 3);
  const double n_hi = sin_over_cos ? s_hi : c_hi, n_lo = sin_over_cos ? s_lo : c_lo;
  const double d_hi = sin_over_cos ? c_hi : s_hi, d_lo = sin_over_cos ? c_lo : s_lo;
  const double q0 = n_hi / d_hi;
  const double rem = std::fma(-q0, d_hi, n_hi);
  const double q1 = ((rem + n_lo) - q0 * d_lo) / d_hi;
  res = q0 + q1;
  if (q & 1) res = -res;
  return neg ? -res : res;
}
double OrcExp(double x) {
  const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return F64FromBits(0x7ff0000000000000ull);
  if (x < -7.45133219101941108420e+02) return 0.0;
  const double ax = x < 0.0 ? -x : x;
  double hi = 0.0, lo = 0.0;
  int32_t k = 0;
  if (ax > 0.34657359027997264) {
    if (ax < 1.0397207708399179) k = x < 0.0 ? -1 : 1;
    else k = static_cast<int32_t>(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
    hi = x - static_cast<double>(k) * ln2hi;
    lo = static_cast<double>(k) * ln2lo;
    x = hi - lo;
  } else if (ax < 3.7252902984619141e-09) {
    return 1.0 + x;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k == 1024) return (y * 2.0) * 8.98846567431157953865e+307;
  if (k >= -1021) return F64FromBits(F64Bits(y) + (static_cast<uint64_t>(static_cast<int64_t>(k)) << 52));
  return F64FromBits(F64Bits(y) + (static_cast<uint64_t>(static_cast<int64_t>(k + 1000)) << 52)) * 9.33263618503218878990e-302;
}
double OrcLog(double x) {
  const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  if (x != x) return x;
  if (x < 0.0) return F64FromBits(0x7ff8000000000000ull);
  if (x == 0.0) return F64FromBits(0xfff0000000000000ull);
  uint64_t bits = F64Bits(x);
  if (bits == 0x7ff0000000000000ull) return x;
  int32_t k = 0;
  if ((bits >> 52) == 0) {
    x = x * 18014398509481984.0;
    bits = F64Bits(x);
    k = -54;
  }
  int32_t hx = static_cast<int32_t>(bits >> 32);
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int32_t i = (hx + 0x95f64) & 0x100000;
  x = F64FromBits((static_cast<uint64_t>(static_cast<uint32_t>(hx | (i ^ 0x3ff00000))) << 32) | (bits & 0xffffffffull));
  k += i >> 20;
  const double f = x - 1.0;
  const double dk = static_cast<double>(k);
  if ((0x000fffff & (2 + hx)) < 3) {
    if (f == 0.0) return k == 0 ? 0.0 : dk * ln2hi + dk * ln2lo;
    const double R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    return dk * ln2hi - ((R - dk * ln2lo) - f);
  }
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  const double R = t2 + t1;
  if (((hx - 0x6147a) | (0x6b851 - hx)) > 0) {
    const double hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2hi - ((hfsq - (s * (hfsq + R) + dk * ln2lo)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * ln2hi - ((s * (f - R) - dk * ln2lo) - f);
}
double OrcLog10(double x) {
  // FreeBSD msun's e_log10: log(1 + f) kept as a hi + lo pair, multiplied by 1 / ln 10 (hi + lo too)
  const double ivln10hi = 4.34294481878168880939e-01, ivln10lo = 2.50829467116452752298e-11;
  const double log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
            Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
            Lg7 = 1.479819860511658591e-01;
  if (x != x) return x;
  if (x < 0.0) return F64FromBits(0x7ff8000000000000ull);
  if (x == 0.0) return F64FromBits(0xfff0000000000000ull);
  uint64_t bits = F64Bits(x);
  if (bits == 0x7ff0000000000000ull) return x;
  if (x == 1.0) return 0.0;
  int32_t k = 0;
  if ((bits >> 52) == 0ull) {
    x = x * 18014398509481984.0;
    bits = F64Bits(x);
    k = -54;
  }
  int32_t hx = (int32_t)(bits >> 32);
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int32_t i = (hx + 0x95f64) & 0x100000;
  x = F64FromBits(((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (bits & 0xffffffffull));
  k += i >> 20;
  const double dk = (double)k;
  const double f = x - 1.0;
  const double hfsq = 0.5 * f * f;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  const double r = s * (hfsq + (t2 + t1));  // log(1 + f) - f + f * f / 2
  double hi = f - hfsq;
  hi = F64FromBits(F64Bits(hi) & 0xffffffff00000000ull);
  const double lo = (f - hi) - hfsq + r;
  double val_hi = hi * ivln10hi;
  const double y2 = dk * log10_2hi;
  double val_lo = dk * log10_2lo + (lo + hi) * ivln10lo + lo * ivln10hi;
  const double ww = y2 + val_hi;
  val_lo += (y2 - ww) + val_hi;
  val_hi = ww;
  return val_lo + val_hi;
}
double OrcCbrt(double x) {
  const double P0 = 1.87595182427177009643, P1 = -1.88497979543377169875, P2 = 1.621429720105354466140,
               P3 = -0.758397934778766047437, P4 = 0.145996192886612446982;
  const uint64_t bits = F64Bits(x);
  const uint64_t sign = bits & 0x8000000000000000ull;
  const uint32_t hx = static_cast<uint32_t>(bits >> 32) & 0x7fffffffu;
  if (hx >= 0x7ff00000u) return x + x;
  double t;
  if (hx < 0x00100000u) {
    if ((bits & 0x7fffffffffffffffull) == 0) return x;
    const double sc = F64FromBits(bits & 0x7fffffffffffffffull) * 18014398509481984.0;
    const uint32_t h2 = static_cast<uint32_t>(F64Bits(sc) >> 32) & 0x7fffffffu;
    t = F64FromBits(sign | (static_cast<uint64_t>(h2 / 3u + 696219795u) << 32));
  } else {
    t = F64FromBits(sign | (static_cast<uint64_t>(hx / 3u + 715094163u) << 32));
  }
  double r = (t * t) * (t / x);
  t = t * ((P0 + r * (P1 + r * P2)) + ((r * r) * r) * (P3 + r * P4));
  t = F64FromBits((F64Bits(t) + 0x80000000ull) & 0xffffffffc0000000ull);
  const double s2 = t * t;
  r = x / s2;
  const double w = t + t;
  r = (r - t) / (w + r);
  return t + t * r;
}

// ---- regexp_matches / regexp_like: RE2's partial-match semantics (the reference's holder), restated as
// a plain backtracking matcher over CODE POINTS -- nothing like the kernel's byte automaton.  '.' is any
// code point but '\n'; \d \w \s are ASCII; '^' / '$' match at the ends of the text only.
struct OrcRe {
  enum K { CHAR, ANY, CLASS, CAT, ALT, REP, BOL, EOL } k = CAT;
  uint32_t cp = 0;
  std::vector<std::pair<uint32_t, uint32_t>> ranges;
  bool neg = false;
  bool icase = false;  // (?i): ASCII letters match in both cases
  std::vector<std::shared_ptr<OrcRe>> kids;
  int lo = 0, hi = -1;
};
using OrcReP = std::shared_ptr<OrcRe>;
std::vector<uint32_t> DecodeUtf8(const std::string& t) {
  std::vector<uint32_t> out;
  for (size_t i = 0; i < t.size();) {
    const unsigned char c = static_cast<unsigned char>(t[i]);
    const int n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
    uint32_t cp = n == 1 ? c : c & (0xffu >> (n + 1));
    for (int k = 1; k < n && i + k < t.size(); ++k) cp = (cp << 6) | (static_cast<unsigned char>(t[i + k]) & 0x3fu);
    out.push_back(cp);
    i += static_cast<size_t>(n);
  }
  return out;
}
struct OrcReParser {
  std::vector<uint32_t> p;
  size_t i = 0;
  bool bad = false;
  bool icase = false;
  OrcReP Node(OrcRe::K k) { auto r = std::make_shared<OrcRe>(); r->k = k; r->icase = icase; return r; }
  bool Posix(OrcRe* cls) {   // at "[:", inside a bracket expression
    size_t e = i + 2;
    std::string name;
    while (e + 1 < p.size() && !(p[e] == ':' && p[e + 1] == ']')) name.push_back(static_cast<char>(p[e++]));
    if (e + 1 >= p.size()) return false;
    auto add = [&](uint32_t lo, uint32_t hi) { cls->ranges.push_back({lo, hi}); };
    if (name == "alpha") { add('a', 'z'); add('A', 'Z'); }
    else if (name == "digit") add('0', '9');
    else if (name == "alnum") { add('a', 'z'); add('A', 'Z'); add('0', '9'); }
    else if (name == "upper") add('A', 'Z');
    else if (name == "lower") add('a', 'z');
    else if (name == "space") { add(9, 13); add(' ', ' '); }
    else if (name == "blank") { add(' ', ' '); add(9, 9); }
    else if (name == "punct") { add('!', '/'); add(':', '@'); add('[', '`'); add('{', '~'); }
    else if (name == "xdigit") { add('0', '9'); add('a', 'f'); add('A', 'F'); }
    else if (name == "word") { add('a', 'z'); add('A', 'Z'); add('0', '9'); add('_', '_'); }
    else if (name == "print") add(' ', '~');
    else if (name == "graph") add('!', '~');
    else if (name == "cntrl") { add(0, 31); add(127, 127); }
    else if (name == "ascii") add(0, 127);
    else return false;
    i = e + 2;
    return true;
  }
  uint32_t Hex2() {   // after "\\x"
    uint32_t v = 0;
    for (int k = 0; k < 2; ++k) {
      if (i >= p.size() || !std::isxdigit(static_cast<int>(p[i]))) { bad = true; return 0; }
      const uint32_t h = p[i++];
      v = v * 16 + (h <= '9' ? h - '0' : (h | 0x20) - 'a' + 10);
    }
    return v;
  }
  void Shorthand(uint32_t c, OrcRe* cls) {
    if (c == 'd' || c == 'D') cls->ranges.push_back({'0', '9'});
    if (c == 'w' || c == 'W') { cls->ranges.push_back({'0', '9'}); cls->ranges.push_back({'a', 'z'}); cls->ranges.push_back({'A', 'Z'}); cls->ranges.push_back({'_', '_'}); }
    if (c == 's' || c == 'S') for (uint32_t ch : {' ', '\t', '\n', '\f', '\r'}) cls->ranges.push_back({ch, ch});
  }
  static uint32_t Unescape(uint32_t c) {
    switch (c) { case 'n': return '\n'; case 't': return '\t'; case 'r': return '\r'; case 'f': return '\f'; case 'v': return '\v'; case 'a': return 7; default: return c; }
  }
  OrcReP Alt() {
    auto alt = Node(OrcRe::ALT);
    alt->kids.push_back(Cat());
    while (i < p.size() && p[i] == '|') { ++i; alt->kids.push_back(Cat()); }
    return alt->kids.size() == 1 ? alt->kids[0] : alt;
  }
  OrcReP Cat() {
    auto cat = Node(OrcRe::CAT);
    while (i < p.size() && p[i] != '|' && p[i] != ')') cat->kids.push_back(Rep());
    return cat;
  }
  bool Num(int* v) {
    if (i >= p.size() || p[i] < '0' || p[i] > '9') return false;
    *v = 0;
    while (i < p.size() && p[i] >= '0' && p[i] <= '9') { *v = std::min(1000, *v * 10 + static_cast<int>(p[i] - '0')); ++i; }
    return true;
  }
  OrcReP Rep() {
    OrcReP a = Atom();
    while (i < p.size()) {
      int lo, hi;
      if (p[i] == '*') { lo = 0; hi = -1; ++i; }
      else if (p[i] == '+') { lo = 1; hi = -1; ++i; }
      else if (p[i] == '?') { lo = 0; hi = 1; ++i; }
      else if (p[i] == '{') {
        const size_t save = i++;
        if (!Num(&lo)) { i = save; break; }
        hi = lo;
        if (i < p.size() && p[i] == ',') { ++i; if (!Num(&hi)) hi = -1; }
        if (i >= p.size() || p[i] != '}') { i = save; break; }
        ++i;
      } else break;
      if (i < p.size() && p[i] == '?') ++i;
      auto r = Node(OrcRe::REP);
      r->kids.push_back(a);
      r->lo = lo;
      r->hi = hi;
      a = r;
    }
    return a;
  }
  OrcReP Atom() {
    const uint32_t c = p[i++];
    if (c == '(') {
      if (i + 1 < p.size() && p[i] == '?' && p[i + 1] == ':') i += 2;
      OrcReP r = Alt();
      if (i >= p.size() || p[i] != ')') bad = true; else ++i;
      return r;
    }
    if (c == '.') return Node(OrcRe::ANY);
    if (c == '^') return Node(OrcRe::BOL);
    if (c == '$') return Node(OrcRe::EOL);
    if (c == '[') {
      auto cls = Node(OrcRe::CLASS);
      if (i < p.size() && p[i] == '^') { cls->neg = true; ++i; }
      bool first = true;
      while (true) {
        if (i >= p.size()) { bad = true; return cls; }
        if (p[i] == '[' && i + 1 < p.size() && p[i + 1] == ':' && !(first && false)) {
          first = false;
          if (!Posix(cls.get())) { bad = true; return cls; }
          continue;
        }
        uint32_t lo = p[i++];
        if (lo == ']' && !first) break;
        first = false;
        if (lo == '\\') {
          if (i >= p.size()) { bad = true; return cls; }
          const uint32_t e = p[i++];
          if (e == 'd' || e == 'w' || e == 's') { Shorthand(e, cls.get()); continue; }
          lo = e == 'x' ? Hex2() : Unescape(e);
        }
        uint32_t hi = lo;
        if (i + 1 < p.size() && p[i] == '-' && p[i + 1] != ']') {
          ++i;
          hi = p[i++];
          if (hi == '\\' && i < p.size()) {
            const uint32_t e2 = p[i++];
            hi = e2 == 'x' ? Hex2() : Unescape(e2);
          }
        }
        cls->ranges.push_back({lo, hi});
      }
      return cls;
    }
    if (c == '\\') {
      if (i >= p.size()) { bad = true; return Node(OrcRe::CAT); }
      const uint32_t e = p[i++];
      if (e == 'd' || e == 'w' || e == 's' || e == 'D' || e == 'W' || e == 'S') {
        auto cls = Node(OrcRe::CLASS);
        Shorthand(e, cls.get());
        cls->neg = e == 'D' || e == 'W' || e == 'S';
        return cls;
      }
      auto ch = Node(OrcRe::CHAR);
      ch->cp = e == 'x' ? Hex2() : Unescape(e);
      return ch;
    }
    auto ch = Node(OrcRe::CHAR);
    ch->cp = c;
    return ch;
  }
};
using OrcReCont = std::function<bool(size_t)>;
bool OrcReMatch(const OrcRe* r, const std::vector<uint32_t>& t, size_t pos, const OrcReCont& k);
bool OrcReCat(const OrcRe* r, size_t idx, const std::vector<uint32_t>& t, size_t pos, const OrcReCont& k) {
  if (idx == r->kids.size()) return k(pos);
  return OrcReMatch(r->kids[idx].get(), t, pos, [&](size_t p2) { return OrcReCat(r, idx + 1, t, p2, k); });
}
bool OrcReRep(const OrcRe* r, int count, const std::vector<uint32_t>& t, size_t pos, const OrcReCont& k) {
  if (count >= r->lo && k(pos)) return true;
  if (r->hi >= 0 && count >= r->hi) return false;
  return OrcReMatch(r->kids[0].get(), t, pos, [&](size_t p2) {
    if (p2 == pos) return count < r->lo ? k(pos) : false;  // an empty iteration: the remaining ones can be empty too
    return OrcReRep(r, count + 1, t, p2, k);
  });
}
bool OrcReMatch(const OrcRe* r, const std::vector<uint32_t>& t, size_t pos, const OrcReCont& k) {
  switch (r->k) {
    case OrcRe::CHAR: {
      if (pos >= t.size()) return false;
      auto low = [](uint32_t c) { return c >= 'A' && c <= 'Z' ? c + 32 : c; };
      return (r->icase ? low(t[pos]) == low(r->cp) : t[pos] == r->cp) && k(pos + 1);
    }
    case OrcRe::ANY: return pos < t.size() && t[pos] != '\n' && k(pos + 1);
    case OrcRe::CLASS: {
      if (pos >= t.size()) return false;
      bool in = false;
      uint32_t other = t[pos];   // the same letter in the other case, under (?i)
      if (r->icase && other >= 'a' && other <= 'z') other -= 32;
      else if (r->icase && other >= 'A' && other <= 'Z') other += 32;
      for (const auto& rg : r->ranges)
        in = in || (t[pos] >= rg.first && t[pos] <= rg.second) || (other >= rg.first && other <= rg.second);
      return in != r->neg && k(pos + 1);
    }
    case OrcRe::CAT: return OrcReCat(r, 0, t, pos, k);
    case OrcRe::ALT:
      for (const auto& kid : r->kids)
        if (OrcReMatch(kid.get(), t, pos, k)) return true;
      return false;
    case OrcRe::REP: return OrcReRep(r, 0, t, pos, k);
    case OrcRe::BOL: return pos == 0 && k(pos);
    case OrcRe::EOL: return pos == t.size() && k(pos);
  }
  return false;
}
bool OrcReSearch(const OrcRe* r, const std::string& text) {
  const std::vector<uint32_t> t = DecodeUtf8(text);
  for (size_t start = 0; start <= t.size(); ++start)
    if (OrcReMatch(r, t, start, [](size_t) { return true; })) return true;
  return false;
}

// ---- castFLOAT8 / castFLOAT4(utf8): m * 10^e10 (m <= 19 digits) carried as X * 2^exp2, X < 2^256:
// times ten, k <= 19 times, after X is cut to its top 192 bits; or shifted up to bit 255 and divided
// by ten k times.  The same cuts as the kernel (gdv_parse_f64), made here one bit / one digit at a
// time on the 32-bit-limb Big; the final round-to-nearest-even goes through ldexp.
int BigTop(const Big& x) {
  int p = 255;
  while (p > 0 && !x.Bit(p)) --p;
  return p;
}
void BigShr1(Big* x, bool* sticky) {
  *sticky = *sticky || (x->w[0] & 1u);
  for (int j = 0; j < 8; ++j) x->w[j] = (x->w[j] >> 1) | (j + 1 < 8 ? x->w[j + 1] << 31 : 0u);
}
void BigShl1(Big* x) {
  for (int j = 7; j >= 0; --j) x->w[j] = (x->w[j] << 1) | (j > 0 ? x->w[j - 1] >> 31 : 0u);
}
double BigToDouble(Big x, bool sticky, int exp2) {  // x != 0: RNE(x * 2^exp2)
  const int p = BigTop(x);
  const int be = p + exp2;
  if (be > 1023) return std::numeric_limits<double>::infinity();
  const int nb = be >= -1022 ? 53 : be + 1075;
  if (nb < 0) return 0.0;
  int e2 = exp2;
  bool round = false;
  for (int k = 0; k < p + 1 - nb; ++k) {  // drop the low bits: the last one dropped is the round bit
    sticky = sticky || round;
    round = x.w[0] & 1u;
    bool unused = false;
    BigShr1(&x, &unused);
    ++e2;
  }
  uint64_t mant = (static_cast<uint64_t>(x.w[1]) << 32) | x.w[0];
  if (round && (sticky || (mant & 1))) ++mant;
  return std::ldexp(static_cast<double>(mant), e2);
}
bool ParseDouble(const std::string& text, double* out) {
  std::string str = text;
  while (!str.empty() && str.front() == ' ') str.erase(str.begin());
  while (!str.empty() && str.back() == ' ') str.pop_back();
  size_t pos = 0;
  bool neg = false;
  if (pos < str.size() && (str[pos] == '-' || str[pos] == '+')) { neg = str[pos] == '-'; ++pos; }
  uint64_t m = 0;
  int sig = 0, ndig = 0;
  int64_t e10 = 0;
  bool point = false, sticky = false;
  for (; pos < str.size(); ++pos) {
    const char ch = str[pos];
    if (ch == '.') { if (point) return false; point = true; continue; }
    if (ch < '0' || ch > '9') break;
    ++ndig;
    if (sig < 19) {
      if (m != 0 || ch != '0') { m = m * 10 + static_cast<uint64_t>(ch - '0'); ++sig; }
      if (point) --e10;
    } else {
      sticky = sticky || ch != '0';
      if (!point) ++e10;
    }
  }
  if (ndig == 0) return false;
  if (pos < str.size() && (str[pos] == 'e' || str[pos] == 'E')) {
    ++pos;
    bool eneg = false;
    if (pos < str.size() && (str[pos] == '-' || str[pos] == '+')) { eneg = str[pos] == '-'; ++pos; }
    int64_t ev = 0;
    int edig = 0;
    for (; pos < str.size() && str[pos] >= '0' && str[pos] <= '9'; ++pos) {
      if (ev < 100000) ev = ev * 10 + (str[pos] - '0');
      ++edig;
    }
    if (edig == 0) return false;
    e10 += eneg ? -ev : ev;
  }
  if (pos != str.size()) return false;
  if (m == 0) { *out = neg ? -0.0 : 0.0; return true; }
  e10 = std::max<int64_t>(-400, std::min<int64_t>(400, e10));
  Big x = Big::From(m);
  int exp2 = 0;
  int64_t rest = e10;
  while (rest > 0) {
    const int k = static_cast<int>(std::min<int64_t>(19, rest));
    while (BigTop(x) > 191) { BigShr1(&x, &sticky); ++exp2; }
    x.MulPow10(k);
    rest -= k;
  }
  while (rest < 0) {
    const int k = static_cast<int>(std::min<int64_t>(19, -rest));
    while (BigTop(x) < 255) { BigShl1(&x); --exp2; }
    for (int j = 0; j < k; ++j) sticky = (x.DivSmall(10) != 0) || sticky;
    rest += k;
  }
  const double d = BigToDouble(x, sticky, exp2);
  *out = neg ? -d : d;
  return true;
}

// ---- power(x, y): the kernel's integer algorithm (device/gdv_device_lib.cuh power_float64_float64) on
// the 32-bit-limb Big: log2 of the significand bit by bit in Q2.126 (120 bits), the exact product with
// y's significand, e^(F ln 2) as a 34-term sum in Q1.127, one final rounding through BigToDouble.
unsigned __int128 BigBits128(const Big& b, int shift) {  // bits [shift, shift + 128) of b
  unsigned __int128 r = 0;
  for (int i = 127; i >= 0; --i) r = (r << 1) | (shift + i < 256 && b.Bit(shift + i) ? 1u : 0u);
  return r;
}
unsigned __int128 MulShr(unsigned __int128 a, unsigned __int128 b, int shift) {
  return BigBits128(Big::Mul(Big::From(a), Big::From(b)), shift);
}
struct OrcBigF { unsigned __int128 mant; int e2; };  // mant * 2^(e2 - 127), mant in [2^127, 2^128)
OrcBigF OrcExp2Q(bool tneg, int ip, unsigned __int128 fq) {  // 2^(+-(ip + fq / 2^128))
  using u128 = unsigned __int128;
  int e2 = ip;
  if (tneg) {
    if (fq != 0) { e2 = -e2 - 1; fq = static_cast<u128>(0) - fq; }
    else e2 = -e2;
  }
  const u128 ln2 = (static_cast<u128>(0x58b90bfbe8e7bcd5ull) << 64) | 0xe4f1d9cc01f97b57ull;
  const u128 z = MulShr(fq, ln2, 128);
  const u128 one = static_cast<u128>(1) << 127;
  u128 acc = 0;
  for (unsigned n = 34; n >= 2; --n) acc = MulShr(one + acc, z, 127) / n;
  return OrcBigF{one + MulShr(one + acc, z, 127), e2};
}
double OrcPow(double x, double y) {
  const uint64_t xb = F64Bits(x), yb = F64Bits(y);
  const uint64_t xa = xb & 0x7fffffffffffffffull, ya = yb & 0x7fffffffffffffffull;
  const uint64_t inf = 0x7ff0000000000000ull, one_bits = 0x3ff0000000000000ull, sign = 0x8000000000000000ull;
  const bool xneg = (xb >> 63) != 0, yneg = (yb >> 63) != 0;
  if (ya == 0 || xb == one_bits) return 1.0;
  if (xa > inf || ya > inf) return F64FromBits(0x7ff8000000000000ull);
  const int yexp = static_cast<int>(ya >> 52);
  const uint64_t my = (ya & 0x000fffffffffffffull) | (yexp == 0 ? 0 : 0x0010000000000000ull);
  const int ey = (yexp == 0 ? 1 : yexp) - 1075;
  bool y_int = false, y_odd = false;
  if (ya < inf) {
    if (ey >= 0) { y_int = true; y_odd = ey == 0 && (my & 1); }
    else if (ey >= -52) {
      y_int = (my & ((1ull << (-ey)) - 1)) == 0;
      y_odd = y_int && ((my >> (-ey)) & 1);
    }
  }
  const bool res_neg = xneg && y_odd;
  const uint64_t sbit = res_neg ? sign : 0;
  if (xa == 0) return F64FromBits(sbit | (yneg ? inf : 0));
  if (ya == inf) {
    if (xa == one_bits) return 1.0;
    return ((xa > one_bits) != yneg) ? F64FromBits(inf) : 0.0;
  }
  if (xa == inf) return F64FromBits(sbit | (yneg ? 0 : inf));
  if (xneg && !y_int) return F64FromBits(0x7ff8000000000000ull);
  const int xexp = static_cast<int>(xa >> 52);
  uint64_t mx = (xa & 0x000fffffffffffffull) | (xexp == 0 ? 0 : 0x0010000000000000ull);
  int k = (xexp == 0 ? 1 : xexp) - 1023;
  while ((mx >> 52) == 0) { mx <<= 1; --k; }
  using u128 = unsigned __int128;
  u128 m = static_cast<u128>(mx) << 74;
  u128 frac = 0;
  for (int i = 0; i < 120; ++i) {
    const u128 sq = MulShr(m, m, 126);
    const bool two = (sq >> 127) != 0;
    m = two ? sq >> 1 : sq;
    frac = (frac << 1) | (two ? 1u : 0u);
  }
  const bool lneg = k < 0;
  uint64_t ip = static_cast<uint64_t>(lneg ? -k : k);
  u128 fp = frac;
  if (lneg && frac != 0) { ip -= 1; fp = (static_cast<u128>(1) << 120) - frac; }
  if (ip == 0 && fp == 0) return res_neg ? -1.0 : 1.0;
  // P = M_y * (ip * 2^120 + fp)
  Big lq = Big::From(static_cast<u128>(ip));
  for (int i = 0; i < 120; ++i) BigShl1(&lq);
  lq.Add(Big::From(fp));
  const Big P = Big::Mul(Big::From(static_cast<u128>(my)), lq);
  const bool tneg = lneg != yneg;
  const int s = 120 - ey;
  if (s > 250) return res_neg ? -1.0 : 1.0;
  const double big = F64FromBits(sbit | (tneg ? 0 : inf));
  if (s <= 0) return big;
  // integer part: bits [s, 256); fraction: the 128 bits below bit s
  for (int i = s + 11; i < 256; ++i) if (P.Bit(i)) return big;
  int e2 = 0;
  for (int i = s + 10; i >= s; --i) e2 = (e2 << 1) | (P.Bit(i) ? 1 : 0);
  if (e2 >= 1100) return big;
  u128 fq = 0;
  for (int i = 1; i <= 128; ++i) fq = (fq << 1) | (s - i >= 0 && P.Bit(s - i) ? 1u : 0u);
  const OrcBigF v = OrcExp2Q(tneg, e2, fq);
  const double r = BigToDouble(Big::From(v.mant), true, v.e2 - 127);
  return res_neg ? -r : r;
}

// sinh / cosh / tanh (the kernel's gdv_hyperbolic): e^|x| and e^-|x| as 128-bit significands, combined in
// 256-bit integers, one rounding.  fn: 0 sinh, 1 cosh, 2 tanh
double OrcHyperbolic(double x, int fn) {
  using u128 = unsigned __int128;
  const uint64_t xb = F64Bits(x), xa = xb & 0x7fffffffffffffffull, inf = 0x7ff0000000000000ull;
  const bool neg = (xb >> 63) != 0 && fn != 1;
  if (xa > inf) return F64FromBits(0x7ff8000000000000ull);
  if (xa < 0x3e30000000000000ull) return fn == 1 ? 1.0 : x;
  const double top = fn == 2 ? 1.0 : F64FromBits(inf);
  if (xa == inf) return neg ? -top : top;
  const uint64_t mx = (xa & 0x000fffffffffffffull) | 0x0010000000000000ull;
  const int ex = static_cast<int>(xa >> 52) - 1075;
  const u128 log2e = (static_cast<u128>(0x5c551d94ae0bf85dull) << 64) | 0xdf43ff68348e9f44ull;
  const Big P = Big::Mul(Big::From(static_cast<u128>(mx)), Big::From(log2e));
  const int s = 126 - ex;
  if (s <= 0) return neg ? -top : top;
  for (int i = s + 11; i < 256; ++i) if (P.Bit(i)) return neg ? -top : top;
  int ip = 0;
  for (int i = s + 10; i >= s; --i) ip = (ip << 1) | (P.Bit(i) ? 1 : 0);
  if (ip >= 1100) return neg ? -top : top;
  u128 fq = 0;
  for (int i = 1; i <= 128; ++i) fq = (fq << 1) | (s - i >= 0 && P.Bit(s - i) ? 1u : 0u);
  const OrcBigF a = OrcExp2Q(false, ip, fq), b = OrcExp2Q(true, ip, fq);
  Big A = Big::From(a.mant), B = Big::From(b.mant);
  for (int i = 0; i < 126; ++i) { BigShl1(&A); BigShl1(&B); }
  bool sticky = true;
  for (int i = 0; i < a.e2 - b.e2; ++i) {
    if (B.IsZero()) break;
    BigShr1(&B, &sticky);
  }
  Big sum = A, dif = A;
  sum.Add(B);
  dif.Sub(B);
  double r;
  if (fn == 2) {
    const Big num = Big::From(BigBits128(dif, 127)), den = Big::From(BigBits128(sum, 127));
    Big shifted = num;
    for (int i = 0; i < 128; ++i) BigShl1(&shifted);
    Big q, rem;
    Big::DivMod(shifted, den, &q, &rem);
    r = q.IsZero() ? 0.0 : BigToDouble(q, true, -128);
  } else {
    r = BigToDouble(fn == 1 ? sum : dif, sticky, a.e2 - 254);
  }
  return neg ? -r : r;
}

// ---- atan / atan2 / asin / acos: the kernel's CORDIC (device/gdv_device_lib.cuh), step for step
// The vector (X, Y) is rotated onto the x axis through the angles atan(2^-i), i = 0..119, accumulated
// in Q2.126 (table derived by tools/derive_trig_constants.py; below i = 42 the angle is 2^-i itself):
// absolute error < 2^-118, relative < 2^-86 for the smallest angles that come here (2^-32; smaller
// ones are y / x).  asin / acos feed it (sqrt(1 - v^2), v) with the square root taken exactly
// (digit by digit) of the 248-bit integer 2^248 - V^2.  One nearest-even rounding of the 128-bit angle.
const uint64_t kAtanTab[43][2] = {
      {0x3243f6a8885a308dull, 0x313198a2e0370734ull},
      {0x1dac670561bb4f68ull, 0xadfc88bd978751a0ull},
      {0x0fadbafc96406eb1ull, 0x56dc79ef5f7a217eull},
      {0x07f56ea6ab0bdb71ull, 0x9644bcc4f9f44477ull},
      {0x03feab76e59fbd38ull, 0xdb2c9e4b7038b835ull},
      {0x01ffd55bba97624aull, 0x84ef3aeedbb518c4ull},
      {0x00fffaaadddb94d5ull, 0xbbe78c564015f760ull},
      {0x007fff5556eeea5cull, 0xb40311a8fddf3057ull},
      {0x003fffeaaab7776eull, 0x52ec4abedadb53dfull},
      {0x001ffffd5555bbbbull, 0xa9729ab7aac08947ull},
      {0x000fffffaaaaadddull, 0xddb94b968067ef3aull},
      {0x0007fffff555556eull, 0xeeeea5ca5d895892ull},
      {0x0003fffffeaaaaabull, 0x777776e52e5356f5ull},
      {0x0001ffffffd55555ull, 0x5bbbbbba972972d0ull},
      {0x0000fffffffaaaaaull, 0xaadddddddb94b94bull},
      {0x00007fffffff5555ull, 0x5556eeeeeeea5ca5ull},
      {0x00003fffffffeaaaull, 0xaaaab77777776e52ull},
      {0x00001ffffffffd55ull, 0x555555bbbbbbbba9ull},
      {0x00000fffffffffaaull, 0xaaaaaaadddddddddull},
      {0x000007fffffffff5ull, 0x555555556eeeeeeeull},
      {0x000003fffffffffeull, 0xaaaaaaaaab777777ull},
      {0x000001ffffffffffull, 0xd5555555555bbbbbull},
      {0x000000ffffffffffull, 0xfaaaaaaaaaaaddddull},
      {0x0000007fffffffffull, 0xff555555555556eeull},
      {0x0000003fffffffffull, 0xffeaaaaaaaaaaab7ull},
      {0x0000001fffffffffull, 0xfffd555555555555ull},
      {0x0000000fffffffffull, 0xffffaaaaaaaaaaaaull},
      {0x00000007ffffffffull, 0xfffff55555555555ull},
      {0x00000003ffffffffull, 0xfffffeaaaaaaaaaaull},
      {0x00000001ffffffffull, 0xffffffd555555555ull},
      {0x00000000ffffffffull, 0xfffffffaaaaaaaaaull},
      {0x000000007fffffffull, 0xffffffff55555555ull},
      {0x000000003fffffffull, 0xffffffffeaaaaaaaull},
      {0x000000001fffffffull, 0xfffffffffd555555ull},
      {0x000000000fffffffull, 0xffffffffffaaaaaaull},
      {0x0000000007ffffffull, 0xfffffffffff55555ull},
      {0x0000000003ffffffull, 0xfffffffffffeaaaaull},
      {0x0000000001ffffffull, 0xffffffffffffd555ull},
      {0x0000000000ffffffull, 0xfffffffffffffaaaull},
      {0x00000000007fffffull, 0xffffffffffffff55ull},
      {0x00000000003fffffull, 0xffffffffffffffeaull},
      {0x00000000001fffffull, 0xfffffffffffffffdull},
      {0x00000000000fffffull, 0xffffffffffffffffull}};
// atan(y0 / x0) in Q2.126 for 0 <= x0, y0 < 2^125, not both zero
__int128 OrcCordicAtan(unsigned __int128 x0, unsigned __int128 y0) {
  __int128 X = (__int128)x0, Y = (__int128)y0, Z = 0;
  for (int32_t i = 0; i < 120; ++i) {
    const __int128 dx = X >> i, dy = Y >> i;
    const __int128 a = i < 43 ? (__int128)(((unsigned __int128)kAtanTab[i][0] << 64) | (unsigned __int128)kAtanTab[i][1]) : (__int128)1 << (126 - i);
    if (Y > 0) {
      X += dy;
      Y -= dx;
      Z += a;
    } else {
      X -= dy;
      Y += dx;
      Z -= a;
    }
  }
  return Z < 0 ? (__int128)0 : Z;
}
unsigned __int128 OrcPiQ126() { return ((unsigned __int128)0xc90fdaa22168c234ull << 64) | (unsigned __int128)0xc4c6628b80dc1cd1ull; }
double OrcAngleToF64(unsigned __int128 z, bool neg) {  // Q2.126 -> double
  if (z == 0) return neg ? -0.0 : 0.0;
  const double r = BigToDouble(Big::From(z), true, -126);
  return neg ? -r : r;
}
// a finite nonzero double as M * 2^e with M in [2^52, 2^53)
void OrcSplitF64(uint64_t abits, uint64_t* m, int32_t* e) {
  const int32_t ex = (int32_t)(abits >> 52);
  uint64_t mm = abits & 0x000fffffffffffffull;
  int32_t ee = (ex == 0 ? 1 : ex) - 1075;
  if (ex != 0) {
    mm |= 0x0010000000000000ull;
  } else {
    while ((mm >> 52) == 0ull) {
      mm <<= 1;
      --ee;
    }
  }
  *m = mm;
  *e = ee;
}
double OrcAtan2(double y, double x) {
  const uint64_t yb = F64Bits(y), xb = F64Bits(x);
  const uint64_t ya = yb & 0x7fffffffffffffffull, xa = xb & 0x7fffffffffffffffull, inf = 0x7ff0000000000000ull;
  const bool yneg = (yb >> 63) != 0ull, xneg = (xb >> 63) != 0ull;
  if (ya > inf || xa > inf) return F64FromBits(0x7ff8000000000000ull);
  const double pi = 3.141592653589793, pi_lo = 1.2246467991473532e-16, pio2 = 1.5707963267948966, pio2_lo = 6.123233995736766e-17;
  double r;
  if (ya == 0ull) r = xneg ? pi : 0.0;
  else if (xa == 0ull) r = pio2;
  else if (ya == inf) r = xa == inf ? (xneg ? 2.356194490192345 : 0.7853981633974483) : pio2;
  else if (xa == inf) r = xneg ? pi : 0.0;
  else {
    uint64_t my, mx;
    int32_t ey, ex;
    OrcSplitF64(ya, &my, &ey);
    OrcSplitF64(xa, &mx, &ex);
    const int32_t d = ey - ex;
    if (d > 70) {
      const double t = x / (yneg ? -y : y);  // signed, tiny
      r = pio2 + (pio2_lo - t);
    } else if (d < -32) {
      const double t = (yneg ? -y : y) / (xneg ? -x : x);
      r = xneg ? pi + (pi_lo - t) : t;
    } else {
      unsigned __int128 X0 = (unsigned __int128)mx << 71, Y0 = (unsigned __int128)my << 71;
      if (d > 0) X0 >>= d;
      else Y0 >>= -d;
      unsigned __int128 z = (unsigned __int128)OrcCordicAtan(X0, Y0);
      if (xneg) z = OrcPiQ126() - z;
      r = OrcAngleToF64(z, false);
    }
  }
  return yneg ? -r : r;
}
double OrcAtan(double v) { return OrcAtan2(v, 1.0); }
// fn: 0 asin, 1 acos
double OrcAsinAcos(double v, int32_t fn) {
  const uint64_t vb = F64Bits(v), va = vb & 0x7fffffffffffffffull;
  const bool neg = (vb >> 63) != 0ull;
  if (va > 0x3ff0000000000000ull) return F64FromBits(0x7ff8000000000000ull);  // |v| > 1, nan
  const double pi = 3.141592653589793, pio2 = 1.5707963267948966, pio2_lo = 6.123233995736766e-17;
  if (va < 0x3e10000000000000ull) return fn == 0 ? v : pio2 + (pio2_lo - v);  // |v| < 2^-30
  uint64_t mv;
  int32_t ev;
  OrcSplitF64(va, &mv, &ev);
  const unsigned __int128 V = (unsigned __int128)mv << (124 + ev);  // |v| in Q0.124
  // W = 2^248 - V^2, S = floor(sqrt(W)): sqrt(1 - v^2) in Q0.124
  const unsigned __int128 sq_hi = MulShr(V, V, 128), sq_lo = MulShr(V, V, 0);
  const unsigned __int128 w_lo = (unsigned __int128)0 - sq_lo;
  const unsigned __int128 w_hi = ((unsigned __int128)1 << 120) - sq_hi - (sq_lo != 0 ? 1u : 0u);
  unsigned __int128 res = 0, rem = 0;
  for (int32_t i = 123; i >= 0; --i) {
    const int32_t bit = 2 * i;  // the pair (bit + 1, bit) of W
    const unsigned __int128 pair = bit >= 128 ? (w_hi >> (bit - 128)) & 3u : (w_lo >> bit) & 3u;
    rem = (rem << 2) | pair;
    const unsigned __int128 trial = (res << 2) | 1u;
    if (rem >= trial) {
      rem -= trial;
      res = (res << 1) | 1u;
    } else {
      res <<= 1;
    }
  }
  if (res == 0) {  // |v| = 1
    if (fn == 0) return neg ? -pio2 : pio2;
    return neg ? pi : 0.0;
  }
  if (fn == 0) return OrcAngleToF64((unsigned __int128)OrcCordicAtan(res, V), neg);
  unsigned __int128 z = (unsigned __int128)OrcCordicAtan(V, res);
  if (neg) z = OrcPiQ126() - z;
  return OrcAngleToF64(z, false);
}
double OrcAsin(double v) { return OrcAsinAcos(v, 0); }
double OrcAcos(double v) { return OrcAsinAcos(v, 1); }

// ---- castVARCHAR(float32 / float64): shortest round-trip digits found with the C library (printf gives
// the nearest p-digit decimal, strtod / strtof read it back) -- nothing like the kernel's fixed-point
// interval test -- then Java's Double.toString layout, which the reference's formatter follows.
std::string OrcFloatText(double v, bool is_float) {
  if (v != v) return "NaN";
  const bool neg = std::signbit(v);
  const double a = std::fabs(v);
  std::string out = neg ? "-" : "";
  if (std::isinf(a)) return out + "Infinity";
  if (a == 0.0) return out + "0.0";
  auto reads_back = [&](const std::string& text) {
    return is_float ? static_cast<double>(std::strtof(text.c_str(), nullptr)) == a : std::strtod(text.c_str(), nullptr) == a;
  };
  auto read = [&](const std::string& text) {
    return is_float ? static_cast<double>(std::strtof(text.c_str(), nullptr)) : std::strtod(text.c_str(), nullptr);
  };
  std::string digits;
  int x10 = 0;
  for (int p = 1; p <= 17 && digits.empty(); ++p) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*e", p - 1, a);
    std::string mant;
    for (const char* q = buf; *q && *q != 'e'; ++q) if (*q != '.') mant.push_back(*q);
    const int x = std::atoi(std::strchr(buf, 'e') + 1);
    if (reads_back(buf)) { digits = mant; x10 = x; break; }
    // the nearest p-digit decimal reads back as a neighbour: try the next one on the other side of v
    unsigned long long d = std::strtoull(mant.c_str(), nullptr, 10);
    unsigned long long top = 1;
    for (int k = 0; k < p; ++k) top *= 10;
    int xo = x;
    if (read(buf) > a) {
      if (d == top / 10) continue;   // would drop to p - 1 digits, already tried
      d -= 1;
    } else {
      d += 1;
      if (d == top) { d = top / 10; xo += 1; }
    }
    char other[64];
    std::snprintf(other, sizeof other, "%llue%d", d, xo - (p - 1));
    if (reads_back(other)) { digits = std::to_string(d); x10 = xo; }
  }
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int nd = static_cast<int>(digits.size());
  if (x10 >= -3 && x10 < 7) {
    if (x10 >= 0) {
      std::string ip = digits.substr(0, std::min(nd, x10 + 1));
      ip.append(static_cast<size_t>(x10 + 1 - static_cast<int>(ip.size())), '0');
      out += ip + "." + (nd > x10 + 1 ? digits.substr(static_cast<size_t>(x10 + 1)) : "0");
    } else {
      out += "0." + std::string(static_cast<size_t>(-x10 - 1), '0') + digits;
    }
  } else {
    out += digits.substr(0, 1) + "." + (nd > 1 ? digits.substr(1) : "0") + "E" + std::to_string(x10);
  }
  return out;
}

// SQL-style short names of the date-part functions (registered as aliases, csrc/gdv_registry.cc).
static const std::string& CanonicalName(const std::string& name) {
  static const std::map<std::string, std::string> kAlias = {
      {"year", "extractYear"},       {"month", "extractMonth"},       {"day", "extractDay"},
      {"dayofmonth", "extractDay"},  {"hour", "extractHour"},         {"minute", "extractMinute"},
      {"second", "extractSecond"},   {"dayofyear", "extractDoy"},     {"dayofweek", "extractDow"},
      {"quarter", "extractQuarter"}, {"weekofyear", "extractWeek"},   {"yearweek", "extractWeek"}};
  const auto it = kAlias.find(name);
  return it == kAlias.end() ? name : it->second;
}

void ApplyFunction(const Node& n, EvalCtx& cx, int64_t row, Val* out) {
  const std::string& f = CanonicalName(n.name);
  const size_t na = n.kids.size();
  if (f == "concat" || f == "concatOperator") {
    // concat: null arguments are empty strings, never null; concatOperator: null if any is null
    out->ok = true;
    out->s.clear();
    for (const auto& k : n.kids) {
      Val v;
      Eval(*k, cx, row, &v);
      if (!v.ok) {
        if (f == "concatOperator") out->ok = false;
        continue;
      }
      out->s += v.s;
    }
    if (!out->ok) out->s.clear();
    return;
  }
  Val a[6];
  const bool is_like = f == "like" || f == "ilike";
  for (size_t k = 0; k < na && k < 6; ++k) {
    // LIKE patterns are literals handled at parse time
    if (is_like && k >= 1) break;
    Eval(*n.kids[k], cx, row, &a[k]);
  }
  const Type& rt = n.type;
  const Type& t0 = n.kids[0]->type;

  // ---- to_date(text, format [, suppress_errors]) ---------------------------------------------------------
  // The reference's holder (from memory; unpinned): translate the SQL-style format into a strptime format,
  // parse with strptime, allow trailing characters, ignore the time of day, year / month / max(day, 1) at
  // midnight.  Restated here as a walk over the FORMAT TEXT itself (no compiled program: that is the product's
  // way, csrc/gdv_datefmt.cc), with glibc's field rules; tests/test_oracle_vs_arrow.py referees against libc.
  if (f == "to_date" && t0.id == T_STRING) {
    out->ok = false;
    if (!a[0].ok) return;
    const std::string& text = a[0].s;
    const std::string& fmt = n.kids[1]->lit.s;
    const bool suppress = na == 3 && n.kids[2]->lit.i != 0;
    size_t pos = 0, fp = 0;
    auto is_space = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    auto number = [&](int lo, int hi, int width, int* v) {
      while (pos < text.size() && is_space(static_cast<unsigned char>(text[pos]))) ++pos;
      if (pos >= text.size() || text[pos] < '0' || text[pos] > '9') return false;
      int val = 0;
      do {
        val = val * 10 + (text[pos] - '0');
        ++pos;
      } while (--width > 0 && val * 10 <= hi && pos < text.size() && text[pos] >= '0' && text[pos] <= '9');
      *v = val;
      return val >= lo && val <= hi;
    };
    auto word = [&](const std::string& w) {   // case-insensitive, does not consume
      if (pos + w.size() > text.size()) return false;
      for (size_t k = 0; k < w.size(); ++k)
        if (std::tolower(static_cast<unsigned char>(text[pos + k])) != w[k]) return false;
      return true;
    };
    auto token = [&](const char* t) {   // case-insensitive match of a format token at fp
      size_t k = 0;
      while (t[k] != 0 && fp + k < fmt.size() && std::toupper(static_cast<unsigned char>(fmt[fp + k])) == t[k]) ++k;
      if (t[k] != 0) return false;
      fp += k;
      return true;
    };
    static const char* kMonths[12] = {"january", "february", "march", "april", "may", "june", "july", "august",
                                      "september", "october", "november", "december"};
    int year = 1900, mon = 1, day = 0, tmp = 0;
    bool good = true;
    while (good && fp < fmt.size()) {
      const unsigned char fc = static_cast<unsigned char>(fmt[fp]);
      if (is_space(fc)) {
        while (pos < text.size() && is_space(static_cast<unsigned char>(text[pos]))) ++pos;
        ++fp;
      } else if (token("YYYY")) {
        good = number(0, 9999, 4, &year);
      } else if (token("YY")) {
        good = number(0, 99, 2, &tmp);
        year = tmp >= 69 ? 1900 + tmp : 2000 + tmp;
      } else if (token("MONTH") || token("MON")) {
        good = false;
        for (int m = 0; m < 12 && !good; ++m) {
          const std::string full = kMonths[m];
          if (word(full)) { pos += full.size(); good = true; }
          else if (word(full.substr(0, 3))) { pos += 3; good = true; }
          if (good) mon = m + 1;
        }
      } else if (token("MM")) {
        good = number(1, 12, 2, &mon);
      } else if (token("MI")) {
        good = number(0, 59, 2, &tmp);
      } else if (token("DD")) {
        good = number(1, 31, 2, &day);
      } else if (token("HH24")) {
        good = number(0, 23, 2, &tmp);
      } else if (token("HH12") || token("HH")) {
        good = number(1, 12, 2, &tmp);
      } else if (token("SS")) {
        good = number(0, 61, 2, &tmp);
      } else if (token("AM") || token("PM")) {
        good = word("am") || word("pm");
        pos += 2;
      } else {
        good = pos < text.size() && static_cast<unsigned char>(text[pos]) == fc;
        ++pos;
        ++fp;
      }
    }
    if (!good) {
      if (!suppress) cx.error = 14;
      return;
    }
    if (day < 1) day = 1;
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    const int64_t y1 = static_cast<int64_t>(year) - 1;
    int64_t dn = y1 * 365 + FloorDiv(y1, 4) - FloorDiv(y1, 100) + FloorDiv(y1, 400) - 719162;
    for (int m = 1; m < mon; ++m) dn += mdays[m - 1] + ((m == 2 && IsLeap(year)) ? 1 : 0);
    dn += day - 1;
    out->ok = true;
    out->i = DaysToMs(dn);
    return;
  }

  // ---- never-null functions -------------------------------------------------------------
  if (f == "nvl") { *out = a[0].ok ? a[0] : a[1]; return; }
  if (f == "isnull") { out->ok = true; out->b = !a[0].ok; return; }
  if (f == "isnotnull") { out->ok = true; out->b = a[0].ok; return; }
  if (f == "istrue") { out->ok = true; out->b = a[0].ok && a[0].b; return; }
  if (f == "isfalse") { out->ok = true; out->b = a[0].ok && !a[0].b; return; }
  if (f == "isnottrue") { out->ok = true; out->b = !(a[0].ok && a[0].b); return; }
  if (f == "isnotfalse") { out->ok = true; out->b = !(a[0].ok && !a[0].b); return; }
  if (f == "is_distinct_from" || f == "is_not_distinct_from") {
    bool distinct;
    if (a[0].ok != a[1].ok) distinct = true;
    else if (!a[0].ok) distinct = false;
    else if (t0.id == T_BOOL) distinct = a[0].b != a[1].b;
    else if (t0.id == T_FLOAT) distinct = a[0].f != a[1].f;
    else if (t0.id == T_DOUBLE) distinct = a[0].d != a[1].d;
    else if (t0.is_unsigned_int()) distinct = a[0].u != a[1].u;
    else distinct = a[0].i != a[1].i;
    out->ok = true;
    out->b = (f == "is_distinct_from") ? distinct : !distinct;
    return;
  }

  if (f == "hash" || f == "hash32" || f == "hash32AsDouble" || f == "hash64" || f == "hash64AsDouble") {
    const bool is64 = f == "hash64" || f == "hash64AsDouble";
    const int64_t seed = (na >= 2 && a[1].ok) ? a[1].i : 0;
    out->ok = true;
    if (!a[0].ok) { out->i = seed; return; }
    std::string bytes;
    if (t0.is_string()) {
      bytes = a[0].s;
    } else {
      double d;
      if (t0.id == T_BOOL) d = a[0].b ? 1.0 : 0.0;
      else if (t0.id == T_FLOAT) d = static_cast<double>(a[0].f);
      else if (t0.id == T_DOUBLE) d = a[0].d;
      else if (t0.is_unsigned_int()) d = static_cast<double>(a[0].u);
      else d = static_cast<double>(a[0].i);
      bytes.assign(reinterpret_cast<const char*>(&d), 8);
    }
    const unsigned char* p = reinterpret_cast<const unsigned char*>(bytes.data());
    const int32_t seed32 = static_cast<int32_t>(static_cast<uint32_t>(static_cast<uint64_t>(seed)));
    if (is64)
      out->i = static_cast<int64_t>(Murmur3_x64_128_lo(p, static_cast<int>(bytes.size()),
                                                       static_cast<uint64_t>(static_cast<int64_t>(seed32))));
    else
      out->i = static_cast<int32_t>(Murmur3_x86_32(p, static_cast<int>(bytes.size()),
                                                   static_cast<uint32_t>(seed32)));
    return;
  }

  // ---- null-if-null functions -------------------------------------------------------------
  bool ok = true;
  for (size_t k = 0; k < na && k < 6; ++k) {
    if (is_like && k >= 1) break;
    ok = ok && a[k].ok;
  }
  out->ok = ok;
  if (!ok) return;

  if (f == "add" || f == "subtract" || f == "multiply") {
    if (rt.id == T_DECIMAL) {
      const Type& t1 = n.kids[1]->type;
      if (f == "multiply") out->dec = DecimalMultiply(a[0].dec, t0.scale, a[1].dec, t1.scale, rt.scale);
      else out->dec = DecimalAddSub(a[0].dec, t0.scale, a[1].dec, t1.scale, rt.scale, f == "subtract");
    } else if (rt.id == T_FLOAT) {
      out->f = f == "add" ? a[0].f + a[1].f : (f == "subtract" ? a[0].f - a[1].f : a[0].f * a[1].f);
    } else if (rt.id == T_DOUBLE) {
      out->d = f == "add" ? a[0].d + a[1].d : (f == "subtract" ? a[0].d - a[1].d : a[0].d * a[1].d);
    } else if (rt.is_unsigned_int()) {
      uint64_t r = f == "add" ? a[0].u + a[1].u : (f == "subtract" ? a[0].u - a[1].u : a[0].u * a[1].u);
      out->u = WrapUnsigned(r, rt.bits());
    } else {
      const uint64_t x = static_cast<uint64_t>(a[0].i), y = static_cast<uint64_t>(a[1].i);
      uint64_t r = f == "add" ? x + y : (f == "subtract" ? x - y : x * y);
      out->i = WrapSigned(static_cast<int64_t>(r), rt.bits());
    }
    return;
  }
  if (f == "divide" && rt.id == T_DECIMAL) {
    bool dz = false;
    out->dec = DecimalDivide(a[0].dec, t0.scale, a[1].dec, n.kids[1]->type.scale, rt.scale, &dz);
    if (dz) cx.error = 1;
    return;
  }
  if ((f == "mod" || f == "modulo") && rt.id == T_DECIMAL) {
    bool dz = false;
    out->dec = DecimalMod(a[0].dec, t0.scale, a[1].dec, n.kids[1]->type.scale, rt.scale, &dz);
    if (dz) cx.error = 1;
    return;
  }
  if (f == "divide") {
    if (rt.id == T_FLOAT) {
      if (a[1].f == 0.0f) { cx.error = 1; return; }
      out->f = a[0].f / a[1].f;
    } else if (rt.id == T_DOUBLE) {
      if (a[1].d == 0.0) { cx.error = 1; return; }
      out->d = a[0].d / a[1].d;
    } else if (rt.is_unsigned_int()) {
      if (a[1].u == 0) { cx.error = 1; return; }
      out->u = a[0].u / a[1].u;
    } else {
      if (a[1].i == 0) { cx.error = 1; return; }
      if (a[1].i == -1) out->i = WrapSigned(static_cast<int64_t>(0 - static_cast<uint64_t>(a[0].i)), rt.bits());
      else out->i = a[0].i / a[1].i;
    }
    return;
  }
  if ((f == "mod" || f == "modulo") && t0.id == T_DOUBLE) {
    if (a[1].d == 0.0) { cx.error = 1; return; }
    const double r = std::fmod(a[0].d, a[1].d);
    out->d = r != r ? F64FromBits(0x7ff8000000000000ull) : r;
    return;
  }
  if (f == "mod" || f == "modulo") {
    const int64_t x = a[0].i, y = a[1].i;
    int64_t r;
    if (y == 0) r = x;
    else if (y == -1) r = 0;
    else r = x % y;
    out->i = WrapSigned(r, rt.bits());
    return;
  }
  if (f == "div") {
    if (a[1].i == 0) { cx.error = 1; return; }
    if (a[1].i == -1) out->i = WrapSigned(static_cast<int64_t>(0 - static_cast<uint64_t>(a[0].i)), rt.bits());
    else out->i = a[0].i / a[1].i;
    return;
  }
  if (f == "pmod") {
    const int64_t x = a[0].i, y = a[1].i;
    int64_t r;
    if (y == 0) r = x;
    else if (y == -1) r = 0;
    else {
      r = x % y;
      if (r != 0 && ((r < 0) != (y < 0))) r += y;
    }
    out->i = WrapSigned(r, rt.bits());
    return;
  }
  if (f == "bround") { out->d = std::nearbyint(a[0].d); return; }   // default rounding mode: to nearest, ties to even
  if (f == "factorial") {
    if (a[0].i < 0) { cx.error = 12; return; }
    if (a[0].i > 20) { cx.error = 13; return; }
    int64_t r = 1;
    for (int64_t k = 2; k <= a[0].i; ++k) r *= k;
    out->i = r;
    return;
  }
  if (f == "sign") {
    if (rt.id == T_FLOAT) out->f = a[0].f > 0.0f ? 1.0f : (a[0].f < 0.0f ? -1.0f : a[0].f);
    else if (rt.id == T_DOUBLE) out->d = a[0].d > 0.0 ? 1.0 : (a[0].d < 0.0 ? -1.0 : a[0].d);
    else out->i = (a[0].i > 0) - (a[0].i < 0);
    return;
  }
  if (f == "greatest" || f == "least") {
    const bool g = f == "greatest";
    *out = a[0];
    for (size_t k = 1; k < na; ++k) {
      bool take;
      if (rt.id == T_FLOAT) take = g ? a[k].f > out->f : a[k].f < out->f;
      else if (rt.id == T_DOUBLE) take = g ? a[k].d > out->d : a[k].d < out->d;
      else take = g ? a[k].i > out->i : a[k].i < out->i;
      if (take) *out = a[k];
    }
    out->ok = true;
    return;
  }
  if (f == "abs") {
    if (rt.id == T_FLOAT) out->f = std::fabs(a[0].f);
    else if (rt.id == T_DOUBLE) out->d = std::fabs(a[0].d);
    else if (rt.id == T_DECIMAL) out->dec = a[0].dec < 0 ? static_cast<i128>(0 - static_cast<u128>(a[0].dec)) : a[0].dec;
    else out->i = a[0].i < 0 ? WrapSigned(static_cast<int64_t>(0 - static_cast<uint64_t>(a[0].i)), rt.bits()) : a[0].i;
    return;
  }
  if (f == "negative") {
    if (rt.id == T_FLOAT) out->f = -a[0].f;
    else if (rt.id == T_DOUBLE) out->d = -a[0].d;
    else if (rt.id == T_DECIMAL) out->dec = static_cast<i128>(0 - static_cast<u128>(a[0].dec));
    else out->i = WrapSigned(static_cast<int64_t>(0 - static_cast<uint64_t>(a[0].i)), rt.bits());
    return;
  }
  if (f == "sqrt") { out->d = std::sqrt(a[0].d); return; }
  if (f == "sin" || f == "cos" || f == "tan" || f == "cot") {
    const double x = (t0.id == T_INT32 || t0.id == T_INT64) ? static_cast<double>(a[0].i)
                     : t0.id == T_FLOAT ? static_cast<double>(a[0].f) : a[0].d;
    out->d = OrcTrig(x, f == "sin" ? 0 : f == "cos" ? 1 : f == "tan" ? 2 : 3);
    return;
  }
  if (f == "power" || f == "pow") { out->d = OrcPow(a[0].d, a[1].d); return; }
  if (f == "atan2") { out->d = OrcAtan2(a[0].d, a[1].d); return; }
  if (f == "atan") { out->d = OrcAtan(a[0].d); return; }
  if (f == "asin") { out->d = OrcAsin(a[0].d); return; }
  if (f == "acos") { out->d = OrcAcos(a[0].d); return; }
  if (f == "sinh" || f == "cosh" || f == "tanh") { out->d = OrcHyperbolic(a[0].d, f == "sinh" ? 0 : f == "cosh" ? 1 : 2); return; }
  if (f == "exp") { out->d = OrcExp(a[0].d); return; }
  if (f == "log" && na == 2) {
    const double lb = OrcLog(a[0].d);
    if (lb == 0.0) { cx.error = 1; return; }
    out->d = OrcLog(a[1].d) / lb;
    return;
  }
  if (f == "log" || f == "ln") { out->d = OrcLog(a[0].d); return; }
  if (f == "log10") { out->d = OrcLog10(a[0].d); return; }
  if (f == "cbrt") { out->d = OrcCbrt(a[0].d); return; }
  if (f == "bitwise_and") { out->i = a[0].i & a[1].i; return; }
  if (f == "bitwise_or") { out->i = a[0].i | a[1].i; return; }
  if (f == "bitwise_xor") { out->i = a[0].i ^ a[1].i; return; }
  if (f == "bitwise_not") { out->i = WrapSigned(~a[0].i, rt.bits()); return; }
  if (f == "not") { out->b = !a[0].b; return; }

  if (IsRelop(f)) {
    if (t0.id == T_BOOL) out->b = RelopNum<int>(f, a[0].b, a[1].b);
    else if (t0.id == T_FLOAT) out->b = RelopNum<float>(f, a[0].f, a[1].f);
    else if (t0.id == T_DOUBLE) out->b = RelopNum<double>(f, a[0].d, a[1].d);
    else if (t0.is_unsigned_int()) out->b = RelopNum<uint64_t>(f, a[0].u, a[1].u);
    else if (t0.is_string()) {
      const int c = a[0].s.compare(a[1].s);  // unsigned-char lexicographic, shorter first
      const std::string& x = a[0].s; const std::string& y = a[1].s;
      int cc = std::memcmp(x.data(), y.data(), std::min(x.size(), y.size()));
      if (cc == 0) cc = x.size() < y.size() ? -1 : (x.size() > y.size() ? 1 : 0);
      (void)c;
      out->b = Relop(f, [&] { return cc < 0 ? -1 : (cc > 0 ? 1 : 0); });
    } else if (t0.id == T_DECIMAL) {
      const int c = DecimalCompare(a[0].dec, t0.scale, a[1].dec, n.kids[1]->type.scale);
      out->b = Relop(f, [&] { return c; });
    } else out->b = RelopNum<int64_t>(f, a[0].i, a[1].i);
    return;
  }

  // ---- casts ---------------------------------------------------------------------------
  if ((f == "castINT" || f == "castBIGINT") && t0.id == T_STRING) {
    // optional spaces, optional sign, digits; anything else or out of range -> error 4
    const std::string& str = a[0].s;
    size_t b = 0, e = str.size();
    while (b < e && str[b] == ' ') ++b;
    while (e > b && str[e - 1] == ' ') --e;
    bool neg = false;
    if (b < e && (str[b] == '-' || str[b] == '+')) { neg = str[b] == '-'; ++b; }
    if (b >= e) { cx.error = 4; return; }
    __int128 v = 0;
    for (size_t k = b; k < e; ++k) {
      if (str[k] < '0' || str[k] > '9') { cx.error = 4; return; }
      v = v * 10 + (str[k] - '0');
      if (v > (static_cast<__int128>(1) << 64)) { cx.error = 4; return; }
    }
    if (neg) v = -v;
    const __int128 lo = f == "castINT" ? -2147483648ll : static_cast<__int128>(INT64_MIN);
    const __int128 hi = f == "castINT" ? 2147483647ll : static_cast<__int128>(INT64_MAX);
    if (v < lo || v > hi) { cx.error = 4; return; }
    out->i = static_cast<int64_t>(v);
    return;
  }
  if ((f == "castBIGINT" || f == "castINT" || f == "to_timestamp" || f == "to_time") && (t0.id == T_FLOAT || t0.id == T_DOUBLE)) {
    // round half away from zero (to_timestamp / to_time: milliseconds, truncated), saturate, NaN -> 0
    double x = t0.id == T_FLOAT ? static_cast<double>(a[0].f) : a[0].d;
    const bool is32 = f == "castINT";
    if (f == "to_timestamp" || f == "to_time") x = std::trunc(x * 1000.0);
    int64_t v;
    if (x != x) v = 0;
    else {
      const double r = std::round(x);
      const double top = is32 ? 2147483647.0 : 9223372036854775808.0, bottom = is32 ? -2147483648.0 : -9223372036854775808.0;
      if (r >= top) v = is32 ? INT32_MAX : INT64_MAX;
      else if (r <= bottom) v = is32 ? INT32_MIN : INT64_MIN;
      else v = static_cast<int64_t>(r);
    }
    out->i = f == "to_time" ? v - FloorDiv(v, 86400000) * 86400000 : v;
    return;
  }
  if (f == "to_timestamp" || f == "to_time") {
    const int64_t ms = static_cast<int64_t>(static_cast<uint64_t>(a[0].i) * 1000ull);
    out->i = f == "to_time" ? ms - FloorDiv(ms, 86400000) * 86400000 : ms;
    return;
  }
  if (f == "castBIT" || f == "castBOOLEAN") {
    std::string t = a[0].s;
    while (!t.empty() && t.front() == ' ') t.erase(t.begin());
    while (!t.empty() && t.back() == ' ') t.pop_back();
    for (auto& ch : t) if (ch >= 'A' && ch <= 'Z') ch = static_cast<char>(ch + 32);
    if (t == "1" || t == "true") out->b = true;
    else if (t == "0" || t == "false") out->b = false;
    else cx.error = 9;
    return;
  }
  if (f == "find_in_set") {
    out->i = 0;
    if (a[0].s.find(',') != std::string::npos) return;
    size_t start = 0;
    for (int64_t item = 1;; ++item) {
      const size_t comma = a[1].s.find(',', start);
      const std::string piece = a[1].s.substr(start, comma == std::string::npos ? std::string::npos : comma - start);
      if (piece == a[0].s) { out->i = item; return; }
      if (comma == std::string::npos) return;
      start = comma + 1;
    }
  }
  if (f == "castBIGINT") {
    if (t0.id == T_DECIMAL) {
      const i128 r = DecimalRescale(a[0].dec, t0.scale, 38, 0);
      out->i = static_cast<int64_t>(static_cast<uint64_t>(static_cast<u128>(r)));
    } else out->i = a[0].i;
    return;
  }
  if (f == "castINT") { out->i = WrapSigned(a[0].i, 32); return; }
  if ((f == "castFLOAT8" || f == "castFLOAT4") && t0.id == T_STRING) {
    double v = 0.0;
    if (!ParseDouble(a[0].s, &v)) { cx.error = 8; return; }
    if (f == "castFLOAT8") out->d = v;
    else out->f = static_cast<float>(v);
    return;
  }
  if (f == "castFLOAT4") {
    if (t0.id == T_DOUBLE) out->f = static_cast<float>(a[0].d);
    else out->f = static_cast<float>(a[0].i);
    return;
  }
  if (f == "castFLOAT8") {
    if (t0.id == T_FLOAT) out->d = static_cast<double>(a[0].f);
    else if (t0.id == T_DECIMAL) out->d = DecimalToDouble(a[0].dec, t0.scale);
    else out->d = static_cast<double>(a[0].i);
    return;
  }
  if ((f == "castDATE" || f == "castTIMESTAMP") && t0.id == T_STRING) {
    // [spaces] [-]Y-M-D [(' '|'T') h:m[:s[.frac]]] [spaces]; sscanf-free, field by field
    std::string str = a[0].s;
    while (!str.empty() && str.front() == ' ') str.erase(str.begin());
    while (!str.empty() && str.back() == ' ') str.pop_back();
    size_t pos = 0;
    bool neg = false;
    if (pos < str.size() && str[pos] == '-') { neg = true; ++pos; }
    auto field = [&](int maxd, int64_t* v) {
      int n = 0;
      *v = 0;
      while (pos < str.size() && n < maxd && str[pos] >= '0' && str[pos] <= '9') { *v = *v * 10 + (str[pos] - '0'); ++pos; ++n; }
      return n >= 1;
    };
    auto expect = [&](const char* any) {
      if (pos >= str.size() || std::strchr(any, str[pos]) == nullptr) return false;
      ++pos;
      return true;
    };
    int64_t y = 0, mo = 0, d = 0, hh = 0, mi = 0, ss = 0, ms = 0;
    bool good = field(9, &y) && expect("-") && field(2, &mo) && expect("-") && field(2, &d);
    if (good && pos < str.size()) {
      good = expect(" T") && field(2, &hh) && expect(":") && field(2, &mi);
      if (good && pos < str.size()) {
        good = expect(":") && field(2, &ss);
        if (good && pos < str.size()) {
          good = expect(".");
          int nd = 0;
          while (good && pos < str.size() && str[pos] >= '0' && str[pos] <= '9') {
            if (nd < 3) ms = ms * 10 + (str[pos] - '0');
            ++nd;
            ++pos;
          }
          good = good && nd >= 1;
          for (; nd < 3; ++nd) ms *= 10;
        }
      }
    }
    if (neg) y = -y;
    good = good && pos == str.size() && mo >= 1 && mo <= 12 && d >= 1 && hh < 24 && mi < 60 && ss < 60;
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if (good) good = d <= mdays[mo - 1] + ((mo == 2 && IsLeap(y)) ? 1 : 0);
    if (!good) { cx.error = 5; return; }
    const int64_t y1 = y - 1;
    int64_t dn = y1 * 365 + FloorDiv(y1, 4) - FloorDiv(y1, 100) + FloorDiv(y1, 400) - 719162;
    for (int m = 1; m < mo; ++m) dn += mdays[m - 1] + ((m == 2 && IsLeap(y)) ? 1 : 0);
    dn += d - 1;
    out->i = DaysToMs(dn);
    if (f == "castTIMESTAMP") out->i += ((hh * 60 + mi) * 60 + ss) * 1000 + ms;
    return;
  }
  if (f == "castDATE") {
    if (t0.id == T_TIMESTAMP) out->i = DaysToMs(FloorDiv(a[0].i, 86400000));
    else out->i = a[0].i;
    return;
  }
  if (f == "castTIMESTAMP") { out->i = a[0].i; return; }
  if (f == "castDECIMAL" && t0.id == T_STRING) {
    // [spaces][+-]digits[.digits][spaces]; half away from zero at the declared scale
    std::string str = a[0].s;
    while (!str.empty() && str.front() == ' ') str.erase(str.begin());
    while (!str.empty() && str.back() == ' ') str.pop_back();
    bool neg = false;
    size_t pos = 0;
    if (pos < str.size() && (str[pos] == '-' || str[pos] == '+')) { neg = str[pos] == '-'; ++pos; }
    std::string ip, fp;
    bool point = false, good = true;
    for (; pos < str.size(); ++pos) {
      if (str[pos] == '.') { good = good && !point; point = true; }
      else if (str[pos] >= '0' && str[pos] <= '9') (point ? fp : ip).push_back(str[pos]);
      else good = false;
    }
    if (!good || ip.size() + fp.size() == 0) { cx.error = 7; return; }
    const size_t scale = static_cast<size_t>(rt.scale);
    const bool up = fp.size() > scale && fp[scale] >= '5';
    fp.resize(scale, '0');   // cut or pad to exactly `scale` fractional digits
    Big mag;
    bool overflow = false;
    for (char ch : ip + fp) {
      if (!mag.MulPow10Checked(1)) overflow = true;
      mag.Add(Big::From(static_cast<u128>(ch - '0')));
    }
    if (up) mag.Add(Big::From(1));
    SignedBig sb{neg, mag};
    out->dec = overflow ? 0 : FromSigned(sb, rt.precision);
    return;
  }
  if (f == "castDECIMAL") {
    if (t0.id == T_DECIMAL) out->dec = DecimalRescale(a[0].dec, t0.scale, rt.precision, rt.scale);
    else if (t0.id == T_DOUBLE) out->dec = DecimalFromDouble(a[0].d, rt.precision, rt.scale);
    else if (t0.id == T_FLOAT) out->dec = DecimalFromDouble(static_cast<double>(a[0].f), rt.precision, rt.scale);
    else out->dec = DecimalRescale(static_cast<i128>(a[0].i), 0, rt.precision, rt.scale);
    return;
  }

  // ---- rounding ---------------------------------------------------------------------------
  if (rt.id == T_DECIMAL && (f == "round" || f == "truncate" || f == "trunc" || f == "ceil" || f == "floor")) {
    const int mode = f == "round" ? 0 : (f == "ceil" ? 2 : (f == "floor" ? 3 : 1));
    out->dec = DecimalRoundTo(a[0].dec, t0.scale, na == 2 ? a[1].i : 0, mode, rt.precision, rt.scale);
    return;
  }
  if ((f == "round" || f == "truncate" || f == "trunc") && na == 2 && rt.id != T_DOUBLE) {
    // integers: to a multiple of 10^-s for s < 0, in 128 bits, wrapped into the output type
    const int64_t sc = a[1].i;
    __int128 x = a[0].i, res;
    if (sc >= 0) res = x;
    else if (sc < -38) res = 0;
    else {
      __int128 p = 1;
      for (int64_t k = 0; k < -sc; ++k) p *= 10;
      const __int128 r = x % p;
      res = x - r;
      if (f == "round") {
        const __int128 ar = r < 0 ? -r : r;
        if (ar >= p - ar) res += x < 0 ? -p : p;
      }
    }
    out->i = WrapSigned(static_cast<int64_t>(static_cast<uint64_t>(static_cast<unsigned __int128>(res))), rt.bits());
    return;
  }
  if ((f == "truncate" || f == "trunc") && na == 2) {
    const int s = static_cast<int>(std::max<int64_t>(-308, std::min<int64_t>(308, a[1].i)));
    double p = 1.0;
    for (int k = 0; k < std::abs(s); ++k) p = p * 10.0;
    if (s >= 0) {
      const double v = a[0].d * p;
      if (!(std::fabs(v) < 1.7976931348623157e308) || v == std::floor(v)) out->d = a[0].d;
      else out->d = std::trunc(v) / p;
    } else {
      const double q = a[0].d / p;
      out->d = (q == std::floor(q)) ? a[0].d : std::trunc(q) * p;
    }
    return;
  }
  if (f == "round") {
    if (na == 2) {
      const int s = static_cast<int>(std::max<int64_t>(-308, std::min<int64_t>(308, a[1].i)));
      double p = 1.0;
      const int e = std::abs(s);
      for (int k = 0; k < e; ++k) p = p * 10.0;
      if (s >= 0) {
        const double v = a[0].d * p;
        if (!(std::fabs(v) < 1.7976931348623157e308) || v == std::floor(v)) out->d = a[0].d;
        else out->d = std::round(v) / p;
      } else {
        const double q = a[0].d / p;
        out->d = (q == std::floor(q)) ? a[0].d : std::round(q) * p;
      }
    } else if (rt.id == T_DOUBLE) out->d = std::round(a[0].d);
    else if (rt.id == T_FLOAT) out->f = std::round(a[0].f);
    else out->i = a[0].i;
    return;
  }
  if (f == "ceil") { out->d = std::ceil(a[0].d); return; }
  if (f == "floor") { out->d = std::floor(a[0].d); return; }
  if (f == "truncate" || f == "trunc") { out->d = std::trunc(a[0].d); return; }

  // ---- date/time arithmetic -----------------------------------------------------------------
  if (f.rfind("timestampadd", 0) == 0 || f == "add_months") {
    const std::string unit = f == "add_months" ? std::string("Month") : f.substr(12);
    // (count, timestamp) or (timestamp, count)
    const bool ts_first = t0.id == T_TIMESTAMP || t0.id == T_DATE64;
    const int64_t n = ts_first ? a[1].i : a[0].i, ts = ts_first ? a[0].i : a[1].i;
    int64_t unit_ms = 0;
    if (unit == "Second") unit_ms = 1000;
    else if (unit == "Minute") unit_ms = 60000;
    else if (unit == "Hour") unit_ms = 3600000;
    else if (unit == "Day") unit_ms = 86400000;
    else if (unit == "Week") unit_ms = 604800000;
    if (unit_ms != 0) {
      out->i = static_cast<int64_t>(static_cast<uint64_t>(ts) + static_cast<uint64_t>(n) * static_cast<uint64_t>(unit_ms));
      return;
    }
    const int64_t months = static_cast<int64_t>(static_cast<uint64_t>(unit == "Month" ? 1 : (unit == "Quarter" ? 3 : 12)) *
                                                static_cast<uint64_t>(n));
    const int64_t days = FloorDiv(ts, 86400000);
    const int64_t in_day = ts - days * 86400000;
    const Ymd c = CivilFromDays(days);
    const int64_t total = c.y * 12 + (c.m - 1) + months;
    const int64_t ny = FloorDiv(total, 12);
    const int nm = static_cast<int>(total - ny * 12) + 1;
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    const int len = mdays[nm - 1] + ((nm == 2 && IsLeap(ny)) ? 1 : 0);
    const int nd = std::min(c.d, len);
    // days since epoch of (ny, nm, nd): walk from the first day of the year
    int64_t y1 = ny - 1;
    int64_t dn = y1 * 365 + FloorDiv(y1, 4) - FloorDiv(y1, 100) + FloorDiv(y1, 400) - 719162;  // Jan 1 of ny
    for (int m = 1; m < nm; ++m) dn += mdays[m - 1] + ((m == 2 && IsLeap(ny)) ? 1 : 0);
    dn += nd - 1;
    out->i = static_cast<int64_t>(static_cast<uint64_t>(DaysToMs(dn)) + static_cast<uint64_t>(in_day));
    return;
  }
  if (f == "date_add" || f == "date_sub") {
    const uint64_t delta = static_cast<uint64_t>(a[1].i) * 86400000ull;
    out->i = static_cast<int64_t>(f == "date_add" ? static_cast<uint64_t>(a[0].i) + delta
                                                  : static_cast<uint64_t>(a[0].i) - delta);
    return;
  }
  if (f == "timestampdiffMonth" || f == "timestampdiffQuarter" || f == "timestampdiffYear" || f == "months_between") {
    // calendar arithmetic on (year, month, day, ms of day) tuples, Python-style
    struct Cal { int64_t y; int m, d; int64_t tod; };
    auto cal = [&](int64_t ms) {
      const int64_t days = FloorDiv(ms, 86400000);
      const Ymd c = CivilFromDays(days);
      return Cal{c.y, c.m, c.d, ms - days * 86400000};
    };
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    auto mlen = [&](int64_t y, int m) { return mdays[m - 1] + ((m == 2 && IsLeap(y)) ? 1 : 0); };
    const Cal A = cal(a[0].i), B = cal(a[1].i);
    if (f == "months_between") {
      const double months = static_cast<double>((A.y - B.y) * 12 + (A.m - B.m));
      if (A.d == B.d || (A.d == mlen(A.y, A.m) && B.d == mlen(B.y, B.m))) { out->d = months; return; }
      const double secs = static_cast<double>(static_cast<int64_t>(A.d - B.d) * 86400) +
                          static_cast<double>(A.tod - B.tod) / 1000.0;
      out->d = months + secs / 2678400.0;
      return;
    }
    // a + k months as a comparable tuple (day clamped), compared with b lexicographically
    auto plus = [&](int64_t k) {
      const int64_t total = A.y * 12 + (A.m - 1) + k;
      const int64_t y = FloorDiv(total, 12);
      const int m = static_cast<int>(total - y * 12) + 1;
      return Cal{y, m, std::min(A.d, mlen(y, m)), A.tod};
    };
    auto less = [](const Cal& x, const Cal& y) {
      if (x.y != y.y) return x.y < y.y;
      if (x.m != y.m) return x.m < y.m;
      if (x.d != y.d) return x.d < y.d;
      return x.tod < y.tod;
    };
    int64_t k = (B.y - A.y) * 12 + (B.m - A.m);
    if (a[1].i >= a[0].i) {
      if (k > 0 && less(B, plus(k))) --k;
      if (k < 0) k = 0;
    } else {
      if (k < 0 && less(plus(k), B)) ++k;
      if (k > 0) k = 0;
    }
    if (f == "timestampdiffQuarter") k /= 3;
    else if (f == "timestampdiffYear") k /= 12;
    out->i = WrapSigned(k, 32);
    return;
  }
  if (f.rfind("timestampdiff", 0) == 0) {
    const std::string unit = f.substr(13);
    int64_t unit_ms = 1000;
    if (unit == "Minute") unit_ms = 60000;
    else if (unit == "Hour") unit_ms = 3600000;
    else if (unit == "Day") unit_ms = 86400000;
    else if (unit == "Week") unit_ms = 604800000;
    const int64_t diff = static_cast<int64_t>(static_cast<uint64_t>(a[1].i) - static_cast<uint64_t>(a[0].i));
    out->i = WrapSigned(diff / unit_ms, 32);
    return;
  }

  // ---- date/time -------------------------------------------------------------------------
  if (t0.id == T_TIME32 && f.rfind("extract", 0) == 0) {
    const int64_t t = a[0].i;
    if (f == "extractHour") out->i = t / 3600000;
    else if (f == "extractMinute") out->i = (t / 60000) % 60;
    else out->i = (t / 1000) % 60;
    return;
  }
  if (f == "datediff") { out->i = WrapSigned(FloorDiv(a[0].i, 86400000) - FloorDiv(a[1].i, 86400000), 32); return; }
  if (f == "degrees") { out->d = a[0].d * 180.0 / 3.14159265358979323846; return; }
  if (f == "radians") { out->d = a[0].d * 3.14159265358979323846 / 180.0; return; }
  if (f == "castTIME") { out->i = a[0].i - FloorDiv(a[0].i, 86400000) * 86400000; return; }
  if (f == "extractWeek" || f == "extractDecade" || f == "extractCentury" || f == "extractMillennium" ||
      f.rfind("date_trunc_", 0) == 0 || f == "last_day") {
    const int64_t ms = a[0].i;
    const int64_t days = FloorDiv(ms, 86400000);
    const Ymd c = CivilFromDays(days);
    auto jan1_of = [&](int64_t y) {  // days since epoch of January 1st of year y
      const int64_t y1 = y - 1;
      return y1 * 365 + FloorDiv(y1, 4) - FloorDiv(y1, 100) + FloorDiv(y1, 400) - 719162;
    };
    auto monday0 = [](int64_t d) { int64_t w = (d + 3) % 7; return w < 0 ? w + 7 : w; };
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    auto first_of_month = [&](int64_t y, int m) {
      int64_t d = jan1_of(y);
      for (int k = 1; k < m; ++k) d += mdays[k - 1] + ((k == 2 && IsLeap(y)) ? 1 : 0);
      return d;
    };
    if (f == "extractDecade") { out->i = c.y / 10; return; }
    if (f == "extractCentury") { out->i = (c.y - 1) / 100 + 1; return; }
    if (f == "extractMillennium") { out->i = (c.y - 1) / 1000 + 1; return; }
    if (f == "extractWeek") {
      // ISO 8601: the week with the year's first Thursday is week 1; walk it out explicitly
      const int64_t thursday = days - monday0(days) + 3;       // Thursday of this row's week
      const Ymd ty = CivilFromDays(thursday);                   // its year owns the week
      out->i = (thursday - jan1_of(ty.y)) / 7 + 1;
      return;
    }
    if (f == "last_day") {
      const int len = mdays[c.m - 1] + ((c.m == 2 && IsLeap(c.y)) ? 1 : 0);
      out->i = DaysToMs(first_of_month(c.y, c.m) + len - 1);
      return;
    }
    const std::string unit = f.substr(11);
    if (unit == "Second") out->i = FloorDiv(ms, 1000) * 1000;
    else if (unit == "Minute") out->i = FloorDiv(ms, 60000) * 60000;
    else if (unit == "Hour") out->i = FloorDiv(ms, 3600000) * 3600000;
    else if (unit == "Day") out->i = DaysToMs(days);
    else if (unit == "Week") out->i = DaysToMs(days - monday0(days));
    else if (unit == "Month") out->i = DaysToMs(first_of_month(c.y, c.m));
    else if (unit == "Quarter") out->i = DaysToMs(first_of_month(c.y, ((c.m - 1) / 3) * 3 + 1));
    else if (unit == "Year") out->i = DaysToMs(jan1_of(c.y));
    else if (unit == "Decade") out->i = DaysToMs(jan1_of((c.y / 10) * 10));
    else if (unit == "Century") out->i = DaysToMs(jan1_of(((c.y - 1) / 100) * 100 + 1));
    else out->i = DaysToMs(jan1_of(((c.y - 1) / 1000) * 1000 + 1));  // Millennium
    return;
  }
  if (f.rfind("extract", 0) == 0) {
    const bool is_d32 = t0.id == T_DATE32;
    const int64_t ms = is_d32 ? a[0].i * 86400000 : a[0].i;
    const int64_t days = FloorDiv(ms, 86400000);
    const int64_t in_day = ms - days * 86400000;
    const Ymd c = CivilFromDays(days);
    if (f == "extractYear") out->i = c.y;
    else if (f == "extractMonth") out->i = c.m;
    else if (f == "extractDay") out->i = c.d;
    else if (f == "extractDoy") out->i = c.doy;
    else if (f == "extractQuarter") out->i = (c.m - 1) / 3 + 1;
    else if (f == "extractDow") { int64_t w = (days + 4) % 7; if (w < 0) w += 7; out->i = w + 1; }
    else if (f == "extractHour") out->i = in_day / 3600000;
    else if (f == "extractMinute") out->i = (in_day / 60000) % 60;
    else if (f == "extractSecond") out->i = (in_day / 1000) % 60;
    else if (f == "extractEpoch") out->i = FloorDiv(ms, 1000);
    return;
  }

  // ---- strings ---------------------------------------------------------------------------
  if (f == "like") { out->b = LikeRec(a[0].s, 0, n.like, 0); return; }
  if (f == "regexp_matches" || f == "regexp_like") { out->b = OrcReSearch(n.regex.get(), a[0].s); return; }
  if (f == "initcap") {
    // ASCII letters only, like upper / lower: upper-case at the start and after a byte that is not part of a
    // word (ASCII letter / digit, or any byte of a multi-byte glyph), lower-case inside a word
    out->s = a[0].s;
    bool in_word = false;
    for (auto& ch : out->s) {
      const unsigned char c = static_cast<unsigned char>(ch);
      const bool letter = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
      if (letter) ch = static_cast<char>(in_word ? (c | 0x20) : (c & ~0x20));
      in_word = letter || (c >= '0' && c <= '9') || c >= 0x80;
    }
    return;
  }
  if (f == "upper" || f == "lower") {
    out->s = a[0].s;
    for (auto& ch : out->s) {
      if (f == "upper" && ch >= 'a' && ch <= 'z') ch = static_cast<char>(ch - 32);
      if (f == "lower" && ch >= 'A' && ch <= 'Z') ch = static_cast<char>(ch + 32);
    }
    return;
  }
  if (f == "substr" || f == "substring") {
    out->s = na == 3 ? Substr(a[0].s, a[1].i, a[2].i)
                     : Substr(a[0].s, a[1].i, static_cast<int64_t>(a[0].s.size()));
    return;
  }
  if (f == "castVARCHAR" && na == 2 && a[1].i < 0) { cx.error = 10; return; }   // "Output buffer length can't be negative"
  if (f == "castVARCHAR" && t0.id == T_DECIMAL) {
    // unscaled digits with the point `scale` places from the right
    u128 m = a[0].dec < 0 ? (~static_cast<u128>(a[0].dec) + 1) : static_cast<u128>(a[0].dec);
    std::string digits;
    do { digits.insert(digits.begin(), static_cast<char>('0' + static_cast<int>(m % 10))); m /= 10; } while (m != 0);
    const size_t sc = static_cast<size_t>(t0.scale);
    if (digits.size() <= sc) digits.insert(0, sc + 1 - digits.size(), '0');
    if (sc > 0) digits.insert(digits.size() - sc, ".");
    if (a[0].dec < 0) digits.insert(digits.begin(), '-');
    out->s = Substr(digits, 1, a[1].i);
    return;
  }
  if (f == "castVARCHAR" && (t0.id == T_FLOAT || t0.id == T_DOUBLE)) {
    out->s = Substr(OrcFloatText(t0.id == T_FLOAT ? static_cast<double>(a[0].f) : a[0].d, t0.id == T_FLOAT), 1, a[1].i);
    return;
  }
  if (f == "castVARCHAR" && t0.id != T_STRING) {
    // numbers / booleans / dates as text, then the first `len` characters
    char buf[64];
    std::string text;
    if (t0.id == T_BOOL) text = a[0].b ? "true" : "false";
    else if (t0.id == T_DATE64 || t0.id == T_TIMESTAMP) {
      const int64_t days = FloorDiv(a[0].i, 86400000);
      const Ymd c = CivilFromDays(days);
      std::snprintf(buf, sizeof(buf), "%s%04lld-%02d-%02d", c.y < 0 ? "-" : "",
                    static_cast<long long>(c.y < 0 ? -c.y : c.y), c.m, c.d);
      text = buf;
      if (t0.id == T_TIMESTAMP) {
        const int64_t in_day = static_cast<int64_t>(static_cast<uint64_t>(a[0].i) - static_cast<uint64_t>(DaysToMs(days)));
        std::snprintf(buf, sizeof(buf), " %02d:%02d:%02d.%03d", static_cast<int>(in_day / 3600000),
                      static_cast<int>((in_day / 60000) % 60), static_cast<int>((in_day / 1000) % 60),
                      static_cast<int>(in_day % 1000));
        text += buf;
      }
    } else text = std::to_string(static_cast<long long>(a[0].i));
    out->s = Substr(text, 1, a[1].i);
    return;
  }
  if (f == "castVARCHAR") { out->s = Substr(a[0].s, 1, a[1].i); return; }
  if (f == "char_length" || f == "length" || f == "lengthUtf8") {
    out->i = static_cast<int64_t>(GlyphStarts(a[0].s).size());
    return;
  }
  if (f == "octet_length") { out->i = static_cast<int64_t>(a[0].s.size()); return; }
  if (f == "bit_length") { out->i = static_cast<int64_t>(a[0].s.size()) * 8; return; }
  if (f == "starts_with") {
    out->b = a[0].s.size() >= a[1].s.size() && a[0].s.compare(0, a[1].s.size(), a[1].s) == 0;
    return;
  }
  if (f == "ends_with") {
    out->b = a[0].s.size() >= a[1].s.size() &&
             a[0].s.compare(a[0].s.size() - a[1].s.size(), a[1].s.size(), a[1].s) == 0;
    return;
  }
  if (f == "is_substr") { out->b = a[0].s.find(a[1].s) != std::string::npos; return; }
  if ((f == "ltrim" || f == "rtrim" || f == "btrim" || f == "trim") && na == 2) {
    // glyph-wise: strip glyphs of a[0] that are one of the glyphs of a[1]
    const std::vector<size_t> st = GlyphStarts(a[0].s), cs = GlyphStarts(a[1].s);
    auto glyph = [](const std::string& str, const std::vector<size_t>& starts, size_t k) {
      const size_t b = starts[k], e = k + 1 < starts.size() ? starts[k + 1] : str.size();
      return str.substr(b, e - b);
    };
    auto in_set = [&](const std::string& g) {
      for (size_t k = 0; k < cs.size(); ++k)
        if (glyph(a[1].s, cs, k) == g) return true;
      return false;
    };
    size_t lo = 0, hi = st.size();
    if (f != "rtrim") while (lo < hi && in_set(glyph(a[0].s, st, lo))) ++lo;
    if (f != "ltrim") while (hi > lo && in_set(glyph(a[0].s, st, hi - 1))) --hi;
    const size_t b = lo < st.size() ? st[lo] : a[0].s.size();
    const size_t e = hi < st.size() ? st[hi] : a[0].s.size();
    out->s = a[0].s.substr(b, e - b);
    return;
  }
  if (f == "split_part") {
    if (a[2].i < 1) { cx.error = 6; return; }
    std::vector<std::string> pieces;
    if (a[1].s.empty()) pieces.push_back(a[0].s);
    else {
      size_t from = 0;
      while (true) {
        const size_t hit = a[0].s.find(a[1].s, from);
        if (hit == std::string::npos) { pieces.push_back(a[0].s.substr(from)); break; }
        pieces.push_back(a[0].s.substr(from, hit - from));
        from = hit + a[1].s.size();
      }
    }
    out->s = static_cast<size_t>(a[2].i) <= pieces.size() ? pieces[static_cast<size_t>(a[2].i) - 1] : std::string();
    return;
  }
  if (f == "hashSHA256" || f == "sha256") { out->s = Sha256Hex(a[0].s); return; }
  if (f == "hashSHA1" || f == "sha1") { out->s = Sha1Hex(a[0].s); return; }
  if (f == "hashMD5" || f == "md5") { out->s = Md5Hex(a[0].s); return; }
  if (f == "replace") {
    out->s.clear();
    if (a[1].s.empty()) { out->s = a[0].s; return; }
    size_t from = 0;
    while (true) {
      const size_t hit = a[0].s.find(a[1].s, from);
      if (hit == std::string::npos) { out->s += a[0].s.substr(from); break; }
      out->s += a[0].s.substr(from, hit - from) + a[2].s;
      from = hit + a[1].s.size();
    }
    return;
  }
  if (f == "repeat") {
    out->s.clear();
    for (int64_t k = 0; k < a[1].i; ++k) out->s += a[0].s;
    return;
  }
  if (f == "space") { out->s.assign(static_cast<size_t>(std::max<int64_t>(a[0].i, 0)), ' '); return; }
  if (f == "reverse") {
    const std::vector<size_t> st = GlyphStarts(a[0].s);
    out->s.clear();
    for (size_t k = st.size(); k-- > 0;) {
      const size_t e = k + 1 < st.size() ? st[k + 1] : a[0].s.size();
      out->s += a[0].s.substr(st[k], e - st[k]);
    }
    return;
  }
  if (f == "lpad" || f == "rpad") {
    const int64_t want = a[1].i;
    const std::string fill = na == 3 ? a[2].s : std::string(" ");
    out->s.clear();
    if (want <= 0) return;
    const std::string text = Substr(a[0].s, 1, want);
    const int64_t have = static_cast<int64_t>(GlyphStarts(a[0].s).size());
    std::string pad;
    const std::vector<size_t> fs = GlyphStarts(fill);
    if (!fs.empty())
      for (int64_t k = 0; k < want - have; ++k) {
        const size_t g = static_cast<size_t>(k) % fs.size();
        const size_t e = g + 1 < fs.size() ? fs[g + 1] : fill.size();
        pad += fill.substr(fs[g], e - fs[g]);
      }
    out->s = f == "lpad" ? pad + text : text + pad;
    return;
  }
  if (f == "crc32") {
    // table-driven here (the kernel is bitwise): same polynomial, different code
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
      for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[n] = c;
      }
      init = true;
    }
    uint32_t crc = 0xffffffffu;
    for (unsigned char ch : a[0].s) crc = table[(crc ^ ch) & 0xffu] ^ (crc >> 8);
    out->i = static_cast<int64_t>(crc ^ 0xffffffffu);
    return;
  }
  if (f == "to_hex") {
    char buf[32];
    const unsigned long long v = t0.id == T_INT32 ? static_cast<unsigned long long>(static_cast<uint32_t>(a[0].i))
                                                  : static_cast<unsigned long long>(a[0].i);
    std::snprintf(buf, sizeof(buf), "%llX", v);
    out->s = buf;
    return;
  }
  if (f == "ltrim" || f == "rtrim" || f == "btrim" || f == "trim") {
    size_t b = 0, e = a[0].s.size();
    if (f != "rtrim") while (b < e && a[0].s[b] == ' ') ++b;
    if (f != "ltrim") while (e > b && a[0].s[e - 1] == ' ') --e;
    out->s = a[0].s.substr(b, e - b);
    return;
  }
  if (f == "ilike") {
    std::string low = a[0].s;
    for (auto& ch : low)
      if (ch >= 'A' && ch <= 'Z') ch = static_cast<char>(ch + 32);
    out->b = LikeRec(low, 0, n.like, 0);
    return;
  }
  if (f == "ascii") { out->i = a[0].s.empty() ? 0 : static_cast<unsigned char>(a[0].s[0]); return; }
  if (f == "left" || f == "right") {
    const int64_t k = a[1].i;
    const int64_t g = static_cast<int64_t>(GlyphStarts(a[0].s).size());
    if (k == 0) out->s.clear();
    else if (f == "left") out->s = k > 0 ? Substr(a[0].s, 1, k) : (g + k <= 0 ? std::string() : Substr(a[0].s, 1, g + k));
    else if (k < 0) out->s = Substr(a[0].s, 1 - k, g + 1);
    else out->s = k >= g ? a[0].s : Substr(a[0].s, g - k + 1, k);
    return;
  }
  if (f == "locate" || f == "position" || f == "strpos" || f == "instr") {
    const bool text_first = f == "strpos" || f == "instr";
    const std::string& sub = text_first ? a[1].s : a[0].s;
    const std::string& str = text_first ? a[0].s : a[1].s;
    const int64_t start = na == 3 ? a[2].i : 1;
    out->i = 0;
    if (start < 1) { cx.error = 11; return; }   // "Start position must be greater than 0"
    const std::vector<size_t> st = GlyphStarts(str);
    const int64_t g = static_cast<int64_t>(st.size());
    if (start > g + 1) return;
    for (int64_t k = start; k <= g + 1; ++k) {   // candidate glyph positions, incl. one past the end
      const size_t pos = k <= g ? st[static_cast<size_t>(k - 1)] : str.size();
      if (pos + sub.size() <= str.size() && str.compare(pos, sub.size(), sub) == 0) { out->i = k; return; }
    }
    return;
  }
  if (f == "byte_substr" || f == "bytesubstring") {
    const int64_t off = a[1].i, len = a[2].i, size = static_cast<int64_t>(a[0].s.size());
    out->s.clear();
    if (len <= 0 || size == 0) return;
    const int64_t from = off > 0 ? off - 1 : (off < 0 ? size + off : 0);
    if (from < 0 || from >= size) return;
    out->s = a[0].s.substr(static_cast<size_t>(from), static_cast<size_t>(std::min(len, size - from)));
    return;
  }
  cx.error = 100;  // unknown function
}

void Eval(const Node& n, EvalCtx& cx, int64_t row, Val* out) {
  switch (n.kind) {
    case K_FIELD: LoadField(n, cx, row, out); return;
    case K_LIT: *out = n.lit; return;
    case K_FN: ApplyFunction(n, cx, row, out); return;
    case K_IF: {
      Val c;
      Eval(*n.kids[0], cx, row, &c);
      Eval(*n.kids[(c.ok && c.b) ? 1 : 2], cx, row, out);
      return;
    }
    case K_AND: case K_OR: {
      // SQL three-valued logic with left-to-right short circuit
      const bool is_and = n.kind == K_AND;
      bool all_ok = true;
      for (const auto& k : n.kids) {
        Val c;
        Eval(*k, cx, row, &c);
        if (c.ok && (is_and ? !c.b : c.b)) {
          out->ok = true;
          out->b = !is_and;
          return;
        }
        all_ok = all_ok && c.ok;
      }
      out->ok = all_ok;
      out->b = is_and;
      return;
    }
    case K_IN: {
      Val c;
      Eval(*n.kids[0], cx, row, &c);
      out->ok = c.ok;
      bool hit = false;
      if (n.kids[0]->type.is_string()) {
        for (const auto& s : n.in_strs) hit = hit || s == c.s;
      } else if (n.kids[0]->type.id == T_FLOAT || n.kids[0]->type.id == T_DOUBLE) {
        const double x = n.kids[0]->type.id == T_FLOAT ? static_cast<double>(c.f) : c.d;
        for (double v : n.in_dbls) hit = hit || v == x;   // IEEE equality: NaN matches nothing, -0.0 == 0.0
      } else {
        for (auto v : n.in_ints) hit = hit || v == c.i;
      }
      out->b = hit;
      return;
    }
  }
}

bool Prepare(Node* n, std::string* err) {
  for (auto& k : n->kids)
    if (!Prepare(k.get(), err)) return false;
  if (n->kind == K_FN && (n->name == "regexp_matches" || n->name == "regexp_like")) {
    if (n->kids.size() != 2 || n->kids[1]->kind != K_LIT) { *err = "regexp_matches needs a literal pattern"; return false; }
    OrcReParser rp;
    rp.p = DecodeUtf8(n->kids[1]->lit.s);
    if (n->kids[1]->lit.s.compare(0, 4, "(?i)") == 0) { rp.icase = true; rp.i = 4; }
    n->regex = rp.Alt();
    if (rp.bad || rp.i != rp.p.size()) { *err = "malformed regular expression"; return false; }
  }
  if (n->kind == K_FN && (n->name == "like" || n->name == "ilike")) {
    if (n->kids.size() < 2 || n->kids[1]->kind != K_LIT) { *err = "like needs a literal pattern"; return false; }
    const bool has_esc = n->kids.size() == 3;
    const char esc = has_esc ? n->kids[2]->lit.s[0] : 0;
    std::string pat = n->kids[1]->lit.s;
    if (n->name == "ilike")  // ASCII case folding of pattern and text
      for (auto& ch : pat)
        if (ch >= 'A' && ch <= 'Z') ch = static_cast<char>(ch + 32);
    n->like = CompileLike(pat, has_esc, esc);
  }
  return true;
}

void SetBitTo(uint8_t* bits, int64_t i, bool v) {
  if (v) bits[i >> 3] |= static_cast<uint8_t>(1u << (i & 7));
  else bits[i >> 3] &= static_cast<uint8_t>(~(1u << (i & 7)));
}

void StoreValue(const Type& t, const Val& v, int64_t i, void* values) {
  uint8_t* p = static_cast<uint8_t*>(values);
  switch (t.id) {
    case T_BOOL: SetBitTo(p, i, v.ok && v.b); break;
    case T_INT8: reinterpret_cast<int8_t*>(p)[i] = static_cast<int8_t>(v.i); break;
    case T_INT16: reinterpret_cast<int16_t*>(p)[i] = static_cast<int16_t>(v.i); break;
    case T_INT32: case T_DATE32: case T_TIME32: reinterpret_cast<int32_t*>(p)[i] = static_cast<int32_t>(v.i); break;
    case T_INT64: case T_DATE64: case T_TIMESTAMP: case T_TIME64: reinterpret_cast<int64_t*>(p)[i] = v.i; break;
    case T_UINT8: p[i] = static_cast<uint8_t>(v.u); break;
    case T_UINT16: reinterpret_cast<uint16_t*>(p)[i] = static_cast<uint16_t>(v.u); break;
    case T_UINT32: reinterpret_cast<uint32_t*>(p)[i] = static_cast<uint32_t>(v.u); break;
    case T_UINT64: reinterpret_cast<uint64_t*>(p)[i] = v.u; break;
    case T_FLOAT: reinterpret_cast<float*>(p)[i] = v.f; break;
    case T_DOUBLE: reinterpret_cast<double*>(p)[i] = v.d; break;
    case T_DECIMAL: std::memcpy(p + 16 * i, &v.dec, 16); break;
    default: break;
  }
}

struct Expr {
  std::unique_ptr<Node> root;
};

template <typename F>
void ParallelFor(int64_t n, int threads, F fn) {
  // ranges are multiples of 64 rows so threads never share a validity byte
  if (threads <= 1 || n < 4096) {
    fn(0, n, 0);
    return;
  }
  int64_t chunk = ((n + threads - 1) / threads + 63) / 64 * 64;
  std::vector<std::thread> ts;
  int idx = 0;
  for (int64_t b = 0; b < n; b += chunk, ++idx) {
    const int64_t e = std::min(n, b + chunk);
    ts.emplace_back([=] { fn(b, e, idx); });
  }
  for (auto& t : ts) t.join();
}

}  // namespace

extern "C" {

void* orc_parse(const char* sexpr, char* err, int errlen) {
  Parser ps;
  ps.p = sexpr;
  std::unique_ptr<Node> n = ps.node();
  std::string e = ps.err;
  if (n && !Prepare(n.get(), &e)) n.reset();
  if (!n) {
    if (err && errlen > 0) std::snprintf(err, static_cast<size_t>(errlen), "%s", e.c_str());
    return nullptr;
  }
  Expr* x = new Expr();
  x->root = std::move(n);
  return x;
}

void orc_free(void* e) { delete static_cast<Expr*>(e); }

// Evaluate one expression over rows [0, n) (or over sel[0..nsel) when sel != null; sel holds
// int64 row indices).  Fixed-width / bool outputs only; strings via orc_project_string.
// out_validity / out_values must be zero-initialised by the caller (bits are OR-ed in).
// Returns 0, or the ExecutionError code (1 = divide by zero).
int orc_project(void* expr, const Column* cols, int64_t n, const int64_t* sel, int64_t nsel,
                uint8_t* out_validity, void* out_values, int threads) {
  const Node& root = *static_cast<Expr*>(expr)->root;
  const int64_t count = sel ? nsel : n;
  std::vector<int> errs(static_cast<size_t>(std::max(threads, 1)) + 1, 0);
  ParallelFor(count, threads, [&](int64_t b, int64_t e, int idx) {
    EvalCtx cx;
    cx.cols = cols;
    for (int64_t i = b; i < e; ++i) {
      Val v;
      Eval(root, cx, sel ? sel[i] : i, &v);
      if (cx.error) break;
      if (out_validity) SetBitTo(out_validity, i, v.ok);
      StoreValue(root.type, v, i, out_values);
    }
    errs[static_cast<size_t>(idx)] = cx.error;
  });
  for (int e : errs) if (e) return e;
  return 0;
}

// String-valued expression: appends bytes to out_data (capacity cap), writes n+1 offsets.
// Returns 0, an error code, or -1 when cap is too small (needed size in *needed).
int orc_project_string(void* expr, const Column* cols, int64_t n, const int64_t* sel, int64_t nsel,
                       uint8_t* out_validity, int32_t* out_offsets, uint8_t* out_data, int64_t cap,
                       int64_t* needed) {
  const Node& root = *static_cast<Expr*>(expr)->root;
  const int64_t count = sel ? nsel : n;
  EvalCtx cx;
  cx.cols = cols;
  int64_t pos = 0;
  out_offsets[0] = 0;
  for (int64_t i = 0; i < count; ++i) {
    Val v;
    Eval(root, cx, sel ? sel[i] : i, &v);
    if (cx.error) return cx.error;
    if (out_validity) SetBitTo(out_validity, i, v.ok);
    const int64_t len = v.ok ? static_cast<int64_t>(v.s.size()) : 0;
    if (pos + len <= cap && len > 0) std::memcpy(out_data + pos, v.s.data(), static_cast<size_t>(len));
    pos += len;
    out_offsets[i + 1] = static_cast<int32_t>(pos);
  }
  *needed = pos;
  return pos <= cap ? 0 : -1;
}

// Filter: ascending indices of rows where the condition is valid and true.
// Two passes like the reference: (1) condition -> bitmap, (2) bitmap -> index list.
int orc_filter(void* expr, const Column* cols, int64_t n, uint64_t* out_idx, int64_t* out_count,
               int threads) {
  const Node& root = *static_cast<Expr*>(expr)->root;
  std::vector<uint8_t> bitmap(static_cast<size_t>((n + 7) / 8 + 8), 0);
  std::vector<int> errs(static_cast<size_t>(std::max(threads, 1)) + 1, 0);
  ParallelFor(n, threads, [&](int64_t b, int64_t e, int idx) {
    EvalCtx cx;
    cx.cols = cols;
    for (int64_t i = b; i < e; ++i) {
      Val v;
      Eval(root, cx, i, &v);
      if (cx.error) break;
      if (v.ok && v.b) bitmap[static_cast<size_t>(i >> 3)] |= static_cast<uint8_t>(1u << (i & 7));
    }
    errs[static_cast<size_t>(idx)] = cx.error;
  });
  for (int e : errs) if (e) return e;
  int64_t c = 0;
  for (int64_t w = 0; w * 64 < n; ++w) {
    uint64_t word;
    std::memcpy(&word, bitmap.data() + w * 8, 8);
    while (word) {
      const int bit = __builtin_ctzll(word);
      const int64_t i = w * 64 + bit;
      if (i < n) out_idx[c++] = static_cast<uint64_t>(i);
      word &= word - 1;
    }
  }
  *out_count = c;
  return 0;
}

// Synthetic lineitem column on the CPU (same stream as the device generator).
void orc_generate_lineitem(int kind, uint64_t seed, int64_t first_row, int64_t num_rows,
                           void* values, uint8_t* validity, int null_permille, int threads) {
  ParallelFor(num_rows, threads, [&](int64_t b, int64_t e, int) {
    gdv_lineitem_fill(kind, seed, first_row, b, e, values, validity, null_permille);
  });
}

int orc_hardware_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n == 0 ? 1 : static_cast<int>(n);
}

}  // extern "C"
