// Regular expression -> position automaton over bytes.  See gdv_regex.h.
#include "gdv_regex.h"

#include <bitset>
#include <cctype>
#include <memory>
#include <vector>

namespace gdv {
namespace {

using ByteSet = std::bitset<256>;

struct Re {
  enum Kind { kEmpty, kSet, kCat, kAlt, kStar, kPlus, kOpt, kBol, kEol } kind = kEmpty;
  ByteSet set;                 // kSet: one position accepting these bytes
  std::shared_ptr<Re> a, b;
};
using ReP = std::shared_ptr<Re>;

ReP Mk(Re::Kind k, ReP a = nullptr, ReP b = nullptr) {
  auto r = std::make_shared<Re>();
  r->kind = k;
  r->a = std::move(a);
  r->b = std::move(b);
  return r;
}
ReP MkSet(const ByteSet& s) {
  auto r = Mk(Re::kSet);
  r->set = s;
  return r;
}
ByteSet Range(int lo, int hi) {
  ByteSet s;
  for (int c = lo; c <= hi; ++c) s.set(static_cast<size_t>(c));
  return s;
}
ReP Cat(ReP a, ReP b) {
  if (a->kind == Re::kEmpty) return b;
  if (b->kind == Re::kEmpty) return a;
  return Mk(Re::kCat, std::move(a), std::move(b));
}
ReP Clone(const ReP& r) {
  if (!r) return nullptr;
  auto c = std::make_shared<Re>(*r);
  c->a = Clone(r->a);
  c->b = Clone(r->b);
  return c;
}
// One code point whose ASCII members are `ascii`; with_multibyte adds every non-ASCII code point.
// A multi-byte character costs two positions, not one per UTF-8 form: a lead byte (folded into the
// ASCII position) followed by any run of continuation bytes.  On well-formed UTF-8 that run is exactly
// the rest of the character -- no atom of the automaton can begin at a continuation byte, so the run
// can neither stop early nor swallow part of the next character.
ReP CodePoint(const ByteSet& ascii, bool with_multibyte) {
  if (!with_multibyte) return MkSet(ascii);  // an empty set never matches
  return Cat(MkSet(ascii | Range(0xc2, 0xf4)), Mk(Re::kStar, MkSet(Range(0x80, 0xbf))));
}
ByteSet Digits() { return Range('0', '9'); }
ByteSet Word() { return Range('0', '9') | Range('a', 'z') | Range('A', 'Z') | Range('_', '_'); }
ByteSet Space() {
  ByteSet s;
  for (char c : {' ', '\t', '\n', '\f', '\r'}) s.set(static_cast<unsigned char>(c));
  return s;
}
ByteSet AsciiNot(const ByteSet& s) { return Range(0, 0x7f) & ~s; }

class Parser {
 public:
  Parser(const std::string& p, std::string* err) : p_(p), err_(err) {}
  int code = 0;  // 0 ok, 1 invalid, 2 not implemented

  ReP Parse() {
    if (p_.compare(0, 4, "(?i)") == 0) {  // ASCII case-insensitive matching for the whole pattern
      icase_ = true;
      i_ = 4;
    }
    ReP r = Alt();
    if (code == 0 && i_ < p_.size()) Fail(1, "unmatched ')'");
    return r;
  }

 private:
  const std::string& p_;
  size_t i_ = 0;
  std::string* err_;
  bool icase_ = false;

  // a position accepting `set` (both cases of its ASCII letters under (?i))
  ReP Pos(ByteSet set) const {
    if (icase_) {
      for (int c = 'a'; c <= 'z'; ++c) {
        const size_t lo = static_cast<size_t>(c), up = static_cast<size_t>(c - 32);
        if (set.test(lo) || set.test(up)) { set.set(lo); set.set(up); }
      }
    }
    return MkSet(set);
  }
  ByteSet Fold(ByteSet set) const {
    if (icase_) {
      for (int c = 'a'; c <= 'z'; ++c) {
        const size_t lo = static_cast<size_t>(c), up = static_cast<size_t>(c - 32);
        if (set.test(lo) || set.test(up)) { set.set(lo); set.set(up); }
      }
    }
    return set;
  }

  void Fail(int c, const std::string& m) {
    if (code == 0) {
      code = c;
      *err_ = m + " in regular expression '" + p_ + "'";
    }
  }
  bool More() const { return i_ < p_.size(); }
  unsigned char Peek() const { return static_cast<unsigned char>(p_[i_]); }

  ReP Alt() {
    ReP r = Concat();
    while (code == 0 && More() && Peek() == '|') {
      ++i_;
      r = Mk(Re::kAlt, r, Concat());
    }
    return r;
  }
  ReP Concat() {
    ReP r = Mk(Re::kEmpty);
    while (code == 0 && More() && Peek() != '|' && Peek() != ')') r = Cat(r, Repeat());
    return r;
  }
  bool Number(int* out) {
    if (!More() || Peek() < '0' || Peek() > '9') return false;
    int v = 0;
    while (More() && Peek() >= '0' && Peek() <= '9') {
      v = v * 10 + (Peek() - '0');
      if (v > 1000) v = 1000;
      ++i_;
    }
    *out = v;
    return true;
  }
  ReP Repeat() {
    ReP atom = Atom();
    while (code == 0 && More()) {
      const unsigned char c = Peek();
      int lo = 0, hi = -1;
      if (c == '*') { ++i_; lo = 0; hi = -1; }
      else if (c == '+') { ++i_; lo = 1; hi = -1; }
      else if (c == '?') { ++i_; lo = 0; hi = 1; }
      else if (c == '{') {
        const size_t save = i_;
        ++i_;
        if (!Number(&lo)) { i_ = save; break; }  // a literal '{'
        hi = lo;
        if (More() && Peek() == ',') {
          ++i_;
          if (!Number(&hi)) hi = -1;
        }
        if (!More() || Peek() != '}') { i_ = save; break; }
        ++i_;
        if (hi >= 0 && hi < lo) { Fail(1, "bad repetition {m,n}"); break; }
      } else {
        break;
      }
      if (atom->kind == Re::kBol || atom->kind == Re::kEol) { Fail(2, "repetition of an anchor"); break; }
      if (More() && Peek() == '?') ++i_;  // lazy: the same set of matching rows
      if (More() && Peek() == '+') { Fail(2, "possessive repetition"); break; }
      atom = Expand(atom, lo, hi);
      // the reference's RE2 does not stack repetition operators: a** / a+* / a{2}* are syntax errors
      if (code == 0 && More()) {
        const unsigned char q = Peek();
        bool again = q == '*' || q == '+' || q == '?';
        if (q == '{') {
          const size_t save = i_;
          int a = 0, b2 = 0;
          ++i_;
          if (Number(&a)) {
            if (More() && Peek() == ',') {
              ++i_;
              (void)Number(&b2);
            }
            again = More() && Peek() == '}';
          }
          i_ = save;
        }
        if (again) { Fail(1, "bad repetition operator: a repetition cannot follow another one"); break; }
      }
    }
    return atom;
  }
  ReP Expand(const ReP& a, int lo, int hi) {
    if (lo == 0 && hi < 0) return Mk(Re::kStar, a);
    if (lo == 1 && hi < 0) return Mk(Re::kPlus, a);
    if (lo == 0 && hi == 1) return Mk(Re::kOpt, a);
    if (lo > 64 || hi > 64) { Fail(2, "repetition count above 64"); return a; }
    ReP r = Mk(Re::kEmpty);
    for (int k = 0; k < lo; ++k) r = Cat(r, Clone(a));
    if (hi < 0) return Cat(r, Mk(Re::kStar, Clone(a)));
    ReP tail = Mk(Re::kEmpty);  // (a(a(a)?)?)?
    for (int k = lo; k < hi; ++k) tail = Mk(Re::kOpt, Cat(Clone(a), tail));
    return Cat(r, tail);
  }
  // one UTF-8 encoded character starting at i_ -> a chain of single-byte sets
  ReP Utf8Char() {
    const unsigned char c = Peek();
    const int n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
    if (n == 0 || i_ + static_cast<size_t>(n) > p_.size()) { Fail(1, "invalid UTF-8"); ++i_; return Mk(Re::kEmpty); }
    ReP r = Mk(Re::kEmpty);
    for (int k = 0; k < n; ++k) {
      const unsigned char b = static_cast<unsigned char>(p_[i_ + static_cast<size_t>(k)]);
      if (k > 0 && (b & 0xc0) != 0x80) { Fail(1, "invalid UTF-8"); break; }
      r = Cat(r, Pos(Range(b, b)));
    }
    i_ += static_cast<size_t>(n);
    return r;
  }
  // after a backslash: a class shorthand (sets *set, *negated) or a literal byte (returns true / sets *lit)
  bool Escape(ByteSet* set, bool* negated, bool* is_class) {
    if (!More()) { Fail(1, "trailing backslash"); return false; }
    const unsigned char c = Peek();
    ++i_;
    *is_class = true;
    *negated = false;
    switch (c) {
      case 'd': *set = Digits(); return true;
      case 'w': *set = Word(); return true;
      case 's': *set = Space(); return true;
      case 'D': *set = Digits(); *negated = true; return true;
      case 'W': *set = Word(); *negated = true; return true;
      case 'S': *set = Space(); *negated = true; return true;
      default: break;
    }
    *is_class = false;
    int lit = -1;
    switch (c) {
      case 'n': lit = '\n'; break;
      case 't': lit = '\t'; break;
      case 'r': lit = '\r'; break;
      case 'f': lit = '\f'; break;
      case 'v': lit = '\v'; break;
      case 'a': lit = 7; break;
      case 'x': {
        int v = 0, nd = 0;
        while (nd < 2 && More() && std::isxdigit(Peek())) {
          const unsigned char h = Peek();
          v = v * 16 + (h <= '9' ? h - '0' : (h | 0x20) - 'a' + 10);
          ++i_;
          ++nd;
        }
        if (nd != 2) { Fail(1, "\\x needs two hexadecimal digits"); return false; }
        if (v >= 0x80) { Fail(2, "\\x escape above 7f"); return false; }
        lit = v;
        break;
      }
      default:
        if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) {
          Fail(2, std::string("escape \\") + static_cast<char>(c));  // \b \B \A \z \1 \p{..} \x.. \Q ...
          return false;
        }
        if (c >= 0x80) { Fail(1, "backslash before a non-ASCII character"); return false; }
        lit = c;  // punctuation stands for itself
    }
    *set = Range(lit, lit);
    return true;
  }
  ReP Class() {  // after '['
    bool neg = false;
    if (More() && Peek() == '^') { neg = true; ++i_; }
    ByteSet ascii;
    std::vector<ReP> wide;  // listed non-ASCII characters
    bool first = true;
    while (true) {
      if (!More()) { Fail(1, "missing ']'"); return Mk(Re::kEmpty); }
      unsigned char c = Peek();
      if (c == ']' && !first) { ++i_; break; }
      first = false;
      if (c == '[' && i_ + 1 < p_.size() && p_[i_ + 1] == ':') {
        const size_t end = p_.find(":]", i_ + 2);
        if (end == std::string::npos) { Fail(1, "missing ':]'"); return Mk(Re::kEmpty); }
        const std::string name = p_.substr(i_ + 2, end - i_ - 2);
        ByteSet cls;
        if (name == "alpha") cls = Range('a', 'z') | Range('A', 'Z');
        else if (name == "digit") cls = Digits();
        else if (name == "alnum") cls = Range('a', 'z') | Range('A', 'Z') | Digits();
        else if (name == "upper") cls = Range('A', 'Z');
        else if (name == "lower") cls = Range('a', 'z');
        else if (name == "space") cls = Space() | Range('\v', '\v');
        else if (name == "blank") cls = Range(' ', ' ') | Range('\t', '\t');
        else if (name == "punct") cls = Range('!', '/') | Range(':', '@') | Range('[', '`') | Range('{', '~');
        else if (name == "xdigit") cls = Digits() | Range('a', 'f') | Range('A', 'F');
        else if (name == "word") cls = Word();
        else if (name == "print") cls = Range(' ', '~');
        else if (name == "graph") cls = Range('!', '~');
        else if (name == "cntrl") cls = Range(0, 31) | Range(127, 127);
        else if (name == "ascii") cls = Range(0, 127);
        else { Fail(1, "unknown class [:" + name + ":]"); return Mk(Re::kEmpty); }
        ascii |= cls;
        i_ = end + 2;
        continue;
      }
      int lo = -1;
      if (c == '\\') {
        ++i_;
        ByteSet s;
        bool n = false, is_class = false;
        if (!Escape(&s, &n, &is_class)) return Mk(Re::kEmpty);
        if (is_class) {
          if (n) { Fail(2, "negated shorthand inside a class"); return Mk(Re::kEmpty); }
          ascii |= s;
          continue;
        }
        for (int b = 0; b < 128; ++b) if (s.test(static_cast<size_t>(b))) lo = b;
      } else if (c >= 0x80) {
        wide.push_back(Utf8Char());
        if (More() && Peek() == '-' && i_ + 1 < p_.size() && p_[i_ + 1] != ']') { Fail(2, "non-ASCII range in a class"); return Mk(Re::kEmpty); }
        continue;
      } else {
        lo = c;
        ++i_;
      }
      int hi = lo;
      if (More() && Peek() == '-' && i_ + 1 < p_.size() && p_[i_ + 1] != ']') {
        ++i_;
        unsigned char h = Peek();
        if (h == '\\') {
          ++i_;
          ByteSet s;
          bool n = false, is_class = false;
          if (!Escape(&s, &n, &is_class)) return Mk(Re::kEmpty);
          if (is_class) { Fail(1, "bad range in a class"); return Mk(Re::kEmpty); }
          for (int b = 0; b < 128; ++b) if (s.test(static_cast<size_t>(b))) hi = b;
        } else if (h >= 0x80) {
          Fail(2, "non-ASCII range in a class");
          return Mk(Re::kEmpty);
        } else {
          hi = h;
          ++i_;
        }
        if (hi < lo) { Fail(1, "bad range in a class"); return Mk(Re::kEmpty); }
      }
      ascii |= Range(lo, hi);
    }
    ascii = Fold(ascii);
    if (neg) {
      if (!wide.empty()) { Fail(2, "negated class with non-ASCII members"); return Mk(Re::kEmpty); }
      return CodePoint(AsciiNot(ascii), true);
    }
    ReP r = ascii.any() ? MkSet(ascii) : nullptr;
    for (auto& w : wide) r = r ? Mk(Re::kAlt, r, w) : w;
    return r ? r : MkSet(ByteSet());
  }
  ReP Atom() {
    const unsigned char c = Peek();
    switch (c) {
      case '(': {
        ++i_;
        if (More() && Peek() == '?') {
          if (i_ + 1 < p_.size() && p_[i_ + 1] == ':') i_ += 2;
          else { Fail(2, "group flags / look-around ((?i) is accepted at the very start only)"); return Mk(Re::kEmpty); }
        }
        ReP r = Alt();
        if (!More() || Peek() != ')') { Fail(1, "missing ')'"); return r; }
        ++i_;
        return r;
      }
      case '[': ++i_; return Class();
      case '.': ++i_; return CodePoint(AsciiNot(Range('\n', '\n')), true);
      case '^': ++i_; return Mk(Re::kBol);
      case '$': ++i_; return Mk(Re::kEol);
      case '*': case '+': case '?': Fail(1, "nothing to repeat"); ++i_; return Mk(Re::kEmpty);
      case '\\': {
        ++i_;
        ByteSet s;
        bool n = false, is_class = false;
        if (!Escape(&s, &n, &is_class)) return Mk(Re::kEmpty);
        if (is_class && n) return CodePoint(AsciiNot(s), true);
        return Pos(s);
      }
      default:
        return Utf8Char();
    }
  }
};

using PosSet = std::bitset<RegexProgram::kMaxPositions>;

struct Info {
  bool nullable = false;
  PosSet first, last;
};

struct Builder {
  RegexProgram* prog;
  bool overflow = false;
  bool inner_anchor = false;
  PosSet follow[RegexProgram::kMaxPositions];

  Info Walk(const ReP& r) {
    Info o;
    switch (r->kind) {
      case Re::kEmpty: o.nullable = true; return o;
      case Re::kBol: case Re::kEol: inner_anchor = true; o.nullable = true; return o;
      case Re::kSet: {
        if (prog->positions >= RegexProgram::kMaxPositions) { overflow = true; return o; }
        const int p = prog->positions++;
        for (int b = 0; b < 256; ++b)
          if (r->set.test(static_cast<size_t>(b))) prog->cls[b][p >> 6] |= 1ull << (p & 63);
        o.first.set(static_cast<size_t>(p));
        o.last = o.first;
        return o;
      }
      case Re::kCat: {
        const Info x = Walk(r->a), y = Walk(r->b);
        Link(x.last, y.first);
        o.nullable = x.nullable && y.nullable;
        o.first = x.first | (x.nullable ? y.first : PosSet());
        o.last = y.last | (y.nullable ? x.last : PosSet());
        return o;
      }
      case Re::kAlt: {
        const Info x = Walk(r->a), y = Walk(r->b);
        o.nullable = x.nullable || y.nullable;
        o.first = x.first | y.first;
        o.last = x.last | y.last;
        return o;
      }
      case Re::kStar: case Re::kPlus: case Re::kOpt: {
        const Info x = Walk(r->a);
        if (r->kind != Re::kOpt) Link(x.last, x.first);
        o.nullable = r->kind == Re::kPlus ? x.nullable : true;
        o.first = x.first;
        o.last = x.last;
        return o;
      }
    }
    return o;
  }
  void Link(const PosSet& from, const PosSet& to) {
    for (int p = 0; p < RegexProgram::kMaxPositions; ++p)
      if (from.test(static_cast<size_t>(p))) follow[p] |= to;
  }
  static void Words(const PosSet& s, uint64_t out[2]) {
    out[0] = out[1] = 0;
    for (int p = 0; p < RegexProgram::kMaxPositions; ++p)
      if (s.test(static_cast<size_t>(p))) out[p >> 6] |= 1ull << (p & 63);
  }
};

}  // namespace

int CompileRegex(const std::string& pattern, RegexProgram* out, std::string* error) {
  *out = RegexProgram();
  Parser parser(pattern, error);
  ReP root = parser.Parse();
  if (parser.code != 0) return parser.code;
  // text anchors: only as the first / last element of the whole pattern
  std::vector<ReP> chain;  // the top-level concatenation, flattened
  std::vector<ReP> todo = {root};
  while (!todo.empty()) {
    ReP r = todo.back();
    todo.pop_back();
    if (r->kind == Re::kCat) {
      todo.push_back(r->b);
      todo.push_back(r->a);
    } else {
      chain.push_back(r);
    }
  }
  size_t lo = 0, hi = chain.size();
  if (lo < hi && chain[lo]->kind == Re::kBol) { out->anchor_start = true; ++lo; }
  if (lo < hi && chain[hi - 1]->kind == Re::kEol) { out->anchor_end = true; --hi; }
  ReP body = Mk(Re::kEmpty);
  for (size_t k = lo; k < hi; ++k) body = Cat(body, chain[k]);
  Builder b;
  b.prog = out;
  const Info info = b.Walk(body);
  if (b.inner_anchor) {
    *error = "'^' / '$' inside the pattern are not supported (regular expression '" + pattern + "')";
    return 2;
  }
  if (b.overflow) {
    *error = "regular expression '" + pattern + "' needs more than 128 automaton positions";
    return 2;
  }
  out->nullable = info.nullable;
  Builder::Words(info.first, out->first);
  Builder::Words(info.last, out->last);
  for (int p = 0; p < out->positions; ++p) Builder::Words(b.follow[p], out->follow[p]);
  return 0;
}

}  // namespace gdv
