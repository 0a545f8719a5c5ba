// regexp_matches / regexp_like: the pattern (a literal, like LIKE's) is compiled at Make() into a
// position automaton (Glushkov construction) over BYTES with at most 128 positions; the kernel runs
// it bit-parallel, one (or two) u64 of live positions per row (device/gdv_device_lib.cuh gdv_regex_match).
// Semantics follow RE2's partial match, which the reference's holder uses: unanchored search,
// '.' is any code point but '\n', \d \w \s are ASCII, '^' / '$' are text (not line) anchors.
#pragma once

#include <cstdint>
#include <string>


namespace gdv {

struct RegexProgram {
  static constexpr int kMaxPositions = 128;
  int positions = 0;              // <= 64: only word 0 of every set is used
  uint64_t first[2] = {};         // positions that can start a match
  uint64_t last[2] = {};          // positions that can end one
  uint64_t follow[128][2] = {};   // follow[p]: positions that may come right after p
  uint64_t cls[256][2] = {};      // cls[b]: positions that accept byte b
  bool nullable = false;          // the pattern matches the empty string
  bool anchor_start = false, anchor_end = false;
};

// Returns 0, or 1 for a malformed pattern, or 2 for syntax outside the subset (back-references,
// look-around, word boundaries, anchors inside the pattern, non-ASCII class ranges) and patterns
// that need more than 128 positions; `error` then says which.
int CompileRegex(const std::string& pattern, RegexProgram* out, std::string* error);

}  // namespace gdv
