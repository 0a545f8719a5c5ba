// regexp_matches / regexp_like: the pattern (a literal, like LIKE's) is compiled at Make() into a
// position automaton (Glushkov construction) over BYTES with at most 64 positions; the kernel runs
// it bit-parallel, one u64 of live positions per row (device/gdv_device_lib.cuh gdv_regex_match).
// Semantics follow RE2's partial match, which the reference's holder uses: unanchored search,
// '.' is any code point but '\n', \d \w \s are ASCII, '^' / '$' are text (not line) anchors.
#pragma once

#include <cstdint>
#include <string>


namespace gdv {

struct RegexProgram {
  int positions = 0;
  uint64_t first = 0;        // positions that can start a match
  uint64_t last = 0;         // positions that can end one
  uint64_t follow[64] = {};  // follow[p]: positions that may come right after p
  uint64_t cls[256] = {};    // cls[b]: positions that accept byte b
  bool nullable = false;     // the pattern matches the empty string
  bool anchor_start = false, anchor_end = false;
};

// Returns 0, or 1 for a malformed pattern, or 2 for syntax outside the subset (back-references,
// look-around, word boundaries, anchors inside the pattern, non-ASCII class ranges) and patterns
// that need more than 64 positions; `error` then says which.
int CompileRegex(const std::string& pattern, RegexProgram* out, std::string* error);

}  // namespace gdv
