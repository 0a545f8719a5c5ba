// to_date(utf8, utf8 pattern [, int32 suppress_errors]): the SQL-style pattern literal is compiled at Make()
// into a small program that the device function gdv_to_date_fmt (device/gdv_device_lib.cuh) interprets.
//
// Reference behaviour being followed (descendant's to_date holder, from memory — unpinned): the
// pattern is translated token by token into a strptime format (YYYY -> %Y, YY -> %y, MM -> %m, MON -> %b,
// MONTH -> %B, DD -> %d, HH24 -> %H, HH / HH12 -> %I, MI -> %M, SS -> %S, AM / PM -> %p, everything else
// verbatim), the text is parsed with strptime, trailing characters are allowed, the time of day is ignored and
// the result is midnight of year / month / max(day, 1) in milliseconds since the epoch.  The program encodes
// exactly glibc's strptime rules for those directives (field widths, the "stop when another digit would
// exceed the maximum" rule, white space in the pattern matching any run of white space, case-insensitive
// month names in either form for both MON and MONTH); tests/test_oracle_vs_arrow.py referees the oracle against
// the C library's strptime itself.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace gdv {

enum DateFmtOp : uint8_t {
  kFmtYear4 = 1, kFmtYear2, kFmtMonth, kFmtDay, kFmtHour24, kFmtHour12, kFmtMinute, kFmtSecond, kFmtMonthName,
  kFmtAmPm, kFmtSpace, kFmtLiteral /* followed by the byte */
};

// 0 = ok; 1 = invalid pattern; 2 = a token this engine does not implement (*why says which).
int CompileDateFormat(const std::string& pattern, std::vector<uint8_t>* prog, std::string* why);

}  // namespace gdv
