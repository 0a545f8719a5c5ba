#include "gdv_datefmt.h"

#include <cctype>

namespace gdv {

int CompileDateFormat(const std::string& pattern, std::vector<uint8_t>* prog, std::string* why) {
  prog->clear();
  if (pattern.empty()) {
    *why = "Invalid pattern: the format of to_date is empty";
    return 1;
  }
  struct Tok {
    const char* text;
    int op;  // DateFmtOp, or -1 = recognised by the reference but not implemented here
  };
  // longest first where one token is a prefix of another
  static const Tok kToks[] = {{"YYYY", kFmtYear4}, {"YY", kFmtYear2},   {"MONTH", kFmtMonthName}, {"MON", kFmtMonthName},
                              {"MM", kFmtMonth},   {"MI", kFmtMinute},  {"DDD", -1},              {"DD", kFmtDay},
                              {"DAY", -1},         {"DY", -1},          {"HH24", kFmtHour24},     {"HH12", kFmtHour12},
                              {"HH", kFmtHour12},  {"SS", kFmtSecond},  {"AM", kFmtAmPm},         {"PM", kFmtAmPm},
                              {"FFF", -1},         {"TZD", -1},         {"TZO", -1},              {"WW", -1},
                              {"CC", -1}};
  size_t i = 0;
  bool after_space = false;  // consecutive white space in the pattern is one "any run of white space"
  while (i < pattern.size()) {
    const unsigned char c = static_cast<unsigned char>(pattern[i]);
    if (std::isspace(c)) {
      if (!after_space) prog->push_back(kFmtSpace);
      after_space = true;
      ++i;
      continue;
    }
    after_space = false;
    if (c == '"') {
      *why = "to_date: quoted text in the format is not supported yet";
      return 2;
    }
    if (c == '%') {
      *why = "Invalid pattern: '%' in the format of to_date";
      return 1;
    }
    bool hit = false;
    if (std::isalpha(c)) {
      for (const Tok& t : kToks) {
        size_t n = 0;
        while (t.text[n] != 0 && i + n < pattern.size() &&
               std::toupper(static_cast<unsigned char>(pattern[i + n])) == t.text[n])
          ++n;
        if (t.text[n] != 0) continue;
        if (t.op < 0) {
          *why = std::string("to_date: format token ") + t.text + " is not supported yet";
          return 2;
        }
        prog->push_back(static_cast<uint8_t>(t.op));
        i += n;
        hit = true;
        break;
      }
    }
    if (hit) continue;
    prog->push_back(kFmtLiteral);
    prog->push_back(c);
    ++i;
  }
  if (prog->size() > 200) {
    *why = "to_date: format too long";
    return 2;
  }
  return 0;
}

}  // namespace gdv
