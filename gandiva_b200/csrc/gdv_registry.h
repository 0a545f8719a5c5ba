// Function registry: name + parameter types + return type -> device function and its
// null behaviour.  Replaces the reference's precompiled-bitcode function_registry
// (named in BASELINE.json north_star; enumeration API P/includes/libgandiva.pxd:258-277).
// Device functions live in device/gdv_device_lib.cuh and are named
// <base>_<suffix of each param>, the reference library's naming scheme.
#pragma once
#include <string>
#include <vector>

#include "gdv_types.h"

namespace gdv {

enum class NullMode {
  kIfNull,    // result is null iff any input is null; function sees values only
  kNever,     // result is never null; function receives (value, valid) pairs
  kInternal,  // function computes validity: receives (value, valid) pairs + bool* out_valid
};

enum FnFlags : uint32_t {
  kCanFail = 1u << 0,      // may raise an ExecutionError -> takes gdv_ctx*, only called on valid rows
  kLikeHolder = 1u << 1,   // second (third) arg must be a literal pattern, compiled at Make()
  kDecimalArgs = 1u << 2,  // decimal params are followed by (precision, scale); out (p, s) appended
  kStringView = 1u << 3,   // returns a view/transform of its first argument (no new bytes)
  kConcat = 1u << 4,       // result = the pieces of its arguments in order (a rope, see the fuser)
  kScratch = 1u << 5,      // writes its result bytes into a per-row scratch slot passed as last argument
  kRegexHolder = 1u << 7,  // second arg is a literal regular expression, compiled at Make() (gdv_regex.h)
  kDateFormat = 1u << 8,   // second arg is a literal date format, compiled at Make() (gdv_datefmt.h); a third
                           // (int32 literal) argument suppresses parse errors: the row is NULL instead
  kVirtual = 1u << 6,      // result is a rope of "virtual" pieces (repeated / reversed bytes) that only the
                           // string write pass can read: projectable, concat-able, if/else-able, nothing else
};

struct FunctionDef {
  std::string name;               // name as written by the user (alias or canonical)
  std::string device_base;        // canonical base used for the device symbol
  std::vector<DataType> params;   // decimal params match any precision/scale
  DataType ret;                   // decimal: precision/scale come from the call site
  NullMode nulls = NullMode::kIfNull;
  uint32_t flags = 0;
  std::string device_name() const;  // e.g. add_int32_int32
  std::string signature() const;    // "int32 add(int32, int32)"
};

class Registry {
 public:
  static const Registry& Get();
  // Exact match on name and parameter types (decimal: id only; timestamp/time: unit too).
  const FunctionDef* Lookup(const std::string& name, const std::vector<DataType>& params) const;
  const std::vector<FunctionDef>& all() const { return defs_; }

 private:
  Registry();
  void Add(const std::string& name, std::vector<DataType> params, DataType ret,
           NullMode nulls = NullMode::kIfNull, uint32_t flags = 0,
           const std::vector<std::string>& aliases = {});
  std::vector<FunctionDef> defs_;
};

}  // namespace gdv
