// Data types of the expression engine.  Type ids are arrow::Type::type values
// (P/include/arrow/type_fwd.h:330-473) so the C++ drop-in layer converts with a cast.
#pragma once
#include <cstdint>
#include <string>

#include "gandiva_b200.h"

namespace gdv {

struct DataType {
  int32_t id = GDV_TYPE_NA;
  int32_t precision = 0;  // decimal precision, or time unit for timestamp/time
  int32_t scale = 0;

  DataType() = default;
  DataType(int32_t i, int32_t p = 0, int32_t s = 0) : id(i), precision(p), scale(s) {}
  explicit DataType(const gdv_type_t& t) : id(t.id), precision(t.precision), scale(t.scale) {}
  gdv_type_t c() const { return gdv_type_t{id, precision, scale}; }

  bool operator==(const DataType& o) const {
    if (id != o.id) return false;
    if (id == GDV_TYPE_DECIMAL128) return precision == o.precision && scale == o.scale;
    if (id == GDV_TYPE_TIMESTAMP || id == GDV_TYPE_TIME32 || id == GDV_TYPE_TIME64)
      return precision == o.precision;
    return true;
  }
  bool operator!=(const DataType& o) const { return !(*this == o); }

  bool is_varlen() const { return id == GDV_TYPE_STRING || id == GDV_TYPE_BINARY; }
  bool is_bool() const { return id == GDV_TYPE_BOOL; }
  bool is_decimal() const { return id == GDV_TYPE_DECIMAL128; }
  bool is_float() const { return id == GDV_TYPE_FLOAT || id == GDV_TYPE_DOUBLE; }
  bool is_integer() const { return id >= GDV_TYPE_UINT8 && id <= GDV_TYPE_INT64; }
  bool is_signed_integer() const {
    return id == GDV_TYPE_INT8 || id == GDV_TYPE_INT16 || id == GDV_TYPE_INT32 ||
           id == GDV_TYPE_INT64;
  }

  // Width in bytes of one value in the Arrow values buffer (0 for bool = bit-packed,
  // 4 for the offsets buffer of utf8/binary).
  int width() const;
  // arrow::DataType::ToString() spelling ("int32", "double", "decimal128(12, 2)", ...).
  std::string ToString() const;
  // C type used for a value of this type inside generated device code.
  const char* ctype() const;
  // Suffix used in precompiled function names ("int32", "float64", "utf8", ...), the
  // naming scheme of the reference's function library (add_int32_int32, ...).
  const char* fn_suffix() const;
};

inline DataType boolean() { return DataType(GDV_TYPE_BOOL); }
inline DataType int32() { return DataType(GDV_TYPE_INT32); }
inline DataType int64() { return DataType(GDV_TYPE_INT64); }
inline DataType float32() { return DataType(GDV_TYPE_FLOAT); }
inline DataType float64() { return DataType(GDV_TYPE_DOUBLE); }
inline DataType utf8() { return DataType(GDV_TYPE_STRING); }
inline DataType binary() { return DataType(GDV_TYPE_BINARY); }
inline DataType date32() { return DataType(GDV_TYPE_DATE32); }
inline DataType date64() { return DataType(GDV_TYPE_DATE64); }
inline DataType timestamp_ms() { return DataType(GDV_TYPE_TIMESTAMP, 1); }
inline DataType time32_ms() { return DataType(GDV_TYPE_TIME32, 1); }
inline DataType decimal128(int p, int s) { return DataType(GDV_TYPE_DECIMAL128, p, s); }

}  // namespace gdv
