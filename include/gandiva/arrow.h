// gandiva/arrow.h — Arrow aliases used by the public API (P/includes/libgandiva.pxd:105-107).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "arrow/array.h"
#include "arrow/memory_pool.h"
#include "arrow/record_batch.h"
#include "arrow/status.h"
#include "arrow/type.h"

#ifndef GANDIVA_EXPORT
#define GANDIVA_EXPORT __attribute__((visibility("default")))
#endif

namespace gandiva {

using arrow::Status;
using ArrayPtr = std::shared_ptr<arrow::Array>;
using DataTypePtr = std::shared_ptr<arrow::DataType>;
using DataTypeVector = std::vector<DataTypePtr>;
using FieldPtr = std::shared_ptr<arrow::Field>;
using FieldVector = std::vector<FieldPtr>;
using RecordBatchPtr = std::shared_ptr<arrow::RecordBatch>;
using SchemaPtr = std::shared_ptr<arrow::Schema>;
using ArrayDataPtr = std::shared_ptr<arrow::ArrayData>;
using ArrayDataVector = std::vector<ArrayDataPtr>;
using ArrayVector = std::vector<ArrayPtr>;

}  // namespace gandiva
