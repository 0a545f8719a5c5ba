// gandiva/filter.h (P/includes/libgandiva.pxd:242-256).
#pragma once
#include "gandiva/condition.h"
#include "gandiva/configuration.h"
#include "gandiva/selection_vector.h"

namespace gandiva {

/// Filter: fills a caller-allocated SelectionVector with the ascending indices of the rows
/// whose condition is true and valid (one fused predicate + ordered-compaction kernel).
class GANDIVA_EXPORT Filter {
 public:
  ~Filter();
  static Status Make(SchemaPtr schema, ConditionPtr condition, std::shared_ptr<Filter>* filter);
  static Status Make(SchemaPtr schema, ConditionPtr condition,
                     std::shared_ptr<Configuration> configuration,
                     std::shared_ptr<Filter>* filter);
  Status Evaluate(const arrow::RecordBatch& batch,
                  std::shared_ptr<SelectionVector> out_selection);
  std::string DumpIR();

 private:
  Filter() = default;
  SchemaPtr schema_;
  ConditionPtr condition_;
  void* handle_ = nullptr;  // gdv_filter_t
  void* schema_handle_ = nullptr;
};

}  // namespace gandiva
