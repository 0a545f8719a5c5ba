// gandiva/projector.h (P/includes/libgandiva.pxd:214-240).
#pragma once
#include "gandiva/configuration.h"
#include "gandiva/expression.h"
#include "gandiva/selection_vector.h"

namespace gandiva {

/// Projection of a RecordBatch through one fused CUDA kernel compiled at Make().
class GANDIVA_EXPORT Projector {
 public:
  ~Projector();
  static Status Make(SchemaPtr schema, const ExpressionVector& exprs,
                     std::shared_ptr<Projector>* projector);
  static Status Make(SchemaPtr schema, const ExpressionVector& exprs,
                     std::shared_ptr<Configuration> configuration,
                     std::shared_ptr<Projector>* projector);
  static Status Make(SchemaPtr schema, const ExpressionVector& exprs,
                     SelectionVector::Mode selection_vector_mode,
                     std::shared_ptr<Configuration> configuration,
                     std::shared_ptr<Projector>* projector);

  Status Evaluate(const arrow::RecordBatch& batch, arrow::MemoryPool* pool,
                  arrow::ArrayVector* output) const;
  Status Evaluate(const arrow::RecordBatch& batch, const SelectionVector* selection_vector,
                  arrow::MemoryPool* pool, arrow::ArrayVector* output) const;
  /// The generated CUDA source (and PTX when Configuration::dump_ir is set).
  std::string DumpIR();

 private:
  Projector() = default;
  SchemaPtr schema_;
  FieldVector output_fields_;
  ExpressionVector exprs_;  // keep the tree alive
  SelectionVector::Mode mode_ = SelectionVector::MODE_NONE;
  void* handle_ = nullptr;  // gdv_projector_t
  void* schema_handle_ = nullptr;
};

}  // namespace gandiva
