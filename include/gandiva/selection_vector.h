// gandiva/selection_vector.h (P/includes/libgandiva.pxd:43-71): ascending row indices.
#pragma once
#include "arrow/buffer.h"
#include "gandiva/arrow.h"

namespace gandiva {

class GANDIVA_EXPORT SelectionVector {
 public:
  enum Mode : int { MODE_NONE, MODE_UINT16, MODE_UINT32, MODE_UINT64, MODE_MAX = MODE_UINT64 };

  SelectionVector(Mode mode, int64_t max_slots, std::shared_ptr<arrow::Buffer> buffer,
                  int64_t num_slots = 0)
      : mode_(mode), max_slots_(max_slots), num_slots_(num_slots), buffer_(std::move(buffer)) {}

  uint64_t GetIndex(int64_t index) const;
  void SetIndex(int64_t index, uint64_t value);
  int64_t GetMaxSlots() const { return max_slots_; }
  int64_t GetNumSlots() const { return num_slots_; }
  void SetNumSlots(int64_t num_slots) { num_slots_ = num_slots; }
  uint64_t GetMaxSupportedValue() const;
  Mode GetMode() const { return mode_; }
  arrow::Buffer& GetBuffer() const { return *buffer_; }
  /// Indices as a uint16/uint32/uint64 Arrow array of length num_slots.
  ArrayPtr ToArray() const;

  static Status MakeInt16(int64_t max_slots, arrow::MemoryPool* pool,
                          std::shared_ptr<SelectionVector>* selection_vector);
  static Status MakeInt32(int64_t max_slots, arrow::MemoryPool* pool,
                          std::shared_ptr<SelectionVector>* selection_vector);
  static Status MakeInt64(int64_t max_slots, arrow::MemoryPool* pool,
                          std::shared_ptr<SelectionVector>* selection_vector);
  static Status MakeInt16(int64_t max_slots, std::shared_ptr<arrow::Buffer> buffer,
                          std::shared_ptr<SelectionVector>* selection_vector);
  static Status MakeInt32(int64_t max_slots, std::shared_ptr<arrow::Buffer> buffer,
                          std::shared_ptr<SelectionVector>* selection_vector);
  static Status MakeInt64(int64_t max_slots, std::shared_ptr<arrow::Buffer> buffer,
                          std::shared_ptr<SelectionVector>* selection_vector);

 private:
  static Status Make(Mode mode, int64_t max_slots, arrow::MemoryPool* pool,
                     std::shared_ptr<SelectionVector>* out);
  Mode mode_;
  int64_t max_slots_;
  int64_t num_slots_;
  std::shared_ptr<arrow::Buffer> buffer_;
};

}  // namespace gandiva
