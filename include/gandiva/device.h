// gandiva/device.h — HBM residency for the reference's C++ API (extension; nothing in the reference
// corresponds to it: Gandiva is a CPU library).
//
// Projector::Evaluate / Filter::Evaluate (P/includes/libgandiva.pxd:218-226, 246-248) accept
// RecordBatches whose buffers live in GPU memory — any arrow::Buffer with is_cpu() == false and
// device_type() == kCUDA on the Configuration's device, e.g. Arrow's own CudaBuffer or the buffers of
// the MemoryManager below.  The kernels then read the batch in place (no PCIe copy), outputs are
// allocated from the MemoryManager of the batch's buffers (the MemoryPool argument is a host pool and is
// not used), and a Filter's SelectionVector must itself be backed by a device buffer
// (SelectionVector::MakeInt32(max_slots, buffer, &out) with a buffer from the same MemoryManager).
// A Filter -> Projector chain over such a batch exchanges eight bytes with the host (the slot count,
// which sizes the Projector's output arrays); the C-ABI's gdv_selection_t.d_num_slots removes even that.
#pragma once
#include "arrow/device.h"
#include "arrow/result.h"
#include "gandiva/arrow.h"

namespace gandiva {

/// The HBM of GPU `device` as an arrow::MemoryManager (device type CUDA), backed by the engine's
/// pooled allocator.  AllocateBuffer gives device buffers; MemoryManager::CopyBuffer /
/// RecordBatch::CopyTo / Array::CopyTo move data between it and the CPU memory manager.
GANDIVA_EXPORT arrow::Result<std::shared_ptr<arrow::MemoryManager>> DeviceMemoryManager(int device = 0);

/// Deep copy of a batch into the HBM of `device` / back into host memory (every buffer, as is).
GANDIVA_EXPORT arrow::Result<std::shared_ptr<arrow::RecordBatch>> CopyToDevice(const arrow::RecordBatch& batch,
                                                                              int device = 0);
GANDIVA_EXPORT arrow::Result<std::shared_ptr<arrow::RecordBatch>> CopyToHost(const arrow::RecordBatch& batch);
GANDIVA_EXPORT arrow::Result<std::shared_ptr<arrow::Array>> CopyToHost(const arrow::Array& array);

}  // namespace gandiva
