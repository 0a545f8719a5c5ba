// Pageable host memory <-> HBM through a ring of pinned staging slots.
//
// Evaluate() with GDV_MEM_HOST batches is the reference's own calling convention (RecordBatches that
// live in an Arrow MemoryPool: plain malloc'ed, pageable memory).  cuMemcpyHtoDAsync on pageable
// memory is staged by the driver through a small internal bounce buffer, one thread, well below what
// PCIe Gen5 moves; this ring does the same job with the machine's cores: worker threads copy piece
// i+1 of the caller's buffer into a pinned slot while the DMA engine moves piece i, so the transfer
// runs at the smaller of the host's memcpy bandwidth and the link.  Buffers that are already pinned
// (cuMemHostAlloc / cuMemHostRegister) are handed to the DMA engine directly.
#pragma once
#include <cstddef>

#include "gdv_cuda.h"

namespace gdv {

class Device;

// true when `p` is ordinary pageable host memory (unknown to the driver).
bool IsPageableHost(const void* p);

// H2D: on return every byte of `src` has been read (the caller may reuse it) and the copies are
// enqueued on `stream`.  Small or pinned buffers go straight to cuMemcpyHtoDAsync.
Status StagedHtoD(Device* dev, CUdeviceptr dst, const void* src, size_t bytes, CUstream stream);
// D2H: stream-ordered after the work already queued on `stream`.  Pageable destinations are complete
// on return (the call synchronises piecewise); pinned ones are only enqueued, like cuMemcpyDtoHAsync.
Status StagedDtoH(Device* dev, void* dst, CUdeviceptr src, size_t bytes, CUstream stream);

// Counters for tests / bench: bytes that went through the ring since load.
long long StagedBytes();

}  // namespace gdv
