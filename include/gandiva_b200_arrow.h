/*
 * gandiva_b200_arrow.h — ingestion through the Arrow C Data / C Device Data interfaces
 * (SURVEY.md §8(f)4).  An upstream producer (cuDF, a GPU scan, another engine) hands a record batch
 * over as one struct-typed `ArrowDeviceArray`; its buffers are used in place — device buffers stay
 * in HBM, nothing is copied — and results go back the same way.  Still plain C: the three structs
 * below are the published Arrow ABI (P/include/arrow/c/abi.h:68-121 ArrowSchema / ArrowArray,
 * :140-287 ArrowDeviceArray) and are only defined here when arrow/c/abi.h has not been included.
 *
 * What this replaces in the reference: arrow::RecordBatch as the argument of
 * Projector::Evaluate / Filter::Evaluate (P/includes/libgandiva.pxd:218-226,246-248); the
 * reference has no device-memory input path at all.
 */
#ifndef GANDIVA_B200_ARROW_H
#define GANDIVA_B200_ARROW_H

#include "gandiva_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif /* ARROW_C_DATA_INTERFACE */

#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE
typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
#define ARROW_DEVICE_CUDA_HOST 3
#define ARROW_DEVICE_CUDA_MANAGED 13
struct ArrowDeviceArray {
  struct ArrowArray array;
  int64_t device_id;
  ArrowDeviceType device_type;
  void* sync_event; /* CUDA: cudaEvent_t* (pointer to the event), or NULL */
  int64_t reserved[3];
};
#endif /* ARROW_C_DEVICE_DATA_INTERFACE */

/* Schema of the engine from a struct-typed ("+s") ArrowSchema; `schema` is only read.
 * Supported child formats: b c C s S i I l L f g u z tdD tdm tss/tsm/tsu/tsn (any tz) tts ttm ttu
 * ttn d:p,s (128 bit).  Anything else -> GDV_NOT_IMPLEMENTED. */
gdv_status gdv_schema_from_arrow(const struct ArrowSchema* schema, gdv_schema_t* out);

/* A record batch imported from a struct-typed ArrowDeviceArray.  Import MOVES the array (the
 * source struct is marked released, as the C Data interface prescribes for consumers) and keeps it
 * alive until gdv_arrow_batch_release.  Device type CPU / CUDA_HOST -> GDV_MEM_HOST (staged by
 * Evaluate), CUDA / CUDA_MANAGED -> GDV_MEM_DEVICE (used in place).  Columns whose null_count is 0
 * are presented without a validity buffer, which selects the no-null kernel variant. */
typedef struct gdv_arrow_batch_s* gdv_arrow_batch_t;
gdv_status gdv_arrow_batch_import(struct ArrowDeviceArray* array, gdv_schema_t schema,
                                  gdv_arrow_batch_t* out);
/* The zero-copy view usable with gdv_projector_evaluate / gdv_filter_evaluate. */
const gdv_batch_t* gdv_arrow_batch_view(gdv_arrow_batch_t batch);
/* Orders `stream` after the producer's sync_event (cuStreamWaitEvent); no-op without one. */
gdv_status gdv_arrow_batch_wait(gdv_arrow_batch_t batch, void* stream);
void gdv_arrow_batch_release(gdv_arrow_batch_t batch);

/* Schema of a projector's outputs ("+s", one child per expression, named after the result
 * fields); the caller owns `out` and releases it through out->release. */
gdv_status gdv_projector_output_schema_arrow(gdv_projector_t p, struct ArrowSchema* out);

/* Projector::Evaluate over an imported batch; the result is exported as a struct-typed
 * ArrowDeviceArray in the memory space of the input: device batches produce device buffers (from
 * the engine's pooled allocator, returned to it by out->array.release), host batches produce host
 * buffers.  The call waits on the input's sync_event, runs on `stream` (NULL = the engine's
 * stream) and returns when the results are complete, so out->sync_event is NULL and
 * ExecutionErrors are reported here.  null_count of the children is -1 (not computed). */
gdv_status gdv_projector_evaluate_arrow(gdv_projector_t p, gdv_arrow_batch_t batch, void* stream,
                                        struct ArrowDeviceArray* out);
/* Filter::Evaluate over an imported batch; `out` is a primitive array (format S / I / L for
 * mode GDV_SEL_UINT16/32/64) of the selected row indices, in the memory space of the input. */
gdv_status gdv_filter_evaluate_arrow(gdv_filter_t f, gdv_arrow_batch_t batch, int32_t mode,
                                     void* stream, struct ArrowDeviceArray* out);

/* Harness helper: copy between host and device memory on the engine's stream, synchronously.
 * kind 1 = host -> device, 2 = device -> host. */
gdv_status gdv_memcpy(int32_t device, void* dst, const void* src, size_t bytes, int32_t kind);

#ifdef __cplusplus
}
#endif
#endif /* GANDIVA_B200_ARROW_H */
