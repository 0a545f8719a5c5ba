// Kernel fuser: lowers a set of validated expression trees to ONE CUDA kernel
// (source text) per Projector / Filter.  Replaces the reference's LLVMGenerator +
// Annotator + BitmapAccumulator (named in BASELINE.json north_star):
//   * Annotator        -> ColumnSlot table (which schema columns the kernel reads) and the
//                         flat gdv_args pointer block the kernel indexes;
//   * LLVMGenerator    -> EmitBody(): straight-line per-row code calling the device library;
//   * BitmapAccumulator-> validity is ANDed in registers and written with one warp ballot
//                         per 32 rows, inside the same kernel.
#pragma once
#include <string>
#include <vector>

#include "gdv_node.h"

namespace gdv {

struct Status {
  int code = GDV_OK;
  std::string msg;
  bool ok() const { return code == GDV_OK; }
  static Status OK() { return Status(); }
  static Status Make(int c, std::string m) {
    Status s;
    s.code = c;
    s.msg = std::move(m);
    return s;
  }
};

// Validate one expression against the schema and the function registry
// (ExpressionValidationError on failure).
Status ValidateExpression(const Schema& schema, const Expression& expr);

enum class KernelKind { kProject, kFilter, kStringSize, kStringWrite };

struct KernelSpec {
  KernelKind kind = KernelKind::kProject;
  int selection_mode = GDV_SEL_NONE;  // project: input selection; filter: output index width
  int rows_per_thread = 0;            // 0 = pick from bytes/row
  int block_threads = 0;              // 0 = pick per kernel shape
  bool large_batch = false;           // filter: the batch has >= 32 M rows (big tiles)
  std::string name;                   // kernel symbol
  bool nullable = true;               // false: specialised for batches where no input has nulls
  int loader = 0;                     // 0 = engine picks, 1 = direct LDG, 2 = TMA bulk -> shared
  int stages = 0;                     // projector, TMA loader: shared-memory stages per CTA (0 = pick);
                                      // fixed-width filter: 1024-row chunks every warp walks per tile (1/2/4/8)
  int string_scan = 0;                // bit 0: LIKE with the per-lane matcher only (no cooperative scan);
                                      // bit 2: row-driven string filter even where the key-driven one applies
};

struct ColumnSlot {
  int schema_index;
  DataType type;
  // Filter kernels: a null in this column makes the condition not-true whatever the other columns
  // hold, so its validity is ANDed into the tile's keep-mask 32 rows at a time after the row loop
  // instead of being read per row (BodyGen::TruthStrict).
  bool hoist = false;
};

struct GeneratedKernel {
  std::string source;   // full translation unit (device library is #include'd by name)
  std::string name;
  KernelKind kind;
  int rows_per_thread;
  int block_threads;
  int selection_mode;
  bool nullable = true;
  std::vector<ColumnSlot> inputs;   // kernel input slot j reads schema column inputs[j]
  std::vector<DataType> outputs;    // project: one per expression; filter: empty
  bool uses_ctx = false;            // some function can raise an ExecutionError
  int in_bytes_per_row = 0;         // algorithmic bytes (values only) read per row
  int out_bytes_per_row = 0;
  size_t args_size = 0;             // sizeof(gdv_args) for this kernel
  int dynamic_smem = 0;             // bytes of dynamic shared memory (string staging)
  int64_t tile_rows = 0;            // filter: rows per CTA tile (one look-back descriptor each)
  bool key_driven = false;          // string filter driven by the occurrences of a literal key in the column's bytes
  bool staged = false;              // project: inputs staged through shared memory by TMA bulk copies
  int stages = 0;                   // staged: shared-memory stages per CTA
  int64_t cta_tile_rows = 0;        // staged: rows per CTA tile (block_threads * rows_per_thread)
};

// Byte offsets inside gdv_args; the host packs the same layout (see EmitArgsStruct()).
struct ArgsLayout {
  int ni, no;  // array extents (>= 1)
  size_t off_n = 0, off_row_base = 8, off_sel = 16, off_out_idx = 24, off_out_count = 32,
         off_tile_state = 40, off_ticket = 48, off_err = 56, off_out_cap = 64, off_n_ptr = 72;
  size_t off_in_val, off_in_vld, off_in_var, off_out_val, off_out_vld, off_out_var, off_in_vsh,
      off_in_dsh, size;
  ArgsLayout(int n_inputs, int n_outputs);
};

Status GenerateKernel(const Schema& schema, const std::vector<ExpressionPtr>& exprs,
                      const KernelSpec& spec, GeneratedKernel* out);

}  // namespace gdv
