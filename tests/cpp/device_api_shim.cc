// device_api_shim.cc — TEST code: drives the namespace-gandiva C++ API (libgandiva.so) the way a C++
// caller would, with the RecordBatch resident in HBM, and hands the results back to pytest through the
// Arrow C Data interface.  tests/test_cpp_device_api.py builds the same trees for the oracle and compares.
//
//   host batch --CopyToDevice--> Filter::Evaluate(batch, device SelectionVector)
//                                 --> Projector::Evaluate(batch, selection, pool, &outputs)   (outputs in HBM)
//                                 --CopyToHost--> returned
// Schema: k int32, d float64, q int64, e float64, x decimal128(15,2).
//   condition: k >= 8766 AND k < 9131 AND d >= 0.05 AND d <= 0.07 AND q < 24
//   outputs:   e * (1.0 - d);  q + q;  if (d > 0.05) e else 0.0;  x * x as decimal128(31,4);
//              if (k <= 10471) q else NULL
#include <cstdio>
#include <cstring>

#include "arrow/api.h"
#include "arrow/c/bridge.h"
#include "gandiva/device.h"
#include "gandiva/filter.h"
#include "gandiva/projector.h"
#include "gandiva/selection_vector.h"
#include "gandiva/tree_expr_builder.h"

using gandiva::NodePtr;
using gandiva::TreeExprBuilder;

namespace {

arrow::Status Run(const std::shared_ptr<arrow::RecordBatch>& host, bool on_device,
                  std::shared_ptr<arrow::Array>* sel_out, std::shared_ptr<arrow::RecordBatch>* proj_out) {
  auto schema = host->schema();
  auto f = [&](const char* name) { return TreeExprBuilder::MakeField(schema->GetFieldByName(name)); };
  auto B = arrow::boolean();
  auto F64 = arrow::float64();
  auto I64 = arrow::int64();
  auto fn = [](const char* name, gandiva::NodeVector kids, gandiva::DataTypePtr t) {
    return TreeExprBuilder::MakeFunction(name, kids, t);
  };
  NodePtr cond = TreeExprBuilder::MakeAnd(
      {fn("greater_than_or_equal_to", {f("k"), TreeExprBuilder::MakeLiteral(int32_t(8766))}, B),
       fn("less_than", {f("k"), TreeExprBuilder::MakeLiteral(int32_t(9131))}, B),
       fn("greater_than_or_equal_to", {f("d"), TreeExprBuilder::MakeLiteral(0.05)}, B),
       fn("less_than_or_equal_to", {f("d"), TreeExprBuilder::MakeLiteral(0.07)}, B),
       fn("less_than", {f("q"), TreeExprBuilder::MakeLiteral(int64_t(24))}, B)});
  auto D31 = arrow::decimal128(31, 4);
  gandiva::ExpressionVector exprs = {
      TreeExprBuilder::MakeExpression(fn("multiply", {f("e"), fn("subtract", {TreeExprBuilder::MakeLiteral(1.0), f("d")}, F64)}, F64),
                                      arrow::field("o0", F64)),
      TreeExprBuilder::MakeExpression(fn("add", {f("q"), f("q")}, I64), arrow::field("o1", I64)),
      TreeExprBuilder::MakeExpression(
          TreeExprBuilder::MakeIf(fn("greater_than", {f("d"), TreeExprBuilder::MakeLiteral(0.05)}, B), f("e"),
                                  TreeExprBuilder::MakeLiteral(0.0), F64),
          arrow::field("o2", F64)),
      TreeExprBuilder::MakeExpression(fn("multiply", {f("x"), f("x")}, D31), arrow::field("o3", D31)),
      TreeExprBuilder::MakeExpression(
          TreeExprBuilder::MakeIf(fn("less_than_or_equal_to", {f("k"), TreeExprBuilder::MakeLiteral(int32_t(10471))}, B),
                                  f("q"), TreeExprBuilder::MakeNull(I64), I64),
          arrow::field("o4", I64))};
  std::shared_ptr<gandiva::Filter> filter;
  ARROW_RETURN_NOT_OK(gandiva::Filter::Make(schema, TreeExprBuilder::MakeCondition(cond), &filter));
  std::shared_ptr<gandiva::Projector> projector;
  ARROW_RETURN_NOT_OK(gandiva::Projector::Make(schema, exprs, gandiva::SelectionVector::MODE_UINT32,
                                               gandiva::ConfigurationBuilder::DefaultConfiguration(), &projector));
  const int64_t n = host->num_rows();
  std::shared_ptr<gandiva::SelectionVector> sel;
  std::shared_ptr<arrow::RecordBatch> batch = host;
  if (on_device) {
    ARROW_ASSIGN_OR_RAISE(batch, gandiva::CopyToDevice(*host, 0));
    for (const auto& col : batch->columns())
      for (const auto& buf : col->data()->buffers)
        if (buf != nullptr && buf->is_cpu()) return arrow::Status::Invalid("CopyToDevice left a buffer on the host");
    ARROW_ASSIGN_OR_RAISE(std::shared_ptr<arrow::MemoryManager> mm, gandiva::DeviceMemoryManager(0));
    ARROW_ASSIGN_OR_RAISE(std::unique_ptr<arrow::Buffer> idx, mm->AllocateBuffer(n * 4));
    ARROW_RETURN_NOT_OK(gandiva::SelectionVector::MakeInt32(n, std::shared_ptr<arrow::Buffer>(std::move(idx)), &sel));
  } else {
    ARROW_RETURN_NOT_OK(gandiva::SelectionVector::MakeInt32(n, arrow::default_memory_pool(), &sel));
  }
  ARROW_RETURN_NOT_OK(filter->Evaluate(*batch, sel));
  arrow::ArrayVector outs;
  ARROW_RETURN_NOT_OK(projector->Evaluate(*batch, sel.get(), arrow::default_memory_pool(), &outs));
  std::shared_ptr<arrow::Array> sel_arr = sel->ToArray();
  if (on_device) {
    if (sel_arr->data()->buffers[1]->is_cpu()) return arrow::Status::Invalid("the selection vector is not in HBM");
    ARROW_ASSIGN_OR_RAISE(sel_arr, gandiva::CopyToHost(*sel_arr));
    for (auto& o : outs) {
      for (const auto& buf : o->data()->buffers)
        if (buf != nullptr && buf->is_cpu()) return arrow::Status::Invalid("a projector output is not in HBM");
      ARROW_ASSIGN_OR_RAISE(o, gandiva::CopyToHost(*o));
    }
  }
  *sel_out = sel_arr;
  arrow::FieldVector fields;
  for (const auto& e : exprs) fields.push_back(e->result());
  *proj_out = arrow::RecordBatch::Make(arrow::schema(fields), sel->GetNumSlots(), outs);
  return arrow::Status::OK();
}

}  // namespace

extern "C" int shim_filter_then_project(struct ArrowArray* in_array, struct ArrowSchema* in_schema, int on_device,
                                        struct ArrowArray* sel_array, struct ArrowSchema* sel_schema,
                                        struct ArrowArray* out_array, struct ArrowSchema* out_schema, char* err,
                                        int err_len) {
  auto fail = [&](const arrow::Status& st) {
    std::snprintf(err, static_cast<size_t>(err_len), "%s", st.ToString().c_str());
    return 1;
  };
  auto imported = arrow::ImportRecordBatch(in_array, in_schema);
  if (!imported.ok()) return fail(imported.status());
  std::shared_ptr<arrow::Array> sel;
  std::shared_ptr<arrow::RecordBatch> out;
  arrow::Status st = Run(*imported, on_device != 0, &sel, &out);
  if (!st.ok()) return fail(st);
  st = arrow::ExportArray(*sel, sel_array, sel_schema);
  if (!st.ok()) return fail(st);
  st = arrow::ExportRecordBatch(*out, out_array, out_schema);
  if (!st.ok()) return fail(st);
  return 0;
}
