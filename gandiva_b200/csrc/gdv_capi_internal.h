// Handle types and error plumbing shared by the extern "C" translation units
// (gdv_capi.cc, gdv_arrow_c.cc).
#pragma once
#include <memory>
#include <string>

#include "gandiva_b200.h"
#include "gdv_node.h"
#include "gdv_runtime.h"

namespace gdv {
namespace capi {

// Records the thread-local message gdv_last_error() returns and passes the code through.
gdv_status Fail(const Status& s);
gdv_status Fail(int code, const std::string& msg);

struct NodeH { NodePtr p; };
struct ExprH { ExpressionPtr p; };
struct CondH { ConditionPtr p; };
struct SchemaH { SchemaPtr p; };
struct ProjH { std::shared_ptr<Projector> p; };
struct FiltH { std::shared_ptr<Filter> p; };

}  // namespace capi
}  // namespace gdv
