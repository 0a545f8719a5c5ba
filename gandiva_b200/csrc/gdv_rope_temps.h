// Consumers of a rope.  concat / repeat / space / lpad / rpad / reverse / replace build their result as a ROPE of
// pieces that only the string write pass can read: projecting it, concatenating it again or choosing
// it with if/else is free, but like(concat(a, b), '%x%'), equal(reverse(s), t), substr(lpad(...), 2) ... need the
// bytes in one place.  When the fuser reports such a consumer, the Projector / Filter is built in two stages
// instead: an internal Projector materialises every consumed rope into a temporary utf8 column in device memory
// (sizing pass + write pass, the ordinary string-output path), and the caller's expressions — with each rope
// replaced by a field of an extended schema — run over the batch plus those columns.  Two extra passes over the
// rope's bytes; nothing changes for expressions the fuser takes directly.
#pragma once

#include <string>
#include <vector>

#include "gdv_codegen.h"
#include "gdv_node.h"

namespace gdv {

struct RopeTemps {
  std::vector<ExpressionPtr> temps;   // one utf8/binary expression per materialised rope
  std::vector<Field> fields;          // the fields that stand for them (appended to the schema)
};

// True for the Status the fuser returns when something other than projection / concat / if-else reads a rope.
bool IsRopeConsumerError(const Status& s);

// Replaces every rope that is read by such a consumer in `root` with a field (`__gdv_rope_<k>`), adding the
// rope's expression to `out`.  `root_is_output`: a rope AT the root of a projector output stays where it is.
// A consumer inside the arguments of a materialised rope is left in the temp's expression: the internal
// Projector for the temps is built through the same path, one level of temporaries per nesting level.
bool ExtractRopes(const NodePtr& root, bool root_is_output, RopeTemps* out, NodePtr* rewritten);

}  // namespace gdv
