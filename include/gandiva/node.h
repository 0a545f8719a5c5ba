// gandiva/node.h — expression-tree node (P/includes/libgandiva.pxd:27-41).
#pragma once
#include "gandiva/arrow.h"

namespace gandiva {

/// A node of the expression tree.  Thin handle over the engine's native node
/// (include/gandiva_b200.h gdv_node_t); built only through TreeExprBuilder.
class GANDIVA_EXPORT Node {
 public:
  Node(DataTypePtr return_type, void* handle);
  ~Node();
  Node(const Node&) = delete;
  Node& operator=(const Node&) = delete;

  const DataTypePtr& return_type() const { return return_type_; }
  std::string ToString() const;
  void* handle() const { return handle_; }

 private:
  DataTypePtr return_type_;
  void* handle_;
};

using NodePtr = std::shared_ptr<Node>;
using NodeVector = std::vector<NodePtr>;

}  // namespace gandiva
