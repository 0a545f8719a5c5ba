// gandiva/expression.h (P/includes/libgandiva.pxd:33-41).
#pragma once
#include "gandiva/node.h"

namespace gandiva {

class GANDIVA_EXPORT Expression {
 public:
  Expression(NodePtr root, FieldPtr result);
  virtual ~Expression();
  const NodePtr& root() const { return root_; }
  const FieldPtr& result() const { return result_; }
  std::string ToString() const;
  void* handle() const { return handle_; }

 protected:
  NodePtr root_;
  FieldPtr result_;
  void* handle_ = nullptr;  // gdv_expression_t
};

using ExpressionPtr = std::shared_ptr<Expression>;
using ExpressionVector = std::vector<ExpressionPtr>;

}  // namespace gandiva
