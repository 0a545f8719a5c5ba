// gandiva/condition.h (P/includes/libgandiva.pxd:98-103): an Expression whose result is the
// boolean field "cond".
#pragma once
#include "gandiva/expression.h"

namespace gandiva {

class GANDIVA_EXPORT Condition : public Expression {
 public:
  explicit Condition(NodePtr root);
  ~Condition() override;
  void* condition_handle() const { return cond_handle_; }

 private:
  void* cond_handle_ = nullptr;  // gdv_condition_t
};

using ConditionPtr = std::shared_ptr<Condition>;

}  // namespace gandiva
