// gandiva/expression_registry.h (P/includes/libgandiva.pxd:274-277).
#pragma once
#include "gandiva/function_signature.h"

namespace gandiva {

GANDIVA_EXPORT std::vector<std::shared_ptr<FunctionSignature>> GetRegisteredFunctionSignatures();

}  // namespace gandiva
