#pragma once
#include <memory>

#include "../frontend/frontend.h"
#include "ir.h"

namespace mm {

// Lowers one filter to optimised, typed, constness-annotated IR.
std::unique_ptr<FilterCode> compile_filter(Module &mod, Filter *filter, bool optimize = true);
void propagate_types(FilterCode &code);
void analyze_constants(FilterCode &code);
// loop-carried values (passes.cpp) on IR loaded from text; before propagate_types / analyze_constants
void carry_loop_values(Module &mod, FilterCode &code);

}  // namespace mm
