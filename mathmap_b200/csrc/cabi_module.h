// Internal definition of the opaque C-ABI handles.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "frontend/frontend.h"
#include "ir/ir.h"

namespace mmbackend { struct ModuleBackend; struct InvocationBackend; }

struct mmb_module {
    std::unique_ptr<mm::Module> mod;
    std::vector<std::unique_ptr<mm::FilterCode>> codes;  // index-aligned with mod->filters (null for native filters)
    mm::Filter *main = nullptr;
    std::string ir_text;
    std::string cuda_source;
    std::shared_ptr<mmbackend::ModuleBackend> backend;
    const mm::FilterCode *code_for(const mm::Filter *f) const { return codes[f->index].get(); }
};
