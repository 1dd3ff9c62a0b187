// C ABI, front-end half: compile, IR text, userval metadata.  The render half
// lives in backend/invocation.cpp.  See include/mathmap_b200.h for the contract.
#include <cstring>
#include <string>

#include "../../include/mathmap_b200.h"
#include "cabi_module.h"
#include "ir/passes.h"

static thread_local std::string g_last_error;

namespace mmbackend {
void set_error(const std::string &msg) { g_last_error = msg; }
std::unique_ptr<mmb_module> load_ir_text(const std::string &text);  // ir/ir_parse.cpp
}  // namespace mmbackend

extern "C" {

const char *mmb_last_error(void) { return g_last_error.c_str(); }
const char *mmb_version(void) { return "mathmap_b200 0.1 (sm_100a)"; }

mmb_module *mmb_compile(const char *source) {
    if (!source) { g_last_error = "mmb_compile: null source"; return nullptr; }
    auto m = std::make_unique<mmb_module>();
    try {
        m->mod = std::make_unique<mm::Module>();
        mm::parse_module(*m->mod, source);
        std::vector<const mm::FilterCode *> ptrs;
        for (auto &f : m->mod->filters) {
            if (f->kind != mm::FILTER_MATHMAP) { m->codes.push_back(nullptr); continue; }
            m->codes.push_back(mm::compile_filter(*m->mod, f.get(), true));
            ptrs.push_back(m->codes.back().get());
        }
        m->main = m->mod->main_filter;
        m->ir_text = mm::dump_module_ir(ptrs, m->main->name);
    } catch (mm::CompileError &e) {
        char buf[64];
        snprintf(buf, sizeof buf, "%d:%d: ", e.line + 1, e.column + 1);
        g_last_error = (e.line >= 0 ? std::string(buf) : std::string()) + e.message;
        return nullptr;
    } catch (std::exception &e) {
        g_last_error = std::string("internal error: ") + e.what();
        return nullptr;
    }
    return m.release();
}

char *mmb_design_to_source(const char *design_text, const char *filter_search_path) {
    if (!design_text || !filter_search_path) { g_last_error = "mmb_design_to_source: null argument"; return nullptr; }
    try {
        std::string src = mm::design_to_source(design_text, filter_search_path);
        char *out = (char *)malloc(src.size() + 1);
        if (!out) { g_last_error = "out of memory"; return nullptr; }
        memcpy(out, src.c_str(), src.size() + 1);
        return out;
    } catch (mm::CompileError &e) {
        g_last_error = e.message;
        return nullptr;
    } catch (std::exception &e) {
        g_last_error = std::string("internal error: ") + e.what();
        return nullptr;
    }
}

void mmb_free_string(char *s) { free(s); }

mmb_module *mmb_compile_design(const char *design_text, const char *filter_search_path) {
    char *src = mmb_design_to_source(design_text, filter_search_path);
    if (!src) return nullptr;
    mmb_module *m = mmb_compile(src);
    free(src);
    return m;
}

mmb_module *mmb_load_ir(const char *ir_text) {
    if (!ir_text) { g_last_error = "mmb_load_ir: null text"; return nullptr; }
    try {
        auto m = mmbackend::load_ir_text(ir_text);
        return m.release();
    } catch (mm::CompileError &e) {
        g_last_error = e.message;
        return nullptr;
    } catch (std::exception &e) {
        g_last_error = std::string("internal error: ") + e.what();
        return nullptr;
    }
}

void mmb_module_free(mmb_module *m) { delete m; }
const char *mmb_module_ir(const mmb_module *m) { return m ? m->ir_text.c_str() : nullptr; }
const char *mmb_module_main_filter_name(const mmb_module *m) { return m ? m->main->name.c_str() : nullptr; }
int mmb_module_num_uservals(const mmb_module *m) { return m ? (int)m->main->uservals.size() : -1; }

int mmb_module_userval_info(const mmb_module *m, int index, char *name, size_t name_len, int *type, float *min_value, float *max_value,
                            float *default_value) {
    if (!m || index < 0 || index >= (int)m->main->uservals.size()) { g_last_error = "userval index out of range"; return -1; }
    const mm::UservalInfo &u = m->main->uservals[index];
    if (name && name_len) {
        strncpy(name, u.name.c_str(), name_len - 1);
        name[name_len - 1] = 0;
    }
    if (type) *type = u.type;
    float lo = 0, hi = 0, d = 0;
    switch (u.type) {
    case mm::UV_INT: lo = (float)u.int_min; hi = (float)u.int_max; d = (float)u.int_default; break;
    case mm::UV_FLOAT: lo = u.float_min; hi = u.float_max; d = u.float_default; break;
    case mm::UV_BOOL: lo = 0; hi = 1; d = (float)u.bool_default; break;
    default: break;
    }
    if (min_value) *min_value = lo;
    if (max_value) *max_value = hi;
    if (default_value) *default_value = d;
    return 0;
}

int mmb_module_userval_index(const mmb_module *m, const char *name) {
    if (!m || !name) return -1;
    for (auto &u : m->main->uservals)
        if (u.name == name) return u.index;
    return -1;
}

}  // extern "C"
