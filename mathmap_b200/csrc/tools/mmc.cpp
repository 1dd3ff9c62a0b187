// mmc: compile a .mm file and print the optimised IR of every filter (debug tool).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../frontend/frontend.h"
#include "../ir/passes.h"

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: mmc file.mm [--no-opt]\n"); return 2; }
    bool opt = !(argc > 2 && std::string(argv[2]) == "--no-opt");
    std::ifstream in(argv[1]);
    if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::stringstream ss;
    ss << in.rdbuf();
    try {
        mm::Module m;
        mm::parse_module(m, ss.str());
        std::vector<std::unique_ptr<mm::FilterCode>> codes;
        std::vector<const mm::FilterCode *> ptrs;
        for (auto &f : m.filters) {
            if (f->kind != mm::FILTER_MATHMAP) continue;
            codes.push_back(mm::compile_filter(m, f.get(), opt));
            ptrs.push_back(codes.back().get());
        }
        std::cout << mm::dump_module_ir(ptrs, m.main_filter->name);
    } catch (mm::CompileError &e) {
        fprintf(stderr, "%s:%d:%d: %s\n", argv[1], e.line + 1, e.column + 1, e.message.c_str());
        return 1;
    }
    return 0;
}
