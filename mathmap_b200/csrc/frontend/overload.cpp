// Tag registry and overload resolution.
//
// Resolution rule restated from reference overload.c:212-279: entries are tried
// in registration order; an entry matches when every argument's (tag,length)
// unifies with the entry's pattern, where a pattern element is a constant, a
// named variable shared across the entry (T, L, ...), or the anonymous wildcard
// "_".  The first match wins and its result pattern is instantiated.
#include <cassert>
#include <cctype>
#include <cstdlib>

#include "frontend.h"

namespace mm {

Module::Module() {
    // tags.c:49-59: the eight well-known tags get numbers 1..8 in this order
    nil_tag = tag_number("nil");
    xy_tag = tag_number("xy");
    ra_tag = tag_number("ra");
    rgba_tag = tag_number("rgba");
    ri_tag = tag_number("ri");
    image_tag = tag_number("image");
    curve_tag = tag_number("curve");
    gradient_tag = tag_number("gradient");
    register_all_builtins(*this);
}

int Module::tag_number(const std::string &name) {
    for (size_t i = 0; i < tags.size(); ++i)
        if (tags[i] == name) return (int)i + 1;
    tags.push_back(name);
    return (int)tags.size();
}

Filter *Module::lookup_filter(const std::string &name) {
    for (auto &f : filters)
        if (f->name == name) return f.get();
    return nullptr;
}

// "TAG:LEN" where TAG is a lower-case tag name, an upper-case variable or "_",
// and LEN is a number, an upper-case variable or "_".
static OverloadPatternElem parse_elem(Module &m, const std::string &s) {
    size_t colon = s.find(':');
    assert(colon != std::string::npos);
    std::string t = s.substr(0, colon), l = s.substr(colon + 1);
    OverloadPatternElem e;
    if (t == "_") e.tag = 0;
    else if (isupper((unsigned char)t[0])) e.tag = -(t[0] - 'A' + 1);
    else e.tag = m.tag_number(t);
    if (l == "_") e.length = 0;
    else if (isupper((unsigned char)l[0])) e.length = -(l[0] - 'A' + 1);
    else e.length = atoi(l.c_str());
    return e;
}

static void parse_spec(Module &m, const char *spec, OverloadEntry &en) {
    std::string s;
    for (const char *p = spec; *p; ++p)
        if (!isspace((unsigned char)*p)) s += *p;
    size_t arrow = s.find("<-");
    assert(arrow != std::string::npos);
    en.result = parse_elem(m, s.substr(0, arrow));
    std::string rest = s.substr(arrow + 2);
    size_t pos = 0;
    while (pos < rest.size()) {
        size_t comma = rest.find(',', pos);
        if (comma == std::string::npos) comma = rest.size();
        en.args.push_back(parse_elem(m, rest.substr(pos, comma - pos)));
        pos = comma + 1;
    }
}

void Module::register_builtin(const char *name, const char *impl, const char *spec, BuiltinGen gen) {
    OverloadEntry en;
    en.name = name;
    en.impl_name = impl;
    en.gen = gen;
    parse_spec(*this, spec, en);
    overloads.push_back(std::move(en));
}

void Module::register_macro(const char *name, const char *spec, std::function<Expr *(std::vector<Expr *> &)> fn) {
    OverloadEntry en;
    en.name = name;
    en.impl_name = "macro";
    en.macro = std::move(fn);
    parse_spec(*this, spec, en);
    overloads.push_back(std::move(en));
}

const OverloadEntry *Module::resolve(const std::string &name, const std::vector<TupleInfo> &args, TupleInfo *result) const {
    for (const OverloadEntry &en : overloads) {
        if (en.name != name || en.args.size() != args.size()) continue;
        int tagvar[27] = {0}, lenvar[27] = {0};
        bool match = true;
        for (size_t i = 0; i < args.size() && match; ++i) {
            const OverloadPatternElem &p = en.args[i];
            if (p.tag > 0) {
                if (p.tag != args[i].tag) match = false;
            } else if (p.tag < 0) {
                int &slot = tagvar[-p.tag];
                if (slot == 0) slot = args[i].tag;
                else if (slot != args[i].tag) match = false;
            }
            if (!match) break;
            if (p.length > 0) {
                if (p.length != args[i].length) match = false;
            } else if (p.length < 0) {
                int &slot = lenvar[-p.length];
                if (slot == 0) slot = args[i].length;
                else if (slot != args[i].length) match = false;
            }
        }
        if (!match) continue;
        TupleInfo r;
        r.tag = en.result.tag > 0 ? en.result.tag : tagvar[-en.result.tag];
        r.length = en.result.length > 0 ? en.result.length : lenvar[-en.result.length];
        assert(r.tag > 0 && r.length > 0);
        *result = r;
        return &en;
    }
    return nullptr;
}

}  // namespace mm
