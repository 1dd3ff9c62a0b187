// Compositions (".mmc" designs): a graph of filter nodes that the reference's composer saves as an
// S-expression and turns into MathMap source before compiling it like any other filter.
//
// WHAT is restated: the file format read by designer/loadsave.c:30-205 (designer_load_design), the source text
// produced by designer_filter.c:28-298 (make_filter_source_from_design: included filter sources, the argument list of
// the composite filter, one assignment per node in dependency order, `root_out(xy)`), and the lookup of a node type
// by the NAME of the main filter of a file found under a directory tree (expression_db.c:65-200 read_expression_db,
// :378-444 fetch_expression_mathmap / get_expression_name / get_expression_args).  HOW is ours: a small reader and
// std::filesystem instead of the reference's lispreader and glib.
#include <algorithm>
#include <filesystem>
#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <sstream>

#include "frontend.h"

namespace mm {

namespace {

// ---------------------------------------------------------------- S-expression reader (strings, symbols, numbers, lists)
struct SNode {
    enum Kind { LIST, STRING, ATOM } kind = ATOM;
    std::string text;
    std::vector<SNode> items;
};

struct SReader {
    const std::string &s;
    size_t pos = 0;
    explicit SReader(const std::string &text) : s(text) {}
    void skip() {
        for (;;) {
            while (pos < s.size() && isspace((unsigned char)s[pos])) ++pos;
            if (pos < s.size() && s[pos] == ';') { while (pos < s.size() && s[pos] != '\n') ++pos; continue; }
            break;
        }
    }
    bool read(SNode &out, std::string &err) {
        skip();
        if (pos >= s.size()) { err = "unexpected end of design"; return false; }
        char c = s[pos];
        if (c == '(') {
            ++pos;
            out.kind = SNode::LIST;
            for (;;) {
                skip();
                if (pos >= s.size()) { err = "unterminated list in design"; return false; }
                if (s[pos] == ')') { ++pos; return true; }
                SNode item;
                if (!read(item, err)) return false;
                out.items.push_back(std::move(item));
            }
        }
        if (c == ')') { err = "unexpected ')' in design"; return false; }
        if (c == '"') {
            ++pos;
            out.kind = SNode::STRING;
            while (pos < s.size() && s[pos] != '"') {
                if (s[pos] == '\\' && pos + 1 < s.size()) ++pos;
                out.text += s[pos++];
            }
            if (pos >= s.size()) { err = "unterminated string in design"; return false; }
            ++pos;
            return true;
        }
        out.kind = SNode::ATOM;
        while (pos < s.size() && !isspace((unsigned char)s[pos]) && s[pos] != '(' && s[pos] != ')' && s[pos] != '"') out.text += s[pos++];
        return true;
    }
};

// value following the keyword `key` in a property list (loadsave.c uses lisp_proplist_lookup_symbol)
const SNode *prop(const std::vector<SNode> &items, size_t from, const char *key) {
    for (size_t i = from; i + 1 < items.size(); ++i)
        if (items[i].kind == SNode::ATOM && items[i].text == key) return &items[i + 1];
    return nullptr;
}

struct DesignNode {
    std::string name, type;
    std::vector<std::pair<std::string, std::string>> inputs;  // (input slot name, source node name)
};
struct Design {
    std::string name = "__untitled_design__", root;
    std::vector<DesignNode> nodes;
    const DesignNode *node(const std::string &n) const {
        for (auto &d : nodes) if (d.name == n) return &d;
        return nullptr;
    }
};

Design parse_design(const std::string &text) {
    SReader rd(text);
    SNode top;
    std::string err;
    if (!rd.read(top, err)) throw CompileError{err};
    if (top.kind != SNode::LIST || top.items.empty() || top.items[0].text != "design") throw CompileError{"not a design: expected (design ...)"};
    Design d;
    size_t i = 1;
    for (; i < top.items.size(); ++i) {
        const SNode &n = top.items[i];
        if (n.kind != SNode::LIST || n.items.empty() || n.items[0].text != "node") break;  // the design's own property list follows
        const SNode *name = prop(n.items, 1, ":name"), *type = prop(n.items, 1, ":type"), *slots = prop(n.items, 1, ":input-slots");
        if (!name || !type || name->kind != SNode::STRING || type->kind != SNode::STRING) throw CompileError{"design node without :name / :type"};
        DesignNode dn;
        dn.name = name->text;
        dn.type = type->text;
        if (slots && slots->kind == SNode::LIST)
            for (const SNode &s : slots->items) {
                if (s.kind != SNode::LIST || s.items.size() != 3) throw CompileError{"malformed :input-slots entry in node " + dn.name};
                if (s.items[2].text != "out") throw CompileError{"node " + s.items[1].text + " has no output slot " + s.items[2].text};
                dn.inputs.emplace_back(s.items[0].text, s.items[1].text);
            }
        if (d.node(dn.name)) throw CompileError{"duplicate node name " + dn.name + " in design"};
        d.nodes.push_back(std::move(dn));
    }
    if (const SNode *root = prop(top.items, i, ":root")) d.root = root->text;
    if (const SNode *name = prop(top.items, i, ":name")) d.name = name->text;
    for (auto &n : d.nodes)
        for (auto &in : n.inputs)
            if (!d.node(in.second)) throw CompileError{"node " + n.name + " is connected to unknown node " + in.second};
    return d;
}

std::string read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw CompileError{"cannot read " + path};
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ---------------------------------------------------------------- the filter database
struct DbEntry {
    std::string path;
    bool is_design = false;
    bool loaded = false;
    std::string filter_name;            // main filter of the file / name of the design
    std::vector<UservalInfo> args;      // its arguments in declaration order
};

struct FilterDb {
    std::vector<DbEntry> entries;
    std::map<std::string, int> by_name;  // main filter name -> entry (first file found wins, in sorted path order)
    std::set<int> loading;               // designs being expanded (cycle guard)

    std::string design_source(const Design &d, const std::string &filter_name, std::set<int> &included);

    // parses a file just far enough to know its main filter and arguments
    void load(int idx) {
        DbEntry &e = entries[idx];
        if (e.loaded) return;
        std::string source;
        if (e.is_design) {
            if (!loading.insert(idx).second) throw CompileError{"design " + e.path + " includes itself"};
            std::set<int> included;
            source = design_source(parse_design(read_file(e.path)), "", included);
            loading.erase(idx);
        } else
            source = read_file(e.path);
        Module m;
        parse_module(m, source);
        e.filter_name = m.main_filter->name;
        e.args = m.main_filter->uservals;
        e.loaded = true;
    }

    void scan(const std::string &root) {
        namespace fs = std::filesystem;
        std::vector<std::string> files;
        std::error_code ec;
        for (fs::recursive_directory_iterator it(root, fs::directory_options::skip_permission_denied, ec), end; !ec && it != end; it.increment(ec)) {
            if (!it->is_regular_file(ec)) continue;
            const std::string ext = it->path().extension().string();
            if (ext == ".mm" || ext == ".mmc") files.push_back(it->path().string());
        }
        if (ec) throw CompileError{"cannot scan " + root + ": " + ec.message()};
        std::sort(files.begin(), files.end());
        for (auto &f : files) {
            DbEntry e;
            e.path = f;
            e.is_design = f.size() > 4 && f.compare(f.size() - 4, 4, ".mmc") == 0;
            entries.push_back(std::move(e));
        }
    }

    // the entry whose main filter is called `type`; files that do not parse are skipped like the reference skips them
    int lookup(const std::string &type) {
        for (size_t i = 0; i < entries.size(); ++i) {
            auto it = by_name.find(type);
            if (it != by_name.end()) return it->second;
            if (entries[i].loaded || loading.count((int)i)) continue;
            try {
                load((int)i);  // loading a design looks up its own node types: entries beyond i may get loaded here
            } catch (CompileError &) {
                entries[i].loaded = true;
                entries[i].filter_name.clear();
                continue;
            }
            if (!entries[i].filter_name.empty() && !by_name.count(entries[i].filter_name)) by_name[entries[i].filter_name] = (int)i;
        }
        auto it = by_name.find(type);
        return it != by_name.end() ? it->second : -1;
    }
};

const char *userval_type_name(int t) {  // userval.c userval_type_name
    switch (t) {
    case UV_INT: return "int";
    case UV_FLOAT: return "float";
    case UV_BOOL: return "bool";
    case UV_COLOR: return "color";
    case UV_CURVE: return "curve";
    case UV_GRADIENT: return "gradient";
    default: return "image";
    }
}

// designer_filter.c:128-298
std::string FilterDb::design_source(const Design &d, const std::string &filter_name_in, std::set<int> &included) {
    if (d.root.empty() || !d.node(d.root)) throw CompileError{"design " + d.name + " has no root node"};
    const std::string filter_name = filter_name_in.empty() ? d.name : filter_name_in;
    auto input_of = [](const DesignNode &n, const std::string &slot) -> const std::string * {
        for (auto &in : n.inputs) if (in.first == slot) return &in.second;
        return nullptr;
    };
    auto entry_of = [&](const DesignNode &n) -> DbEntry & {
        int idx = lookup(n.type);
        if (idx < 0) throw CompileError{"design " + d.name + ": no filter named " + n.type + " under the filter path"};
        return entries[idx];
    };
    // nodes reachable from the root, breadth first, and the node types in order of discovery (newest first, like
    // the reference's g_slist_prepend)
    std::vector<const DesignNode *> nodes{d.node(d.root)};
    std::vector<std::string> types{nodes[0]->type};
    for (size_t i = 0; i < nodes.size(); ++i) {
        DbEntry &e = entry_of(*nodes[i]);
        for (auto &a : e.args)
            if (const std::string *src = input_of(*nodes[i], a.name)) {
                const DesignNode *partner = d.node(*src);
                if (std::find(nodes.begin(), nodes.end(), partner) == nodes.end()) {
                    nodes.push_back(partner);
                    if (std::find(types.begin(), types.end(), partner->type) == types.end()) types.insert(types.begin(), partner->type);
                }
            }
    }
    std::ostringstream out;
    for (auto &t : types) {
        int idx = lookup(t);
        if (!included.insert(idx).second) continue;
        DbEntry &e = entries[idx];
        if (e.is_design) {
            if (!loading.insert(idx).second) throw CompileError{"design " + e.path + " includes itself"};
            out << design_source(parse_design(read_file(e.path)), "", included) << "\n\n";
            loading.erase(idx);
        } else
            out << read_file(e.path) << "\n\n";
    }
    out << "filter " << filter_name << " (";
    bool first = true;
    for (const DesignNode *n : nodes)
        for (auto &a : entry_of(*n).args) {
            if (input_of(*n, a.name)) continue;
            if (!first) out << ", ";
            first = false;
            out << userval_type_name(a.type) << " " << n->name << "_" << a.name;
            char buf[160];
            switch (a.type) {  // append_limits_and_defaults, designer_filter.c:78-111
            case UV_INT: snprintf(buf, sizeof buf, " : %d - %d (%d)", a.int_min, a.int_max, a.int_default); out << buf; break;
            case UV_FLOAT: snprintf(buf, sizeof buf, " : %f - %f (%f)", a.float_min, a.float_max, a.float_default); out << buf; break;
            case UV_BOOL: out << " (" << (a.bool_default ? '1' : '0') << ")"; break;
            default: break;
            }
        }
    out << ")\n";
    std::set<const DesignNode *> computed, visiting;
    std::function<void(const DesignNode *)> compute = [&](const DesignNode *n) {
        if (computed.count(n)) return;
        if (!visiting.insert(n).second) throw CompileError{"design " + d.name + " contains a cycle through node " + n->name};
        for (auto &in : n->inputs) compute(d.node(in.second));
        out << "    " << n->name << "_out = " << n->type << "(";
        bool f = true;
        for (auto &a : entry_of(*n).args) {
            if (!f) out << ", ";
            f = false;
            if (const std::string *src = input_of(*n, a.name)) out << *src << "_out";
            else out << n->name << "_" << a.name;
        }
        out << ");\n";
        visiting.erase(n);
        computed.insert(n);
    };
    compute(nodes[0]);
    out << "    " << nodes[0]->name << "_out(xy)\nend\n";
    return out.str();
}

}  // namespace

std::string design_to_source(const std::string &design_text, const std::string &filter_search_path) {
    FilterDb db;
    db.scan(filter_search_path);
    std::set<int> included;
    return db.design_source(parse_design(design_text), "", included);
}

}  // namespace mm
