// ir_text.cpp — reader and printer of the `.kazan-ir` text format.
//
// Follows (kazan-3d/kazan): shader-compiler-ir/src/text.rs:1133-1326 (lexer), :1417-1500 (names), :1544-1613 (printer
// name suffixes, definitions before uses, 4-space indent); module.rs:47-91; interface.rs:199-252,490-544;
// function.rs:166-193,357-434; block.rs:46-63,251-300,455-476,623-684,750-767; instructions_impl.rs:369-386,514-526;
// values.rs:141-200; types.rs:1049-1204; consts.rs:372-447,533-567; target_properties.rs.
// Checked against every text vector of the reference's unit tests (tests/golden/ir_goldens.json).
#include <cctype>
#include <cstring>
#include <set>
#include <sstream>

#include "ir.h"

namespace kir {

const ExtOp kExtOps[] = {
    {"i_sub", ExtKind::IntBinary, "sub"},
    {"i_mul", ExtKind::IntBinary, "mul.lo"},
    {"bitwise_and", ExtKind::IntBinary, "and"},
    {"bitwise_or", ExtKind::IntBinary, "or"},
    {"bitwise_xor", ExtKind::IntBinary, "xor"},
    {"shift_left_logical", ExtKind::IntBinary, "shl"},
    {"shift_right_logical", ExtKind::IntBinary, "shr.u"},
    {"shift_right_arithmetic", ExtKind::IntBinary, "shr.s"},
    {"s_negate", ExtKind::IntUnary, "neg"},
    {"not", ExtKind::IntUnary, "not"},
    {"f_add", ExtKind::FloatBinary, "add.rn"},
    {"f_sub", ExtKind::FloatBinary, "sub.rn"},
    {"f_mul", ExtKind::FloatBinary, "mul.rn"},
    {"f_div", ExtKind::FloatBinary, "div.rn"},
    {"f_min", ExtKind::FloatBinary, "min"},
    {"f_max", ExtKind::FloatBinary, "max"},
    {"f_negate", ExtKind::FloatUnary, "neg"},
    {"f_abs", ExtKind::FloatUnary, "abs"},
    {"fma", ExtKind::FloatTernary, "fma.rn"},
    {"i_equal", ExtKind::IntCompare, "eq"},
    {"i_not_equal", ExtKind::IntCompare, "ne"},
    {"u_less_than", ExtKind::IntCompare, "lt.u"},
    {"s_less_than", ExtKind::IntCompare, "lt.s"},
    {"u_less_than_equal", ExtKind::IntCompare, "le.u"},
    {"s_less_than_equal", ExtKind::IntCompare, "le.s"},
    {"u_greater_than", ExtKind::IntCompare, "gt.u"},
    {"s_greater_than", ExtKind::IntCompare, "gt.s"},
    {"u_greater_than_equal", ExtKind::IntCompare, "ge.u"},
    {"s_greater_than_equal", ExtKind::IntCompare, "ge.s"},
    {"f_ord_equal", ExtKind::FloatCompare, "eq"},
    {"f_ord_less_than", ExtKind::FloatCompare, "lt"},
    {"f_ord_less_than_equal", ExtKind::FloatCompare, "le"},
    {"f_ord_greater_than", ExtKind::FloatCompare, "gt"},
    {"f_ord_greater_than_equal", ExtKind::FloatCompare, "ge"},
    {"f_unord_not_equal", ExtKind::FloatCompare, "neu"},
    {"logical_and", ExtKind::BoolBinary, "and"},
    {"logical_or", ExtKind::BoolBinary, "or"},
    {"logical_not", ExtKind::BoolUnary, "not"},
    {"select", ExtKind::Select, "selp"},
    {"convert_f_to_u", ExtKind::Convert, "f2u"},
    {"convert_f_to_s", ExtKind::Convert, "f2s"},
    {"convert_s_to_f", ExtKind::Convert, "s2f"},
    {"convert_u_to_f", ExtKind::Convert, "u2f"},
    {"u_convert", ExtKind::Convert, "u2u"},
    {"s_convert", ExtKind::Convert, "s2s"},
    {"f_convert", ExtKind::Convert, "f2f"},
    {"bitcast", ExtKind::Convert, "bits"},
    {"store", ExtKind::Store, "st"},
    {"ptr_add", ExtKind::PtrAdd, "add"},
    {"group_non_uniform_elect", ExtKind::SubgroupElect, "elect"},
    {"group_non_uniform_all", ExtKind::SubgroupVote, "all"},
    {"group_non_uniform_any", ExtKind::SubgroupVote, "any"},
    {"group_non_uniform_ballot", ExtKind::SubgroupBallot, "ballot"},
    {"group_non_uniform_broadcast_first", ExtKind::SubgroupBroadcast, "first"},
    {"group_non_uniform_shuffle", ExtKind::SubgroupShuffle, "idx"},
    {"group_non_uniform_shuffle_xor", ExtKind::SubgroupShuffle, "bfly"},
    {"group_non_uniform_i_add", ExtKind::SubgroupReduce, "add"},
};
const int kExtOpCount = (int)(sizeof(kExtOps) / sizeof(kExtOps[0]));
// Built-in kinds of version 2 (the reference has `vertex_position` only, interface.rs:61-67), spelled like the SPIR-V
// BuiltIn they carry: inputs global_invocation_id <3 x i32>, subgroup_local_invocation_id i32, subgroup_size i32,
// frag_coord <4 x f32>; push_constants / descriptor_table: a data_ptr to the bound block (vkCmdPushConstants,
// vkCmdBindDescriptorSets).
bool is_v2_built_in(const std::string& kind) {
    static const char* const kinds[] = {"global_invocation_id", "subgroup_local_invocation_id", "subgroup_size", "frag_coord",
                                        "push_constants", "descriptor_table"};
    for (auto* k : kinds)
        if (kind == k) return true;
    return false;
}
int ext_op_index(const std::string& name) {
    for (int i = 0; i < kExtOpCount; ++i)
        if (name == kExtOps[i].name) return i;
    return -1;
}

int int_bits(IntKind k) {
    switch (k) {
    case IntKind::I8: return 8;
    case IntKind::I16: return 16;
    case IntKind::I64: return 64;
    default: return 32;
    }
}
int float_bits(FloatKind k) {
    switch (k) {
    case FloatKind::F16: return 16;
    case FloatKind::F64: return 64;
    default: return 32;
    }
}
const char* int_name(IntKind k) {
    switch (k) {
    case IntKind::I8: return "i8";
    case IntKind::I16: return "i16";
    case IntKind::I32: return "i32";
    case IntKind::RI32: return "ri32";
    default: return "i64";
    }
}
const char* float_name(FloatKind k) {
    switch (k) {
    case FloatKind::F16: return "f16";
    case FloatKind::F32: return "f32";
    case FloatKind::RF32: return "rf32";
    default: return "f64";
    }
}

static bool int_kind_from(const std::string& s, IntKind* k) {
    static const std::pair<const char*, IntKind> t[] = {
        {"i8", IntKind::I8}, {"i16", IntKind::I16}, {"i32", IntKind::I32}, {"ri32", IntKind::RI32}, {"i64", IntKind::I64}};
    for (auto& e : t)
        if (s == e.first) {
            *k = e.second;
            return true;
        }
    return false;
}
static bool float_kind_from(const std::string& s, FloatKind* k) {
    static const std::pair<const char*, FloatKind> t[] = {
        {"f16", FloatKind::F16}, {"f32", FloatKind::F32}, {"rf32", FloatKind::RF32}, {"f64", FloatKind::F64}};
    for (auto& e : t)
        if (s == e.first) {
            *k = e.second;
            return true;
        }
    return false;
}

static const std::set<std::string>& keywords() {  // text.rs:493-529
    static const std::set<std::string> k = {
        "align", "bool", "const", "data_ptr", "dont_inline", "entry_point", "f16", "f32", "f64", "false", "fixed",
        "fn", "i16", "i32", "i64", "i8", "inline", "kind", "module", "none", "normal", "null", "pure", "rf32", "ri32",
        "size", "struct", "true", "undef", "variable", "vscale", "x"};
    return k;
}

// ---- module arenas ----------------------------------------------------------------------------------------------

static std::string type_key(const Type& t) {
    std::ostringstream o;
    switch (t.kind) {
    case TypeKind::Int: o << int_name(t.int_kind); break;
    case TypeKind::Float: o << float_name(t.float_kind); break;
    case TypeKind::Bool: o << "bool"; break;
    case TypeKind::DataPtr: o << "data_ptr"; break;
    case TypeKind::Vector: o << "<" << (t.scalable ? "v" : "") << t.length << "x" << t.element->key << ">"; break;
    case TypeKind::Fn:
        o << "fn[";
        for (auto* a : t.args) o << a->key << ",";
        o << "]->";
        if (t.diverges) o << "!";
        else {
            o << "[";
            for (auto* r : t.results) o << r->key << ",";
            o << "]";
        }
        break;
    case TypeKind::Struct:
        o << "struct{" << (t.size_variable ? "v" : "f") << t.size << "," << t.align << ";";
        for (auto& m : t.members) o << m.offset << ":" << m.type->key << ",";
        o << "}";
        break;
    }
    return o.str();
}

const Type* Module::intern(Type t) {
    t.key = type_key(t);
    auto it = type_table.find(t.key);
    if (it != type_table.end()) return it->second;
    types.push_back(std::make_unique<Type>(std::move(t)));
    const Type* p = types.back().get();
    type_table[p->key] = p;
    return p;
}
const Type* Module::int_type(IntKind k) {
    Type t{TypeKind::Int};
    t.int_kind = k;
    return intern(t);
}
const Type* Module::float_type(FloatKind k) {
    Type t{TypeKind::Float};
    t.float_kind = k;
    return intern(t);
}
const Type* Module::bool_type() { return intern(Type{TypeKind::Bool}); }
const Type* Module::data_ptr() { return intern(Type{TypeKind::DataPtr}); }
Value* Module::new_value(const std::string& name, const Type* type) {
    values.push_back(std::make_unique<Value>());
    Value* v = values.back().get();
    v->name = name;
    v->type = type;
    v->id = (int)values.size() - 1;
    return v;
}
Block* Module::new_block(const std::string& name) {
    blocks.push_back(std::make_unique<Block>());
    blocks.back()->name = name;
    blocks.back()->id = (int)blocks.size() - 1;
    return blocks.back().get();
}
Loop* Module::new_loop(const std::string& name) {
    loops.push_back(std::make_unique<Loop>());
    loops.back()->name = name;
    loops.back()->id = (int)loops.size() - 1;
    return loops.back().get();
}
Instr* Module::new_instr(Op op) {
    instrs.push_back(std::make_unique<Instr>());
    instrs.back()->op = op;
    return instrs.back().get();
}
Function* Module::new_function(const std::string& name) {
    function_arena.push_back(std::make_unique<Function>());
    function_arena.back()->name = name;
    return function_arena.back().get();
}

uint64_t type_size(const Type* t) {
    switch (t->kind) {
    case TypeKind::Int: return int_bits(t->int_kind) / 8;
    case TypeKind::Float: return float_bits(t->float_kind) / 8;
    case TypeKind::Bool: return 1;
    case TypeKind::DataPtr:
    case TypeKind::Fn: return 8;
    case TypeKind::Vector:
        if (t->scalable) throw Error("vscale vectors have no size on a SIMT target");
        return t->length * type_size(t->element);
    case TypeKind::Struct:
        if (t->size_variable) throw Error("variable-size struct has no fixed size");
        return t->size;
    }
    return 0;
}

// ---- lexer -----------------------------------------------------------------------------------------------------------

namespace {

enum class Tk { Id, Int, Str, Punct, Eof };
struct Token {
    Tk kind;
    std::string text;  // identifier / punctuation / decoded string
    uint64_t value = 0;
    std::string suffix;
    size_t pos = 0;
};

[[noreturn]] void fail(const std::string& msg, size_t pos) { throw Error(msg + " at offset " + std::to_string(pos)); }

std::vector<Token> lex(const std::string& s) {
    std::vector<Token> out;
    size_t i = 0, n = s.size();
    auto is_id_start = [](char c) { return std::isalpha((unsigned char)c) || c == '_'; };
    auto is_id = [](char c) { return std::isalnum((unsigned char)c) || c == '_'; };
    while (i < n) {
        char c = s[i];
        if (c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v') {
            ++i;
            continue;
        }
        if (c == '#') {  // comment to end of line (text.rs:876,1133-1144)
            while (i < n && s[i] != '\n') ++i;
            continue;
        }
        Token t;
        t.pos = i;
        if (is_id_start(c)) {
            size_t j = i;
            while (j < n && is_id(s[j])) ++j;
            t.kind = Tk::Id;
            t.text = s.substr(i, j - i);
            i = j;
        } else if (std::isdigit((unsigned char)c)) {
            int base = 10;
            size_t j = i;
            if (c == '0' && j + 1 < n && (s[j + 1] == 'x' || s[j + 1] == 'X')) base = 16, j += 2;
            else if (c == '0' && j + 1 < n && (s[j + 1] == 'o' || s[j + 1] == 'O')) base = 8, j += 2;
            else if (c == '0' && j + 1 < n && (s[j + 1] == 'b' || s[j + 1] == 'B')) base = 2, j += 2;
            else if (c == '0' && j + 1 < n && std::isdigit((unsigned char)s[j + 1])) fail("octal numbers must start with 0o", i);
            size_t d0 = j;
            uint64_t v = 0;
            auto digit = [&](char ch) -> int {
                if (ch >= '0' && ch <= '9') return ch - '0';
                if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
                if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
                return 99;
            };
            while (j < n && digit(s[j]) < base) {
                uint64_t nv = v * base + (uint64_t)digit(s[j]);
                if (nv / base != v && v != 0 && (nv - digit(s[j])) / base != v) fail("integer too large", i);
                if (v > (~0ull - (uint64_t)digit(s[j])) / (uint64_t)base) fail("integer too large", i);
                v = nv;
                ++j;
            }
            if (j == d0) fail("number has no digits", i);
            t.kind = Tk::Int;
            t.value = v;
            // optional type suffix: longest of ri32 i8 i16 i32 i64 (text.rs:531-540)
            for (const char* suf : {"ri32", "i16", "i32", "i64", "i8"}) {
                size_t L = std::strlen(suf);
                if (s.compare(j, L, suf) == 0 && !(j + L < n && is_id(s[j + L]))) {
                    t.suffix = suf;
                    j += L;
                    break;
                }
            }
            if (j < n && is_id(s[j])) fail("invalid number suffix", j);
            i = j;
        } else if (c == '"') {
            size_t j = i + 1;
            std::string v;
            for (;;) {
                if (j >= n) fail("unterminated string", i);
                char ch = s[j];
                if (ch == '"') break;
                if (ch == '\\') {
                    if (j + 1 >= n) fail("unterminated string", i);
                    char e = s[j + 1];
                    j += 2;
                    switch (e) {
                    case '0': v.push_back('\0'); break;
                    case 'n': v.push_back('\n'); break;
                    case 'r': v.push_back('\r'); break;
                    case 't': v.push_back('\t'); break;
                    case '\'': v.push_back('\''); break;
                    case '"': v.push_back('"'); break;
                    case '\\': v.push_back('\\'); break;
                    case 'u': {
                        if (j >= n || s[j] != '{') fail("bad \\u escape", j);
                        size_t k = s.find('}', j);
                        if (k == std::string::npos) fail("bad \\u escape", j);
                        uint32_t cp = (uint32_t)std::stoul(s.substr(j + 1, k - j - 1), nullptr, 16);
                        // UTF-8 encode
                        if (cp < 0x80) v.push_back((char)cp);
                        else if (cp < 0x800) { v.push_back((char)(0xC0 | (cp >> 6))); v.push_back((char)(0x80 | (cp & 63))); }
                        else if (cp < 0x10000) { v.push_back((char)(0xE0 | (cp >> 12))); v.push_back((char)(0x80 | ((cp >> 6) & 63))); v.push_back((char)(0x80 | (cp & 63))); }
                        else { v.push_back((char)(0xF0 | (cp >> 18))); v.push_back((char)(0x80 | ((cp >> 12) & 63))); v.push_back((char)(0x80 | ((cp >> 6) & 63))); v.push_back((char)(0x80 | (cp & 63))); }
                        j = k + 1;
                        break;
                    }
                    default: fail("unknown escape", j);
                    }
                    continue;
                }
                v.push_back(ch);
                ++j;
            }
            t.kind = Tk::Str;
            t.text = v;
            i = j + 1;
        } else {
            t.kind = Tk::Punct;
            if (c == '-' && i + 1 < n && s[i + 1] == '>') {
                t.text = "->";
                i += 2;
            } else if (std::strchr("{}[]()<>,;:!@*=", c)) {
                t.text = std::string(1, c);
                ++i;
            } else {
                fail(std::string("unexpected character '") + c + "'", i);
            }
        }
        out.push_back(std::move(t));
    }
    Token e;
    e.kind = Tk::Eof;
    e.pos = n;
    out.push_back(e);
    return out;
}

using Key = std::pair<std::string, uint64_t>;  // (name, suffix)

// ---- parser ------------------------------------------------------------------------------------------------------------

struct Parser {
    Module& m;
    std::vector<Token> toks;
    size_t i = 0;
    std::map<uint64_t, const Type*> structs;
    std::vector<std::map<Key, Value*>> values{1};
    std::vector<std::map<Key, Block*>> blocks{1};
    std::vector<std::map<Key, Loop*>> loops{1};
    std::map<Key, Function*> functions;
    int version = 1;  // 2: version 2 instructions and built-in kinds are accepted (ir.h)

    Parser(Module& mod, const std::string& text, int ver = 1) : m(mod), toks(lex(text)), version(ver) {}

    const Token& t() const { return toks[i]; }
    Token next() { return toks[i++]; }
    bool at(const char* s) const { return (t().kind == Tk::Punct || t().kind == Tk::Id) && t().text == s; }
    bool accept(const char* s) {
        if (at(s)) {
            ++i;
            return true;
        }
        return false;
    }
    void expect(const char* s) {
        if (!accept(s)) fail(std::string("expected '") + s + "', found '" + t().text + "'", t().pos);
    }
    void finish() {
        if (t().kind != Tk::Eof) fail("unexpected token '" + t().text + "'", t().pos);
    }
    uint64_t integer() {
        Token k = next();
        if (k.kind != Tk::Int || !k.suffix.empty()) fail("expected an unsuffixed integer", k.pos);
        return k.value;
    }
    std::string word() {
        Token k = next();
        if (k.kind != Tk::Id) fail("expected a word", k.pos);
        return k.text;
    }
    Key name() {
        Token k = next();
        if (k.kind == Tk::Id && !keywords().count(k.text)) return {k.text, 0};
        if (k.kind == Tk::Str) return {k.text, integer()};
        fail("expected a name, found '" + k.text + "'", k.pos);
    }
    void push() {
        values.emplace_back();
        blocks.emplace_back();
        loops.emplace_back();
    }
    void pop() {
        values.pop_back();
        blocks.pop_back();
        loops.pop_back();
    }
    template <typename T>
    T* lookup(std::vector<std::map<Key, T*>>& stack, const Key& k, const char* what) {
        for (size_t s = stack.size(); s-- > 0;) {
            auto it = stack[s].find(k);
            if (it != stack[s].end()) return it->second;
        }
        throw Error(std::string(what) + " '" + k.first + "' is not defined");
    }
    void define_value(const Key& k, Value* v, bool root = false) {
        auto& scope = root ? values.front() : values.back();
        if (!scope.emplace(k, v).second) throw Error("value '" + k.first + "' defined twice");
    }

    uint64_t alignment() {
        expect("align");
        expect(":");
        uint64_t a = integer();
        if (a == 0 || (a & (a - 1))) throw Error("alignment must be a power of two");
        return a;
    }

    const Type* type() {
        const Token& k = t();
        IntKind ik;
        FloatKind fk;
        if (k.kind == Tk::Id && int_kind_from(k.text, &ik)) {
            ++i;
            return m.int_type(ik);
        }
        if (k.kind == Tk::Id && float_kind_from(k.text, &fk)) {
            ++i;
            return m.float_type(fk);
        }
        if (accept("bool")) return m.bool_type();
        if (accept("data_ptr")) return m.data_ptr();
        if (accept("(")) {
            const Type* ty = type();
            expect(")");
            return ty;
        }
        if (accept("<")) {
            Type v{TypeKind::Vector};
            if (accept("vscale")) {
                expect("x");
                v.scalable = true;
            }
            v.length = integer();
            expect("x");
            v.element = type();
            expect(">");
            return m.intern(v);
        }
        if (accept("fn")) return fn_type_tail();
        if (accept("struct")) {
            uint64_t sid = integer();
            expect("{");
            if (accept("}")) {
                auto it = structs.find(sid);
                if (it == structs.end()) throw Error("struct " + std::to_string(sid) + " referenced before its definition");
                return it->second;
            }
            Type s{TypeKind::Struct};
            expect("size");
            expect(":");
            std::string kind = word();
            if (kind != "fixed" && kind != "variable") throw Error("struct size must be fixed or variable");
            s.size_variable = kind == "variable";
            s.size = integer();
            expect(",");
            s.align = alignment();
            expect(",");
            while (!at("}")) {
                StructMember mem;
                mem.offset = integer();
                expect(":");
                mem.type = type();
                expect(",");
                s.members.push_back(mem);
            }
            expect("}");
            const Type* ty = m.intern(s);
            auto it = structs.find(sid);
            if (it != structs.end() && it->second != ty) throw Error("struct " + std::to_string(sid) + " redefined");
            structs[sid] = ty;
            return ty;
        }
        fail("expected a type, found '" + k.text + "'", k.pos);
    }

    const Type* fn_type_tail() {
        Type f{TypeKind::Fn};
        expect("[");
        while (!at("]")) {
            f.args.push_back(type());
            if (!accept(",")) break;
        }
        expect("]");
        expect("->");
        if (accept("!")) {
            f.diverges = true;
        } else {
            expect("[");
            while (!at("]")) {
                f.results.push_back(type());
                if (!accept(",")) break;
            }
            expect("]");
        }
        return m.intern(f);
    }

    Const constant() {
        const Token& k = t();
        Const c{};
        FloatKind fk;
        if (k.kind == Tk::Int) {
            Token v = next();
            if (v.suffix.empty()) fail("integer constant needs a type suffix", v.pos);
            IntKind ik = IntKind::I32;
            int_kind_from(v.suffix, &ik);
            int bits = int_bits(ik);
            if (bits < 64 && (v.value >> bits)) fail("constant does not fit its type", v.pos);
            c.kind = ConstKind::Int;
            c.type = m.int_type(ik);
            c.bits = v.value;
            return c;
        }
        if (k.kind == Tk::Id && float_kind_from(k.text, &fk)) {
            ++i;
            c.kind = ConstKind::Float;
            c.type = m.float_type(fk);
            c.bits = integer();
            if (float_bits(fk) < 64 && (c.bits >> float_bits(fk))) throw Error("float bit pattern too wide");
            return c;
        }
        if (at("true") || at("false")) {
            c.kind = ConstKind::Bool;
            c.type = m.bool_type();
            c.bits = next().text == "true" ? 1 : 0;
            return c;
        }
        if (accept("<")) {
            c.kind = ConstKind::Vector;
            c.elements.push_back(constant());
            while (accept(",")) c.elements.push_back(constant());
            expect(">");
            for (auto& e : c.elements)
                if (e.type != c.elements[0].type) throw Error("vector constant elements differ in type");
            Type v{TypeKind::Vector};
            v.length = c.elements.size();
            v.element = c.elements[0].type;
            c.type = m.intern(v);
            return c;
        }
        if (accept("undef")) {
            c.kind = ConstKind::Undef;
            c.type = type();
            return c;
        }
        if (accept("null")) {
            c.kind = ConstKind::Null;
            c.type = type();
            if (c.type->kind != TypeKind::DataPtr && c.type->kind != TypeKind::Fn) throw Error("null needs a pointer type");
            return c;
        }
        if (accept("fn")) {
            Key key = name();
            auto it = functions.find(key);
            if (it == functions.end()) throw Error("function '" + key.first + "' is not defined");
            c.kind = ConstKind::Fn;
            c.function = it->second;
            c.type = it->second->type;
            return c;
        }
        fail("expected a constant, found '" + k.text + "'", k.pos);
    }

    Value* value_def() {
        Key k = name();
        expect(":");
        Value* v = m.new_value(k.first, type());
        define_value(k, v);
        return v;
    }
    Value* value_use() {
        Key k = name();
        if (accept(":")) {
            Const c = constant();
            Value* v = m.new_value(k.first, c.type);
            v->has_const = true;
            v->constant = c;
            define_value(k, v, true);
            return v;
        }
        return lookup(values, k, "value");
    }
    std::vector<Value*> uses() {  // after '[' ... consumes ']'
        std::vector<Value*> out;
        while (!at("]")) {
            out.push_back(value_use());
            if (!accept(",")) break;
        }
        expect("]");
        return out;
    }
    std::vector<Value*> defs() {
        expect("[");
        std::vector<Value*> out;
        while (!at("]")) {
            out.push_back(value_def());
            if (!accept(",")) break;
        }
        expect("]");
        return out;
    }
    Variable variable() {
        Variable v;
        v.type = type();
        expect(",");
        v.align = alignment();
        expect("->");
        v.pointer = value_def();
        if (v.pointer->type->kind != TypeKind::DataPtr) throw Error("a variable's pointer must have type data_ptr");
        return v;
    }
    Location location() {
        Location l;
        if (accept("@")) {
            Token f = next();
            if (f.kind != Tk::Str) fail("expected a file name string", f.pos);
            l.present = true;
            l.file = f.text;
            expect(":");
            l.line = integer();
            expect(":");
            l.column = integer();
        }
        return l;
    }

    static bool diverges(const Instr* ins) {
        if (ins->op == Op::Break || ins->op == Op::Continue) return true;
        if (ins->op == Op::BlockI) return ins->block->diverges;
        if (ins->op == Op::LoopI) return ins->loop->diverges;
        return false;
    }

    Instr* break_tail() {
        Key k = name();
        Block* b = lookup(blocks, k, "block");
        expect("[");
        Instr* ins = m.new_instr(Op::Break);
        ins->block = b;
        ins->args = uses();
        if (!b->diverges && ins->args.size() != b->results.size()) throw Error("break " + b->name + ": wrong number of results");
        return ins;
    }

    // argument / result counts and types of a version 2 instruction
    static void check_ext(const Instr* ins) {
        const ExtOp& op = kExtOps[ins->ext];
        auto bad = [&](const char* what) { throw Error(std::string(op.name) + ": " + what); };
        auto counts = [&](size_t a, size_t r) {
            if (ins->args.size() != a || ins->results.size() != r)
                bad(("takes " + std::to_string(a) + " argument(s) and " + std::to_string(r) + " result(s)").c_str());
        };
        auto is = [](const Value* v, TypeKind k) { return v->type->kind == k; };
        auto int32 = [](const Value* v) { return v->type->kind == TypeKind::Int && int_bits(v->type->int_kind) == 32; };
        auto scalar = [](const Value* v) {
            TypeKind k = v->type->kind;
            return k == TypeKind::Int || k == TypeKind::Float || k == TypeKind::Bool || k == TypeKind::DataPtr;
        };
        switch (op.kind) {
        case ExtKind::IntBinary:
            counts(2, 1);
            if (!is(ins->results[0], TypeKind::Int) || ins->args[0]->type != ins->results[0]->type || ins->args[1]->type != ins->results[0]->type)
                bad("operands and result must have one integer type");
            break;
        case ExtKind::IntUnary:
            counts(1, 1);
            if (!is(ins->results[0], TypeKind::Int) || ins->args[0]->type != ins->results[0]->type) bad("operand and result must have one integer type");
            break;
        case ExtKind::FloatBinary:
        case ExtKind::FloatUnary:
        case ExtKind::FloatTernary: {
            const size_t n = op.kind == ExtKind::FloatBinary ? 2 : op.kind == ExtKind::FloatUnary ? 1 : 3;
            counts(n, 1);
            if (!is(ins->results[0], TypeKind::Float) || float_bits(ins->results[0]->type->float_kind) == 16) bad("the result must be f32 or f64");
            for (auto* a : ins->args)
                if (a->type != ins->results[0]->type) bad("operands and result must have one float type");
            break;
        }
        case ExtKind::IntCompare:
        case ExtKind::FloatCompare:
            counts(2, 1);
            if (!is(ins->results[0], TypeKind::Bool)) bad("the result must be bool");
            if (ins->args[0]->type != ins->args[1]->type) bad("operands must have one type");
            if (op.kind == ExtKind::IntCompare ? !(is(ins->args[0], TypeKind::Int) || is(ins->args[0], TypeKind::DataPtr))
                                               : !(is(ins->args[0], TypeKind::Float) && float_bits(ins->args[0]->type->float_kind) != 16))
                bad(op.kind == ExtKind::IntCompare ? "operands must be integers" : "operands must be f32 or f64");
            break;
        case ExtKind::BoolBinary:
        case ExtKind::BoolUnary:
            counts(op.kind == ExtKind::BoolBinary ? 2 : 1, 1);
            if (!is(ins->results[0], TypeKind::Bool)) bad("the result must be bool");
            for (auto* a : ins->args)
                if (!is(a, TypeKind::Bool)) bad("operands must be bool");
            break;
        case ExtKind::Select:
            counts(3, 1);
            if (!is(ins->args[0], TypeKind::Bool)) bad("the condition must be bool");
            if (!scalar(ins->results[0]) || ins->args[1]->type != ins->results[0]->type || ins->args[2]->type != ins->results[0]->type)
                bad("both values and the result must have one scalar type");
            break;
        case ExtKind::Convert: {
            counts(1, 1);
            const Value *a = ins->args[0], *r = ins->results[0];
            const std::string n = op.name;
            auto flt = [&](const Value* v) { return is(v, TypeKind::Float) && float_bits(v->type->float_kind) != 16; };
            bool ok = true;
            if (n == "convert_f_to_u" || n == "convert_f_to_s") ok = flt(a) && is(r, TypeKind::Int);
            else if (n == "convert_s_to_f" || n == "convert_u_to_f") ok = is(a, TypeKind::Int) && flt(r);
            else if (n == "u_convert" || n == "s_convert") ok = is(a, TypeKind::Int) && is(r, TypeKind::Int);
            else if (n == "f_convert") ok = flt(a) && flt(r);
            else ok = scalar(a) && scalar(r) && !is(a, TypeKind::Bool) && !is(r, TypeKind::Bool) && type_size(a->type) == type_size(r->type);
            if (!ok) bad("operand / result types do not fit the conversion");
            break;
        }
        case ExtKind::Store:
            counts(2, 0);
            if (!is(ins->args[0], TypeKind::DataPtr)) bad("the first argument must be a data_ptr");
            if (!scalar(ins->args[1])) bad("only scalar values are stored");
            break;
        case ExtKind::PtrAdd:
            counts(2, 1);
            if (!is(ins->args[0], TypeKind::DataPtr) || !is(ins->results[0], TypeKind::DataPtr)) bad("pointer and result must be data_ptr");
            if (!is(ins->args[1], TypeKind::Int) || int_bits(ins->args[1]->type->int_kind) != 64) bad("the offset must be i64");
            break;
        case ExtKind::SubgroupElect:
            counts(0, 1);
            if (!is(ins->results[0], TypeKind::Bool)) bad("the result must be bool");
            break;
        case ExtKind::SubgroupVote:
            counts(1, 1);
            if (!is(ins->args[0], TypeKind::Bool) || !is(ins->results[0], TypeKind::Bool)) bad("operand and result must be bool");
            break;
        case ExtKind::SubgroupBallot:
            counts(1, 1);
            if (!is(ins->args[0], TypeKind::Bool) || !int32(ins->results[0])) bad("takes a bool, returns an i32");
            break;
        case ExtKind::SubgroupBroadcast:
            counts(1, 1);
            if (ins->args[0]->type != ins->results[0]->type || !(int32(ins->args[0]) || (is(ins->args[0], TypeKind::Float) && float_bits(ins->args[0]->type->float_kind) == 32)))
                bad("operand and result must be one 32-bit type");
            break;
        case ExtKind::SubgroupShuffle:
            counts(2, 1);
            if (ins->args[0]->type != ins->results[0]->type || !(int32(ins->args[0]) || (is(ins->args[0], TypeKind::Float) && float_bits(ins->args[0]->type->float_kind) == 32)))
                bad("value and result must be one 32-bit type");
            if (!int32(ins->args[1])) bad("the lane selector must be i32");
            break;
        case ExtKind::SubgroupReduce:
            counts(1, 1);
            if (!int32(ins->args[0]) || ins->args[0]->type != ins->results[0]->type) bad("operand and result must be i32");
            break;
        }
    }

    Instr* instruction() {
        Token k = next();
        if (k.kind != Tk::Id) fail("expected an instruction, found '" + k.text + "'", k.pos);
        Instr* ins = nullptr;
        if (k.text == "add") {
            ins = m.new_instr(Op::Add);
            expect("[");
            ins->args = uses();
            expect("->");
            ins->results = defs();
            if (ins->args.size() != 2 || ins->results.size() != 1) throw Error("add takes two arguments and one result");
        } else if (k.text == "load") {
            ins = m.new_instr(Op::Load);
            expect("[");
            ins->args = uses();
            expect(",");
            ins->align = alignment();
            expect("->");
            ins->results = defs();
            if (ins->args.size() != 1 || ins->results.size() != 1) throw Error("load takes one argument and one result");
        } else if (k.text == "break") {
            ins = break_tail();
        } else if (k.text == "continue") {
            Key key = name();
            ins = m.new_instr(Op::Continue);
            ins->loop = lookup(loops, key, "loop");
            expect("[");
            ins->args = uses();
            if (ins->args.size() != ins->loop->params.size()) throw Error("continue " + ins->loop->name + ": wrong number of arguments");
        } else if (k.text == "branch") {
            ins = m.new_instr(Op::Branch);
            expect("[");
            ins->args = uses();
            if (ins->args.size() != 1) throw Error("branch takes one argument");
            expect(",");
            expect("{");
            while (!at("}")) {
                BranchCase bc;
                bc.value = constant();
                expect("->");
                expect("break");
                Instr* b = break_tail();
                expect(";");
                bc.target = b->block;
                bc.results = b->args;
                ins->cases.push_back(std::move(bc));
            }
            expect("}");
            expect("->");
            expect("[");
            expect("]");
        } else if (k.text == "block") {
            ins = m.new_instr(Op::BlockI);
            ins->block = block_instruction();
        } else if (k.text == "loop") {
            ins = m.new_instr(Op::LoopI);
            ins->loop = loop_instruction();
        } else if (version >= 2 && ext_op_index(k.text) >= 0) {
            ins = m.new_instr(Op::Ext);
            ins->ext = ext_op_index(k.text);
            expect("[");
            ins->args = uses();
            if (kExtOps[ins->ext].kind == ExtKind::Store) {
                expect(",");
                ins->align = alignment();
            }
            expect("->");
            ins->results = defs();
            check_ext(ins);
        } else {
            fail("unknown instruction '" + k.text + "'", k.pos);
        }
        ins->location = location();
        return ins;
    }

    void body(Block* b) {
        expect("{");
        while (!at("}")) {
            if (!b->body.empty() && diverges(b->body.back())) throw Error("instructions after a diverging instruction");  // block.rs:291-300
            b->body.push_back(instruction());
            expect(";");
        }
        expect("}");
        if (b->body.empty() || !diverges(b->body.back())) throw Error("block '" + b->name + "' must end in break/continue or a diverging block");
    }

    void results(bool* div, std::vector<Value*>* out) {
        if (accept("!")) {
            *div = true;
            return;
        }
        *div = false;
        *out = defs();
    }

    Block* block_instruction() {  // after "block"
        Key k = name();
        expect("->");
        Block* b = m.new_block(k.first);
        results(&b->diverges, &b->results);  // result definitions live in the enclosing scope
        push();
        blocks.back()[k] = b;
        body(b);
        pop();
        return b;
    }
    Block* named_body(bool div, const std::vector<Value*>& res) {
        Key k = name();
        Block* b = m.new_block(k.first);
        b->diverges = div;
        b->results = res;
        push();
        blocks.back()[k] = b;
        body(b);
        pop();
        return b;
    }
    Loop* loop_instruction() {  // after "loop"
        Key k = name();
        expect("[");
        Loop* lp = m.new_loop(k.first);
        lp->arguments = uses();
        expect("->");
        results(&lp->diverges, &lp->results);
        expect("{");
        push();
        loops.back()[k] = lp;
        expect("->");
        lp->params = defs();
        expect(";");
        if (lp->params.size() != lp->arguments.size()) throw Error("loop argument count differs from its parameter count");
        lp->body = named_body(lp->diverges, lp->results);
        pop();
        expect("}");
        return lp;
    }

    Function* function() {
        expect("fn");
        Key k = name();
        push();
        Function* f = m.new_function(k.first);
        f->params = defs();
        expect("->");
        if (accept("!")) {
            f->diverges = true;
        } else {
            expect("[");
            while (!at("]")) {
                f->returns.push_back(type());
                if (!accept(",")) break;
            }
            expect("]");
        }
        Type ft{TypeKind::Fn};
        for (auto* p : f->params) ft.args.push_back(p->type);
        ft.diverges = f->diverges;
        ft.results = f->returns;
        f->type = m.intern(ft);
        functions[k] = f;  // visible inside its own body (function.rs golden)
        expect("{");
        expect("hints");
        expect("{");
        while (!at("}")) {
            std::string h = word();
            expect(":");
            std::string v = word();
            accept(",");
            if (h == "inlining_hint" && (v == "none" || v == "inline" || v == "dont_inline")) f->inlining_hint = v;
            else if (h == "side_effects" && (v == "normal" || v == "pure" || v == "const")) f->side_effects = v;
            else throw Error("bad hint " + h + ": " + v);
        }
        expect("}");
        expect("{");
        while (!at("}")) {
            f->locals.push_back(variable());
            expect(";");
        }
        expect("}");
        std::vector<Value*> res;
        for (auto* rt : f->returns) res.push_back(m.new_value("", rt));
        f->body = named_body(f->diverges, res);
        expect("}");
        pop();
        return f;
    }

    void target_properties(IntKind* dp, IntKind* fp) {
        expect("target_properties");
        expect("{");
        *dp = *fp = IntKind::I64;
        while (!at("}")) {
            std::string k = word();
            expect(":");
            std::string v = word();
            IntKind ik = IntKind::I32;
            if (!int_kind_from(v, &ik)) throw Error("pointer underlying type must be an integer type");
            if (k == "data_pointer_underlying_type") *dp = ik;
            else if (k == "function_pointer_underlying_type") *fp = ik;
            else throw Error("unknown target property " + k);
            if (!accept(",")) break;
        }
        expect("}");
    }

    InterfaceBlock interface_block() {
        InterfaceBlock b;
        expect("{");
        expect("->");
        b.pointer = value_def();
        expect(";");
        expect("size");
        expect(":");
        std::string kind = word();
        if (kind != "fixed" && kind != "variable") throw Error("interface block size must be fixed or variable");
        b.size_variable = kind == "variable";
        b.size = integer();
        expect(";");
        b.align = alignment();
        expect(";");
        while (!at("}")) {
            StructMember mem;
            mem.offset = integer();
            expect(":");
            mem.type = type();
            expect(":");
            expect("{");
            expect("}");
            expect(";");
            b.members.push_back(mem);
        }
        expect("}");
        return b;
    }

    std::vector<Variable> interface_vars() {
        std::vector<Variable> out;
        expect("{");
        while (!at("}")) {
            Variable v = variable();
            expect(":");
            expect("{");
            expect("kind");
            expect(":");
            v.built_in_kind = word();
            if (v.built_in_kind != "vertex_position" && !(version >= 2 && is_v2_built_in(v.built_in_kind)))
                throw Error("unknown built-in kind " + v.built_in_kind);  // interface.rs:61-67
            expect("}");
            expect(";");
            out.push_back(v);
        }
        expect("}");
        return out;
    }

    void module() {
        expect("module");
        expect("{");
        target_properties(&m.data_ptr_type, &m.fn_ptr_type);
        expect("built_in_inputs");
        m.built_in_inputs = interface_vars();
        expect("user_inputs_block");
        m.user_inputs = interface_block();
        expect("built_in_outputs");
        m.built_in_outputs = interface_vars();
        expect("user_outputs_block");
        m.user_outputs = interface_block();
        expect("invocation_global_variables");
        expect("{");
        while (!at("}")) {
            m.globals.push_back(variable());
            expect(";");
        }
        expect("}");
        while (at("fn")) m.functions.push_back(function());
        expect("entry_point");
        expect(":");
        Key k = name();
        auto it = functions.find(k);
        if (it == functions.end()) throw Error("entry point '" + k.first + "' is not defined");
        m.entry_point = it->second;
        expect(";");
        expect("}");
    }
};

// ---- printer -------------------------------------------------------------------------------------------------------------

std::string hex(uint64_t v) {
    char buf[32];
    snprintf(buf, sizeof buf, "0x%llX", (unsigned long long)v);
    return buf;
}

bool is_plain_ident(const std::string& s) {
    if (s.empty() || !(std::isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
    for (char c : s)
        if (!(std::isalnum((unsigned char)c) || c == '_')) return false;
    return !keywords().count(s);
}

std::string escape(const std::string& s) {
    std::string o;
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        case '\0': o += "\\0"; break;
        default:
            if (c < 0x20 || c == 0x7f) {
                char b[16];
                snprintf(b, sizeof b, "\\u{%X}", c);
                o += b;
            } else {
                o.push_back((char)c);
            }
        }
    }
    return o;
}

struct Printer {
    std::string out;
    int indent = 0;
    std::map<std::string, std::vector<const void*>> names[4];  // value, block, loop, fn
    std::map<const Type*, size_t> struct_ids;
    std::set<const Value*> printed_consts;

    void w(const std::string& s) { out += s; }
    void nl() {
        out.push_back('\n');
        out.append((size_t)indent * 4, ' ');
    }
    std::string name(int kind, const void* obj, const std::string& n) {
        auto& lst = names[kind][n];
        size_t idx = lst.size();
        for (size_t k = 0; k < lst.size(); ++k)
            if (lst[k] == obj) idx = k;
        if (idx == lst.size()) lst.push_back(obj);
        if (idx == 0 && is_plain_ident(n)) return n;
        return "\"" + escape(n) + "\"" + std::to_string(idx);
    }
    std::string vname(const Value* v) { return name(0, v, v->name); }
    std::string bname(const Block* b) { return name(1, b, b->name); }
    std::string lname(const Loop* l) { return name(2, l, l->name); }
    std::string fname(const Function* f) { return name(3, f, f->name); }

    void type(const Type* t) {
        switch (t->kind) {
        case TypeKind::Int: w(int_name(t->int_kind)); break;
        case TypeKind::Float: w(float_name(t->float_kind)); break;
        case TypeKind::Bool: w("bool"); break;
        case TypeKind::DataPtr: w("data_ptr"); break;
        case TypeKind::Vector:
            w(t->scalable ? "<vscale x " : "<");
            w(std::to_string(t->length) + " x ");
            type(t->element);
            w(">");
            break;
        case TypeKind::Fn:
            w("fn[");
            for (size_t k = 0; k < t->args.size(); ++k) {
                if (k) w(", ");
                type(t->args[k]);
            }
            w("] -> ");
            if (t->diverges) w("!");
            else {
                w("[");
                for (size_t k = 0; k < t->results.size(); ++k) {
                    if (k) w(", ");
                    type(t->results[k]);
                }
                w("]");
            }
            break;
        case TypeKind::Struct: {
            auto it = struct_ids.find(t);
            if (it != struct_ids.end()) {
                w("struct " + std::to_string(it->second) + " {}");
                break;
            }
            size_t sid = struct_ids.size();
            struct_ids[t] = sid;
            w("struct " + std::to_string(sid) + " {");
            ++indent;
            nl();
            w(std::string("size: ") + (t->size_variable ? "variable " : "fixed ") + hex(t->size) + ",");
            nl();
            w("align: " + hex(t->align) + ",");
            for (auto& mem : t->members) {
                nl();
                w(hex(mem.offset) + ": ");
                type(mem.type);
                w(",");
            }
            --indent;
            nl();
            w("}");
            break;
        }
        }
    }

    void constant(const Const& c) {
        switch (c.kind) {
        case ConstKind::Int: w(hex(c.bits) + int_name(c.type->int_kind)); break;
        case ConstKind::Float: w(std::string(float_name(c.type->float_kind)) + " " + hex(c.bits)); break;
        case ConstKind::Bool: w(c.bits ? "true" : "false"); break;
        case ConstKind::Vector:
            w("<");
            for (size_t k = 0; k < c.elements.size(); ++k) {
                if (k) w(", ");
                constant(c.elements[k]);
            }
            w(">");
            break;
        case ConstKind::Undef: w("undef "); type(c.type); break;
        case ConstKind::Null: w("null "); type(c.type); break;
        case ConstKind::Fn: w("fn " + fname(c.function)); break;
        }
    }

    void value_def(const Value* v) {
        w(vname(v) + ": ");
        type(v->type);
    }
    void value_use(const Value* v) {
        w(vname(v));
        if (v->has_const && printed_consts.insert(v).second) {
            w(": ");
            constant(v->constant);
        }
    }
    void use_list(const std::vector<Value*>& vs) {
        for (size_t k = 0; k < vs.size(); ++k) {
            if (k) w(", ");
            value_use(vs[k]);
        }
    }
    void def_list(const std::vector<Value*>& vs) {
        w("[");
        for (size_t k = 0; k < vs.size(); ++k) {
            if (k) w(", ");
            value_def(vs[k]);
        }
        w("]");
    }
    void results(bool div, const std::vector<Value*>& vs) {
        if (div) w("!");
        else def_list(vs);
    }
    void variable(const Variable& v) {
        type(v.type);
        w(", align: " + hex(v.align) + " -> ");
        value_def(v.pointer);
    }
    void brk(const Block* b, const std::vector<Value*>& vals) {
        w("break " + bname(b) + "[");
        use_list(vals);
        w("]");
    }
    void body(const Block* b) {
        w("{");
        ++indent;
        for (auto* ins : b->body) {
            nl();
            instr(ins);
            w(";");
        }
        --indent;
        nl();
        w("}");
    }
    void instr(const Instr* ins) {
        switch (ins->op) {
        case Op::Add:
            w("add [");
            use_list(ins->args);
            w("] -> ");
            def_list(ins->results);
            break;
        case Op::Load:
            w("load [");
            use_list(ins->args);
            w("], align: " + hex(ins->align) + " -> ");
            def_list(ins->results);
            break;
        case Op::Ext:
            w(std::string(kExtOps[ins->ext].name) + " [");
            use_list(ins->args);
            w("]");
            if (kExtOps[ins->ext].kind == ExtKind::Store) w(", align: " + hex(ins->align));
            w(" -> ");
            def_list(ins->results);
            break;
        case Op::Break: brk(ins->block, ins->args); break;
        case Op::Continue:
            w("continue " + lname(ins->loop) + "[");
            use_list(ins->args);
            w("]");
            break;
        case Op::Branch:
            w("branch [");
            use_list(ins->args);
            w("], {");
            ++indent;
            for (auto& c : ins->cases) {
                nl();
                constant(c.value);
                w(" -> ");
                brk(c.target, c.results);
                w(";");
            }
            --indent;
            nl();
            w("} -> []");
            break;
        case Op::BlockI:
            w("block " + bname(ins->block) + " -> ");
            results(ins->block->diverges, ins->block->results);
            w(" ");
            body(ins->block);
            break;
        case Op::LoopI: {
            const Loop* lp = ins->loop;
            w("loop " + lname(lp) + "[");
            use_list(lp->arguments);
            w("] -> ");
            results(lp->diverges, lp->results);
            w(" {");
            ++indent;
            nl();
            w("-> ");
            def_list(lp->params);
            w(";");
            nl();
            w(bname(lp->body) + " ");
            body(lp->body);
            --indent;
            nl();
            w("}");
            break;
        }
        }
        if (ins->location.present)
            w(" @ \"" + escape(ins->location.file) + "\":" + std::to_string(ins->location.line) + ":" + std::to_string(ins->location.column));
    }
    void function(const Function* f) {
        w("fn " + fname(f));
        def_list(f->params);
        w(" -> ");
        if (f->diverges) w("!");
        else {
            w("[");
            for (size_t k = 0; k < f->returns.size(); ++k) {
                if (k) w(", ");
                type(f->returns[k]);
            }
            w("]");
        }
        w(" {");
        ++indent;
        nl();
        w("hints {");
        ++indent;
        nl();
        w("inlining_hint: " + f->inlining_hint + ",");
        nl();
        w("side_effects: " + f->side_effects + ",");
        --indent;
        nl();
        w("}");
        nl();
        w("{");
        ++indent;
        for (auto& v : f->locals) {
            nl();
            variable(v);
            w(";");
        }
        --indent;
        nl();
        w("}");
        nl();
        w(bname(f->body) + " ");
        body(f->body);
        --indent;
        nl();
        w("}");
    }
    void target_properties(IntKind dp, IntKind fp) {
        w("target_properties {");
        ++indent;
        nl();
        w(std::string("data_pointer_underlying_type: ") + int_name(dp) + ",");
        nl();
        w(std::string("function_pointer_underlying_type: ") + int_name(fp) + ",");
        --indent;
        nl();
        w("}");
    }
    void interface_block(const char* title, const InterfaceBlock& b) {
        nl();
        w(std::string(title) + " {");
        ++indent;
        nl();
        w("-> ");
        value_def(b.pointer);
        w(";");
        nl();
        w(std::string("size: ") + (b.size_variable ? "variable " : "fixed ") + hex(b.size) + ";");
        nl();
        w("align: " + hex(b.align) + ";");
        for (auto& mem : b.members) {
            nl();
            w(hex(mem.offset) + ": ");
            type(mem.type);
            w(": {};");
        }
        --indent;
        nl();
        w("}");
    }
    void interface_vars(const char* title, const std::vector<Variable>& vs) {
        nl();
        w(std::string(title) + " {");
        ++indent;
        for (auto& v : vs) {
            nl();
            variable(v);
            w(": { kind: " + v.built_in_kind + " };");
        }
        --indent;
        nl();
        w("}");
    }
    void module(const Module& m) {
        w("module {");
        ++indent;
        nl();
        target_properties(m.data_ptr_type, m.fn_ptr_type);
        interface_vars("built_in_inputs", m.built_in_inputs);
        interface_block("user_inputs_block", m.user_inputs);
        interface_vars("built_in_outputs", m.built_in_outputs);
        interface_block("user_outputs_block", m.user_outputs);
        nl();
        w("invocation_global_variables {");
        ++indent;
        for (auto& v : m.globals) {
            nl();
            variable(v);
            w(";");
        }
        --indent;
        nl();
        w("}");
        for (auto* f : m.functions) {
            nl();
            function(f);
        }
        nl();
        w("entry_point: " + fname(m.entry_point) + ";");
        --indent;
        nl();
        w("}");
    }
};

}  // namespace

std::unique_ptr<Module> parse_module(const std::string& text, int version) {
    auto m = std::make_unique<Module>();
    m->version = version;
    Parser p(*m, text, version);
    p.module();
    p.finish();
    return m;
}

std::string module_to_text(const Module& m) {
    Printer pr;
    pr.module(m);
    return pr.out + "\n";
}

std::string round_trip(const std::string& kind, const std::string& text) {
    if (kind == "module") return module_to_text(*parse_module(text));
    Module m;
    Parser p(m, text);
    Printer pr;
    if (kind == "type") {
        const Type* t = p.type();
        p.finish();
        pr.type(t);
    } else if (kind == "const") {
        Const c = p.constant();
        p.finish();
        pr.constant(c);
    } else if (kind == "target_properties") {
        IntKind dp, fp;
        p.target_properties(&dp, &fp);
        p.finish();
        pr.target_properties(dp, fp);
    } else if (kind == "block") {
        p.expect("block");
        Instr ins;
        ins.op = Op::BlockI;
        ins.block = p.block_instruction();
        p.finish();
        pr.instr(&ins);
    } else if (kind == "function") {
        Function* f = p.function();
        p.finish();
        pr.function(f);
    } else if (kind == "definitions_then_instruction") {
        auto defs = p.defs();
        Instr* ins = p.instruction();
        p.finish();
        pr.def_list(defs);
        pr.w("\n");
        pr.instr(ins);
    } else {
        throw Error("unknown syntactic category '" + kind + "'");
    }
    return pr.out;
}

}  // namespace kir
