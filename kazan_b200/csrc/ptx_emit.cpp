// ptx_emit.cpp — shader-compiler-ir -> PTX for sm_100a: the code generator of `shader-compiler-backend-ptx`.
//
// Replaces, for the B200, what shader-compiler-backend-llvm-7/src/backend.rs:553-690 (`LLVM7Compiler::run`) is for the
// host CPU.  Input is the structured SSA of shader-compiler-ir (instructions_impl.rs:395-433: add, load, break, branch,
// continue, block, loop) plus, for modules read as version 2, the instructions of ir.h (`ExtOp`: float arithmetic,
// compares, select, conversions, store, ptr_add, subgroup operations).  Two forms of output:
//   emit_ptx             one `.visible .entry` per module, one invocation per thread (BASELINE north_star): compute
//                        dispatch.  Subgroup operations map to warp votes and shuffles (a subgroup is a warp).
//   emit_fragment_inline the entry point of a fragment shader as a PTX scope that reads its inputs from, and leaves its
//                        outputs in, registers of the surrounding code: fragment_jit.cpp puts one copy at every place the
//                        raster kernels (raster.cu, compiled to PTX at build time) call the fragment program.
//
// Invocation ABI of a kernel (include/kazan_b200_backend.h): parameters (user_inputs_base, user_outputs_base,
// invocation_count, first_invocation, global_size_x, global_size_y, descriptor_table, global_base[16], push_constants[128]).  Thread g of the
// grid (2-D folded) runs invocation first+g; its interface blocks (interface.rs:199-252) are base + (first+g) * block.size.
// The values the entry point's body block is left with (`break <body>[v...]`, block.rs:16-23) are stored to the members of
// user_outputs_block, in order.
//
// Lowering:
//   value                    -> one virtual register of its class (.pred bool, .b16 i8/i16/f16, .b32, .b64)
//   block name -> [r..] {..} -> fall-through region closed by a label; `break` = moves into r.. + bra to the label
//   loop                     -> moves arguments into the parameters, head label, body block; `continue` = parallel move + bra
//   branch                   -> setp.eq chain, each case a predicated skip around (moves + bra)
//   load / store, align A    -> one access of natural width when A >= size, otherwise A-sized pieces
//   variables (function.rs:155-193, module.rs:145-147, interface.rs:61-67): a variable whose pointer is only ever used as
//   `pointer + constant` in loads and stores of disjoint scalar slots lives in registers, one per slot (built-in inputs
//   arrive that way: global_invocation_id is three registers, not memory); any other use of the pointer gives it a
//   .local array (address through cvta.local).
#include <algorithm>
#include <cstdio>
#include <map>
#include <set>
#include <sstream>

#include "ptx_emit.h"

namespace kir {
namespace {

enum class RC { Pred, B16, B32, B64 };
enum class Mode { Kernel, Inline };

// what a data_ptr value is known to point at: `base` + a constant
struct Base {
    enum Kind { UserIn, UserOut, Var, PushConstants } kind = Var;
    const Variable* var = nullptr;
    std::string built_in;  // kind of a built-in variable
    uint64_t size = 0;
    bool scalarized = true;  // lives in registers: one per accessed slot
    bool conflicted = false;
    std::map<uint64_t, unsigned> slots;      // offset -> bytes, from the analysis
    std::map<uint64_t, std::string> cells;   // offset -> register
    std::map<uint64_t, std::string> cells_b; // pair mode: the registers of the second fragment while the first one's are in `cells`
    std::string address;                     // register (or symbol expression) of the base when it is memory
};
struct Prov {
    int base;
    int64_t offset;
};

struct Emitter {
    const Module& m;
    Mode mode;
    std::string tag;  // suffix that makes labels and symbols of one inline copy unique
    std::ostringstream body;
    std::map<const Value*, std::string> reg;
    std::map<const Block*, const std::vector<Value*>*> dests;  // where `break block[...]` writes
    std::map<const Block*, std::string> end_label;
    std::map<const Loop*, std::string> head_label;
    int n_pred = 0, n16 = 0, n32 = 0, n64 = 0, n_label = 0, n_var = 0;
    std::ostringstream locals;
    std::vector<Base> bases;
    std::map<const Value*, Prov> prov;
    std::string submask;      // kernel: member mask of the in-range threads of the warp (subgroup operations)
    bool uses_subgroup = false;
    // ---- pair mode (fragment shaders): ONE copy of the code shades TWO fragments.  Values that differ between the two
    // ("varying": everything computed from frag_coord or the interpolated inputs) exist twice; an f32 one lives packed in a
    // 64-bit register and its add / sub / mul / fma are the packed fma.rn.f32x2 forms (FFMA2: two IEEE operations per issue
    // slot, each identical to the scalar instruction), every other varying instruction is emitted once per fragment.
    // Values that do not depend on the fragment (push constants, loop counters) and all control flow stay single.
    bool pair = false;
    std::set<const Value*> varying;
    std::set<std::pair<int, uint64_t>> varying_cells;
    struct PairReg {
        std::string packed, s0, s1;
    };
    std::map<const Value*, PairReg> preg;
    // f32 values made by a packed multiply (or a sign change of one).  ptxas contracts add.rn.f32x2 / sub.rn.f32x2 of a
    // mul.rn.f32x2 result into one FFMA2 although all of them carry .rn (it does not do that to the scalar forms): a sum with
    // such an operand is therefore emitted per fragment, as two scalar instructions.
    std::set<const Value*> packed_products;
    std::string pc_param;     // symbol (+ offset) of the push constant block in parameter space, when it is there
    bool reads_user_inputs = false, reads_frag_coord = false, reads_frag_coord_zw = false;

    Emitter(const Module& mod, Mode md, const std::string& suffix = "") : m(mod), mode(md), tag(suffix) {}

    [[noreturn]] static void unsupported(const std::string& what) { throw Unsupported("PTX back end: " + what); }

    static RC reg_class(const Type* t) {
        switch (t->kind) {
            case TypeKind::Bool: return RC::Pred;
            case TypeKind::Int: return int_bits(t->int_kind) <= 16 ? RC::B16 : int_bits(t->int_kind) == 32 ? RC::B32 : RC::B64;
            case TypeKind::Float: return float_bits(t->float_kind) == 16 ? RC::B16 : float_bits(t->float_kind) == 32 ? RC::B32 : RC::B64;
            case TypeKind::DataPtr:
            case TypeKind::Fn: return RC::B64;
            default: unsupported("vector / struct typed SSA values (only scalar values are lowered)");
        }
    }
    static int scalar_bytes(const Type* t) {
        switch (t->kind) {
            case TypeKind::Bool: return 1;
            case TypeKind::Int: return int_bits(t->int_kind) / 8;
            case TypeKind::Float: return float_bits(t->float_kind) / 8;
            case TypeKind::DataPtr:
            case TypeKind::Fn: return 8;
            default: unsupported("vector / struct typed SSA values (only scalar values are lowered)");
        }
    }
    static RC class_of_bytes(int n) { return n <= 2 ? RC::B16 : n == 4 ? RC::B32 : RC::B64; }
    std::string fresh(RC c) {
        char b[48];
        const char* t = tag.c_str();
        switch (c) {
            case RC::Pred: snprintf(b, sizeof b, "%%p%s%d", t, n_pred++); break;
            case RC::B16: snprintf(b, sizeof b, "%%h%s%d", t, n16++); break;
            case RC::B32: snprintf(b, sizeof b, "%%r%s%d", t, n32++); break;
            case RC::B64: snprintf(b, sizeof b, "%%rd%s%d", t, n64++); break;
        }
        return b;
    }
    static const char* mov_type(RC c) { return c == RC::Pred ? "pred" : c == RC::B16 ? "b16" : c == RC::B32 ? "b32" : "b64"; }
    std::string label(const char* stem) {
        char b[96];
        snprintf(b, sizeof b, "$%s%s_%d", stem, tag.c_str(), n_label++);
        return b;
    }
    void line(const std::string& s) { body << "    " << s << "\n"; }
    void put_label(const std::string& l) { body << l << ":\n"; }
    static bool is_immediate(const std::string& s) { return s.rfind("0x", 0) == 0 || s == "0" || s.rfind("0f", 0) == 0 || s.rfind("0d", 0) == 0; }
    // an operand that must be a register: immediates are moved into one
    std::string in_register(const std::string& operand, RC c) {
        if (!is_immediate(operand)) return operand;
        std::string r = fresh(c);
        line(std::string("mov.") + mov_type(c) + " " + r + ", " + operand + ";");
        return r;
    }

    // register that holds a definition
    const std::string& def(const Value* v) {
        auto it = reg.find(v);
        if (it != reg.end()) return it->second;
        return reg.emplace(v, fresh(reg_class(v->type))).first->second;
    }
    // the register holding `base + offset` of a pointer the analysis knows (memory-backed bases only)
    std::string materialize(const Prov& p) {
        Base& b = bases[p.base];
        if (b.scalarized) throw Error("internal: the address of a register-resident variable was asked for");
        if (p.offset == 0) return b.address;
        std::string r = fresh(RC::B64);
        line("add.s64 " + r + ", " + b.address + ", " + std::to_string(p.offset) + ";");
        return r;
    }
    // operand text for a use: registers for definitions, immediates for inline constants (values.rs:168-200)
    std::string use(const Value* v) {
        if (!v->has_const) {
            auto it = reg.find(v);
            auto pr = prov.find(v);
            // a known place used as a value: base + constant, made where it is used (a register made inside one block
            // would not be defined on the paths around it)
            if (pr != prov.end() && it == reg.end()) return materialize(pr->second);
            if (it == reg.end()) throw Error("value '" + v->name + "' is used before it is defined");
            return it->second;
        }
        const Const& c = v->constant;
        RC rc = reg_class(v->type);
        char b[64];
        switch (c.kind) {
            case ConstKind::Int: {
                uint64_t bits = c.bits;
                int w = int_bits(v->type->int_kind);
                if (w < 64) bits &= (1ull << w) - 1;
                snprintf(b, sizeof b, "0x%llx", (unsigned long long)bits);
                return b;
            }
            case ConstKind::Float: {
                // bit pattern moved as an untyped immediate: no decimal round trip
                snprintf(b, sizeof b, "0x%llx", (unsigned long long)c.bits);
                return b;
            }
            case ConstKind::Bool: {
                std::string t = fresh(RC::B32), p = fresh(RC::Pred);
                line("mov.b32 " + t + ", " + (c.bits ? "1" : "0") + ";");
                line("setp.ne.u32 " + p + ", " + t + ", 0;");
                return p;
            }
            case ConstKind::Null: return "0";
            case ConstKind::Undef: {  // any value will do; a defined register keeps ptxas quiet
                std::string r = fresh(rc);
                if (rc == RC::Pred) {
                    std::string t = fresh(RC::B32);
                    line("mov.b32 " + t + ", 0;");
                    line("setp.ne.u32 " + r + ", " + t + ", 0;");
                } else {
                    line(std::string("mov.") + mov_type(rc) + " " + r + ", 0;");
                }
                return r;
            }
            case ConstKind::Fn: unsupported("function-pointer constants (indirect calls are outside the hot path)");
            case ConstKind::Vector: unsupported("vector constants");
        }
        unsupported("constant kind");
    }
    // a use that an arithmetic instruction reads as a typed operand: float constants become typed immediates
    // (an untyped 0x.. literal is read as an integer by floating-point instructions)
    std::string use_float(const Value* v) {
        std::string s = use(v);
        if (!is_immediate(s)) return s;
        return in_register(s, reg_class(v->type));
    }
    std::string use_reg(const Value* v) { return in_register(use(v), reg_class(v->type)); }
    void move(const Value* dst, const std::string& src) {
        RC c = reg_class(dst->type);
        line(std::string("mov.") + mov_type(c) + " " + def(dst) + ", " + src + ";");
    }
    // simultaneous assignment dsts <- srcs (continue re-binds loop parameters that may appear among the sources)
    void parallel_move(const std::vector<Value*>& dsts, const std::vector<Value*>& srcs, const char* what) {
        if (dsts.size() != srcs.size()) throw Error(std::string(what) + ": value count does not match the target's definition count");
        std::vector<std::string> ops;
        if (pair) {
            bool any = false;
            for (const Value* d : dsts) any = any || is_varying(d);
            if (any) {
                // sources first (a destination may be among them), then the assignments
                std::vector<std::pair<std::string, std::string>> vals;
                for (size_t i = 0; i < srcs.size(); ++i) {
                    if (reg_class(dsts[i]->type) != reg_class(srcs[i]->type)) throw Error(std::string(what) + ": type mismatch");
                    auto v = scalars_of(srcs[i]);
                    if (!srcs[i]->has_const) {
                        const RC c = reg_class(srcs[i]->type);
                        std::string t0 = fresh(c), t1 = fresh(c);
                        line(std::string("mov.") + mov_type(c) + " " + t0 + ", " + v.first + ";");
                        line(std::string("mov.") + mov_type(c) + " " + t1 + ", " + v.second + ";");
                        v = {t0, t1};
                    }
                    vals.push_back(v);
                }
                for (size_t i = 0; i < dsts.size(); ++i) {
                    if (is_varying(dsts[i])) pair_assign(dsts[i], vals[i].first, vals[i].second);
                    else move(dsts[i], vals[i].first);
                }
                return;
            }
        }
        for (size_t i = 0; i < srcs.size(); ++i) {
            if (reg_class(dsts[i]->type) != reg_class(srcs[i]->type)) throw Error(std::string(what) + ": type mismatch");
            std::string s = use(srcs[i]);
            if (srcs.size() > 1 && !srcs[i]->has_const) {
                RC c = reg_class(srcs[i]->type);
                std::string t = fresh(c);
                line(std::string("mov.") + mov_type(c) + " " + t + ", " + s + ";");
                s = t;
            }
            ops.push_back(s);
        }
        for (size_t i = 0; i < dsts.size(); ++i) move(dsts[i], ops[i]);
    }

    // ---- pointer analysis -----------------------------------------------------------------------------------------
    int add_base(Base b) {
        bases.push_back(std::move(b));
        return (int)bases.size() - 1;
    }
    void pointer_escapes(const Value* v) {
        auto it = prov.find(v);
        if (it != prov.end()) bases[it->second.base].scalarized = false;
    }
    void note_access(const Value* pointer, int bytes) {
        auto it = prov.find(pointer);
        if (it == prov.end()) return;
        Base& b = bases[it->second.base];
        if (b.kind == Base::UserIn) reads_user_inputs = true;
        if (b.built_in == "frag_coord") reads_frag_coord = true;
        const int64_t off = it->second.offset;
        if (b.built_in == "frag_coord" && off + bytes > 8) reads_frag_coord_zw = true;
        if (off < 0 || (uint64_t)off + (uint64_t)bytes > b.size) {
            b.conflicted = true;  // outside the object: only real memory gives that a meaning
            return;
        }
        auto s = b.slots.find((uint64_t)off);
        if (s != b.slots.end() && s->second != (unsigned)bytes) b.conflicted = true;
        b.slots[(uint64_t)off] = (unsigned)bytes;
    }
    void analyze_values(const std::vector<Value*>& vs) {
        for (const Value* v : vs) pointer_escapes(v);
    }
    void analyze(const Block* b) {
        for (const Instr* ins : b->body) {
            switch (ins->op) {
                case Op::Load:
                    if (ins->args.size() == 1 && ins->results.size() == 1) {
                        note_access(ins->args[0], scalar_bytes(ins->results[0]->type));
                        // the pointer a push_constants variable holds is itself a known place
                        auto it = prov.find(ins->args[0]);
                        if (it != prov.end() && bases[it->second.base].built_in == "push_constants" && it->second.offset == 0 &&
                            ins->results[0]->type->kind == TypeKind::DataPtr)
                            prov[ins->results[0]] = Prov{push_constants_base(), 0};
                    }
                    break;
                case Op::Ext: {
                    const ExtKind k = kExtOps[ins->ext].kind;
                    if (k == ExtKind::PtrAdd && ins->args.size() == 2 && ins->results.size() == 1) {
                        auto it = prov.find(ins->args[0]);
                        if (it != prov.end()) {
                            const Value* off = ins->args[1];
                            if (off->has_const && off->constant.kind == ConstKind::Int) prov[ins->results[0]] = Prov{it->second.base, it->second.offset + (int64_t)off->constant.bits};
                            else bases[it->second.base].scalarized = false;
                        }
                    } else if (k == ExtKind::Store && ins->args.size() == 2) {
                        note_access(ins->args[0], scalar_bytes(ins->args[1]->type));
                        pointer_escapes(ins->args[1]);
                    } else {
                        analyze_values(ins->args);
                    }
                    break;
                }
                case Op::Add: analyze_values(ins->args); break;
                case Op::Break:
                case Op::Continue: analyze_values(ins->args); break;
                case Op::Branch:
                    analyze_values(ins->args);
                    for (auto& c : ins->cases) analyze_values(c.results);
                    break;
                case Op::BlockI: analyze(ins->block); break;
                case Op::LoopI:
                    analyze_values(ins->loop->arguments);
                    analyze(ins->loop->body);
                    break;
            }
        }
    }
    int pc_base = -1;
    int push_constants_base() {
        if (pc_base < 0) {
            Base b;
            b.kind = Base::PushConstants;
            b.size = 128;
            b.scalarized = false;
            pc_base = add_base(std::move(b));
        }
        return pc_base;
    }
    void finish_analysis() {
        for (Base& b : bases) {
            if (b.conflicted) b.scalarized = false;
            if (!b.scalarized) continue;
            uint64_t end = 0;
            for (auto& s : b.slots) {  // ordered by offset: slots must not overlap
                if (s.first < end) b.scalarized = false;
                end = s.first + s.second;
            }
        }
    }

    // ---- pair mode: which values differ between the two fragments ---------------------------------------------------------
    bool is_varying(const Value* v) const { return pair && !v->has_const && varying.count(v) != 0; }
    static bool is_f32(const Type* t) { return t->kind == TypeKind::Float && float_bits(t->float_kind) == 32; }
    bool taint_values(const std::vector<Value*>& dsts, const std::vector<Value*>& srcs) {
        bool changed = false;
        for (size_t i = 0; i < dsts.size() && i < srcs.size(); ++i)
            if (!srcs[i]->has_const && varying.count(srcs[i]) && varying.insert(dsts[i]).second) changed = true;
        return changed;
    }
    // one sweep; returns true when something became varying.  `eligible` is cleared when the code cannot be shared by two
    // fragments: control flow that depends on the fragment, or memory that both would write.
    bool taint(const Block* b, bool* eligible) {
        bool changed = false;
        auto var = [&](const Value* v) { return !v->has_const && varying.count(v) != 0; };
        auto mark = [&](const Value* v) {
            if (varying.insert(v).second) changed = true;
        };
        for (const Instr* ins : b->body) {
            switch (ins->op) {
                case Op::Load: {
                    const Value* pv = ins->args[0];
                    auto it = prov.find(pv);
                    bool v = var(pv);
                    if (it != prov.end()) {
                        const Base& bs = bases[it->second.base];
                        if (bs.scalarized && varying_cells.count({it->second.base, (uint64_t)it->second.offset})) v = true;
                    }
                    if (v) mark(ins->results[0]);
                    break;
                }
                case Op::Ext:
                    if (kExtOps[ins->ext].kind == ExtKind::Store) {
                        auto it = prov.find(ins->args[0]);
                        if (it != prov.end() && bases[it->second.base].scalarized) {
                            if (var(ins->args[1]) && varying_cells.insert({it->second.base, (uint64_t)it->second.offset}).second) changed = true;
                        } else {
                            *eligible = false;  // a store to real memory would be done once for two fragments
                        }
                        break;
                    }
                    [[fallthrough]];
                case Op::Add: {
                    bool v = false;
                    for (const Value* a : ins->args) v = v || var(a);
                    if (v)
                        for (const Value* r : ins->results) mark(r);
                    break;
                }
                case Op::Break: {
                    auto it = dests_all.find(ins->block);
                    if (it != dests_all.end() && taint_values(*it->second, ins->args)) changed = true;
                    break;
                }
                case Op::Continue:
                    if (taint_values(ins->loop->params, ins->args)) changed = true;
                    break;
                case Op::Branch:
                    if (var(ins->args[0])) *eligible = false;
                    for (auto& c : ins->cases) {
                        auto it = dests_all.find(c.target);
                        if (it != dests_all.end() && taint_values(*it->second, c.results)) changed = true;
                    }
                    break;
                case Op::BlockI:
                    dests_all[ins->block] = &ins->block->results;
                    if (taint(ins->block, eligible)) changed = true;
                    break;
                case Op::LoopI:
                    dests_all[ins->loop->body] = &ins->loop->results;
                    if (taint_values(ins->loop->params, ins->loop->arguments)) changed = true;
                    if (taint(ins->loop->body, eligible)) changed = true;
                    break;
            }
        }
        return changed;
    }
    std::map<const Block*, const std::vector<Value*>*> dests_all;
    // can one copy of the entry point's code shade two fragments?  (fills `varying` / `varying_cells`)
    bool analyze_pair(const Function* f, int in_base, int frag_coord_base) {
        bool eligible = true;
        for (const Base& b : bases)
            if (b.kind == Base::Var && !b.scalarized) eligible = false;  // a .local variable would be shared by the two
        for (auto& s : bases[in_base].slots) varying_cells.insert({in_base, s.first});
        if (frag_coord_base >= 0)
            for (auto& s : bases[frag_coord_base].slots) varying_cells.insert({frag_coord_base, s.first});
        dests_all[f->body] = &f->body->results;
        while (taint(f->body, &eligible)) {
        }
        while (mark_products(f->body)) {
        }
        return eligible;
    }
    // packed_products to a fixed point: results of f_mul on varying f32 values, sign changes of those, and the block
    // results / loop parameters they are passed to
    bool mark_products(const Block* b) {
        bool changed = false;
        auto pass = [&](const std::vector<Value*>& dsts, const std::vector<Value*>& srcs) {
            for (size_t i = 0; i < dsts.size() && i < srcs.size(); ++i)
                if (packed_products.count(srcs[i]) && packed_products.insert(dsts[i]).second) changed = true;
        };
        for (const Instr* ins : b->body) {
            switch (ins->op) {
                case Op::Ext: {
                    const ExtOp& op = kExtOps[ins->ext];
                    if (ins->results.size() != 1 || !is_f32(ins->results[0]->type) || !varying.count(ins->results[0])) break;
                    const bool mul = op.kind == ExtKind::FloatBinary && std::string(op.ptx) == "mul.rn";
                    const bool sign = op.kind == ExtKind::FloatUnary && packed_products.count(ins->args[0]);
                    const bool sel = op.kind == ExtKind::Select && (packed_products.count(ins->args[1]) || packed_products.count(ins->args[2]));
                    if ((mul || sign || sel) && packed_products.insert(ins->results[0]).second) changed = true;
                    break;
                }
                case Op::Break: {
                    auto it = dests_all.find(ins->block);
                    if (it != dests_all.end()) pass(*it->second, ins->args);
                    break;
                }
                case Op::Continue: pass(ins->loop->params, ins->args); break;
                case Op::Branch:
                    for (auto& c : ins->cases) {
                        auto it = dests_all.find(c.target);
                        if (it != dests_all.end()) pass(*it->second, c.results);
                    }
                    break;
                case Op::BlockI:
                    if (mark_products(ins->block)) changed = true;
                    break;
                case Op::LoopI:
                    pass(ins->loop->params, ins->loop->arguments);
                    if (mark_products(ins->loop->body)) changed = true;
                    break;
                default: break;
            }
        }
        return changed;
    }
    // a varying value's registers: f32 values are kept packed, everything else as two scalars
    PairReg& pair_def(const Value* v) {
        auto it = preg.find(v);
        if (it != preg.end()) return it->second;
        PairReg r;
        if (is_f32(v->type)) {
            r.packed = fresh(RC::B64);
        } else {
            r.s0 = fresh(reg_class(v->type));
            r.s1 = fresh(reg_class(v->type));
        }
        return preg.emplace(v, r).first->second;
    }
    // the packed form of an f32 value (a value that is the same for both fragments is duplicated)
    std::string packed_of(const Value* v) {
        if (is_varying(v)) {
            auto it = preg.find(v);
            if (it == preg.end()) throw Error("value '" + v->name + "' is used before it is defined");
            return it->second.packed;
        }
        const std::string u = in_register(use(v), RC::B32);
        std::string r = fresh(RC::B64);
        line("mov.b64 " + r + ", {" + u + ", " + u + "};");
        return r;
    }
    // the two scalar forms of a value
    std::pair<std::string, std::string> scalars_of(const Value* v) {
        if (!is_varying(v)) {
            const std::string u = use(v);
            return {u, u};
        }
        auto it = preg.find(v);
        if (it == preg.end()) throw Error("value '" + v->name + "' is used before it is defined");
        if (!it->second.packed.empty()) {
            std::string a = fresh(RC::B32), b = fresh(RC::B32);
            line("mov.b64 {" + a + ", " + b + "}, " + it->second.packed + ";");
            return {a, b};
        }
        return {it->second.s0, it->second.s1};
    }
    void pair_assign(const Value* dst, const std::string& v0, const std::string& v1) {
        PairReg& d = pair_def(dst);
        if (!d.packed.empty()) {
            line("mov.b64 " + d.packed + ", {" + in_register(v0, RC::B32) + ", " + in_register(v1, RC::B32) + "};");
        } else {
            const RC c = reg_class(dst->type);
            line(std::string("mov.") + mov_type(c) + " " + d.s0 + ", " + v0 + ";");
            line(std::string("mov.") + mov_type(c) + " " + d.s1 + ", " + v1 + ";");
        }
    }
    bool instruction_is_varying(const Instr* ins) {
        for (const Value* r : ins->results)
            if (is_varying(r)) return true;
        if (ins->op == Op::Ext && kExtOps[ins->ext].kind == ExtKind::Store) {
            auto it = prov.find(ins->args[0]);
            if (it != prov.end() && varying_cells.count({it->second.base, (uint64_t)it->second.offset})) return true;
        }
        return false;
    }
    void scalar_instruction(const Instr* ins) {
        switch (ins->op) {
            case Op::Add: add(ins); break;
            case Op::Load: load(ins); break;
            case Op::Ext: ext(ins); break;
            default: throw Error("internal: not a value instruction");
        }
    }
    // one varying instruction for both fragments
    void pair_instruction(const Instr* ins) {
        if (ins->op == Op::Ext && ins->results.size() == 1 && is_f32(ins->results[0]->type)) {
            const ExtOp& op = kExtOps[ins->ext];
            const std::string ptx = op.ptx;
            bool sum_of_product = false;
            if (op.kind == ExtKind::FloatBinary && (ptx == "add.rn" || ptx == "sub.rn"))
                for (const Value* a : ins->args) sum_of_product = sum_of_product || packed_products.count(a) != 0;
            if (!sum_of_product &&
                ((op.kind == ExtKind::FloatBinary && (ptx == "add.rn" || ptx == "sub.rn" || ptx == "mul.rn")) || op.kind == ExtKind::FloatTernary)) {
                std::string ops;
                for (const Value* a : ins->args) ops += ", " + packed_of(a);
                line(ptx + ".f32x2 " + pair_def(ins->results[0]).packed + ops + ";");
                return;
            }
        }
        // any other instruction: once per fragment, on that fragment's registers and cells
        std::vector<std::pair<std::string, std::string>> args;
        for (const Value* a : ins->args) args.push_back(is_varying(a) ? scalars_of(a) : std::pair<std::string, std::string>());
        std::string result[2];
        for (int frag = 0; frag < 2; ++frag) {
            std::map<const Value*, std::string> saved;
            for (size_t i = 0; i < ins->args.size(); ++i)
                if (is_varying(ins->args[i])) {
                    auto it = reg.find(ins->args[i]);
                    if (it != reg.end()) saved[ins->args[i]] = it->second;
                    reg[ins->args[i]] = frag ? args[i].second : args[i].first;
                }
            for (const Value* r : ins->results) reg.erase(r);
            if (frag)
                for (Base& b : bases) std::swap(b.cells, b.cells_b);
            scalar_instruction(ins);
            if (frag)
                for (Base& b : bases) std::swap(b.cells, b.cells_b);
            if (!ins->results.empty()) result[frag] = reg.at(ins->results[0]);
            for (const Value* r : ins->results) reg.erase(r);
            for (size_t i = 0; i < ins->args.size(); ++i)
                if (is_varying(ins->args[i])) {
                    reg.erase(ins->args[i]);
                    auto it = saved.find(ins->args[i]);
                    if (it != saved.end()) reg[ins->args[i]] = it->second;
                }
        }
        if (!ins->results.empty()) pair_assign(ins->results[0], result[0], result[1]);
    }

    // ---- memory -------------------------------------------------------------------------------------------------
    // text of `[register + constant]`
    static std::string address(const std::string& r, int64_t off) {
        const size_t plus = r.find('+');  // "symbol+constant": the constants add up
        if (plus != std::string::npos) return r.substr(0, plus) + "+" + std::to_string(std::stoll(r.substr(plus + 1)) + off);
        if (off == 0) return r;
        return r + (off > 0 ? "+" : "") + std::to_string(off);
    }
    // dst register <- size bytes at [addr + off], the access split into `piece`-byte naturally aligned loads
    void load_scalar(const Value* result, const std::string& addr, int64_t off, uint64_t align, const char* space = "") {
        const Type* t = result->type;
        int size = scalar_bytes(t);
        RC rc = reg_class(t);
        if (align == 0) align = 1;
        uint64_t piece = 1;
        while (piece * 2 <= align && piece * 2 <= (uint64_t)size) piece *= 2;
        std::string value;
        const std::string ld = std::string("ld") + space;
        if (piece == (uint64_t)size) {
            value = rc == RC::Pred ? fresh(RC::B16) : def(result);
            line(ld + ".u" + std::to_string(size * 8) + " " + value + ", [" + address(addr, off) + "];");
        } else {
            // little-endian assembly in a 64-bit temporary
            std::string acc = fresh(RC::B64), part = fresh(RC::B64);
            for (int o = 0; o < size; o += (int)piece) {
                std::string dstp = o == 0 ? acc : part;
                line(ld + ".u" + std::to_string(piece * 8) + " " + dstp + ", [" + address(addr, off + o) + "];");
                if (o) {
                    line("shl.b64 " + part + ", " + part + ", " + std::to_string(o * 8) + ";");
                    line("or.b64 " + acc + ", " + acc + ", " + part + ";");
                }
            }
            if (rc == RC::B64) {
                value = def(result);
                line("mov.b64 " + value + ", " + acc + ";");
            } else if (rc == RC::B32) {
                value = def(result);
                line("cvt.u32.u64 " + value + ", " + acc + ";");
            } else {
                value = rc == RC::Pred ? fresh(RC::B16) : def(result);
                line("cvt.u16.u64 " + value + ", " + acc + ";");
            }
        }
        if (rc == RC::Pred) line("setp.ne.u16 " + def(result) + ", " + value + ", 0;");
    }
    void store_scalar(const Type* t, const std::string& src_operand, const std::string& addr, int64_t offset, uint64_t align, const char* space = "") {
        int size = scalar_bytes(t);
        RC rc = reg_class(t);
        std::string src = src_operand;
        if (rc == RC::Pred) {
            std::string h = fresh(RC::B16);
            line("selp.u16 " + h + ", 1, 0, " + src + ";");
            src = h;
        } else {
            src = in_register(src, rc);  // immediates cannot be store sources
        }
        if (align == 0) align = 1;
        uint64_t piece = 1;
        while (piece * 2 <= align && piece * 2 <= (uint64_t)size) piece *= 2;
        const std::string st = std::string("st") + space;
        if (piece == (uint64_t)size) {
            line(st + ".u" + std::to_string(size * 8) + " [" + address(addr, offset) + "], " + src + ";");
            return;
        }
        std::string wide = fresh(RC::B64), part = fresh(RC::B64);
        line(std::string("cvt.u64.u") + (rc == RC::B64 ? "64 " : rc == RC::B32 ? "32 " : "16 ") + wide + ", " + src + ";");
        for (int off = 0; off < size; off += (int)piece) {
            line("shr.u64 " + part + ", " + wide + ", " + std::to_string(off * 8) + ";");
            line(st + ".u" + std::to_string(piece * 8) + " [" + address(addr, offset + off) + "], " + part + ";");
        }
    }
    // the register of slot `off` of a register-resident base (created on first use, holding 0)
    const std::string& cell(Base& b, uint64_t off) {
        auto it = b.cells.find(off);
        if (it != b.cells.end()) return it->second;
        auto s = b.slots.find(off);
        if (s == b.slots.end()) throw Error("internal: access to a slot the analysis did not see");
        const int bi = (int)(&b - &bases[0]);
        if (pair && !varying_cells.count({bi, off})) {  // one register for both fragments
            auto o = b.cells_b.find(off);
            if (o != b.cells_b.end()) return b.cells.emplace(off, o->second).first->second;
        }
        RC c = class_of_bytes((int)s->second);
        std::string r = fresh(c);
        prologue << "    mov." << mov_type(c) << " " << r << ", 0;\n";
        if (pair && !varying_cells.count({bi, off})) b.cells_b.emplace(off, r);
        return b.cells.emplace(off, r).first->second;
    }
    std::ostringstream prologue;  // register cells are zeroed ahead of the body

    // a variable: registers (nothing to emit: cells appear on demand) or a .local array
    void variable(const Variable& v, int base_index) {
        if (!v.pointer) return;
        Base& b = bases[base_index];
        if (b.scalarized) return;
        uint64_t size = type_size(v.type), align = v.align ? v.align : 1;
        if (align & (align - 1)) throw Error("variable alignment is not a power of two");
        std::string sym = "__kzn_var" + tag + std::to_string(n_var++);
        locals << "    .local .align " << align << " .b8 " << sym << "[" << (size ? size : 1) << "];\n";
        std::string r = fresh(RC::B64);
        line("mov.u64 " + r + ", " + sym + ";");
        line("cvta.local.u64 " + r + ", " + r + ";");
        b.address = r;
    }
    int declare_variable(const Variable& v) {
        Base b;
        b.kind = Base::Var;
        b.var = &v;
        b.built_in = v.built_in_kind;
        b.size = type_size(v.type);
        int i = add_base(std::move(b));
        if (v.pointer) prov[v.pointer] = Prov{i, 0};
        return i;
    }
    // give slot `off` (bytes) of a built-in input its value: a register cell or a store into the variable's memory
    void set_built_in(Base& b, uint64_t off, int bytes, const std::string& value) {
        RC c = class_of_bytes(bytes);
        if (b.scalarized) {
            if (!b.slots.count(off)) return;  // never read
            if (b.slots[off] != (unsigned)bytes) unsupported("built-in '" + b.built_in + "' is read with another width than it has");
            std::string r = fresh(c);
            line(std::string("mov.") + mov_type(c) + " " + r + ", " + value + ";");
            b.cells[off] = r;
        } else {
            std::string v = in_register(value, c);
            line("st.u" + std::to_string(bytes * 8) + " [" + address(b.address, (int64_t)off) + "], " + v + ";");
        }
    }

    // ---- instructions -------------------------------------------------------------------------------------------
    void comment_location(const Instr* ins) {
        if (ins->location.present) {
            std::string f = ins->location.file;
            for (auto& ch : f)
                if (ch == '\n' || ch == '\r') ch = ' ';
            line("// @ \"" + f + "\":" + std::to_string(ins->location.line) + ":" + std::to_string(ins->location.column));
        }
    }
    static const char* int_suffix(int w) { return w <= 16 ? "16" : w == 32 ? "32" : "64"; }
    void add(const Instr* ins) {
        if (ins->args.size() != 2 || ins->results.size() != 1) throw Error("add takes two arguments and one result");
        const Type* t = ins->results[0]->type;
        if (t->kind != TypeKind::Int) unsupported("add on non-integer types (the reference defines no float semantics, SURVEY.md §8 a12; version 2 has f_add)");
        if (ins->args[0]->type != t || ins->args[1]->type != t) throw Error("add: operand types differ from the result type");
        std::string a = use_reg(ins->args[0]), b = use(ins->args[1]);  // first source operand of add must be a register
        const std::string& d = def(ins->results[0]);
        int w = int_bits(t->int_kind);
        line(std::string("add.s") + int_suffix(w) + " " + d + ", " + a + ", " + b + ";");
        if (w == 8) line("and.b16 " + d + ", " + d + ", 0xff;");  // i8 lives zero-extended in a 16-bit register
    }
    void load(const Instr* ins) {
        if (ins->args.size() != 1 || ins->results.size() != 1) throw Error("load takes one pointer and one result");
        if (ins->args[0]->type->kind != TypeKind::DataPtr) throw Error("load: the argument is not a data_ptr");
        const Value* res = ins->results[0];
        auto it = prov.find(ins->args[0]);
        if (it != prov.end()) {
            Base& b = bases[it->second.base];
            if (b.scalarized) {  // a register cell
                const std::string& c = cell(b, (uint64_t)it->second.offset);
                if (reg_class(res->type) == RC::Pred) line("setp.ne.u16 " + def(res) + ", " + c + ", 0;");
                else move(res, c);
                return;
            }
            if (b.kind == Base::PushConstants && !pc_param.empty()) {
                load_scalar(res, pc_param, it->second.offset, ins->align, ".param");
                return;
            }
            load_scalar(res, b.address, it->second.offset, ins->align);
            return;
        }
        load_scalar(res, in_register(use(ins->args[0]), RC::B64), 0, ins->align);
    }
    void store(const Instr* ins) {
        const Value *p = ins->args[0], *v = ins->args[1];
        auto it = prov.find(p);
        if (it != prov.end()) {
            Base& b = bases[it->second.base];
            if (b.kind == Base::PushConstants) unsupported("stores to the push constant block");
            if (b.scalarized) {
                const std::string& c = cell(b, (uint64_t)it->second.offset);
                RC rc = reg_class(v->type);
                if (rc == RC::Pred) line("selp.u16 " + c + ", 1, 0, " + use(v) + ";");
                else line(std::string("mov.") + mov_type(rc) + " " + c + ", " + use(v) + ";");
                return;
            }
            store_scalar(v->type, use(v), b.address, it->second.offset, ins->align);
            return;
        }
        store_scalar(v->type, use(v), in_register(use(p), RC::B64), 0, ins->align);
    }
    // i8 values live zero-extended in 16-bit registers: a signed reading needs the sign first
    std::string signed16(const std::string& r8) {
        std::string t = fresh(RC::B16);
        line("cvt.s16.s8 " + t + ", " + r8 + ";");
        return t;
    }
    void ext(const Instr* ins) {
        const ExtOp& op = kExtOps[ins->ext];
        const std::string name = op.name, ptx = op.ptx;
        auto res = [&]() -> const std::string& { return def(ins->results[0]); };
        switch (op.kind) {
        case ExtKind::IntBinary: {
            const Type* t = ins->results[0]->type;
            const int w = int_bits(t->int_kind);
            const std::string sfx = int_suffix(w);
            std::string a = use_reg(ins->args[0]), b = use_reg(ins->args[1]);
            const bool shift = ptx.rfind("sh", 0) == 0;
            if (shift) {
                // PTX takes a 32-bit amount and clamps amounts past the width, which is this IR's meaning of an over-wide shift
                std::string amt = fresh(RC::B32);
                if (w == 64) {
                    std::string c = fresh(RC::B64);
                    line("min.u64 " + c + ", " + b + ", 64;");
                    line("cvt.u32.u64 " + amt + ", " + c + ";");
                } else if (w == 32) {
                    line("mov.b32 " + amt + ", " + b + ";");
                } else {
                    line("cvt.u32.u16 " + amt + ", " + b + ";");
                }
                if (ptx == "shr.s" && w == 8) a = signed16(a);
                std::string mn = ptx == "shl" ? "shl.b" + sfx : ptx + sfx;
                line(mn + " " + res() + ", " + a + ", " + amt + ";");
            } else if (ptx == "sub" || ptx == "mul.lo") {
                line(ptx + ".s" + sfx + " " + res() + ", " + a + ", " + b + ";");
            } else {
                line(ptx + ".b" + sfx + " " + res() + ", " + a + ", " + b + ";");
            }
            if (w == 8) line("and.b16 " + res() + ", " + res() + ", 0xff;");
            break;
        }
        case ExtKind::IntUnary: {
            const int w = int_bits(ins->results[0]->type->int_kind);
            std::string a = use_reg(ins->args[0]);
            line(ptx + (ptx == "neg" ? ".s" : ".b") + int_suffix(w) + " " + res() + ", " + a + ";");
            if (w == 8) line("and.b16 " + res() + ", " + res() + ", 0xff;");
            break;
        }
        case ExtKind::FloatBinary: {
            const std::string ft = float_bits(ins->results[0]->type->float_kind) == 64 ? ".f64" : ".f32";
            line(ptx + ft + " " + res() + ", " + use_float(ins->args[0]) + ", " + use_float(ins->args[1]) + ";");
            break;
        }
        case ExtKind::FloatUnary: {  // neg / abs: the sign bit only (ptxas folds them into operand modifiers); a NaN stays a NaN
            const std::string ft = float_bits(ins->results[0]->type->float_kind) == 64 ? ".f64" : ".f32";
            line(ptx + ft + " " + res() + ", " + use_float(ins->args[0]) + ";");
            break;
        }
        case ExtKind::FloatTernary: {
            const std::string ft = float_bits(ins->results[0]->type->float_kind) == 64 ? ".f64" : ".f32";
            line(ptx + ft + " " + res() + ", " + use_float(ins->args[0]) + ", " + use_float(ins->args[1]) + ", " + use_float(ins->args[2]) + ";");
            break;
        }
        case ExtKind::IntCompare: {
            const Type* t = ins->args[0]->type;
            const int w = t->kind == TypeKind::DataPtr ? 64 : int_bits(t->int_kind);
            std::string a = use_reg(ins->args[0]), b = use_reg(ins->args[1]);
            std::string cmp = ptx;  // "eq" "ne" or "lt.u" ...
            if (cmp == "eq" || cmp == "ne") cmp += ".b";
            else if (w == 8 && cmp.back() == 's') {
                a = signed16(a);
                b = signed16(b);
            }
            line("setp." + cmp + int_suffix(w) + " " + res() + ", " + a + ", " + b + ";");
            break;
        }
        case ExtKind::FloatCompare: {
            const std::string ft = float_bits(ins->args[0]->type->float_kind) == 64 ? ".f64" : ".f32";
            line("setp." + ptx + ft + " " + res() + ", " + use_float(ins->args[0]) + ", " + use_float(ins->args[1]) + ";");
            break;
        }
        case ExtKind::BoolBinary: line(ptx + ".pred " + res() + ", " + use(ins->args[0]) + ", " + use(ins->args[1]) + ";"); break;
        case ExtKind::BoolUnary: line("not.pred " + res() + ", " + use(ins->args[0]) + ";"); break;
        case ExtKind::Select: {
            const RC rc = reg_class(ins->results[0]->type);
            const std::string c = use(ins->args[0]);
            if (rc == RC::Pred) {
                std::string a = use(ins->args[1]), b = use(ins->args[2]), t1 = fresh(RC::Pred), t2 = fresh(RC::Pred), nc = fresh(RC::Pred);
                line("and.pred " + t1 + ", " + c + ", " + a + ";");
                line("not.pred " + nc + ", " + c + ";");
                line("and.pred " + t2 + ", " + nc + ", " + b + ";");
                line("or.pred " + res() + ", " + t1 + ", " + t2 + ";");
            } else {
                pointer_value_check(ins);
                line(std::string("selp.") + mov_type(rc) + " " + res() + ", " + use(ins->args[1]) + ", " + use(ins->args[2]) + ", " + c + ";");
            }
            break;
        }
        case ExtKind::Convert: convert(ins); break;
        case ExtKind::Store: store(ins); break;
        case ExtKind::PtrAdd: {
            if (prov.count(ins->results[0])) break;  // known place: no code until it is used as a value
            std::string p = in_register(use(ins->args[0]), RC::B64);
            line("add.s64 " + res() + ", " + p + ", " + use(ins->args[1]) + ";");
            break;
        }
        case ExtKind::SubgroupElect:
        case ExtKind::SubgroupVote:
        case ExtKind::SubgroupBallot:
        case ExtKind::SubgroupBroadcast:
        case ExtKind::SubgroupShuffle:
        case ExtKind::SubgroupReduce: subgroup(ins); break;
        }
    }
    void pointer_value_check(const Instr*) {}
    void convert(const Instr* ins) {
        const std::string n = kExtOps[ins->ext].name;
        const Value *a = ins->args[0], *r = ins->results[0];
        const std::string src = use_reg(a);
        const std::string& d = def(r);
        auto ibits = [](const Value* v) { return int_bits(v->type->int_kind); };
        auto fty = [](const Value* v) { return float_bits(v->type->float_kind) == 64 ? std::string("f64") : std::string("f32"); };
        if (n == "convert_f_to_u" || n == "convert_f_to_s") {
            const int w = ibits(r);
            const char s = n.back();  // 'u' or 's'
            // round toward zero; PTX saturates out-of-range values and turns NaN into 0, which is this IR's definition
            line(std::string("cvt.rzi.") + s + std::to_string(w) + "." + fty(a) + " " + d + ", " + src + ";");
            if (w == 8) line("and.b16 " + d + ", " + d + ", 0xff;");
        } else if (n == "convert_s_to_f" || n == "convert_u_to_f") {
            const char s = n[8];  // convert_S_to_f / convert_U_to_f
            line(std::string("cvt.rn.") + fty(r) + "." + s + std::to_string(ibits(a)) + " " + d + ", " + src + ";");
        } else if (n == "u_convert" || n == "s_convert") {
            const char s = n[0];
            const int wa = ibits(a), wr = ibits(r);
            if (wa == wr) line(std::string("mov.b") + int_suffix(wr) + " " + d + ", " + src + ";");
            else line(std::string("cvt.") + s + std::to_string(wr) + "." + s + std::to_string(wa) + " " + d + ", " + src + ";");
            if (wr == 8) line("and.b16 " + d + ", " + d + ", 0xff;");
        } else if (n == "f_convert") {
            if (fty(a) == fty(r)) line("mov.b" + std::string(fty(r) == "f64" ? "64 " : "32 ") + d + ", " + src + ";");
            else if (fty(r) == "f64") line("cvt.f64.f32 " + d + ", " + src + ";");
            else line("cvt.rn.f32.f64 " + d + ", " + src + ";");
        } else {  // bitcast: same size, same register class
            if (reg_class(a->type) != reg_class(r->type)) throw Error("bitcast between values of different sizes");
            line(std::string("mov.") + mov_type(reg_class(r->type)) + " " + d + ", " + src + ";");
        }
    }
    void subgroup(const Instr* ins) {
        if (mode != Mode::Kernel) unsupported("subgroup operations outside compute shaders (fragment invocations are not laid out in subgroups by the rasteriser)");
        uses_subgroup = true;
        const ExtOp& op = kExtOps[ins->ext];
        const std::string ptx = op.ptx;
        const std::string& d = def(ins->results[0]);
        const std::string mask = "%kzn_submask";
        switch (op.kind) {
        case ExtKind::SubgroupElect: {  // the in-range lanes of a warp start at lane 0
            std::string l = fresh(RC::B32);
            line("mov.u32 " + l + ", %laneid;");
            line("setp.eq.u32 " + d + ", " + l + ", 0;");
            break;
        }
        case ExtKind::SubgroupVote: line("vote.sync." + ptx + ".pred " + d + ", " + use(ins->args[0]) + ", " + mask + ";"); break;
        case ExtKind::SubgroupBallot: line("vote.sync.ballot.b32 " + d + ", " + use(ins->args[0]) + ", " + mask + ";"); break;
        case ExtKind::SubgroupBroadcast: line("shfl.sync.idx.b32 " + d + ", " + use_reg(ins->args[0]) + ", 0, 31, " + mask + ";"); break;
        case ExtKind::SubgroupShuffle: line("shfl.sync." + ptx + ".b32 " + d + ", " + use_reg(ins->args[0]) + ", " + use(ins->args[1]) + ", 31, " + mask + ";"); break;
        case ExtKind::SubgroupReduce: line("redux.sync.add.s32 " + d + ", " + use_reg(ins->args[0]) + ", " + mask + ";"); break;
        default: break;
        }
    }
    void do_break(const Instr* ins) {
        auto it = dests.find(ins->block);
        if (it == dests.end()) throw Error("break targets block '" + (ins->block ? ins->block->name : "?") + "' which does not enclose it");
        parallel_move(*it->second, ins->args, "break");
        line("bra " + end_label.at(ins->block) + ";");
    }
    void do_continue(const Instr* ins) {
        auto it = head_label.find(ins->loop);
        if (it == head_label.end()) throw Error("continue targets a loop that does not enclose it");
        parallel_move(ins->loop->params, ins->args, "continue");
        line("bra " + it->second + ";");
    }
    void branch(const Instr* ins) {
        if (ins->args.size() != 1) throw Error("branch takes one selector");
        const Type* t = ins->args[0]->type;
        std::string x = use(ins->args[0]);
        for (auto& c : ins->cases) {
            auto it = dests.find(c.target);
            if (it == dests.end()) throw Error("branch case targets a block that does not enclose it");
            std::string skip = label("case_skip");
            if (t->kind == TypeKind::Bool) {
                if (c.value.kind != ConstKind::Bool) throw Error("branch: case constant type differs from the selector type");
                line(std::string("@") + (c.value.bits ? "!" : "") + x + " bra " + skip + ";");
            } else if (t->kind == TypeKind::Int) {
                if (c.value.kind != ConstKind::Int) throw Error("branch: case constant type differs from the selector type");
                int w = int_bits(t->int_kind);
                uint64_t bits = c.value.bits;
                if (w < 64) bits &= (1ull << w) - 1;
                std::string xr = in_register(x, reg_class(t));
                std::string p = fresh(RC::Pred);
                char imm[32];
                snprintf(imm, sizeof imm, "0x%llx", (unsigned long long)bits);
                line(std::string("setp.ne.b") + int_suffix(w) + " " + p + ", " + xr + ", " + imm + ";");
                line("@" + p + " bra " + skip + ";");
            } else {
                unsupported("branch on a selector that is neither an integer nor a bool");
            }
            parallel_move(*it->second, c.results, "branch case");
            line("bra " + end_label.at(c.target) + ";");
            put_label(skip);
        }
    }
    void block_body(const Block* b) {
        for (const Instr* ins : b->body) {
            comment_location(ins);
            if (pair && (ins->op == Op::Add || ins->op == Op::Load || ins->op == Op::Ext) && instruction_is_varying(ins)) {
                pair_instruction(ins);
                continue;
            }
            switch (ins->op) {
                case Op::Add: add(ins); break;
                case Op::Load: load(ins); break;
                case Op::Break: do_break(ins); break;
                case Op::Continue: do_continue(ins); break;
                case Op::Branch: branch(ins); break;
                case Op::BlockI: nested_block(ins->block, &ins->block->results); break;
                case Op::LoopI: loop(ins->loop); break;
                case Op::Ext: ext(ins); break;
            }
        }
        // a block is left only through `break` (block.rs:16-23); control reaching here is what the interpreter rejects
        line("trap;");
    }
    void define(const Value* v) {
        if (is_varying(v)) pair_def(v);
        else def(v);
    }
    void nested_block(const Block* b, const std::vector<Value*>* result_defs) {
        for (Value* r : *result_defs) define(r);
        std::string end = label(("end_" + sanitize(b->name)).c_str());
        end_label[b] = end;
        dests[b] = result_defs;
        line("// block " + sanitize(b->name));
        block_body(b);
        dests.erase(b);
        put_label(end);
    }
    void loop(const Loop* l) {
        if (l->arguments.size() != l->params.size()) throw Error("loop: argument count differs from parameter count");
        for (Value* p : l->params) define(p);
        parallel_move(l->params, l->arguments, "loop entry");
        std::string head = label(("head_" + sanitize(l->name)).c_str());
        head_label[l] = head;
        line("// loop " + sanitize(l->name));
        put_label(head);
        nested_block(l->body, &l->results);  // leaving the body block leaves the loop (block.rs:492-503)
        head_label.erase(l);
    }
    static std::string sanitize(const std::string& s) {
        std::string o;
        for (char ch : s) o += (isalnum((unsigned char)ch) || ch == '_') ? ch : '_';
        if (o.size() > 40) o.resize(40);
        return o.empty() ? "anon" : o;
    }

    std::string register_declarations() const {
        std::ostringstream o;
        const char* t = tag.c_str();
        if (n_pred) o << "    .reg .pred %p" << t << "<" << n_pred << ">;\n";
        if (n16) o << "    .reg .b16 %h" << t << "<" << n16 << ">;\n";
        if (n32) o << "    .reg .b32 %r" << t << "<" << n32 << ">;\n";
        if (n64) o << "    .reg .b64 %rd" << t << "<" << n64 << ">;\n";
        return o.str();
    }
};

void check_entry(const Module& m) {
    const Function* f = m.entry_point;
    if (!f) throw Error("module has no entry point");
    if (!f->body) throw Error("entry point has no body");
    if (!f->params.empty()) Emitter::unsupported("entry points with parameters");
    if (m.user_inputs.size_variable || m.user_outputs.size_variable) Emitter::unsupported("variable-size interface blocks");
    const auto& outs = m.user_outputs.members;
    const auto& rets = f->body->results;
    if (!rets.empty() && rets.size() != outs.size())
        throw Error("entry point leaves " + std::to_string(rets.size()) + " values but user_outputs_block has " +
                    std::to_string(outs.size()) + " members");
}

}  // namespace

std::string ptx_mangle(const std::string& name) {
    // PTX identifiers: [a-zA-Z][a-zA-Z0-9_$]* — backend.rs:444-459 additionally allows '.', '-' and a leading '$'/'_'
    std::string o = "kzn_";
    for (unsigned char ch : name) {
        if (isalnum(ch) || ch == '_') {
            o += (char)ch;
        } else {
            char b[8];
            snprintf(b, sizeof b, "$%02X", ch);
            o += b;
        }
    }
    return o;
}

std::string emit_ptx(const Module& m, EntryInfo* info) {
    check_entry(m);
    const Function* f = m.entry_point;
    const auto& outs = m.user_outputs.members;
    const auto& rets = f->body->results;

    Emitter e(m, Mode::Kernel);
    std::string entry = ptx_mangle(f->name);
    // ---- what every pointer is: interface blocks are memory, variables are registers unless their address is needed ----
    int in_base, out_base;
    {
        Base b;
        b.kind = Base::UserIn;
        b.size = m.user_inputs.size;
        b.scalarized = false;
        in_base = e.add_base(b);
        b.kind = Base::UserOut;
        b.size = m.user_outputs.size;
        out_base = e.add_base(b);
    }
    if (m.user_inputs.pointer) e.prov[m.user_inputs.pointer] = Prov{in_base, 0};
    if (m.user_outputs.pointer) e.prov[m.user_outputs.pointer] = Prov{out_base, 0};
    std::vector<std::pair<const Variable*, int>> vars;
    for (auto* list : {&m.built_in_inputs, &m.built_in_outputs, &m.globals, &f->locals})
        for (auto& v : *list) vars.emplace_back(&v, e.declare_variable(v));
    e.analyze(f->body);
    // Built-in OUTPUT variables (vertex_position) leave the invocation after its user outputs: the output slot of an
    // invocation is [user_outputs_block | built-in outputs in declaration order, each at a multiple of 16 bytes], every
    // component written (0 where the shader wrote nothing).
    uint64_t out_slot_size = m.user_outputs.size;
    std::vector<std::pair<int, uint64_t>> built_in_outs;  // (base, offset in the slot)
    for (auto& v : vars) {
        Base& b = e.bases[v.second];
        bool is_out = false;
        for (auto& o : m.built_in_outputs) is_out = is_out || &o == v.first;
        if (!is_out) continue;
        if (v.first->type->kind != TypeKind::Vector || type_size(v.first->type->element) != 4) Emitter::unsupported("built-in outputs other than vectors of 32-bit components");
        out_slot_size = (out_slot_size + 15) / 16 * 16;
        built_in_outs.push_back({v.second, out_slot_size});
        for (uint64_t c = 0; c < v.first->type->length; ++c)
            if (!b.slots.count(4 * c)) b.slots[4 * c] = 4;
        out_slot_size += (type_size(v.first->type) + 15) / 16 * 16;
    }
    e.finish_analysis();

    // ---- prologue: linear invocation index over a 2-D folded grid ----
    std::string gid = e.fresh(RC::B64), t64 = e.fresh(RC::B64), count = e.fresh(RC::B64), first = e.fresh(RC::B64), local = e.fresh(RC::B64);
    std::string cx = e.fresh(RC::B32), cy = e.fresh(RC::B32), ncx = e.fresh(RC::B32), nt = e.fresh(RC::B32), tx = e.fresh(RC::B32);
    std::string oob = e.fresh(RC::Pred);
    e.line("mov.u32 " + cx + ", %ctaid.x;");
    e.line("mov.u32 " + cy + ", %ctaid.y;");
    e.line("mov.u32 " + ncx + ", %nctaid.x;");
    e.line("mov.u32 " + nt + ", %ntid.x;");
    e.line("mov.u32 " + tx + ", %tid.x;");
    e.line("mul.wide.u32 " + gid + ", " + cy + ", " + ncx + ";");
    e.line("cvt.u64.u32 " + t64 + ", " + cx + ";");
    e.line("add.u64 " + gid + ", " + gid + ", " + t64 + ";");
    e.line("cvt.u64.u32 " + t64 + ", " + nt + ";");
    e.line("mul.lo.u64 " + gid + ", " + gid + ", " + t64 + ";");
    e.line("cvt.u64.u32 " + t64 + ", " + tx + ";");
    e.line("add.u64 " + gid + ", " + gid + ", " + t64 + ";");
    e.line("ld.param.u64 " + count + ", [kzn_invocation_count];");
    e.line("setp.ge.u64 " + oob + ", " + gid + ", " + count + ";");
    e.line("@" + oob + " ret;");
    e.line("mov.b64 " + local + ", " + gid + ";");  // index within this launch (subgroups are cut from it)
    e.line("ld.param.u64 " + first + ", [kzn_first_invocation];");
    e.line("add.u64 " + gid + ", " + gid + ", " + first + ";");
    // interface block pointers of this invocation
    auto bind_block = [&](uint64_t size, const char* param, int base) -> std::string {
        std::string r = e.fresh(RC::B64);
        e.line(std::string("ld.param.u64 ") + r + ", [" + param + "];");
        if (size) {
            std::string o = e.fresh(RC::B64);
            e.line("mul.lo.u64 " + o + ", " + gid + ", " + std::to_string(size) + ";");
            e.line("add.u64 " + r + ", " + r + ", " + o + ";");
        }
        e.bases[base].address = r;
        return r;
    };
    bind_block(m.user_inputs.size, "kzn_user_inputs_base", in_base);
    std::string out_ptr = bind_block(out_slot_size, "kzn_user_outputs_base", out_base);
    e.pc_param = "kzn_push_constants";
    if (e.pc_base >= 0) e.bases[e.pc_base].address = "kzn_push_constants";
    for (auto& v : vars) e.variable(*v.first, v.second);
    // built-in inputs (version 2)
    for (auto& v : vars) {
        Base& b = e.bases[v.second];
        if (b.built_in.empty() || b.built_in == "vertex_position") continue;
        if (b.built_in == "global_invocation_id") {
            // x fastest: id = (g % sx, (g / sx) % sy, g / (sx * sy)) with sx, sy the global size of the dispatch in x and y
            std::string sx = e.fresh(RC::B64), sy = e.fresh(RC::B64), q = e.fresh(RC::B64), x = e.fresh(RC::B64), y = e.fresh(RC::B64), z = e.fresh(RC::B64);
            std::string x32 = e.fresh(RC::B32), y32 = e.fresh(RC::B32), z32 = e.fresh(RC::B32);
            e.line("ld.param.u64 " + sx + ", [kzn_global_size_x];");
            e.line("ld.param.u64 " + sy + ", [kzn_global_size_y];");
            e.line("rem.u64 " + x + ", " + gid + ", " + sx + ";");
            e.line("div.u64 " + q + ", " + gid + ", " + sx + ";");
            e.line("rem.u64 " + y + ", " + q + ", " + sy + ";");
            e.line("div.u64 " + z + ", " + q + ", " + sy + ";");
            e.line("cvt.u32.u64 " + x32 + ", " + x + ";");
            e.line("cvt.u32.u64 " + y32 + ", " + y + ";");
            e.line("cvt.u32.u64 " + z32 + ", " + z + ";");
            // vkCmdDispatchBase: the first workgroup's id times LocalSize
            std::string bx = e.fresh(RC::B32), by = e.fresh(RC::B32), bz = e.fresh(RC::B32);
            e.line("ld.param.u32 " + bx + ", [kzn_global_base];");
            e.line("ld.param.u32 " + by + ", [kzn_global_base+4];");
            e.line("ld.param.u32 " + bz + ", [kzn_global_base+8];");
            e.line("add.u32 " + x32 + ", " + x32 + ", " + bx + ";");
            e.line("add.u32 " + y32 + ", " + y32 + ", " + by + ";");
            e.line("add.u32 " + z32 + ", " + z32 + ", " + bz + ";");
            e.set_built_in(b, 0, 4, x32);
            e.set_built_in(b, 4, 4, y32);
            e.set_built_in(b, 8, 4, z32);
        } else if (b.built_in == "subgroup_local_invocation_id") {
            std::string l = e.fresh(RC::B32);
            e.line("mov.u32 " + l + ", %laneid;");
            e.set_built_in(b, 0, 4, l);
        } else if (b.built_in == "subgroup_size") {
            e.set_built_in(b, 0, 4, "32");
        } else if (b.built_in == "descriptor_table") {
            std::string r = e.fresh(RC::B64);
            e.line("ld.param.u64 " + r + ", [kzn_descriptor_table];");
            e.set_built_in(b, 0, 8, r);
        } else if (b.built_in == "push_constants") {
            if (!b.scalarized) Emitter::unsupported("the push_constants variable may only be loaded from");
            for (auto& s : b.slots)
                if (s.first != 0 || s.second != 8) Emitter::unsupported("the push_constants variable holds one data_ptr");
        } else if (b.built_in == "frag_coord") {
            Emitter::unsupported("frag_coord in a compute entry point");
        }
    }
    std::string body_prologue = e.body.str();
    e.body.str("");

    e.nested_block(f->body, &f->body->results);
    // epilogue: the body's results are the invocation's outputs
    for (size_t i = 0; i < rets.size(); ++i) {
        const Type* t = outs[i].type;
        if (Emitter::reg_class(t) != Emitter::reg_class(rets[i]->type) || Emitter::scalar_bytes(t) != Emitter::scalar_bytes(rets[i]->type))
            throw Error("entry point result " + std::to_string(i) + " does not match the type of its user_outputs_block member");
        uint64_t align = m.user_outputs.align ? m.user_outputs.align : 1;
        uint64_t low = outs[i].offset ? (outs[i].offset & (~outs[i].offset + 1)) : align;
        uint64_t eff = low < align ? low : align;
        if (m.user_outputs.size % align) eff = 1;
        e.store_scalar(t, e.def(rets[i]), out_ptr, (int64_t)outs[i].offset, eff);
    }
    for (auto& bo : built_in_outs) {
        Base& b = e.bases[bo.first];
        for (uint64_t c = 0; c * 4 < b.size; ++c) {
            std::string v;
            if (b.scalarized) {
                v = e.cell(b, 4 * c);
            } else {
                v = e.fresh(RC::B32);
                e.line("ld.u32 " + v + ", [" + Emitter::address(b.address, (int64_t)(4 * c)) + "];");
            }
            e.line("st.u32 [" + Emitter::address(out_ptr, (int64_t)(bo.second + 4 * c)) + "], " + v + ";");
        }
    }
    e.line("ret;");

    std::ostringstream sub;
    if (e.uses_subgroup) {
        // member mask of the subgroup operations: the lanes of this warp that are inside the launch (lanes 0 .. n-1)
        std::string l = e.fresh(RC::B32), l64 = e.fresh(RC::B64), base = e.fresh(RC::B64), left = e.fresh(RC::B64), n = e.fresh(RC::B32), full = e.fresh(RC::Pred);
        sub << "    mov.u32 " << l << ", %laneid;\n    cvt.u64.u32 " << l64 << ", " << l << ";\n    sub.u64 " << base << ", " << local << ", " << l64 << ";\n"
            << "    sub.u64 " << left << ", " << count << ", " << base << ";\n    min.u64 " << left << ", " << left << ", 32;\n    cvt.u32.u64 " << n << ", " << left << ";\n"
            << "    setp.ge.u32 " << full << ", " << n << ", 32;\n    shl.b32 %kzn_submask, 1, " << n << ";\n    sub.u32 %kzn_submask, %kzn_submask, 1;\n"
            << "    @" << full << " mov.b32 %kzn_submask, 0xffffffff;\n";
    }

    std::ostringstream o;
    o << "//\n// Generated by kazan_b200 shader-compiler-backend-ptx from shader-compiler-ir\n//\n";
    o << ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
    o << ".visible .entry " << entry
      << "(\n    .param .u64 kzn_user_inputs_base,\n    .param .u64 kzn_user_outputs_base,\n"
         "    .param .u64 kzn_invocation_count,\n    .param .u64 kzn_first_invocation,\n"
         "    .param .u64 kzn_global_size_x,\n    .param .u64 kzn_global_size_y,\n    .param .u64 kzn_descriptor_table,\n"
         "    .param .align 16 .b8 kzn_global_base[16],\n    .param .align 16 .b8 kzn_push_constants[128]\n)\n{\n";
    o << e.register_declarations();
    if (e.uses_subgroup) o << "    .reg .b32 %kzn_submask;\n";
    o << e.locals.str() << "\n" << body_prologue << sub.str() << e.prologue.str() << e.body.str() << "}\n";
    if (info) {
        info->entry_name = entry;
        info->inputs_size = m.user_inputs.size;
        info->outputs_size = out_slot_size;
        info->uses_global_invocation_id = false;
        for (auto& v : vars)
            if (e.bases[v.second].built_in == "global_invocation_id") info->uses_global_invocation_id = true;
    }
    return o.str();
}

// ---- fragment shaders ------------------------------------------------------------------------------------------------

FragmentInterface fragment_interface(const Module& m) {
    check_entry(m);
    FragmentInterface fi;
    if (m.user_inputs.size > 16) Emitter::unsupported("fragment shaders with more than four 32-bit interpolated inputs (the rasteriser interpolates one vec4)");
    for (auto& mem : m.user_inputs.members)
        if (mem.type->kind != TypeKind::Float || float_bits(mem.type->float_kind) != 32 || mem.offset % 4)
            Emitter::unsupported("interpolated fragment inputs other than f32 at 4-byte offsets");
    if (m.user_outputs.size != 16 || m.user_outputs.members.size() != 4) Emitter::unsupported("fragment shaders must write one vec4 colour (four f32 members of user_outputs_block)");
    for (size_t i = 0; i < 4; ++i)
        if (m.user_outputs.members[i].type->kind != TypeKind::Float || float_bits(m.user_outputs.members[i].type->float_kind) != 32 || m.user_outputs.members[i].offset != 4 * i)
            Emitter::unsupported("fragment output members must be f32 at offsets 0, 4, 8, 12");
    // one dry run tells what the shader reads
    FragmentSite site;
    site.id = 0;
    site.in.assign(10, "0");
    site.out = {"%o0", "%o1", "%o2", "%o3"};
    FragmentInterface probe;
    emit_fragment_inline(m, site, &probe);
    fi = probe;
    return fi;
}

namespace {

// one copy of the entry point for `site` (site.fragments fragments); returns false — nothing emitted — when two fragments
// were asked for and the code cannot be shared
bool fragment_copy(const Module& m, const FragmentSite& site, const std::string& tag, FragmentInterface* info, std::string* text) {
    const Function* f = m.entry_point;
    const int nf = site.fragments;
    Emitter e(m, Mode::Inline, tag);
    int in_base, out_base, frag_coord_base = -1;
    {
        Base b;
        b.kind = Base::UserIn;
        b.size = 16;
        in_base = e.add_base(b);
        b.kind = Base::UserOut;
        b.size = 16;
        out_base = e.add_base(b);
    }
    if (m.user_inputs.pointer) e.prov[m.user_inputs.pointer] = Prov{in_base, 0};
    if (m.user_outputs.pointer) e.prov[m.user_outputs.pointer] = Prov{out_base, 0};
    std::vector<std::pair<const Variable*, int>> vars;
    for (auto* list : {&m.built_in_inputs, &m.built_in_outputs, &m.globals, &f->locals})
        for (auto& v : *list) vars.emplace_back(&v, e.declare_variable(v));
    for (auto& v : vars)
        if (e.bases[v.second].built_in == "frag_coord") frag_coord_base = v.second;
    e.analyze(f->body);
    e.finish_analysis();
    if (!e.bases[in_base].scalarized || !e.bases[out_base].scalarized)
        Emitter::unsupported("fragment shaders must access their interface blocks as pointer + constant (the blocks live in registers)");
    for (auto& s : e.bases[in_base].slots)
        if (s.second != 4 || s.first % 4) Emitter::unsupported("interpolated inputs are read as 32-bit values");
    for (auto& s : e.bases[out_base].slots)
        if (s.second != 4 || s.first % 4) Emitter::unsupported("fragment outputs are written as 32-bit values");
    if (frag_coord_base >= 0 && !e.bases[frag_coord_base].scalarized && nf == 2) return false;
    if (nf == 2) {
        e.pair = true;
        if (!e.analyze_pair(f, in_base, frag_coord_base)) return false;
    }
    if (info) {
        info->reads_inputs = e.reads_user_inputs;
        info->reads_frag_coord = e.reads_frag_coord;
        info->reads_frag_coord_zw = e.reads_frag_coord_zw;
        for (auto& b : e.bases)  // a frag_coord whose address is taken is read in ways the analysis cannot see
            if (b.built_in == "frag_coord" && !b.scalarized) info->reads_frag_coord_zw = true;
    }

    // inputs of fragment k: site.in[8k .. 8k+7] = frag_coord x y z w, the four interpolated inputs
    for (int k = 0; k < nf; ++k)
        for (auto& s : e.bases[in_base].slots) {
            std::string r = e.fresh(RC::B32);
            e.line("mov.b32 " + r + ", " + site.in[8 * k + 4 + s.first / 4] + ";");
            (k ? e.bases[in_base].cells_b : e.bases[in_base].cells)[s.first] = r;
        }
    e.pc_param = site.push_constants_param;
    if (e.pc_base >= 0 && e.pc_param.empty()) e.bases[e.pc_base].address = e.in_register(site.in[8 * nf], RC::B64);
    for (auto& v : vars) e.variable(*v.first, v.second);
    for (auto& v : vars) {
        Base& b = e.bases[v.second];
        if (b.built_in.empty() || b.built_in == "vertex_position") continue;
        if (b.built_in == "frag_coord") {
            for (int c = 0; c < 4; ++c) e.set_built_in(b, 4 * c, 4, site.in[c]);
            if (nf == 2) {
                std::swap(b.cells, b.cells_b);
                for (int c = 0; c < 4; ++c) e.set_built_in(b, 4 * c, 4, site.in[8 + c]);
                std::swap(b.cells, b.cells_b);
            }
        } else if (b.built_in == "descriptor_table") {
            e.set_built_in(b, 0, 8, site.in[8 * nf + 1]);
            if (nf == 2 && b.scalarized && b.cells.count(0)) b.cells_b[0] = b.cells[0];
        } else if (b.built_in == "push_constants") {
            if (!b.scalarized) Emitter::unsupported("the push_constants variable may only be loaded from");
        } else {
            Emitter::unsupported("built-in '" + b.built_in + "' in a fragment shader");
        }
    }
    std::string body_prologue = e.body.str();
    e.body.str("");
    e.nested_block(f->body, &f->body->results);
    // outputs: the cells of user_outputs_block, overridden by the body's results when it returns values
    const auto& rets = f->body->results;
    for (int c = 0; c < 4; ++c) {
        if (c < (int)rets.size()) {
            if (Emitter::reg_class(rets[c]->type) != RC::B32) throw Error("fragment shader result is not a 32-bit value");
            if (nf == 2) {
                auto v = e.scalars_of(rets[c]);
                if (!site.out[c].empty()) e.line("mov.b32 " + site.out[c] + ", " + v.first + ";");
                if (!site.out[4 + c].empty()) e.line("mov.b32 " + site.out[4 + c] + ", " + v.second + ";");
            } else if (!site.out[c].empty()) {
                e.line("mov.b32 " + site.out[c] + ", " + e.def(rets[c]) + ";");
            }
        } else {
            Base& ob = e.bases[out_base];
            for (int k = 0; k < nf; ++k) {
                if (site.out[4 * k + c].empty()) continue;
                auto& cells = k ? ob.cells_b : ob.cells;
                auto it = cells.find(4 * c);
                e.line("mov.b32 " + site.out[4 * k + c] + ", " + (it != cells.end() ? it->second : std::string("0")) + ";");
            }
        }
    }
    std::ostringstream o;
    o << "    {\n" << e.register_declarations() << e.locals.str() << e.prologue.str() << body_prologue << e.body.str() << "    }\n";
    *text = o.str();
    return true;
}

}  // namespace

std::string emit_fragment_inline(const Module& m, const FragmentSite& site, FragmentInterface* info) {
    check_entry(m);
    const int nf = site.fragments;
    if ((nf != 1 && nf != 2) || (int)site.in.size() != 8 * nf + 2 || (int)site.out.size() != 4 * nf)
        throw Error("internal: a fragment call site has 8 inputs and 4 outputs per fragment, and two pointers");
    const std::string tag = "_fs" + std::to_string(site.id) + "_";
    std::string text;
    if (fragment_copy(m, site, tag, info, &text)) {
        if (info) info->pairs_share_code = nf == 2;
        return text;
    }
    // two fragments whose code cannot be shared: one copy each
    std::string both;
    for (int k = 0; k < 2; ++k) {
        FragmentSite one;
        one.id = site.id;
        one.fragments = 1;
        one.in.assign(site.in.begin() + 8 * k, site.in.begin() + 8 * k + 8);
        one.in.push_back(site.in[16]);
        one.in.push_back(site.in[17]);
        one.out.assign(site.out.begin() + 4 * k, site.out.begin() + 4 * k + 4);
        one.push_constants_param = site.push_constants_param;
        std::string t;
        fragment_copy(m, one, tag + (k ? "b_" : "a_"), k == 0 ? info : nullptr, &t);
        both += t;
    }
    if (info) info->pairs_share_code = false;
    return both;
}

}  // namespace kir
