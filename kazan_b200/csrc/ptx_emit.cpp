// ptx_emit.cpp — shader-compiler-ir -> PTX for sm_100a: the code generator of `shader-compiler-backend-ptx`.
//
// Replaces, for the B200, what shader-compiler-backend-llvm-7/src/backend.rs:553-690 (`LLVM7Compiler::run`) is for the
// host CPU.  Input is the structured SSA of shader-compiler-ir (instructions_impl.rs:395-433: add, load, break, branch,
// continue, block, loop); output is one `.visible .entry` per module, one invocation per thread (BASELINE north_star).
//
// Invocation ABI (include/kazan_b200_backend.h): parameters (user_inputs_base, user_outputs_base, invocation_count,
// first_invocation).  Thread g of the grid (2-D folded) runs invocation first+g; its interface blocks
// (interface.rs:199-252) are base + (first+g) * block.size.  The values the entry point's body block is left with
// (`break <body>[v...]`, block.rs:16-23) are stored to the members of user_outputs_block, in order.
//
// Lowering:
//   value                    -> one virtual register of its class (.pred bool, .b16 i8/i16/f16, .b32, .b64)
//   block name -> [r..] {..} -> fall-through region closed by a label; `break` = moves into r.. + bra to the label
//   loop                     -> moves arguments into the parameters, head label, body block; `continue` = parallel move + bra
//   branch                   -> setp.eq chain, each case a predicated skip around (moves + bra)
//   load, align A            -> one ld of natural width when A >= size, otherwise A-sized pieces combined with shl/or
//   variables (function.rs:155-193, module.rs:145-147, interface.rs:61-67) -> .local arrays, address through cvta.local
#include <cstdio>
#include <map>
#include <set>
#include <sstream>

#include "ptx_emit.h"

namespace kir {
namespace {

enum class RC { Pred, B16, B32, B64 };

struct Emitter {
    const Module& m;
    std::ostringstream body;
    std::map<const Value*, std::string> reg;
    std::map<const Block*, const std::vector<Value*>*> dests;  // where `break block[...]` writes
    std::map<const Block*, std::string> end_label;
    std::map<const Loop*, std::string> head_label;
    int n_pred = 0, n16 = 0, n32 = 0, n64 = 0, n_label = 0, n_var = 0;
    std::ostringstream locals;

    explicit Emitter(const Module& mod) : m(mod) {}

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
    std::string fresh(RC c) {
        char b[32];
        switch (c) {
            case RC::Pred: snprintf(b, sizeof b, "%%p%d", n_pred++); break;
            case RC::B16: snprintf(b, sizeof b, "%%h%d", n16++); break;
            case RC::B32: snprintf(b, sizeof b, "%%r%d", n32++); break;
            case RC::B64: snprintf(b, sizeof b, "%%rd%d", n64++); break;
        }
        return b;
    }
    static const char* mov_type(RC c) { return c == RC::Pred ? "pred" : c == RC::B16 ? "b16" : c == RC::B32 ? "b32" : "b64"; }
    std::string label(const char* stem) {
        char b[48];
        snprintf(b, sizeof b, "$%s_%d", stem, n_label++);
        return b;
    }
    void line(const std::string& s) { body << "    " << s << "\n"; }
    void put_label(const std::string& l) { body << l << ":\n"; }

    // register that holds a definition
    const std::string& def(const Value* v) {
        auto it = reg.find(v);
        if (it != reg.end()) return it->second;
        return reg.emplace(v, fresh(reg_class(v->type))).first->second;
    }
    // operand text for a use: registers for definitions, immediates for inline constants (values.rs:168-200)
    std::string use(const Value* v) {
        if (!v->has_const) {
            auto it = reg.find(v);
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
    void move(const Value* dst, const std::string& src) {
        RC c = reg_class(dst->type);
        line(std::string("mov.") + mov_type(c) + " " + def(dst) + ", " + src + ";");
    }
    // simultaneous assignment dsts <- srcs (continue re-binds loop parameters that may appear among the sources)
    void parallel_move(const std::vector<Value*>& dsts, const std::vector<Value*>& srcs, const char* what) {
        if (dsts.size() != srcs.size()) throw Error(std::string(what) + ": value count does not match the target's definition count");
        std::vector<std::string> ops;
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

    // ---- memory -------------------------------------------------------------------------------------------------
    // dst register <- size bytes at [addr], the access split into `piece`-byte naturally aligned loads
    void load_scalar(const Value* result, const std::string& addr, uint64_t align) {
        const Type* t = result->type;
        int size = scalar_bytes(t);
        RC rc = reg_class(t);
        if (align == 0) align = 1;
        uint64_t piece = 1;
        while (piece * 2 <= align && piece * 2 <= (uint64_t)size) piece *= 2;
        std::string value;
        if (piece == (uint64_t)size) {
            value = rc == RC::Pred ? fresh(RC::B16) : def(result);
            line("ld.u" + std::to_string(size * 8) + " " + value + ", [" + addr + "];");
        } else {
            // little-endian assembly in a 64-bit temporary
            std::string acc = fresh(RC::B64), part = fresh(RC::B64);
            for (int off = 0; off < size; off += (int)piece) {
                std::string dstp = off == 0 ? acc : part;
                line("ld.u" + std::to_string(piece * 8) + " " + dstp + ", [" + addr + "+" + std::to_string(off) + "];");
                if (off) {
                    line("shl.b64 " + part + ", " + part + ", " + std::to_string(off * 8) + ";");
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
    void store_scalar(const Type* t, const std::string& src_operand, const std::string& addr, uint64_t offset, uint64_t align) {
        int size = scalar_bytes(t);
        RC rc = reg_class(t);
        std::string src = src_operand;
        if (rc == RC::Pred) {
            std::string h = fresh(RC::B16);
            line("selp.u16 " + h + ", 1, 0, " + src + ";");
            src = h;
        } else if (src.rfind("0x", 0) == 0 || src == "0") {  // immediates cannot be store sources
            std::string r = fresh(rc);
            line(std::string("mov.") + mov_type(rc) + " " + r + ", " + src + ";");
            src = r;
        }
        if (align == 0) align = 1;
        uint64_t piece = 1;
        while (piece * 2 <= align && piece * 2 <= (uint64_t)size) piece *= 2;
        if (piece == (uint64_t)size) {
            line("st.u" + std::to_string(size * 8) + " [" + addr + "+" + std::to_string(offset) + "], " + src + ";");
            return;
        }
        std::string wide = fresh(RC::B64), part = fresh(RC::B64);
        line(std::string("cvt.u64.u") + (rc == RC::B64 ? "64 " : rc == RC::B32 ? "32 " : "16 ") + wide + ", " + src + ";");
        for (int off = 0; off < size; off += (int)piece) {
            line("shr.u64 " + part + ", " + wide + ", " + std::to_string(off * 8) + ";");
            line("st.u" + std::to_string(piece * 8) + " [" + addr + "+" + std::to_string(offset + off) + "], " + part + ";");
        }
    }
    void variable(const Variable& v) {
        if (!v.pointer) return;
        uint64_t size = type_size(v.type), align = v.align ? v.align : 1;
        if (align & (align - 1)) throw Error("variable alignment is not a power of two");
        std::string sym = "__kzn_var" + std::to_string(n_var++);
        locals << "    .local .align " << align << " .b8 " << sym << "[" << (size ? size : 1) << "];\n";
        const std::string& r = def(v.pointer);
        line("mov.u64 " + r + ", " + sym + ";");
        line("cvta.local.u64 " + r + ", " + r + ";");
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
    void add(const Instr* ins) {
        if (ins->args.size() != 2 || ins->results.size() != 1) throw Error("add takes two arguments and one result");
        const Type* t = ins->results[0]->type;
        if (t->kind != TypeKind::Int) unsupported("add on non-integer types (the reference defines no float semantics, SURVEY.md §8 a12)");
        if (ins->args[0]->type != t || ins->args[1]->type != t) throw Error("add: operand types differ from the result type");
        std::string a = use(ins->args[0]), b = use(ins->args[1]);
        if (a.rfind("0x", 0) == 0) {  // first source operand of add must be a register
            RC c = reg_class(t);
            std::string r = fresh(c);
            line(std::string("mov.") + mov_type(c) + " " + r + ", " + a + ";");
            a = r;
        }
        const std::string& d = def(ins->results[0]);
        int w = int_bits(t->int_kind);
        line(std::string("add.s") + (w <= 16 ? "16 " : w == 32 ? "32 " : "64 ") + d + ", " + a + ", " + b + ";");
        if (w == 8) line("and.b16 " + d + ", " + d + ", 0xff;");  // i8 lives zero-extended in a 16-bit register
    }
    void load(const Instr* ins) {
        if (ins->args.size() != 1 || ins->results.size() != 1) throw Error("load takes one pointer and one result");
        if (ins->args[0]->type->kind != TypeKind::DataPtr) throw Error("load: the argument is not a data_ptr");
        std::string addr = use(ins->args[0]);
        if (addr.rfind("0x", 0) == 0 || addr == "0") {
            std::string r = fresh(RC::B64);
            line("mov.b64 " + r + ", " + addr + ";");
            addr = r;
        }
        load_scalar(ins->results[0], addr, ins->align);
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
                std::string xr = x;
                if (xr.rfind("0x", 0) == 0) {
                    RC rc = reg_class(t);
                    xr = fresh(rc);
                    line(std::string("mov.") + mov_type(rc) + " " + xr + ", " + x + ";");
                }
                std::string p = fresh(RC::Pred);
                char imm[32];
                snprintf(imm, sizeof imm, "0x%llx", (unsigned long long)bits);
                line(std::string("setp.ne.b") + (w <= 16 ? "16 " : w == 32 ? "32 " : "64 ") + p + ", " + xr + ", " + imm + ";");
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
            switch (ins->op) {
                case Op::Add: add(ins); break;
                case Op::Load: load(ins); break;
                case Op::Break: do_break(ins); break;
                case Op::Continue: do_continue(ins); break;
                case Op::Branch: branch(ins); break;
                case Op::BlockI: nested_block(ins->block, &ins->block->results); break;
                case Op::LoopI: loop(ins->loop); break;
            }
        }
        // a block is left only through `break` (block.rs:16-23); control reaching here is what the interpreter rejects
        line("trap;");
    }
    void nested_block(const Block* b, const std::vector<Value*>* result_defs) {
        for (Value* r : *result_defs) def(r);
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
        for (Value* p : l->params) def(p);
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
        return o.empty() ? "anon" : o;
    }
};

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

    Emitter e(m);
    std::string entry = ptx_mangle(f->name);
    // prologue: linear invocation index over a 2-D folded grid
    std::string gid = e.fresh(RC::B64), t64 = e.fresh(RC::B64), count = e.fresh(RC::B64), first = e.fresh(RC::B64);
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
    e.line("ld.param.u64 " + first + ", [kzn_first_invocation];");
    e.line("add.u64 " + gid + ", " + gid + ", " + first + ";");
    // interface block pointers of this invocation
    auto bind_block = [&](const InterfaceBlock& b, const char* param) -> std::string {
        std::string r = b.pointer ? e.def(b.pointer) : e.fresh(RC::B64);
        e.line(std::string("ld.param.u64 ") + r + ", [" + param + "];");
        if (b.size) {
            std::string o = e.fresh(RC::B64);
            e.line("mul.lo.u64 " + o + ", " + gid + ", " + std::to_string(b.size) + ";");
            e.line("add.u64 " + r + ", " + r + ", " + o + ";");
        }
        return r;
    };
    bind_block(m.user_inputs, "kzn_user_inputs_base");
    std::string out_ptr = bind_block(m.user_outputs, "kzn_user_outputs_base");
    for (auto& v : m.built_in_inputs) e.variable(v);
    for (auto& v : m.built_in_outputs) e.variable(v);
    for (auto& v : m.globals) e.variable(v);
    for (auto& v : f->locals) e.variable(v);

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
        e.store_scalar(t, e.def(rets[i]), out_ptr, outs[i].offset, eff);
    }
    e.line("ret;");

    std::ostringstream o;
    o << "//\n// Generated by kazan_b200 shader-compiler-backend-ptx from shader-compiler-ir\n//\n";
    o << ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
    o << ".visible .entry " << entry
      << "(\n    .param .u64 kzn_user_inputs_base,\n    .param .u64 kzn_user_outputs_base,\n"
         "    .param .u64 kzn_invocation_count,\n    .param .u64 kzn_first_invocation\n)\n{\n";
    if (e.n_pred) o << "    .reg .pred %p<" << e.n_pred << ">;\n";
    if (e.n16) o << "    .reg .b16 %h<" << e.n16 << ">;\n";
    if (e.n32) o << "    .reg .b32 %r<" << e.n32 << ">;\n";
    if (e.n64) o << "    .reg .b64 %rd<" << e.n64 << ">;\n";
    o << e.locals.str() << "\n" << e.body.str() << "}\n";
    if (info) {
        info->entry_name = entry;
        info->inputs_size = m.user_inputs.size;
        info->outputs_size = m.user_outputs.size;
    }
    return o.str();
}

}  // namespace kir
