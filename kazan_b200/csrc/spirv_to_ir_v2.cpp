// spirv_to_ir_v2.cpp — SPIR-V -> shader-compiler-ir (version 2 of the IR, ir.h) for shaders with a body.
//
// The reference's translator (shader-compiler-translate-spirv-to-ir) stops at the first instruction that needs an IR
// instruction it does not have: OpLoad parses and then `todo!()`s, OpStore / OpAccessChain / every composite, conversion and
// branch instruction are `todo!()` (parse/memory_instructions.rs:99-141, parse/unimplemented_instructions.rs:14-39,
// parse/translate_structure_tree.rs:258-291), and its golden for the second test shader is the placeholder text
// "not yet entered" (src/lib.rs:260-287, test_data/simple_test.kazan-ir).  This file is what SURVEY.md §8f row 2 asks for:
// those instructions translated into the version 2 instructions of the C++ IR mirror.  It follows the reference where the
// reference has decided something:
//   names            OpName or `id_<n>` (parse/debug_names.rs:25-31); a function's body block is named after its entry label
//   interface blocks 16 bytes per Location, 4 per Component (io_layout.rs:5-13), both blocks 16-byte aligned
//   built-ins        a member decorated BuiltIn Position becomes the built-in output variable `vertex_position`
//                    (shader-compiler-ir/src/interface.rs:61-67), <4 x f32>
//   structure        structured control flow becomes nested blocks that are left by `break` (docs/Shader Compiler IR
//                    Example.md): a selection construct `OpSelectionMerge %m; OpSwitch / OpBranchConditional` is
//                        block m { block t1 { block t0 { branch [selector], { case -> break t_k[]; ... } -> []; break default[]; };
//                                             <code of t0> };  <code of t1> };  <code of m>
//                    with the targets nested in the order their labels appear, so that leaving block t jumps to label t.
// Values are scalarised, like everything the PTX back end lowers: a SPIR-V vector is a list of scalar IR values, OpLoad /
// OpStore of a composite are one access per scalar, OpAccessChain with constant indices is `ptr_add` with a constant.
// Loops: `OpLoopMerge %merge %continue` becomes the IR's `loop` (see loop_construct): header OpPhis are the loop's parameters,
// the back edge `continue`, a jump to %merge a `break` out of the loop's body block, a jump to %continue a `break` out of a
// block that the continue block's code follows; conditional branches without a selection merge (the `while` test, `do ..
// while`, conditional break / continue) nest their targets in a block that nothing follows.
// Not translated (KZN_ERROR_FEATURE_NOT_PRESENT, by opcode): function calls and parameters, images, matrices, dynamic
// indexing, values defined inside a loop and used after it other than through OpPhi or variables.
#include <algorithm>
#include <functional>
#include <map>
#include <set>

#include "ptx_emit.h"

namespace kir {
namespace {

constexpr uint32_t kMagic = 0x07230203;

enum : uint32_t {
    OpNop = 0, OpUndef = 1, OpSourceContinued = 2, OpSource = 3, OpSourceExtension = 4, OpName = 5, OpMemberName = 6, OpString = 7, OpLine = 8,
    OpExtension = 10, OpExtInstImport = 11, OpExtInst = 12, OpMemoryModel = 14, OpEntryPoint = 15, OpExecutionMode = 16, OpCapability = 17,
    OpTypeVoid = 19, OpTypeBool = 20, OpTypeInt = 21, OpTypeFloat = 22, OpTypeVector = 23, OpTypeArray = 28, OpTypeStruct = 30, OpTypePointer = 32,
    OpTypeFunction = 33, OpConstantTrue = 41, OpConstantFalse = 42, OpConstant = 43, OpConstantComposite = 44, OpFunction = 54,
    OpFunctionParameter = 55, OpFunctionEnd = 56, OpFunctionCall = 57, OpVariable = 59, OpLoad = 61, OpStore = 62, OpAccessChain = 65, OpInBoundsAccessChain = 66,
    OpDecorate = 71, OpMemberDecorate = 72, OpCompositeConstruct = 80, OpCompositeExtract = 81, OpConvertFToU = 109, OpConvertFToS = 110,
    OpConvertSToF = 111, OpConvertUToF = 112, OpUConvert = 113, OpSConvert = 114, OpFConvert = 115, OpBitcast = 124, OpSNegate = 126, OpFNegate = 127,
    OpIAdd = 128, OpFAdd = 129, OpISub = 130, OpFSub = 131, OpIMul = 132, OpFMul = 133, OpFDiv = 136, OpLogicalOr = 166, OpLogicalAnd = 167,
    OpLogicalNot = 168, OpSelect = 169, OpIEqual = 170, OpINotEqual = 171, OpUGreaterThan = 172, OpSGreaterThan = 173, OpUGreaterThanEqual = 174,
    OpSGreaterThanEqual = 175, OpULessThan = 176, OpSLessThan = 177, OpULessThanEqual = 178, OpSLessThanEqual = 179, OpFOrdEqual = 180,
    OpFUnordNotEqual = 183, OpFOrdLessThan = 184, OpFOrdGreaterThan = 186, OpFOrdLessThanEqual = 188, OpFOrdGreaterThanEqual = 190,
    OpShiftRightLogical = 194, OpShiftRightArithmetic = 195, OpShiftLeftLogical = 196, OpBitwiseOr = 197, OpBitwiseXor = 198, OpBitwiseAnd = 199,
    OpNot = 200, OpPhi = 245, OpLoopMerge = 246, OpSelectionMerge = 247, OpLabel = 248, OpBranch = 249, OpBranchConditional = 250, OpSwitch = 251,
    OpKill = 252, OpReturn = 253, OpReturnValue = 254, OpUnreachable = 255, OpNoLine = 317, OpModuleProcessed = 330,
};
enum : uint32_t { DecBlock = 2, DecBuiltIn = 11, DecLocation = 30, DecComponent = 31, DecOffset = 35 };
enum : uint32_t { ScInput = 1, ScOutput = 3, ScPrivate = 6, ScFunction = 7 };
enum : uint32_t { BuiltInPosition = 0 };

std::string literal_string(const uint32_t* w, size_t n) {
    std::string s;
    for (size_t i = 0; i < n; ++i)
        for (int b = 0; b < 4; ++b) {
            char ch = (char)((w[i] >> (8 * b)) & 0xff);
            if (!ch) return s;
            s += ch;
        }
    throw Error("SPIR-V: unterminated literal string");
}

struct SType {
    enum Kind { Void, Bool, Int, Float, Vector, Array, Struct, Pointer, Function } kind = Void;
    uint32_t bits = 0, count = 0, element = 0, storage = 0;
    bool is_signed = false;
    std::vector<uint32_t> members;
};

// what a SPIR-V id stands for while a function body is translated
struct SValue {
    std::vector<Value*> scalars;  // a scalar (1) or the components of a vector, in order
    // pointers: `base` + byte offset (constant indices only), of `pointee` type
    Value* base = nullptr;
    uint64_t offset = 0;
    uint32_t pointee = 0;
    bool is_pointer = false;
};

struct Inst {
    uint32_t op;
    std::vector<uint32_t> w;  // operands (without the first word)
    Location loc;
};
struct BasicBlock {
    uint32_t label = 0;
    size_t order = 0;  // position of the label in the function
    std::vector<Inst> body;
    Inst terminator;
    uint32_t merge = 0;  // OpSelectionMerge target, 0 = none
    bool loop = false;   // OpLoopMerge: this block is a loop header
    uint32_t loop_merge = 0, loop_continue = 0;
};

struct FunctionBody {  // one OpFunction .. OpFunctionEnd
    uint32_t id = 0, control = 0, return_type = 0;
    std::vector<uint32_t> params;  // OpFunctionParameter result ids
    std::vector<BasicBlock> blocks;
};

struct Translator {
    Module& m;
    uint32_t bound;
    std::map<uint32_t, const FunctionBody*> spv_functions;  // callable functions by id (OpFunctionCall inlines them)
    std::vector<uint32_t> call_stack;                    // functions being inlined right now (recursion is refused)
    std::map<uint32_t, std::string> names, strings;
    std::map<uint32_t, SType> types;
    std::map<uint32_t, std::map<uint32_t, uint32_t>> decorations;                         // id -> decoration -> literal
    std::map<uint32_t, std::map<uint32_t, std::map<uint32_t, uint32_t>>> member_decorations;  // struct id -> member -> decoration -> literal
    std::map<uint32_t, SValue> values;                                                    // constants and module-scope variables, then function-local ids
    std::map<uint32_t, Function*> functions;
    int n_const = 0;

    explicit Translator(Module& mod, uint32_t b) : m(mod), bound(b) {}

    [[noreturn]] static void unsupported(const std::string& what) { throw Unsupported("SPIR-V -> IR: " + what); }
    std::string name_of(uint32_t id) const {
        auto it = names.find(id);
        return it != names.end() && !it->second.empty() ? it->second : "id_" + std::to_string(id);
    }
    const SType& type(uint32_t id) const {
        auto it = types.find(id);
        if (it == types.end()) throw Error("SPIR-V: id " + std::to_string(id) + " is not a type");
        return it->second;
    }
    const Type* scalar_ir_type(const SType& t) {
        switch (t.kind) {
            case SType::Bool: return m.bool_type();
            case SType::Int: return m.int_type(t.bits == 8 ? IntKind::I8 : t.bits == 16 ? IntKind::I16 : t.bits == 64 ? IntKind::I64 : IntKind::I32);
            case SType::Float:
                if (t.bits == 16) unsupported("16-bit floats");
                return m.float_type(t.bits == 64 ? FloatKind::F64 : FloatKind::F32);
            case SType::Pointer: return m.data_ptr();
            default: unsupported("a composite type where a scalar is needed");
        }
    }
    // scalar leaves of a type with their byte offsets (tightly packed scalars; struct members at their Offset decoration or
    // packed; every scalar naturally aligned)
    void leaves(uint32_t type_id, uint64_t at, std::vector<std::pair<uint64_t, const Type*>>* out) {
        const SType& t = type(type_id);
        if (out->size() > 65536) unsupported("a value of more than 65536 scalars (nested arrays)");
        switch (t.kind) {
            case SType::Vector:
            case SType::Array: {
                const uint64_t esz = size_of(t.element);
                for (uint32_t i = 0; i < t.count; ++i) leaves(t.element, at + i * esz, out);
                break;
            }
            case SType::Struct: {
                uint64_t off = 0;
                for (size_t k = 0; k < t.members.size(); ++k) {
                    off = member_offset(type_id, (uint32_t)k);
                    leaves(t.members[k], at + off, out);
                }
                break;
            }
            default: out->push_back({at, scalar_ir_type(t)});
        }
    }
    uint64_t size_of(uint32_t type_id) {
        const SType& t = type(type_id);
        switch (t.kind) {
            case SType::Bool: return 1;
            case SType::Int:
            case SType::Float: return t.bits / 8;
            case SType::Pointer: return 8;
            case SType::Vector:
            case SType::Array: return size_of(t.element) * t.count;
            case SType::Struct: {
                uint64_t end = 0;
                for (size_t k = 0; k < t.members.size(); ++k) end = std::max(end, member_offset(type_id, (uint32_t)k) + size_of(t.members[k]));
                return end;
            }
            default: unsupported("the size of a void / function type");
        }
    }
    uint64_t align_of(uint32_t type_id) {
        const SType& t = type(type_id);
        if (t.kind == SType::Vector) return std::min<uint64_t>(16, size_of(t.element) * (t.count == 3 ? 4 : t.count));
        if (t.kind == SType::Array) return align_of(t.element);
        if (t.kind == SType::Struct) {
            uint64_t a = 1;
            for (uint32_t mem : t.members) a = std::max(a, align_of(mem));
            return a;
        }
        return std::max<uint64_t>(1, size_of(type_id));
    }
    uint64_t member_offset(uint32_t struct_id, uint32_t member) {
        auto sd = member_decorations.find(struct_id);
        if (sd != member_decorations.end()) {
            auto md = sd->second.find(member);
            if (md != sd->second.end() && md->second.count(DecOffset)) return md->second.at(DecOffset);
        }
        const SType& t = type(struct_id);
        uint64_t off = 0;
        for (uint32_t k = 0; k <= member; ++k) {
            const uint64_t a = align_of(t.members[k]);
            off = (off + a - 1) / a * a;
            if (k == member) return off;
            off += size_of(t.members[k]);
        }
        return off;
    }
    const Type* ir_type_of(uint32_t type_id) {  // IR type of a variable's pointee
        const SType& t = type(type_id);
        if (t.kind == SType::Vector) {
            Type v{TypeKind::Vector};
            v.length = t.count;
            v.element = scalar_ir_type(type(t.element));
            return m.intern(v);
        }
        if (t.kind == SType::Array || t.kind == SType::Struct) {
            Type s{TypeKind::Struct};
            s.size = size_of(type_id);
            s.align = align_of(type_id);
            std::vector<std::pair<uint64_t, const Type*>> ls;
            leaves(type_id, 0, &ls);
            for (auto& l : ls) s.members.push_back({l.first, l.second});
            return m.intern(s);
        }
        return scalar_ir_type(t);
    }
    Value* constant(const Type* t, uint64_t bits, ConstKind kind) {
        Value* v = m.new_value("c" + std::to_string(n_const++), t);
        v->has_const = true;
        v->constant.kind = kind;
        v->constant.type = t;
        v->constant.bits = bits;
        return v;
    }
};

struct FunctionTranslator {
    Translator& T;
    Module& m;
    Function* fn;
    std::map<uint32_t, BasicBlock> blocks;
    std::vector<uint32_t> order;
    std::map<uint32_t, SValue> values;      // function-local ids
    std::map<uint32_t, Block*> open;        // labels that are the exit of an IR block that is open right now
    std::map<uint32_t, std::vector<std::pair<uint32_t, uint32_t>>> phis;  // merge label -> [(result id, type id)] (+ operands kept in the Inst)
    std::map<uint32_t, std::vector<Inst>> phi_insts;
    std::map<uint32_t, Loop*> loop_open;    // headers of the loops that are open right now: a jump to one is a `continue`
    std::set<uint32_t> emitted;
    std::set<uint32_t> loops_built;

    FunctionTranslator(Translator& t, Function* f) : T(t), m(t.m), fn(f) {}

    void load_blocks(const FunctionBody& fb) {
        for (const BasicBlock& b : fb.blocks) {
            BasicBlock copy = b;
            std::vector<Inst> rest;  // OpPhis belong to the block they merge into
            for (const Inst& in : copy.body) {
                if (in.op == OpPhi) {
                    if (in.w.size() < 2) throw Error("SPIR-V: OpPhi is missing operands");
                    phi_insts[b.label].push_back(in);
                }
                else rest.push_back(in);
            }
            copy.body = rest;
            blocks[b.label] = copy;
            order.push_back(b.label);
        }
    }

    const SValue& value(uint32_t id) {
        auto it = values.find(id);
        if (it != values.end()) return it->second;
        auto g = T.values.find(id);
        if (g == T.values.end()) throw Error("SPIR-V: id " + std::to_string(id) + " is used before it is defined");
        return g->second;
    }
    Value* scalar(uint32_t id) {
        const SValue& v = value(id);
        if (v.is_pointer || v.scalars.size() != 1) Translator::unsupported("a composite or pointer operand where a scalar is needed (id " + std::to_string(id) + ")");
        return v.scalars[0];
    }
    Instr* push(Block* b, Instr* ins, const Location& loc) {
        ins->location = loc;
        b->body.push_back(ins);
        return ins;
    }
    Value* ext(Block* b, const char* op, std::vector<Value*> args, const Type* result_type, const std::string& name, const Location& loc, uint64_t align = 0) {
        Instr* ins = m.new_instr(Op::Ext);
        ins->ext = ext_op_index(op);
        if (ins->ext < 0) throw Error(std::string("internal: unknown version 2 instruction ") + op);
        ins->args = std::move(args);
        ins->align = align;
        Value* r = nullptr;
        if (result_type) {
            r = m.new_value(name, result_type);
            ins->results.push_back(r);
        }
        push(b, ins, loc);
        return r;
    }
    // the data_ptr value of `base + offset`
    Value* address(Block* b, const SValue& p, uint64_t extra, const std::string& name, const Location& loc) {
        const uint64_t off = p.offset + extra;
        if (off == 0) return p.base;
        return ext(b, "ptr_add", {p.base, T.constant(m.int_type(IntKind::I64), off, ConstKind::Int)}, m.data_ptr(), name, loc);
    }

    void instruction(Block* b, const Inst& in) {
        const auto& w = in.w;
        auto result_scalar_type = [&](uint32_t type_id) -> const Type* {
            const SType& t = T.type(type_id);
            if (t.kind == SType::Vector) return T.scalar_ir_type(T.type(t.element));
            return T.scalar_ir_type(t);
        };
        auto width = [&](uint32_t type_id) -> size_t {
            const SType& t = T.type(type_id);
            return t.kind == SType::Vector ? t.count : 1;
        };
        auto comp_name = [&](uint32_t id, size_t k, size_t n) { return n == 1 ? T.name_of(id) : T.name_of(id) + "_" + std::to_string(k); };
        // component-wise instruction `op` over the operands ids[...]
        auto elementwise = [&](const char* op, std::vector<uint32_t> ids, const Type* rt = nullptr) {
            const size_t n = width(w[0]);
            SValue r;
            for (size_t k = 0; k < n; ++k) {
                std::vector<Value*> args;
                for (uint32_t id : ids) {
                    const SValue& v = value(id);
                    if (v.is_pointer) Translator::unsupported("pointer operands of arithmetic");
                    args.push_back(v.scalars.size() == 1 ? v.scalars[0] : v.scalars.at(k));
                }
                r.scalars.push_back(ext(b, op, args, rt ? rt : result_scalar_type(w[0]), comp_name(w[1], k, n), in.loc));
            }
            values[w[1]] = r;
        };
        {   // operands every form of the instruction has (a truncated instruction must not be read past its end)
            size_t least = 0;
            switch (in.op) {
                case OpStore: case OpCompositeConstruct: least = 2; break;
                case OpVariable: case OpLoad: case OpAccessChain: case OpInBoundsAccessChain: case OpFunctionCall: case OpConvertFToU:
                case OpConvertFToS: case OpConvertSToF: case OpConvertUToF: case OpUConvert: case OpSConvert: case OpFConvert: case OpBitcast:
                case OpSNegate: case OpFNegate: case OpNot: case OpLogicalNot:
                    least = 3;
                    break;
                case OpSelect: least = 5; break;
                case OpExtInst: least = 5; break;  // result type, result, set, instruction, at least one operand
                default: least = 4; break;         // result type, result, two operands (OpCompositeExtract: composite and one index)
            }
            if (w.size() < least) throw Error("SPIR-V: opcode " + std::to_string(in.op) + " is missing operands");
        }
        switch (in.op) {
            case OpFunctionCall: {
                // The IR has no call instruction (instructions_impl.rs:395-433): the callee is inlined into a block of its own —
                //   block <callee>_call -> [the return value's scalars] { <callee's body; OpReturn(Value) = break out of the block> }
                // parameters stand for the arguments (pointers to the caller's variables included, which is how glslang
                // passes everything), the callee's Function variables become more locals of this IR function.
                if (w.size() < 3) throw Error("SPIR-V: OpFunctionCall is missing operands");
                auto f = T.spv_functions.find(w[2]);
                if (f == T.spv_functions.end()) throw Error("SPIR-V: OpFunctionCall of an id that is not a function");
                const FunctionBody& callee = *f->second;
                if (std::find(T.call_stack.begin(), T.call_stack.end(), callee.id) != T.call_stack.end())
                    Translator::unsupported("recursion (a function that calls itself cannot be inlined)");
                if (T.call_stack.size() >= 32) Translator::unsupported("calls nested more than 32 deep");
                if (w.size() - 3 != callee.params.size()) throw Error("SPIR-V: OpFunctionCall argument count differs from the function's parameters");
                if (callee.return_type != w[0]) throw Error("SPIR-V: OpFunctionCall result type differs from the function's return type");
                FunctionTranslator g(T, fn);
                g.load_blocks(callee);
                for (size_t k = 0; k < callee.params.size(); ++k) g.values[callee.params[k]] = value(w[3 + k]);
                Block* cb = new_child_block(T.name_of(callee.id) + "_call", b);
                SValue ret;
                if (T.type(w[0]).kind != SType::Void) {
                    std::vector<std::pair<uint64_t, const Type*>> ls;
                    T.leaves(w[0], 0, &ls);
                    for (size_t k = 0; k < ls.size(); ++k) {
                        Value* r = m.new_value(comp_name(w[1], k, ls.size()), ls[k].second);
                        cb->results.push_back(r);
                        ret.scalars.push_back(r);
                    }
                }
                T.call_stack.push_back(callee.id);
                g.region(callee.blocks.front().label, cb, cb);
                T.call_stack.pop_back();
                values[w[1]] = ret;
                break;
            }
            case OpVariable: {  // a Function variable: a local of the IR function
                if (w[2] != ScFunction) Translator::unsupported("OpVariable of a storage class other than Function inside a function");
                if (w.size() > 3) Translator::unsupported("variable initialisers");
                const SType& pt = T.type(w[0]);
                Variable v;
                v.type = T.ir_type_of(pt.element);
                v.align = T.align_of(pt.element);
                v.pointer = m.new_value(T.name_of(w[1]), m.data_ptr());
                fn->locals.push_back(v);
                SValue sv;
                sv.is_pointer = true;
                sv.base = v.pointer;
                sv.pointee = pt.element;
                values[w[1]] = sv;
                break;
            }
            case OpAccessChain:
            case OpInBoundsAccessChain: {
                SValue p = value(w[2]);
                if (!p.is_pointer) throw Error("SPIR-V: OpAccessChain base is not a pointer");
                uint32_t ty = p.pointee;
                for (size_t k = 3; k < w.size(); ++k) {
                    const SValue& iv = value(w[k]);
                    if (iv.is_pointer || iv.scalars.size() != 1 || !iv.scalars[0]->has_const) Translator::unsupported("OpAccessChain with a dynamic index");
                    const uint64_t idx = iv.scalars[0]->constant.bits;
                    const SType& t = T.type(ty);
                    if (t.kind == SType::Struct) {
                        if (idx >= t.members.size()) throw Error("SPIR-V: struct member index out of range");
                        p.offset += T.member_offset(ty, (uint32_t)idx);
                        ty = t.members[idx];
                    } else if (t.kind == SType::Vector || t.kind == SType::Array) {
                        if (idx >= t.count) throw Error("SPIR-V: constant index out of range");
                        p.offset += idx * T.size_of(t.element);
                        ty = t.element;
                    } else {
                        throw Error("SPIR-V: OpAccessChain indexes into a scalar");
                    }
                }
                p.pointee = ty;
                values[w[1]] = p;
                break;
            }
            case OpLoad: {
                const SValue& p = value(w[2]);
                if (!p.is_pointer) throw Error("SPIR-V: OpLoad from a value that is not a pointer");
                std::vector<std::pair<uint64_t, const Type*>> ls;
                T.leaves(p.pointee, 0, &ls);
                SValue r;
                for (size_t k = 0; k < ls.size(); ++k) {
                    Value* a = address(b, p, ls[k].first, T.name_of(w[1]) + "_p" + std::to_string(k), in.loc);
                    Instr* ld = m.new_instr(Op::Load);
                    ld->args = {a};
                    ld->align = std::max<uint64_t>(1, type_size(ls[k].second));
                    Value* v = m.new_value(comp_name(w[1], k, ls.size()), ls[k].second);
                    ld->results = {v};
                    push(b, ld, in.loc);
                    r.scalars.push_back(v);
                }
                values[w[1]] = r;
                break;
            }
            case OpStore: {
                const SValue& p = value(w[0]);
                const SValue& v = value(w[1]);
                if (!p.is_pointer) throw Error("SPIR-V: OpStore to a value that is not a pointer");
                std::vector<std::pair<uint64_t, const Type*>> ls;
                T.leaves(p.pointee, 0, &ls);
                if (ls.size() != v.scalars.size()) throw Error("SPIR-V: OpStore object does not match the pointee type");
                for (size_t k = 0; k < ls.size(); ++k) {
                    Value* a = address(b, p, ls[k].first, "st" + std::to_string(T.n_const++), in.loc);
                    ext(b, "store", {a, v.scalars[k]}, nullptr, "", in.loc, std::max<uint64_t>(1, type_size(ls[k].second)));
                }
                break;
            }
            case OpCompositeExtract: {
                const SValue& c = value(w[2]);
                if (w.size() != 4) Translator::unsupported("OpCompositeExtract through more than one level");
                SValue r;
                r.scalars.push_back(c.scalars.at(w[3]));
                values[w[1]] = r;
                break;
            }
            case OpCompositeConstruct: {
                SValue r;
                for (size_t k = 2; k < w.size(); ++k)
                    for (Value* s : value(w[k]).scalars) r.scalars.push_back(s);
                if (r.scalars.size() != width(w[0])) throw Error("SPIR-V: OpCompositeConstruct component count does not match the result type");
                values[w[1]] = r;
                break;
            }
            case OpConvertFToU: elementwise("convert_f_to_u", {w[2]}); break;
            case OpConvertFToS: elementwise("convert_f_to_s", {w[2]}); break;
            case OpConvertSToF: elementwise("convert_s_to_f", {w[2]}); break;
            case OpConvertUToF: elementwise("convert_u_to_f", {w[2]}); break;
            case OpUConvert: elementwise("u_convert", {w[2]}); break;
            case OpSConvert: elementwise("s_convert", {w[2]}); break;
            case OpFConvert: elementwise("f_convert", {w[2]}); break;
            case OpBitcast: elementwise("bitcast", {w[2]}); break;
            case OpSNegate: elementwise("s_negate", {w[2]}); break;
            case OpFNegate: elementwise("f_negate", {w[2]}); break;
            case OpNot: elementwise("not", {w[2]}); break;
            case OpLogicalNot: elementwise("logical_not", {w[2]}); break;
            case OpIAdd: {  // the reference's `add`
                const size_t n = width(w[0]);
                SValue r;
                for (size_t k = 0; k < n; ++k) {
                    Instr* ins = m.new_instr(Op::Add);
                    ins->args = {value(w[2]).scalars.at(k), value(w[3]).scalars.at(k)};
                    Value* v = m.new_value(comp_name(w[1], k, n), result_scalar_type(w[0]));
                    ins->results = {v};
                    push(b, ins, in.loc);
                    r.scalars.push_back(v);
                }
                values[w[1]] = r;
                break;
            }
#define KZN_BIN(OPC, NAME) case OPC: elementwise(NAME, {w[2], w[3]}); break;
            KZN_BIN(OpFAdd, "f_add") KZN_BIN(OpISub, "i_sub") KZN_BIN(OpFSub, "f_sub") KZN_BIN(OpIMul, "i_mul") KZN_BIN(OpFMul, "f_mul")
            KZN_BIN(OpFDiv, "f_div") KZN_BIN(OpLogicalOr, "logical_or") KZN_BIN(OpLogicalAnd, "logical_and") KZN_BIN(OpIEqual, "i_equal")
            KZN_BIN(OpINotEqual, "i_not_equal") KZN_BIN(OpUGreaterThan, "u_greater_than") KZN_BIN(OpSGreaterThan, "s_greater_than")
            KZN_BIN(OpUGreaterThanEqual, "u_greater_than_equal") KZN_BIN(OpSGreaterThanEqual, "s_greater_than_equal")
            KZN_BIN(OpULessThan, "u_less_than") KZN_BIN(OpSLessThan, "s_less_than") KZN_BIN(OpULessThanEqual, "u_less_than_equal")
            KZN_BIN(OpSLessThanEqual, "s_less_than_equal") KZN_BIN(OpFOrdEqual, "f_ord_equal") KZN_BIN(OpFUnordNotEqual, "f_unord_not_equal")
            KZN_BIN(OpFOrdLessThan, "f_ord_less_than") KZN_BIN(OpFOrdGreaterThan, "f_ord_greater_than")
            KZN_BIN(OpFOrdLessThanEqual, "f_ord_less_than_equal") KZN_BIN(OpFOrdGreaterThanEqual, "f_ord_greater_than_equal")
            KZN_BIN(OpShiftRightLogical, "shift_right_logical") KZN_BIN(OpShiftRightArithmetic, "shift_right_arithmetic")
            KZN_BIN(OpShiftLeftLogical, "shift_left_logical") KZN_BIN(OpBitwiseOr, "bitwise_or") KZN_BIN(OpBitwiseXor, "bitwise_xor")
            KZN_BIN(OpBitwiseAnd, "bitwise_and")
#undef KZN_BIN
            case OpSelect: elementwise("select", {w[2], w[3], w[4]}); break;
            case OpExtInst: {
                // GLSL.std.450: the instructions that have a version 2 counterpart
                const uint32_t inst = w[3];
                if (w.size() < (inst == 50 ? 7u : (inst == 37 || inst == 40) ? 6u : 5u)) throw Error("SPIR-V: OpExtInst is missing operands");
                if (inst == 50) elementwise("fma", {w[4], w[5], w[6]});       // Fma
                else if (inst == 4) elementwise("f_abs", {w[4]});             // FAbs
                else if (inst == 37) elementwise("f_min", {w[4], w[5]});      // FMin
                else if (inst == 40) elementwise("f_max", {w[4], w[5]});      // FMax
                else Translator::unsupported("GLSL.std.450 instruction " + std::to_string(inst));
                break;
            }
            default: Translator::unsupported("opcode " + std::to_string(in.op) + " (todo!() in the reference's parse/unimplemented_instructions.rs)");
        }
    }

    void break_to(Block* b, Block* target, const std::vector<Value*>& vals, const Location& loc) {
        Instr* br = m.new_instr(Op::Break);
        br->block = target;
        br->args = vals;
        push(b, br, loc);
    }
    // values the predecessor `from` passes to the OpPhis of `label`
    std::vector<Value*> phi_arguments(uint32_t label, uint32_t from) {
        std::vector<Value*> out;
        auto it = phi_insts.find(label);
        if (it == phi_insts.end()) return out;
        for (const Inst& p : it->second) {
            bool found = false;
            for (size_t k = 2; k + 1 < p.w.size(); k += 2)
                if (p.w[k + 1] == from) {
                    out.push_back(scalar(p.w[k]));
                    found = true;
                }
            if (!found) throw Error("SPIR-V: OpPhi has no operand for a predecessor");
        }
        return out;
    }

    // emits the code of SPIR-V block `label` (and what follows it linearly) into IR block `b`; `body` is the function's body block
    Block* new_child_block(const std::string& name, Block* parent) {
        Block* nb = m.new_block(name);
        Instr* bi = m.new_instr(Op::BlockI);
        bi->block = nb;
        parent->body.push_back(bi);
        return nb;
    }
    // the OpPhis of `label` as the results of IR block `nb` (what a `break nb[...]` binds); `bind`: they are the phis' values from now on
    void phis_as_results(uint32_t label, Block* nb, bool bind) {
        auto it = phi_insts.find(label);
        if (it == phi_insts.end()) return;
        for (const Inst& p : it->second) {
            Value* r = m.new_value(T.name_of(p.w[1]), T.scalar_ir_type(T.type(p.w[0])));
            nb->results.push_back(r);
            if (bind) {
                SValue sv;
                sv.scalars = {r};
                values[p.w[1]] = sv;
            }
        }
    }
    void continue_to(Block* b, Loop* lp, const std::vector<Value*>& vals, const Location& loc) {
        Instr* c = m.new_instr(Op::Continue);
        c->loop = lp;
        c->args = vals;
        push(b, c, loc);
    }

    // A structured loop (header `label` with OpLoopMerge %merge %continue), entered from SPIR-V block `from`:
    //   loop L[the header's OpPhi operands for `from`] -> [the merge block's OpPhis] { -> [the header's OpPhis]; L {
    //       block <continue> -> [the continue block's OpPhis] { <header and body; a jump to %continue breaks out of this block,
    //                                                            a jump to %merge out of L> }
    //       <continue block ... its jump back to the header is `continue L[the header's OpPhi operands for it]`>
    //   } }
    // (a loop whose continue target is its own header has no inner block).  A value defined inside the loop is in scope inside
    // it only: what a shader carries out of a loop goes through OpPhis of the merge block or through variables, which is
    // what glslang emits.
    void loop_construct(uint32_t label, Block* b, Block* body, uint32_t from) {
        BasicBlock& bb = blocks.at(label);
        const uint32_t merge = bb.loop_merge, cont = bb.loop_continue;
        if (!blocks.count(merge) || !blocks.count(cont)) throw Error("SPIR-V: OpLoopMerge names a label that is not in the function");
        if (merge == label || merge == cont) throw Error("SPIR-V: a loop's merge block is its header or its continue target");
        if (!loops_built.insert(label).second) Translator::unsupported("a loop header reached twice (unstructured control flow)");
        if (open.count(merge) || loop_open.count(label)) Translator::unsupported("a loop whose merge block is the exit of an enclosing construct");
        Loop* lp = m.new_loop(T.name_of(label));
        if (phi_insts.count(label)) {
            if (!from) Translator::unsupported("OpPhi in a loop header that is the function's first block");
            lp->arguments = phi_arguments(label, from);
            for (const Inst& p : phi_insts[label]) {
                Value* r = m.new_value(T.name_of(p.w[1]), T.scalar_ir_type(T.type(p.w[0])));
                lp->params.push_back(r);
                SValue sv;
                sv.scalars = {r};
                values[p.w[1]] = sv;
            }
        }
        Block* lbody = m.new_block(T.name_of(label));
        phis_as_results(merge, lbody, true);
        lp->results = lbody->results;
        lp->body = lbody;
        Instr* li = m.new_instr(Op::LoopI);
        li->loop = lp;
        push(b, li, bb.terminator.loc);
        const std::map<uint32_t, Block*> saved_open = open;
        const std::map<uint32_t, Loop*> saved_loops = loop_open;
        open[merge] = lbody;
        loop_open[label] = lp;
        Block* inner = lbody;
        if (cont != label) {
            inner = new_child_block(T.name_of(cont), lbody);
            phis_as_results(cont, inner, true);
            open[cont] = inner;
        }
        region(label, inner, body, from, true);
        if (cont != label) {
            open.erase(cont);
            region(cont, lbody, body, 0, false);
        }
        open = saved_open;
        loop_open = saved_loops;
    }

    // emits the code of SPIR-V block `label` (and what follows it linearly) into IR block `b`; `body` is the function's body block;
    // `from`: the SPIR-V block control comes from (0: function entry); `header_inside`: `label` is a loop header whose loop is open already
    void region(uint32_t label, Block* b, Block* body, uint32_t from = 0, bool header_inside = false) {
        for (;;) {
            auto it = blocks.find(label);
            if (it == blocks.end()) throw Error("SPIR-V: branch to a label that is not in the function");
            BasicBlock& bb = it->second;
            if (bb.loop && !header_inside) {
                loop_construct(label, b, body, from);
                from = 0;  // (the merge block's own OpPhis were bound as the loop's results)
                label = bb.loop_merge;
                continue;
            }
            header_inside = false;
            if (!emitted.insert(label).second) Translator::unsupported("a basic block reached twice without a merge (unstructured control flow)");
            for (const Inst& in : bb.body) instruction(b, in);
            const Inst& t = bb.terminator;
            auto jump = [&](uint32_t target, Block* from_block) -> bool {  // true: the jump left through an open construct
                auto o = open.find(target);
                if (o != open.end()) {
                    break_to(from_block, o->second, phi_arguments(target, label), t.loc);
                    return true;
                }
                auto lo = loop_open.find(target);
                if (lo != loop_open.end()) {  // the back edge
                    continue_to(from_block, lo->second, phi_arguments(target, label), t.loc);
                    return true;
                }
                return false;
            };
            if (t.op == OpReturn || t.op == OpReturnValue) {  // (`body`: the function's body block, or the block an inlined call stands in)
                std::vector<Value*> vals;
                if (t.op == OpReturnValue) {
                    if (t.w.empty()) throw Error("SPIR-V: OpReturnValue without a value");
                    const SValue& rv = value(t.w[0]);
                    if (rv.is_pointer) Translator::unsupported("returning a pointer");
                    vals = rv.scalars;
                }
                if (vals.size() != body->results.size()) throw Error("SPIR-V: a return does not match the function's return type");
                break_to(b, body, vals, t.loc);
                return;
            }
            if (t.op == OpKill || t.op == OpUnreachable) Translator::unsupported("terminator opcode " + std::to_string(t.op));
            if (t.w.size() < (t.op == OpBranch ? 1u : t.op == OpBranchConditional ? 3u : t.op == OpSwitch ? 2u : 0u))
                throw Error("SPIR-V: a branch is missing operands");
            if (t.op == OpBranch) {
                if (jump(t.w[0], b)) return;
                from = label;
                label = t.w[0];  // a straight-line successor: keep going in the same IR block
                continue;
            }
            if (t.op != OpSwitch && t.op != OpBranchConditional) throw Error("SPIR-V: a basic block ends without a terminator");
            // A conditional branch without OpSelectionMerge is legal where every way out leaves through an enclosing construct:
            // the branch of a loop header (`while`), of a continue block (`do .. while`), a conditional break / continue.  Its
            // targets are nested in a block that nothing follows.
            const bool no_merge = !bb.merge;
            if (no_merge && loop_open.empty()) Translator::unsupported("a conditional branch without OpSelectionMerge outside a loop");
            // the construct's targets: every label the branch can reach that is not an already open construct's exit; the
            // header of an open loop is one too (a `branch` case can only `break`: it breaks to a block that holds the `continue`)
            std::vector<uint32_t> targets;
            auto add_target = [&](uint32_t l) {
                if (l != bb.merge && !open.count(l) && std::find(targets.begin(), targets.end(), l) == targets.end()) targets.push_back(l);
            };
            if (t.op == OpSwitch) {
                add_target(t.w[1]);
                const Value* sel = scalar(t.w[0]);
                if (sel->type->kind != TypeKind::Int) throw Error("SPIR-V: OpSwitch selector is not an integer");
                const size_t lit = int_bits(sel->type->int_kind) > 32 ? 2 : 1;  // 64-bit case literals take two words
                for (size_t k = 2; k + lit < t.w.size(); k += lit + 1) add_target(t.w[k + lit]);
            } else {
                add_target(t.w[1]);
                add_target(t.w[2]);
            }
            std::sort(targets.begin(), targets.end(), [&](uint32_t x, uint32_t y) { return blocks.at(x).order < blocks.at(y).order; });
            // block merge { block t_last { ... block t_first { branch } <t_first> } ... <t_last> }
            Block* merge_block;
            if (no_merge) {
                merge_block = new_child_block(T.name_of(label) + "_branch", b);
                merge_block->diverges = true;
            } else {
                merge_block = new_child_block(T.name_of(bb.merge), b);
                phis_as_results(bb.merge, merge_block, true);  // the merge block's OpPhis are the results of its IR block
            }
            const std::map<uint32_t, Block*> saved_open = open;
            if (!no_merge) open[bb.merge] = merge_block;
            std::vector<Block*> nest(targets.size());
            std::vector<Loop*> back_edge(targets.size(), nullptr);
            Block* parent = merge_block;
            for (size_t k = targets.size(); k-- > 0;) {
                auto lo = loop_open.find(targets[k]);
                if (lo != loop_open.end()) {  // the header of an open loop: the block's results are what `continue` passes on
                    back_edge[k] = lo->second;
                    nest[k] = new_child_block(T.name_of(targets[k]) + "_continue", parent);
                    phis_as_results(targets[k], nest[k], false);
                } else {
                    nest[k] = new_child_block(T.name_of(targets[k]), parent);
                }
                parent = nest[k];
            }
            for (size_t k = 0; k < targets.size(); ++k) open[targets[k]] = nest[k];
            Block* inner = targets.empty() ? merge_block : nest[0];
            // the branch itself, in the innermost block
            Instr* br = m.new_instr(Op::Branch);
            br->args = {scalar(t.w[0])};
            auto add_case = [&](Const c, uint32_t l) {
                BranchCase bc;
                bc.value = c;
                bc.target = open.at(l);
                bc.results = phi_arguments(l, label);
                br->cases.push_back(bc);
            };
            uint32_t fallthrough;
            if (t.op == OpSwitch) {
                const Value* sel = scalar(t.w[0]);
                const size_t lit = int_bits(sel->type->int_kind) > 32 ? 2 : 1;
                for (size_t k = 2; k + lit < t.w.size(); k += lit + 1) {
                    Const c;
                    c.kind = ConstKind::Int;
                    c.type = sel->type;
                    c.bits = lit == 2 ? ((uint64_t)t.w[k + 1] << 32) | t.w[k] : t.w[k];
                    add_case(c, t.w[k + lit]);
                }
                fallthrough = t.w[1];
            } else {
                Const c;
                c.kind = ConstKind::Bool;
                c.type = m.bool_type();
                c.bits = 1;
                add_case(c, t.w[1]);
                fallthrough = t.w[2];
            }
            push(inner, br, t.loc);
            break_to(inner, open.at(fallthrough), phi_arguments(fallthrough, label), t.loc);
            // the targets' code: label t_k follows the block that is left to reach it
            for (size_t k = 0; k < targets.size(); ++k) {
                open.erase(targets[k]);
                Block* where = k + 1 < targets.size() ? nest[k + 1] : merge_block;
                if (back_edge[k]) continue_to(where, back_edge[k], nest[k]->results, t.loc);
                else region(targets[k], where, body, label, false);
            }
            open = saved_open;
            if (no_merge) return;  // nothing follows: every path left through an enclosing construct
            from = 0;
            label = bb.merge;  // the merge block's code follows the construct, in the block the construct lives in
        }
    }
};

// ---- values that leave the block they are defined in -------------------------------------------------------------------
// SPIR-V lets an id be used wherever its definition dominates the use; the IR scopes a value to the block it is defined in
// (text.rs scope stack).  After translation a use can therefore sit outside its definition's block — a value of a loop body
// used in the continue block, a value of the header used after the loop.  Such a value is handed out through the results of
// every block between the definition and the use: each `break` to that block passes the value where it is in scope (and
// `undef` where it is not: by dominance such a break cannot lead to the use), and the uses outside are renamed to the result.
struct ScopeFixer {
    Translator& T;
    Module& m;
    std::map<const Block*, Block*> parent;      // enclosing block (for a loop's body block: the block that holds the loop)
    std::map<const Block*, Loop*> loop_of_body; // body block -> its loop
    explicit ScopeFixer(Translator& t) : T(t), m(t.m) {}

    void index(Block* b) {
        for (Instr* in : b->body) {
            if (in->op == Op::BlockI) { parent[in->block] = b; index(in->block); }
            if (in->op == Op::LoopI) { parent[in->loop->body] = b; loop_of_body[in->loop->body] = in->loop; index(in->loop->body); }
        }
    }
    static bool has(const std::vector<Value*>& v, const Value* x) { return std::find(v.begin(), v.end(), x) != v.end(); }

    // first use of a value that is not in scope: the value (or null) and the block that holds the use
    Value* offender = nullptr;
    Block* offender_home = nullptr;
    void check_use(Value* v, const std::set<const Value*>& visible, Block* where) {
        if (offender || !v || v->has_const || visible.count(v)) return;
        offender = v;
        offender_home = where;  // (the block that holds the use)
    }
    void scan(Block* b, std::set<const Value*> visible) {
        for (Instr* in : b->body) {
            if (offender) return;
            for (Value* a : in->args) check_use(a, visible, b);
            for (BranchCase& c : in->cases)
                for (Value* a : c.results) check_use(a, visible, b);
            if (in->op == Op::BlockI) {
                scan(in->block, visible);
                for (Value* r : in->block->results) visible.insert(r);
            } else if (in->op == Op::LoopI) {
                for (Value* a : in->loop->arguments) check_use(a, visible, b);
                std::set<const Value*> inner = visible;
                for (Value* pv : in->loop->params) inner.insert(pv);
                scan(in->loop->body, inner);
                for (Value* r : in->loop->results) visible.insert(r);
            } else {
                for (Value* r : in->results) visible.insert(r);
            }
        }
    }
    // the block whose body holds the definition of v (an instruction result, or the results of a nested block / loop, which
    // are defined in the block that holds the nested one); null: v is a loop parameter or a function-level value
    Block* home_of(Block* b, const Value* v) {
        for (Instr* in : b->body) {
            if (has(in->results, v)) return b;
            if (in->op == Op::BlockI) {
                if (has(in->block->results, v)) return b;
                if (Block* h = home_of(in->block, v)) return h;
            }
            if (in->op == Op::LoopI) {
                if (has(in->loop->results, v)) return b;
                if (has(in->loop->params, v)) return in->loop->body;
                if (Block* h = home_of(in->loop->body, v)) return h;
            }
        }
        return nullptr;
    }
    // every `break x` under b gets one more value: v where it is in scope, undef elsewhere
    void pass_out(Block* b, Block* x, Value* v, Value* undef, bool in_scope) {
        for (Instr* in : b->body) {
            if (in->op == Op::Break && in->block == x) in->args.push_back(in_scope ? v : undef);
            for (BranchCase& c : in->cases)
                if (c.target == x) c.results.push_back(in_scope ? v : undef);
            if (in->op == Op::BlockI) {
                pass_out(in->block, x, v, undef, in_scope);
                if (has(in->block->results, v)) in_scope = true;
            } else if (in->op == Op::LoopI) {
                pass_out(in->loop->body, x, v, undef, in_scope || has(in->loop->params, v));
                if (has(in->loop->results, v)) in_scope = true;
            } else if (has(in->results, v)) {
                in_scope = true;
            }
        }
    }
    // uses of v anywhere except inside block x (and its nested blocks) become r
    void rename_outside(Block* b, Block* x, Value* v, Value* r) {
        if (b == x) return;
        for (Instr* in : b->body) {
            for (Value*& a : in->args)
                if (a == v) a = r;
            for (BranchCase& c : in->cases)
                for (Value*& a : c.results)
                    if (a == v) a = r;
            if (in->op == Op::BlockI) rename_outside(in->block, x, v, r);
            if (in->op == Op::LoopI) {
                for (Value*& a : in->loop->arguments)
                    if (a == v) a = r;
                rename_outside(in->loop->body, x, v, r);
            }
        }
    }
    void run(Function* fn) {
        index(fn->body);
        for (int round = 0; round < 10000; ++round) {
            offender = nullptr;
            std::set<const Value*> visible;
            for (Value* pv : fn->params) visible.insert(pv);
            for (const Variable& var : fn->locals) visible.insert(var.pointer);
            for (const Variable& var : m.built_in_inputs) visible.insert(var.pointer);
            for (const Variable& var : m.built_in_outputs) visible.insert(var.pointer);
            for (const Variable& var : m.globals) visible.insert(var.pointer);
            visible.insert(m.user_inputs.pointer);
            visible.insert(m.user_outputs.pointer);
            scan(fn->body, visible);
            if (!offender) return;
            Value* v = offender;
            // x: the block to hand v out of — the outermost block around the definition that does not also hold the use
            Block* x = home_of(fn->body, v);
            std::set<const Block*> around_use;
            for (const Block* u = offender_home; u; u = parent.count(u) ? parent[u] : nullptr) around_use.insert(u);
            if (!x || around_use.count(x)) throw Error("internal: SPIR-V -> IR left a value used before its definition ('" + v->name + "')");
            while (parent.count(x) && !around_use.count(parent[x])) x = parent[x];
            if (x->diverges) throw Error("internal: SPIR-V -> IR: a value would have to leave a block that nothing follows");
            Value* r = m.new_value(v->name, v->type);  // (the printer tells equal names apart by a suffix)
            Value* undef = T.constant(v->type, 0, ConstKind::Undef);
            pass_out(x, x, v, undef, has_param(x, v));
            x->results.push_back(r);
            auto lb = loop_of_body.find(x);
            if (lb != loop_of_body.end()) lb->second->results.push_back(r);
            rename_outside(fn->body, x, v, r);
        }
        throw Error("internal: SPIR-V -> IR scope repair does not terminate");
    }
    bool has_param(Block* x, const Value* v) {
        auto lb = loop_of_body.find(x);
        return lb != loop_of_body.end() && has(lb->second->params, v);
    }
};

}  // namespace

std::unique_ptr<Module> translate_spirv_v2(const uint32_t* words, size_t count, const std::string& entry_point, uint32_t execution_model,
                                           uint32_t local_size[3]) {
    if (count < 5) throw Error("SPIR-V: module is shorter than its header");
    if (words[0] != kMagic) throw Error("SPIR-V: bad magic number (byte-swapped modules are not accepted)");
    const uint32_t bound = words[3];
    auto mod = std::make_unique<Module>();
    Module& m = *mod;
    m.version = 2;
    m.user_inputs.pointer = m.new_value("user_inputs_block", m.data_ptr());
    m.user_inputs.align = 16;
    m.user_outputs.pointer = m.new_value("user_outputs_block", m.data_ptr());
    m.user_outputs.align = 16;
    if (local_size) local_size[0] = local_size[1] = local_size[2] = 1;
    Translator T(m, bound);

    struct EntryPoint {
        uint32_t model, id;
        std::string name;
    };
    std::vector<EntryPoint> entries;
    std::map<uint32_t, std::vector<uint32_t>> local_sizes;
    struct GlobalVar {
        uint32_t id, type, storage;
    };
    std::vector<GlobalVar> globals;
    std::vector<FunctionBody> bodies;
    FunctionBody* cur = nullptr;
    BasicBlock* bb = nullptr;
    Location loc;
    auto check_id = [&](uint32_t id) {
        if (id == 0 || id >= bound) throw Error("SPIR-V: id " + std::to_string(id) + " is out of bounds");
    };

    size_t i = 5;
    while (i < count) {
        const uint32_t wc = words[i] >> 16, op = words[i] & 0xffff;
        if (wc == 0 || i + wc > count) throw Error("SPIR-V: instruction at word " + std::to_string(i) + " overruns the module");
        const uint32_t* o = words + i + 1;
        const size_t n = wc - 1;
        auto need = [&](size_t k) {
            if (n < k) throw Error("SPIR-V: opcode " + std::to_string(op) + " is missing operands");
        };
        if (cur && bb && op != OpLine && op != OpNoLine && op != OpLabel && op != OpFunctionEnd) {
            // inside a basic block: keep the instruction for the function translator
            Inst in{op, std::vector<uint32_t>(o, o + n), loc};
            if (op == OpSelectionMerge) {
                need(2);
                bb->merge = o[0];
            } else if (op == OpLoopMerge) {
                need(2);
                bb->loop = true;
                bb->loop_merge = o[0];
                bb->loop_continue = o[1];
            } else if (op == OpBranch || op == OpBranchConditional || op == OpSwitch || op == OpReturn || op == OpReturnValue || op == OpKill ||
                       op == OpUnreachable) {
                bb->terminator = in;
                bb = nullptr;
                loc = Location();  // a terminator ends the line scope (parse/debug_locations.rs:41-44)
            } else {
                bb->body.push_back(in);
            }
            i += wc;
            continue;
        }
        if (op >= OpTypeVoid && op <= OpTypeFunction) {  // (a type defined twice could be made to contain itself)
            need(1);
            if (T.types.count(o[0])) throw Error("SPIR-V: type id " + std::to_string(o[0]) + " is defined twice");
        }
        switch (op) {
            case OpNop: case OpSourceContinued: case OpSource: case OpSourceExtension: case OpMemberName: case OpExtension:
            case OpExtInstImport: case OpMemoryModel: case OpCapability: case OpModuleProcessed:
                break;
            case OpName: need(2); check_id(o[0]); T.names[o[0]] = literal_string(o + 1, n - 1); break;
            case OpString: need(2); check_id(o[0]); T.strings[o[0]] = literal_string(o + 1, n - 1); break;
            case OpLine: {
                need(3);
                auto it = T.strings.find(o[0]);
                if (it == T.strings.end()) throw Error("SPIR-V: OpLine file is not an OpString");
                loc.present = true; loc.file = it->second; loc.line = o[1]; loc.column = o[2];
                break;
            }
            case OpNoLine: loc = Location(); break;
            case OpEntryPoint: need(3); check_id(o[1]); entries.push_back({o[0], o[1], literal_string(o + 2, n - 2)}); break;
            case OpExecutionMode:
                need(2);
                if (o[1] == 17 /* LocalSize */) {
                    need(5);
                    local_sizes[o[0]] = {o[2], o[3], o[4]};
                }
                break;
            case OpDecorate: need(2); T.decorations[o[0]][o[1]] = n > 2 ? o[2] : 1; break;
            case OpMemberDecorate: need(3); T.member_decorations[o[0]][o[1]][o[2]] = n > 3 ? o[3] : 1; break;
            case OpTypeVoid: need(1); T.types[o[0]] = SType{SType::Void}; break;
            case OpTypeBool: need(1); T.types[o[0]] = SType{SType::Bool}; break;
            case OpTypeInt: {
                need(3);
                if (o[1] != 8 && o[1] != 16 && o[1] != 32 && o[1] != 64) throw Error("SPIR-V: integer width " + std::to_string(o[1]));
                SType t{SType::Int}; t.bits = o[1]; t.is_signed = o[2] != 0; T.types[o[0]] = t; break;
            }
            case OpTypeFloat: {
                need(2);
                if (o[1] != 16 && o[1] != 32 && o[1] != 64) throw Error("SPIR-V: float width " + std::to_string(o[1]));
                SType t{SType::Float}; t.bits = o[1]; T.types[o[0]] = t; break;
            }
            case OpTypeVector: {
                need(3);
                if (!T.types.count(o[1]) || o[1] == o[0]) throw Error("SPIR-V: a type is used before it is defined");
                if (o[2] < 2 || o[2] > 4) throw Error("SPIR-V: vectors have 2, 3 or 4 components (Vector16 is not a Shader capability)");
                SType t{SType::Vector}; t.element = o[1]; t.count = o[2]; T.types[o[0]] = t; break;
            }
            case OpTypeArray: {
                need(3);
                if (!T.types.count(o[1]) || o[1] == o[0]) throw Error("SPIR-V: a type is used before it is defined");
                SType t{SType::Array};
                t.element = o[1];
                auto c = T.values.find(o[2]);
                if (c == T.values.end() || c->second.scalars.size() != 1) throw Error("SPIR-V: array length is not a constant");
                t.count = (uint32_t)c->second.scalars[0]->constant.bits;
                // every element becomes scalars of its own here (no dynamic indexing): an array is as long as a shader's interface
                // needs, not as long as a buffer
                if (c->second.scalars[0]->constant.bits == 0 || c->second.scalars[0]->constant.bits > 1024) Translator::unsupported("arrays of more than 1024 elements");
                T.types[o[0]] = t;
                break;
            }
            case OpTypeStruct: {
                need(1);
                for (size_t k = 1; k < n; ++k)
                    if (!T.types.count(o[k]) || o[k] == o[0]) throw Error("SPIR-V: a type is used before it is defined");
                SType t{SType::Struct}; t.members.assign(o + 1, o + n); T.types[o[0]] = t; break;
            }
            case OpTypePointer: {
                need(3);
                if (!T.types.count(o[2]) || o[2] == o[0]) throw Error("SPIR-V: a type is used before it is defined (OpTypeForwardPointer is not supported)");
                SType t{SType::Pointer}; t.storage = o[1]; t.element = o[2]; T.types[o[0]] = t; break;
            }
            case OpTypeFunction: {
                need(2);
                // (parameters: parse/functions.rs:236-268 is todo!() upstream; here functions other than the entry point are inlined)
                SType t{SType::Function};
                t.element = o[1];
                T.types[o[0]] = t;
                break;
            }
            case OpConstantTrue:
            case OpConstantFalse: {
                need(2);
                SValue v;
                v.scalars = {T.constant(m.bool_type(), op == OpConstantTrue, ConstKind::Bool)};
                T.values[o[1]] = v;
                break;
            }
            case OpConstant: {
                need(3);
                const SType& t = T.type(o[0]);
                uint64_t bits = o[2];
                if (t.bits > 32) {
                    need(4);
                    bits |= (uint64_t)o[3] << 32;
                }
                if (t.kind == SType::Int && t.bits < 64) bits &= (1ull << t.bits) - 1;
                SValue v;
                v.scalars = {T.constant(T.scalar_ir_type(t), bits, t.kind == SType::Float ? ConstKind::Float : ConstKind::Int)};
                T.values[o[1]] = v;
                break;
            }
            case OpConstantComposite: {
                need(2);
                SValue v;
                for (size_t k = 2; k < n; ++k) {
                    auto c = T.values.find(o[k]);
                    if (c == T.values.end()) throw Error("SPIR-V: OpConstantComposite constituent is not a constant");
                    for (Value* s : c->second.scalars) v.scalars.push_back(s);
                }
                T.values[o[1]] = v;
                break;
            }
            case OpVariable:
                need(3);
                if (cur) throw Error("SPIR-V: OpVariable outside a basic block");
                globals.push_back({o[1], o[0], o[2]});
                break;
            case OpFunction: {
                need(4);
                if (cur) throw Error("SPIR-V: OpFunction inside a function");
                bodies.emplace_back();
                cur = &bodies.back();
                cur->id = o[1];
                cur->control = o[2];
                cur->return_type = o[0];
                break;
            }
            case OpFunctionParameter:
                need(2);
                if (!cur || !cur->blocks.empty()) throw Error("SPIR-V: OpFunctionParameter outside a function's head");
                check_id(o[1]);
                cur->params.push_back(o[1]);
                break;
            case OpLabel:
                need(1);
                if (!cur) throw Error("SPIR-V: OpLabel outside a function");
                if (bb) throw Error("SPIR-V: OpLabel inside a basic block");
                cur->blocks.emplace_back();
                bb = &cur->blocks.back();
                bb->label = o[0];
                bb->order = cur->blocks.size() - 1;
                break;
            case OpFunctionEnd:
                if (!cur) throw Error("SPIR-V: OpFunctionEnd outside a function");
                if (bb) throw Error("SPIR-V: basic block is missing its termination instruction");
                if (cur->blocks.empty()) throw Error("SPIR-V: function must have a body");
                cur = nullptr;
                break;
            default: Translator::unsupported("opcode " + std::to_string(op) + " at module scope");
        }
        i += wc;
    }
    if (cur) throw Error("SPIR-V: missing OpFunctionEnd");

    const EntryPoint* entry = nullptr;
    for (auto& e : entries)
        if (e.name == entry_point && e.model == execution_model) entry = &e;
    if (!entry) throw Error("SPIR-V: no entry point named '" + entry_point + "' with the requested execution model");

    // ---- interface: Input / Output variables by Location (io_layout.rs: 16 bytes per location, 4 per component), built-ins,
    // module-scope Private variables
    auto lay_out = [&](InterfaceBlock& block, const GlobalVar& g, uint32_t pointee) {
        const auto& dec = T.decorations[g.id];
        if (!dec.count(DecLocation)) Translator::unsupported("an interface variable without a Location");
        const uint64_t at = 16ull * dec.at(DecLocation) + 4ull * (dec.count(DecComponent) ? dec.at(DecComponent) : 0);
        std::vector<std::pair<uint64_t, const Type*>> ls;
        T.leaves(pointee, at, &ls);
        for (auto& l : ls) block.members.push_back({l.first, l.second});
        block.size = std::max<uint64_t>(block.size, (at + T.size_of(pointee) + 15) / 16 * 16);
        SValue v;
        v.is_pointer = true;
        v.base = block.pointer;
        v.offset = at;
        v.pointee = pointee;
        T.values[g.id] = v;
    };
    auto built_in_variable = [&](std::vector<Variable>& list, const char* kind, const Type* type, uint64_t align, const std::string& name) {
        Variable v;
        v.type = type;
        v.align = align;
        v.pointer = m.new_value(name, m.data_ptr());
        v.built_in_kind = kind;
        list.push_back(v);
        return v.pointer;
    };
    auto vec = [&](const Type* e, uint64_t n) {
        Type v{TypeKind::Vector};
        v.length = n;
        v.element = e;
        return m.intern(v);
    };
    for (const GlobalVar& g : globals) {
        const SType& pt = T.type(g.type);
        if (pt.kind != SType::Pointer) throw Error("SPIR-V: OpVariable result type must be a pointer");
        const uint32_t pointee = pt.element;
        const auto& dec = T.decorations[g.id];
        if (g.storage == ScPrivate) {
            Variable v;
            v.type = T.ir_type_of(pointee);
            v.align = T.align_of(pointee);
            v.pointer = m.new_value(T.name_of(g.id), m.data_ptr());
            m.globals.push_back(v);
            SValue sv;
            sv.is_pointer = true;
            sv.base = v.pointer;
            sv.pointee = pointee;
            T.values[g.id] = sv;
        } else if (g.storage == ScInput || g.storage == ScOutput) {
            const bool in = g.storage == ScInput;
            const SType& st = T.type(pointee);
            Value* base = nullptr;
            if (st.kind == SType::Struct && T.decorations[pointee].count(DecBlock)) {
                // gl_PerVertex: the member decorated BuiltIn Position is the vertex_position variable (interface.rs:61-67)
                if (in) Translator::unsupported("built-in input blocks");
                const auto& md = T.member_decorations[pointee];
                if (st.members.size() != 1 || !md.count(0) || !md.at(0).count(DecBuiltIn) || md.at(0).at(DecBuiltIn) != BuiltInPosition)
                    Translator::unsupported("built-in output blocks other than { vec4 Position }");
                base = built_in_variable(m.built_in_outputs, "vertex_position", vec(m.float_type(FloatKind::F32), 4), 16, T.name_of(g.id));
            } else if (dec.count(DecBuiltIn)) {
                const uint32_t bi = dec.at(DecBuiltIn);
                const Type* i32 = m.int_type(IntKind::I32);
                if (!in && bi == BuiltInPosition) base = built_in_variable(m.built_in_outputs, "vertex_position", vec(m.float_type(FloatKind::F32), 4), 16, T.name_of(g.id));
                else if (in && bi == 28) base = built_in_variable(m.built_in_inputs, "global_invocation_id", vec(i32, 3), 16, T.name_of(g.id));
                else if (in && bi == 15) base = built_in_variable(m.built_in_inputs, "frag_coord", vec(m.float_type(FloatKind::F32), 4), 16, T.name_of(g.id));
                else if (in && bi == 41) base = built_in_variable(m.built_in_inputs, "subgroup_local_invocation_id", i32, 4, T.name_of(g.id));
                else if (in && bi == 36) base = built_in_variable(m.built_in_inputs, "subgroup_size", i32, 4, T.name_of(g.id));
                else Translator::unsupported("BuiltIn " + std::to_string(bi));
            } else {
                lay_out(in ? m.user_inputs : m.user_outputs, g, pointee);
                continue;
            }
            SValue sv;
            sv.is_pointer = true;
            sv.base = base;
            sv.pointee = pointee;
            T.values[g.id] = sv;
        } else {
            Translator::unsupported("variables of storage class " + std::to_string(g.storage) + " (resources bound through descriptors)");
        }
    }
    for (InterfaceBlock* blk : {&m.user_inputs, &m.user_outputs})
        std::sort(blk->members.begin(), blk->members.end(), [](const StructMember& a, const StructMember& b) { return a.offset < b.offset; });

    // ---- the entry point's function
    const FunctionBody* fb = nullptr;
    for (auto& b : bodies)
        if (b.id == entry->id) fb = &b;
    if (!fb) throw Error("SPIR-V: entry point is not a function");
    if ((fb->control & 1) && (fb->control & 2)) throw Error("SPIR-V: Inline and DontInline are not allowed together");
    Function* fn = m.new_function(T.name_of(fb->id));
    fn->inlining_hint = (fb->control & 1) ? "inline" : (fb->control & 2) ? "dont_inline" : "none";
    fn->side_effects = (fb->control & 4) ? "pure" : (fb->control & 8) ? "const" : "normal";
    Type ft{TypeKind::Fn};
    fn->type = m.intern(ft);
    if (!fb->params.empty()) throw Error("SPIR-V: an entry point takes no parameters");
    for (const FunctionBody& b : bodies) T.spv_functions[b.id] = &b;
    T.call_stack.push_back(fb->id);
    FunctionTranslator F(T, fn);
    F.load_blocks(*fb);
    Block* body = m.new_block(T.name_of(fb->blocks.front().label));
    fn->body = body;
    F.region(fb->blocks.front().label, body, body);
    ScopeFixer(T).run(fn);
    m.functions.push_back(fn);
    m.entry_point = fn;
    auto ls = local_sizes.find(entry->id);
    if (ls != local_sizes.end() && local_size)
        for (int k = 0; k < 3; ++k) local_size[k] = ls->second[k];
    return mod;
}

}  // namespace kir
