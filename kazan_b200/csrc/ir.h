// ir.h — in-memory mirror of shader-compiler-ir (the input language of the PTX back end).
//
// Mirrors (kazan-3d/kazan): shader-compiler-ir/src/module.rs:128-150 `Module`, function.rs:217-231 `FunctionData`,
// block.rs:116-124 `BlockData`, block.rs:492-503 `LoopData`, instructions_impl.rs:395-433 (the 7 instruction kinds),
// types.rs:524-538 `Type`, consts.rs:277-293 `Const`, values.rs `ValueDefinition`/`ValueUse`,
// target_properties.rs:12-17, interface.rs:61-67,134-141.
// Text format: ir_text.cpp (reader + printer).
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace kir {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class TypeKind { Int, Float, Bool, DataPtr, Fn, Vector, Struct };
enum class IntKind { I8, I16, I32, RI32, I64 };
enum class FloatKind { F16, F32, RF32, F64 };

struct Type;
struct StructMember {
    uint64_t offset;
    const Type* type;
};

struct Type {  // interned: pointer equality == structural equality (types.rs interning)
    TypeKind kind;
    IntKind int_kind = IntKind::I32;
    FloatKind float_kind = FloatKind::F32;
    std::vector<const Type*> args, results;  // Fn
    bool diverges = false;                   // Fn: "-> !"
    bool scalable = false;                   // Vector: "vscale x"
    uint64_t length = 0;
    const Type* element = nullptr;
    bool size_variable = false;  // Struct
    uint64_t size = 0, align = 1;
    std::vector<StructMember> members;
    std::string key;  // canonical structural key
};

int int_bits(IntKind k);
int float_bits(FloatKind k);
const char* int_name(IntKind k);
const char* float_name(FloatKind k);

struct Function;

enum class ConstKind { Int, Float, Bool, Vector, Undef, Null, Fn };
struct Const {
    ConstKind kind;
    const Type* type = nullptr;
    uint64_t bits = 0;            // Int / Float bit pattern / Bool
    std::vector<Const> elements;  // Vector
    Function* function = nullptr;
};

struct Value {  // a ValueDefinition; `constant` set for inline constant definitions (values.rs:168-200)
    std::string name;
    const Type* type = nullptr;
    bool has_const = false;
    Const constant;
    int id = 0;  // dense per module, assigned at creation
};

struct Variable {  // function.rs:155-193
    const Type* type = nullptr;
    uint64_t align = 1;
    Value* pointer = nullptr;
    std::string built_in_kind;  // interface variables only ("vertex_position")
};

struct Instr;
struct Block {
    std::string name;
    bool diverges = false;        // results: "!"
    std::vector<Value*> results;  // bound by `break name[...]`
    std::vector<Instr*> body;
    int id = 0;
};

struct Loop {
    std::string name;
    std::vector<Value*> arguments;  // uses
    std::vector<Value*> params;     // definitions re-bound by `continue`
    bool diverges = false;
    std::vector<Value*> results;
    Block* body = nullptr;
    int id = 0;
};

enum class Op { Add, Load, Break, Continue, Branch, BlockI, LoopI, Ext };

// ---- version 2 of the instruction set ----------------------------------------------------------------
// The reference's IR has seven instructions (instructions_impl.rs:395-433) and its translator stops at the first SPIR-V
// instruction with a body (473 `todo!()`s, parse/unimplemented_instructions.rs).  What a shader needs beyond them lives here,
// in the C++ mirror, behind a version flag: a reader accepts these instructions only when it is asked for version 2
// (parse_module(text, 2) / kzn_ir_parse_ex(..., KZN_IR_VERSION_2)); version 1 is the reference's grammar to the letter.
// Each instruction is spelled like the SPIR-V instruction it translates (OpFMul -> `f_mul`, OpConvertFToU ->
// `convert_f_to_u`, OpGroupNonUniformBallot -> `group_non_uniform_ballot`) in the reference's instruction syntax
// `name [arguments] -> [results]`, and means what the SPIR-V specification says for scalar operands, with the
// implementation-defined corners fixed (oracle/ir_simt.py is the checker):
//   floating point: IEEE round-to-nearest-even, no contraction (`fma` is the only fused form), denormals kept;
//   f_min / f_max return the other operand when one is NaN; float -> integer conversions round toward zero and saturate
//   (NaN -> 0); shifts by the operand width or more give 0 (logical) or the sign (arithmetic);
//   `store [pointer, value], align: A -> []` and `ptr_add [pointer, offset: i64] -> [q: data_ptr]` (the byte-offset form of
//   OpAccessChain once the layout is known) complete `load`;
//   group_non_uniform_*: subgroup = 32 consecutive invocations (one warp), defined in subgroup-uniform control flow only;
//   ballot returns the low word of the SPIR-V uvec4; i_add is the Reduce group operation.
enum class ExtKind {
    IntBinary,     // [a, b] -> [r], one integer type
    IntUnary,      // [a] -> [r]
    FloatBinary,   // [a, b] -> [r], one float type
    FloatUnary,
    FloatTernary,  // fma
    IntCompare,    // [a, b: iN] -> [r: bool]
    FloatCompare,  // [a, b: fN] -> [r: bool]
    BoolBinary,    // [a, b: bool] -> [r: bool]
    BoolUnary,
    Select,        // [c: bool, a, b] -> [r]
    Convert,       // [a] -> [r], kinds checked per instruction
    Store,         // [p: data_ptr, v], align: A -> []
    PtrAdd,        // [p: data_ptr, off: i64] -> [q: data_ptr]
    SubgroupElect,      // [] -> [bool]
    SubgroupVote,       // [bool] -> [bool]
    SubgroupBallot,     // [bool] -> [i32]
    SubgroupBroadcast,  // [x] -> [x]
    SubgroupShuffle,    // [x, i32] -> [x]
    SubgroupReduce,     // [i32] -> [i32]
};
struct ExtOp {
    const char* name;
    ExtKind kind;
    const char* ptx;  // mnemonic stem used by ptx_emit.cpp
};
extern const ExtOp kExtOps[];
extern const int kExtOpCount;
int ext_op_index(const std::string& name);  // -1 when `name` is not a version 2 instruction
bool is_v2_built_in(const std::string& kind);

struct BranchCase {
    Const value;
    Block* target;
    std::vector<Value*> results;
};

struct Location {
    bool present = false;
    std::string file;
    uint64_t line = 0, column = 0;
};

struct Instr {
    Op op;
    std::vector<Value*> args;
    std::vector<Value*> results;
    uint64_t align = 0;          // Load, Ext store
    int ext = -1;                // Ext: index into kExtOps
    Block* block = nullptr;      // Break target / BlockI body
    Loop* loop = nullptr;        // Continue target / LoopI body
    std::vector<BranchCase> cases;
    Location location;
};

struct Function {
    std::string name;
    std::vector<Value*> params;
    bool diverges = false;
    std::vector<const Type*> returns;
    std::string inlining_hint = "none", side_effects = "normal";
    std::vector<Variable> locals;
    Block* body = nullptr;
    const Type* type = nullptr;
};

struct InterfaceBlock {  // interface.rs:199-252
    Value* pointer = nullptr;
    bool size_variable = false;
    uint64_t size = 0, align = 1;
    std::vector<StructMember> members;
};

struct Module {
    IntKind data_ptr_type = IntKind::I64, fn_ptr_type = IntKind::I64;  // target_properties.rs:12-17
    std::vector<Variable> built_in_inputs, built_in_outputs, globals;
    InterfaceBlock user_inputs, user_outputs;
    std::vector<Function*> functions;
    Function* entry_point = nullptr;

    // arenas
    std::vector<std::unique_ptr<Type>> types;
    std::map<std::string, const Type*> type_table;
    std::vector<std::unique_ptr<Value>> values;
    std::vector<std::unique_ptr<Block>> blocks;
    std::vector<std::unique_ptr<Loop>> loops;
    std::vector<std::unique_ptr<Instr>> instrs;
    std::vector<std::unique_ptr<Function>> function_arena;
    int version = 1;  // 2: the module may hold Ext instructions and version 2 built-in kinds

    const Type* intern(Type t);
    const Type* int_type(IntKind k);
    const Type* float_type(FloatKind k);
    const Type* bool_type();
    const Type* data_ptr();
    Value* new_value(const std::string& name, const Type* type);
    Block* new_block(const std::string& name);
    Loop* new_loop(const std::string& name);
    Instr* new_instr(Op op);
    Function* new_function(const std::string& name);
};

// text.rs FromText / ToText
std::unique_ptr<Module> parse_module(const std::string& text, int version = 1);
std::string module_to_text(const Module& m);
// parse `text` as one syntactic category and print it canonically — what the reference's test_from_to_text /
// test_type / test_const macros check.  kind: type const target_properties block function
// definitions_then_instruction module
std::string round_trip(const std::string& kind, const std::string& text);

uint64_t type_size(const Type* t);  // types.rs:62-81,109-126,308-327 (pointers are 8 bytes)

}  // namespace kir
