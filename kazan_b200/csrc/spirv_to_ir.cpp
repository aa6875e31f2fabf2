// spirv_to_ir.cpp — SPIR-V -> shader-compiler-ir for the shaders the reference translator handles.
//
// Mirrors shader-compiler-translate-spirv-to-ir (kazan-3d/kazan): TranslatedSPIRVShader::new src/lib.rs:190-215; names
// are the OpName of an id or `id_<n>` (parse/debug_names.rs:25-31); a function's body block is named after its entry
// label (parse/functions.rs:287); FunctionControl bits become the hints (parse/functions.rs:203-224); OpLine / OpNoLine
// attach `@ "file":line:col` to what follows until the block terminator (parse/debug_locations.rs:19-51); OpReturn is
// `break <body>[]` (parse/translate_structure_tree.rs:236-256); both user interface blocks are empty with the
// 16-byte location alignment (io_layout.rs:11-13).  Everything the reference leaves todo!()
// (parse/unimplemented_instructions.rs:14-39, memory_instructions.rs:99-141, translate_structure_tree.rs:258-291) is
// reported as unsupported, by opcode.  Golden: test_data/trivial_test.spv -> trivial_test.kazan-ir (lib.rs:232-258).
#include <map>

#include "ptx_emit.h"

namespace kir {
namespace {

constexpr uint32_t kMagic = 0x07230203;

std::string literal_string(const uint32_t* w, size_t n, size_t* used) {
    std::string s;
    for (size_t i = 0; i < n; ++i) {
        for (int b = 0; b < 4; ++b) {
            char ch = (char)((w[i] >> (8 * b)) & 0xff);
            if (!ch) {
                if (used) *used = i + 1;
                return s;
            }
            s += ch;
        }
    }
    throw Error("SPIR-V: unterminated literal string");
}

enum : uint32_t {
    OpNop = 0, OpSourceContinued = 2, OpSource = 3, OpSourceExtension = 4, OpName = 5, OpMemberName = 6, OpString = 7, OpLine = 8,
    OpExtension = 10, OpExtInstImport = 11, OpMemoryModel = 14, OpEntryPoint = 15, OpExecutionMode = 16, OpCapability = 17,
    OpTypeVoid = 19, OpTypeFunction = 33, OpFunction = 54, OpFunctionParameter = 55, OpFunctionEnd = 56, OpLabel = 248,
    OpReturn = 253, OpNoLine = 317, OpModuleProcessed = 330,
};

}  // namespace

std::unique_ptr<Module> translate_spirv(const uint32_t* words, size_t count, const std::string& entry_point,
                                        uint32_t execution_model, uint32_t local_size[3]) {
    if (count < 5) throw Error("SPIR-V: module is shorter than its header");
    if (words[0] != kMagic) throw Error("SPIR-V: bad magic number (byte-swapped modules are not accepted)");
    const uint32_t bound = words[3];
    auto m = std::make_unique<Module>();
    m->user_inputs.pointer = m->new_value("user_inputs_block", m->data_ptr());
    m->user_inputs.align = 16;
    m->user_outputs.pointer = m->new_value("user_outputs_block", m->data_ptr());
    m->user_outputs.align = 16;
    if (local_size) local_size[0] = local_size[1] = local_size[2] = 1;

    std::map<uint32_t, std::string> names, strings;
    std::map<uint32_t, bool> void_types;                   // id -> is OpTypeVoid
    std::map<uint32_t, uint32_t> fn_types;                 // id -> return type id (parameterless only)
    struct EntryPoint {
        uint32_t model, id;
        std::string name;
    };
    std::vector<EntryPoint> entries;
    std::map<uint32_t, std::vector<uint32_t>> local_sizes;  // function id -> LocalSize
    std::map<uint32_t, Function*> functions;

    auto name_of = [&](uint32_t id) {
        auto it = names.find(id);
        return it != names.end() ? it->second : "id_" + std::to_string(id);
    };
    auto check_id = [&](uint32_t id) {
        if (id == 0 || id >= bound) throw Error("SPIR-V: id " + std::to_string(id) + " is out of bounds");
    };

    Function* fn = nullptr;
    Block* body = nullptr;
    bool terminated = false;
    Location loc;
    size_t i = 5;
    while (i < count) {
        uint32_t wc = words[i] >> 16, op = words[i] & 0xffff;
        if (wc == 0 || i + wc > count) throw Error("SPIR-V: instruction at word " + std::to_string(i) + " overruns the module");
        const uint32_t* o = words + i + 1;
        size_t n = wc - 1;
        auto need = [&](size_t k) {
            if (n < k) throw Error("SPIR-V: opcode " + std::to_string(op) + " is missing operands");
        };
        switch (op) {
            case OpNop: case OpSourceContinued: case OpSource: case OpSourceExtension: case OpMemberName: case OpExtension:
            case OpExtInstImport: case OpMemoryModel: case OpCapability: case OpModuleProcessed:
                break;
            case OpName: need(2); check_id(o[0]); names[o[0]] = literal_string(o + 1, n - 1, nullptr); break;
            case OpString: need(2); check_id(o[0]); strings[o[0]] = literal_string(o + 1, n - 1, nullptr); break;
            case OpLine: {
                need(3);
                auto it = strings.find(o[0]);
                if (it == strings.end()) throw Error("SPIR-V: OpLine file is not an OpString");
                loc.present = true; loc.file = it->second; loc.line = o[1]; loc.column = o[2];
                break;
            }
            case OpNoLine: loc = Location(); break;
            case OpEntryPoint: {
                need(3);
                check_id(o[1]);
                entries.push_back({o[0], o[1], literal_string(o + 2, n - 2, nullptr)});
                break;
            }
            case OpExecutionMode:
                need(2);
                if (o[1] == 17 /* LocalSize */) {
                    need(5);
                    local_sizes[o[0]] = {o[2], o[3], o[4]};
                }
                break;  // other modes carry nothing the IR records (parse/execution_mode.rs keeps them in translator state)
            case OpTypeVoid: need(1); check_id(o[0]); void_types[o[0]] = true; break;
            case OpTypeFunction:
                need(2);
                check_id(o[0]);
                if (n > 2) throw Unsupported("SPIR-V -> IR: function parameters (parse/functions.rs:236-268 is todo!() upstream)");
                if (!void_types.count(o[1])) throw Unsupported("SPIR-V -> IR: non-void return types");
                fn_types[o[0]] = o[1];
                break;
            case OpFunction: {
                need(4);
                if (fn) throw Error("SPIR-V: OpFunction inside a function");
                check_id(o[1]);
                if (!fn_types.count(o[3])) throw Error("SPIR-V: OpFunction's type is not an OpTypeFunction");
                if (fn_types[o[3]] != o[0]) throw Error("SPIR-V: OpFunction's result type must match its function type's return type");
                uint32_t control = o[2];
                if ((control & 1) && (control & 2)) throw Error("SPIR-V: Inline and DontInline are not allowed together");
                if ((control & 4) && (control & 8)) throw Error("SPIR-V: Const and Pure are not allowed together");
                fn = m->new_function(name_of(o[1]));
                fn->inlining_hint = (control & 1) ? "inline" : (control & 2) ? "dont_inline" : "none";
                fn->side_effects = (control & 4) ? "pure" : (control & 8) ? "const" : "normal";
                Type ft{TypeKind::Fn};
                fn->type = m->intern(ft);
                functions[o[1]] = fn;
                body = nullptr;
                terminated = false;
                break;
            }
            case OpFunctionParameter: throw Unsupported("SPIR-V -> IR: OpFunctionParameter (todo!() upstream)");
            case OpLabel:
                need(1);
                if (!fn) throw Error("SPIR-V: OpLabel outside a function");
                if (body) throw Unsupported("SPIR-V -> IR: more than one basic block (branches are todo!() upstream, translate_structure_tree.rs:258-291)");
                check_id(o[0]);
                body = m->new_block(name_of(o[0]));
                fn->body = body;
                break;
            case OpReturn: {
                if (!body || terminated) throw Error("SPIR-V: OpReturn outside a basic block");
                Instr* br = m->new_instr(Op::Break);
                br->block = body;
                br->location = loc;
                body->body.push_back(br);
                terminated = true;
                loc = Location();  // a terminator ends the line scope (parse/debug_locations.rs:41-44)
                break;
            }
            case OpFunctionEnd:
                if (!fn) throw Error("SPIR-V: OpFunctionEnd outside a function");
                if (!body) throw Error("SPIR-V: function must have a body");
                if (!terminated) throw Error("SPIR-V: basic block is missing its termination instruction");
                m->functions.push_back(fn);
                fn = nullptr;
                body = nullptr;
                break;
            default:
                throw Unsupported("SPIR-V -> IR: opcode " + std::to_string(op) +
                                  " is not translated (todo!() in the reference's parse/unimplemented_instructions.rs)");
        }
        i += wc;
    }
    if (fn) throw Error("SPIR-V: missing OpFunctionEnd");
    for (auto& e : entries) {
        if (e.name == entry_point && e.model == execution_model) {
            auto it = functions.find(e.id);
            if (it == functions.end()) throw Error("SPIR-V: entry point is not a function");
            m->entry_point = it->second;
            auto ls = local_sizes.find(e.id);
            if (ls != local_sizes.end() && local_size)
                for (int k = 0; k < 3; ++k) local_size[k] = ls->second[k];
        }
    }
    if (!m->entry_point) throw Error("SPIR-V: no entry point named '" + entry_point + "' with the requested execution model");
    return m;
}

}  // namespace kir
