// ptx_emit.h — shader-compiler-ir -> PTX (sm_100a) and SPIR-V -> shader-compiler-ir, internal interfaces.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "ir.h"

namespace kir {

// a construct the IR can express but this back end does not lower (maps to VK_ERROR_FEATURE_NOT_PRESENT)
struct Unsupported : Error {
    using Error::Error;
};

struct EntryInfo {
    std::string entry_name;  // PTX symbol of the kernel
    uint64_t inputs_size = 0, outputs_size = 0;
};

// PTX symbol for an IR / backend function name (backend.rs:444-459 allows '.', '-', '$' which PTX does not)
std::string ptx_mangle(const std::string& name);
// the module's entry point as one `.visible .entry`; ABI in include/kazan_b200_backend.h
std::string emit_ptx(const Module& m, EntryInfo* info);

// shader-compiler-translate-spirv-to-ir/src/lib.rs:190-215 for the shaders that crate translates
// (control-flow-free entry points); throws Unsupported for what the reference leaves todo!()
std::unique_ptr<Module> translate_spirv(const uint32_t* words, size_t count, const std::string& entry_point,
                                        uint32_t execution_model, uint32_t local_size[3]);

// spirv-parser's parse-and-dump (spirv-parser/src/lib.rs:26-44): header block + one line per instruction with its byte
// offset, in the reference's exact text format; throws Error with the reference's messages for malformed modules
std::string disassemble_spirv(const uint32_t* words, size_t count);

}  // namespace kir
