// ptx_emit.h — shader-compiler-ir -> PTX (sm_100a) and SPIR-V -> shader-compiler-ir, internal interfaces.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "ir.h"

namespace kir {

// a construct the IR can express but this back end does not lower (maps to VK_ERROR_FEATURE_NOT_PRESENT)
struct Unsupported : Error {
    using Error::Error;
};

struct EntryInfo {
    std::string entry_name;  // PTX symbol of the kernel
    uint64_t inputs_size = 0, outputs_size = 0;
    bool uses_global_invocation_id = false;
};

// One place where the raster kernels call the fragment program: operands of the surrounding PTX code.
struct FragmentSite {
    int id = 0;                    // makes the copy's registers and labels unique
    int fragments = 1;             // 1, or 2: a call that shades two fragments at once (packed f32x2 arithmetic where the shader allows)
    // per fragment 8 operands: frag_coord x y z w, interpolated inputs 0..3; then the push constant pointer and the
    // descriptor table pointer: 8 * fragments + 2
    std::vector<std::string> in;
    std::vector<std::string> out;  // 4 * fragments registers that receive the colours ("" = unused)
    // when the call is inside a kernel entry: "<entry>_param_0+<offset>" of the push constant block in the kernel's
    // parameter space, so that the shader's push constant loads are ld.param (constant bank operands); otherwise the block
    // is read through the generic pointer in[8]
    std::string push_constants_param;
};
struct FragmentInterface {
    bool reads_inputs = false;      // the shader loads from user_inputs_block (the rasteriser must interpolate)
    bool reads_frag_coord = false;
    bool reads_frag_coord_zw = false;  // depth or 1/w: the rasteriser must evaluate the fragment's barycentrics
    bool pairs_share_code = false;     // two fragments run as one copy of the code (no fragment-dependent control flow, no memory writes)
};

// PTX symbol for an IR / backend function name (backend.rs:444-459 allows '.', '-', '$' which PTX does not)
std::string ptx_mangle(const std::string& name);
// the module's entry point as one `.visible .entry`; ABI in include/kazan_b200_backend.h
std::string emit_ptx(const Module& m, EntryInfo* info);
// a fragment shader's entry point as one PTX scope for `site` (see ptx_emit.cpp); fragment_interface checks that the module
// has the shape of a fragment shader (<= 4 f32 inputs, one vec4 colour output) and tells what it reads
std::string emit_fragment_inline(const Module& m, const FragmentSite& site, FragmentInterface* info);
FragmentInterface fragment_interface(const Module& m);

// shader-compiler-translate-spirv-to-ir/src/lib.rs:190-215 for the shaders that crate translates
// (control-flow-free entry points); throws Unsupported for what the reference leaves todo!()
std::unique_ptr<Module> translate_spirv(const uint32_t* words, size_t count, const std::string& entry_point,
                                        uint32_t execution_model, uint32_t local_size[3]);

// ... and, into version 2 of the IR, shaders with a body (spirv_to_ir_v2.cpp): loads, stores, access chains, composites,
// conversions, arithmetic, selection constructs (OpSwitch / OpBranchConditional)
std::unique_ptr<Module> translate_spirv_v2(const uint32_t* words, size_t count, const std::string& entry_point,
                                           uint32_t execution_model, uint32_t local_size[3]);

// spirv-parser's parse-and-dump (spirv-parser/src/lib.rs:26-44): header block + one line per instruction with its byte
// offset, in the reference's exact text format; throws Error with the reference's messages for malformed modules
std::string disassemble_spirv(const uint32_t* words, size_t count);

}  // namespace kir
