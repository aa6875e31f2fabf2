/* kazan_b200_backend.h — C ABI of `shader-compiler-backend-ptx`: the shader-compiler-ir -> PTX (sm_100a) back end.
 *
 * It replaces, for the B200, what shader-compiler-backend-llvm-7 is for the host CPU:
 *   - the IR it consumes is shader-compiler-ir (module.rs:128-150), exchanged in the `.kazan-ir` text format
 *     (text.rs FromText/ToText) because a C ABI cannot carry Rust's arena types;
 *   - kzn_translate_spirv_to_ir mirrors TranslatedSPIRVShader::new
 *     (shader-compiler-translate-spirv-to-ir/src/lib.rs:190-215) for the shaders that crate can translate;
 *   - kzn_compute_pipeline_* / kzn_cmd_dispatch replace ComputePipeline::new (shader-compiler/src/lib.rs:753-821) and
 *     vkCmdDispatch / vkCmdDispatchBase (vulkan-driver/src/api_impl.rs:4677-4684, :5079-5089);
 *   - kzn_backend_* mirrors the LLVM-builder-shaped trait surface of shader-compiler-backend/src/lib.rs:15-257
 *     (Context, Module, Function, BasicBlock, Builder with its two instructions alloca/ret, TypeBuilder, verify,
 *     Compiler::run, CompiledCode::get) verb for verb.
 * Code is loaded with the CUDA driver API (cuModuleLoadDataEx / cuModuleGetFunction / cuLaunchKernel), which the library
 * dlopen()s on first use, so everything that only produces PTX text works on a machine without a GPU.
 *
 * Invocation ABI of a compiled entry point (defined by this back end; the reference never got that far):
 *   one invocation == one thread, 256 invocations per CTA whatever the SPIR-V LocalSize;
 *   kernel parameters (user_inputs_base, user_outputs_base, invocation_count, first_invocation);
 *   invocation g (0 <= g < invocation_count) sees  user_inputs_block  = user_inputs_base  + (first+g) * inputs.size
 *                                                  user_outputs_block = user_outputs_base + (first+g) * outputs.size;
 *   the values the entry point returns (`break <body>[v...]`) are stored to user_outputs_block at the offsets of the
 *   block's members, in order — the IR has no store instruction (instructions_impl.rs:395-433), this is its only output.
 */
#ifndef KAZAN_B200_BACKEND_H
#define KAZAN_B200_BACKEND_H

#include "kazan_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shader-compiler-ir text (text.rs:1392 FromText, :1987 ToText) ------------------------------------------------ */
typedef struct KznIrModule KznIrModule;
KZN_API KznResult kzn_ir_parse(const char* text, size_t length, KznIrModule** module);
KZN_API void kzn_ir_module_destroy(KznIrModule* module);
/* canonical text; *text is malloc()ed, release with kzn_free */
KZN_API KznResult kzn_ir_to_text(const KznIrModule* module, char** text, size_t* length);
/* parse one syntactic category and print it canonically, exactly what the reference's round-trip unit tests do.
 * kind: "type" "const" "target_properties" "block" "function" "definitions_then_instruction" "module" */
KZN_API KznResult kzn_ir_round_trip(const char* kind, const char* text, size_t length, char** out, size_t* out_length);
KZN_API void kzn_free(void* ptr);

/* ---- SPIR-V -> IR (shader-compiler-translate-spirv-to-ir/src/lib.rs:190-215) ---------------------------------------- */
#define KZN_EXECUTION_MODEL_VERTEX 0u
#define KZN_EXECUTION_MODEL_FRAGMENT 4u
#define KZN_EXECUTION_MODEL_GLCOMPUTE 5u
/* Translates what the reference can translate: control-flow-free entry points (OpLabel .. OpReturn).  Anything the
 * reference leaves `todo!()` (parse/unimplemented_instructions.rs:14-39) returns KZN_ERROR_FEATURE_NOT_PRESENT.
 * local_size receives the LocalSize execution mode (1,1,1 when absent). */
KZN_API KznResult kzn_translate_spirv_to_ir(const uint32_t* words, size_t word_count, const char* entry_point,
                                            uint32_t execution_model, KznIrModule** module, uint32_t local_size[3]);

/* ---- SPIR-V reader (spirv-parser crate: Parser::start / iteration / Display, spirv-parser/src/lib.rs:26-44) --------- */
/* Parses a SPIR-V module and writes its disassembly in the reference's text format: "; SPIR-V / ; Version / ; Generator
 * / ; Bound / ; Schema" header lines, then one line per instruction ("%id = " right-aligned in 15 columns, operands,
 * " ; 0x%08x" byte offset).  Reproduces the reference's goldens for test_inputs/test{,2,3}.spv byte for byte
 * (spirv-parser/src/lib.rs:46-215).  Malformed input returns KZN_ERROR_VALIDATION_FAILED with the reference's message in
 * kzn_last_error().  *text is malloc()ed, release with kzn_free. */
KZN_API KznResult kzn_spirv_disassemble(const uint32_t* words, size_t word_count, char** text, size_t* length);

/* ---- IR -> PTX ------------------------------------------------------------------------------------------------------ */
/* PTX text for sm_100a of the module's entry point; needs no GPU.  *ptx is malloc()ed, release with kzn_free. */
KZN_API KznResult kzn_ptx_from_ir(const KznIrModule* module, char** ptx, size_t* length);

/* ---- compute pipelines and dispatch ----------------------------------------------------------------------------------- */
typedef struct KznPipeline KznPipeline; /* vulkan-driver/src/pipeline.rs:40-43 `ComputePipeline` */
KZN_API KznResult kzn_compute_pipeline_create(KznDevice* device, const KznIrModule* module, const uint32_t local_size[3],
                                              KznPipeline** pipeline);
KZN_API void kzn_pipeline_destroy(KznPipeline* pipeline);
/* vkCmdDispatchBase (api_impl.rs:5079-5089); vkCmdDispatch is base = {0,0,0}.  Invocations = product(group_count) *
 * product(local_size); `first_invocation` offsets the interface-block slots (used when a dispatch is sharded over GPUs). */
KZN_API KznResult kzn_cmd_dispatch(KznQueue* queue, KznPipeline* pipeline, const uint32_t base_group[3],
                                   const uint32_t group_count[3], uint64_t user_inputs_address,
                                   uint64_t user_outputs_address, uint64_t first_invocation);
KZN_API uint64_t kzn_pipeline_inputs_size(const KznPipeline* pipeline);  /* bytes per invocation */
KZN_API uint64_t kzn_pipeline_outputs_size(const KznPipeline* pipeline);

/* ---- the backend trait surface (shader-compiler-backend/src/lib.rs) ------------------------------------------------------ */
typedef struct KznBackendContext KznBackendContext;   /* trait Context  lib.rs:162-189 */
typedef struct KznBackendModule KznBackendModule;     /* trait Module   lib.rs:132-145 */
typedef struct KznBackendFunction KznBackendFunction; /* trait Function lib.rs:113-124 */
typedef struct KznBackendBlock KznBackendBlock;       /* trait BasicBlock / BuildableBasicBlock lib.rs:96-111 */
typedef struct KznBackendBuilder KznBackendBuilder;   /* traits DetachedBuilder / AttachedBuilder lib.rs:15-41 */
typedef struct KznBackendType KznBackendType;         /* trait Type, built by TypeBuilder types.rs:113-153 */
typedef struct KznBackendValue KznBackendValue;       /* trait Value lib.rs:80-94 */
typedef struct KznCompiledCode KznCompiledCode;       /* trait CompiledCode lib.rs:200-205 */

typedef struct KznLaunchDims {
    uint32_t grid[3], block[3];
} KznLaunchDims;
/* CompiledCode::get returns `unsafe extern "C" fn()` that "needs to be cast to the correct type" (lib.rs:204): here the
 * callable is a device kernel, so the host-side function pointer is a trampoline taking the launch shape and the
 * kernel's arguments; `self` carries the kernel. */
typedef struct KznCallable {
    KznResult (*call)(void* self, KznQueue* queue, const KznLaunchDims* dims, void** arguments);
    void* self;
} KznCallable;

/* Compiler::name (lib.rs:249): "PTX (sm_100a)" where the reference's only back end answers "LLVM 7" */
KZN_API const char* kzn_backend_compiler_name(void);
KZN_API KznResult kzn_backend_context_create(KznBackendContext** context); /* the Context handed to CompilerUser::run */
KZN_API void kzn_backend_context_destroy(KznBackendContext* context);
/* TypeBuilder (types.rs:113-153) */
KZN_API KznBackendType* kzn_backend_type_bool(KznBackendContext* c);
KZN_API KznBackendType* kzn_backend_type_int(KznBackendContext* c, uint32_t bits, int32_t is_signed); /* i8..i64, u8..u64 */
KZN_API KznBackendType* kzn_backend_type_f32(KznBackendContext* c);
KZN_API KznBackendType* kzn_backend_type_f64(KznBackendContext* c);
KZN_API KznBackendType* kzn_backend_type_pointer(KznBackendContext* c, KznBackendType* target);
KZN_API KznBackendType* kzn_backend_type_array(KznBackendContext* c, KznBackendType* element, uint64_t count);
KZN_API KznBackendType* kzn_backend_type_vector(KznBackendContext* c, KznBackendType* element, uint64_t length);
KZN_API KznBackendType* kzn_backend_type_struct(KznBackendContext* c, KznBackendType* const* members, uint64_t count);
/* return_type NULL = void */
KZN_API KznBackendType* kzn_backend_type_function(KznBackendContext* c, KznBackendType* return_type,
                                                  KznBackendType* const* arguments, uint64_t count);
KZN_API uint64_t kzn_backend_type_size(const KznBackendType* type);
/* Context::create_module / Module::add_function / Function::append_new_basic_block / Context::create_builder */
KZN_API KznBackendModule* kzn_backend_create_module(KznBackendContext* c, const char* name);
KZN_API KznResult kzn_backend_module_set_source_file_name(KznBackendModule* m, const char* source_file_name); /* lib.rs:136 */
/* names must match [A-Za-z_.$-][A-Za-z0-9_.$-]* and be unique, as backend.rs:444-459 checks; NULL otherwise */
KZN_API KznBackendFunction* kzn_backend_add_function(KznBackendModule* m, const char* name, KznBackendType* function_type);
KZN_API KznBackendBlock* kzn_backend_append_new_basic_block(KznBackendFunction* f, const char* name); /* name may be NULL */
KZN_API uint64_t kzn_backend_function_parameter_count(const KznBackendFunction* f);                   /* Function::parameters lib.rs:123 */
KZN_API KznBackendValue* kzn_backend_function_parameter(KznBackendFunction* f, uint64_t index);
KZN_API KznBackendValue* kzn_backend_function_as_value(KznBackendFunction* f); /* Function::as_value lib.rs:116 */
KZN_API KznBackendValue* kzn_backend_block_as_value(KznBackendBlock* b);       /* BasicBlock::as_value lib.rs:99 */
KZN_API KznBackendBuilder* kzn_backend_create_builder(KznBackendContext* c);
KZN_API KznResult kzn_backend_builder_attach(KznBackendBuilder* b, KznBackendBlock* block);
KZN_API KznBackendBlock* kzn_backend_builder_current_basic_block(KznBackendBuilder* b); /* lib.rs:19; NULL when detached */
/* AttachedBuilder::build_alloca (lib.rs:25) / build_return (lib.rs:26-29); value NULL = `ret void` */
KZN_API KznBackendValue* kzn_backend_build_alloca(KznBackendBuilder* b, KznBackendType* variable_type);
KZN_API KznResult kzn_backend_build_return(KznBackendBuilder* b, KznBackendValue* value);
/* Module::verify (lib.rs:141): structural checks + every block terminated; KZN_ERROR_VALIDATION_FAILED with text in
 * kzn_last_error otherwise */
KZN_API KznResult kzn_backend_module_verify(KznBackendModule* m);
/* PTX text of the module (malloc()ed) — no GPU needed */
KZN_API KznResult kzn_backend_module_to_ptx(KznBackendModule* m, char** ptx, size_t* length);
/* Compiler::run (lib.rs:252-256): compile the module for `device` and make `callable_functions` available by key */
KZN_API KznResult kzn_backend_compile(KznDevice* device, KznBackendModule* m, const char* const* keys,
                                      KznBackendFunction* const* callable_functions, uint64_t count, KznCompiledCode** code);
KZN_API KznResult kzn_compiled_code_get(KznCompiledCode* code, const char* key, KznCallable* callable); /* lib.rs:204 */
KZN_API void kzn_compiled_code_destroy(KznCompiledCode* code);

#ifdef __cplusplus
}
#endif
#endif
