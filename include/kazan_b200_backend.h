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
 *   kernel parameters (user_inputs_base, user_outputs_base, invocation_count, first_invocation, global_size_x,
 *   global_size_y, descriptor_table, global_base[16], push_constants[128]);
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
/* Version 1 is the reference's instruction set to the letter (instructions_impl.rs:395-433: add, load, break, branch,
 * continue, block, loop).  Version 2 adds, in the C++ mirror only, what a shader with a body needs:
 * instructions spelled like the SPIR-V instruction they translate, in the reference's `name [args] -> [results]` syntax —
 *   i_sub i_mul bitwise_and bitwise_or bitwise_xor shift_left_logical shift_right_logical shift_right_arithmetic s_negate not
 *   f_add f_sub f_mul f_div f_min f_max f_negate f_abs fma
 *   i_equal i_not_equal u_less_than s_less_than u_less_than_equal s_less_than_equal u_greater_than s_greater_than
 *   u_greater_than_equal s_greater_than_equal f_ord_equal f_ord_less_than f_ord_less_than_equal f_ord_greater_than
 *   f_ord_greater_than_equal f_unord_not_equal logical_and logical_or logical_not select
 *   convert_f_to_u convert_f_to_s convert_s_to_f convert_u_to_f u_convert s_convert f_convert bitcast
 *   store [pointer, value], align: A -> []      ptr_add [pointer, offset: i64] -> [q: data_ptr]
 *   group_non_uniform_elect / _all / _any / _ballot / _broadcast_first / _shuffle / _shuffle_xor / _i_add
 *     (subgroup = 32 consecutive invocations = one warp: vote.sync / shfl.sync / redux.sync; subgroup-uniform control flow)
 * and the built-in kinds global_invocation_id <3 x i32>, subgroup_local_invocation_id i32, subgroup_size i32,
 * frag_coord <4 x f32>, push_constants data_ptr, descriptor_table data_ptr.  Semantics: csrc/ir.h; checker: oracle/ir_simt.py. */
#define KZN_IR_VERSION_1 1u
#define KZN_IR_VERSION_2 2u
KZN_API KznResult kzn_ir_parse_ex(const char* text, size_t length, uint32_t version, KznIrModule** module);
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
/* version KZN_IR_VERSION_2: shaders with a body — what the reference leaves todo!() (OpLoad / OpStore / OpAccessChain,
 * composites, conversions, arithmetic, compares, OpSelect, selection constructs with OpSwitch / OpBranchConditional, OpPhi at
 * their merge blocks, structured loops (OpLoopMerge -> the IR's `loop`), function calls (inlined: the IR has no call;
 * recursion returns VK_ERROR_FEATURE_NOT_PRESENT)) translated into the version 2 instructions of the IR
 * (csrc/spirv_to_ir_v2.cpp), values scalarised.  Vectors have 2 .. 4 components, arrays at most 1024 elements.
 * Input / Output variables are laid out by Location (16 bytes each, io_layout.rs:5-13) in the two interface blocks, a
 * `BuiltIn Position` output is the `vertex_position` built-in variable.  The reference's own second test shader
 * (test_data/simple_test.spv, src/lib.rs:260-287: #[ignore], golden "not yet entered") translates. */
KZN_API KznResult kzn_translate_spirv_to_ir_ex(const uint32_t* words, size_t word_count, const char* entry_point,
                                               uint32_t execution_model, uint32_t version, KznIrModule** module,
                                               uint32_t local_size[3]);

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

/* ---- descriptor sets and push constants -------------------------------------------------------------- */
/* vulkan-driver/src/descriptor_set.rs: DescriptorLayout :12-49, Descriptor :230-267, Descriptor::write :460-527;
 * vkAllocateDescriptorSets api_impl.rs:4218, vkUpdateDescriptorSets :4339, vkCmdBindDescriptorSets :4597,
 * vkCmdPushConstants :4932 (the two commands are unimplemented!() upstream); PipelineLayout shader-compiler/src/lib.rs:706-730.
 * Kernel-argument ABI: a bound set reaches the shader as ONE pointer (built-in kind `descriptor_table`) to a table of
 * 16-byte slots {uint64 address, uint64 range}, one per descriptor, bindings in layout order, array elements consecutive
 * (kzn_descriptor_set_slot gives a binding's first slot); the push constant block (128 bytes, maxPushConstantsSize) travels
 * in the launch's parameter space and reaches the shader through the built-in kind `push_constants`. */
#define KZN_DESCRIPTOR_TYPE_UNIFORM_BUFFER 6u /* VkDescriptorType */
#define KZN_DESCRIPTOR_TYPE_STORAGE_BUFFER 7u
#define KZN_PIPELINE_BIND_POINT_GRAPHICS 0u /* VkPipelineBindPoint */
#define KZN_PIPELINE_BIND_POINT_COMPUTE 1u
typedef struct KznDescriptorSet KznDescriptorSet; /* descriptor_set.rs `DescriptorSet` + its layout */
typedef struct KznDescriptorSetLayoutBinding {   /* VkDescriptorSetLayoutBinding without samplers and stage flags */
    uint32_t binding, descriptor_type, descriptor_count;
} KznDescriptorSetLayoutBinding;
typedef struct KznDescriptorBufferInfo { /* VkDescriptorBufferInfo with the buffer's device address for the handle */
    uint64_t address, offset, range;
} KznDescriptorBufferInfo;
KZN_API KznResult kzn_descriptor_set_create(KznDevice* device, uint32_t binding_count, const KznDescriptorSetLayoutBinding* bindings,
                                            KznDescriptorSet** set);
KZN_API void kzn_descriptor_set_destroy(KznDescriptorSet* set);
/* vkUpdateDescriptorSets with pBufferInfo: elements first_element .. first_element + count - 1 of `binding` */
KZN_API KznResult kzn_descriptor_set_write_buffers(KznDescriptorSet* set, uint32_t binding, uint32_t first_element, uint32_t count,
                                                   const KznDescriptorBufferInfo* buffers);
KZN_API uint64_t kzn_descriptor_set_table_address(const KznDescriptorSet* set);
KZN_API int64_t kzn_descriptor_set_slot(const KznDescriptorSet* set, uint32_t binding); /* -1: not in the layout */
/* state of the queue's command stream, consumed by the next kzn_cmd_dispatch (COMPUTE) / kzn_cmd_draw* (GRAPHICS) */
KZN_API KznResult kzn_cmd_bind_descriptor_set(KznQueue* queue, uint32_t pipeline_bind_point, const KznDescriptorSet* set);
KZN_API KznResult kzn_cmd_push_constants(KznQueue* queue, uint32_t offset, uint32_t size, const void* values);

/* ---- fragment programs: the fragment stage of a graphics pipeline ------------------------------------------------------------ */
/* vkCreateGraphicsPipelines api_impl.rs:4018-4032 -> GraphicsPipeline::create pipeline.rs:167-183 (unimplemented!() upstream),
 * vkCmdBindPipeline :4506.  The module is a fragment shader in shader-compiler-ir: user_inputs_block = up to four f32
 * interpolated inputs (the passthrough vertex stage's colour, perspective correct), user_outputs_block = four f32 (the
 * colour of attachment 0), built-in inputs frag_coord / push_constants / descriptor_table.  Creation generates the shader's
 * PTX and builds it INTO the raster kernels (their PTX is part of this library; every call of the fragment program is
 * replaced by the shader's code), so a compiled program costs what a built-in one costs.  A draw whose state says
 * fragment_program == KZN_FS_COMPILED runs the program bound to the queue. */
typedef struct KznFragmentProgram KznFragmentProgram;
KZN_API KznResult kzn_fragment_program_create(KznDevice* device, const KznIrModule* module, KznFragmentProgram** program);
KZN_API void kzn_fragment_program_destroy(KznFragmentProgram* program);
KZN_API KznResult kzn_cmd_bind_fragment_program(KznQueue* queue, KznFragmentProgram* program);
/* the PTX of the fragment shader alone as it is placed at one call site (malloc()ed; needs no GPU) */
KZN_API KznResult kzn_fragment_program_ptx(const KznIrModule* module, char** ptx, size_t* length);
/* ... and of the whole module kzn_fragment_program_create loads: the raster kernels with the shader at every call site */
KZN_API KznResult kzn_fragment_program_module_ptx(const KznIrModule* module, char** ptx, size_t* length);

/* ---- vertex programs: the vertex stage of a graphics pipeline ---------------------------------------------------------------- */
/* vkCreateGraphicsPipelines api_impl.rs:4018-4032 (vertex stage of GraphicsPipeline::create, pipeline.rs:167-183, unimplemented!()
 * upstream).  The module is a vertex shader in version 2 of the IR — what kzn_translate_spirv_to_ir_ex makes of a SPIR-V
 * vertex shader: user_inputs_block = the attributes of the vertex binding in the shader's Location layout (16 bytes per
 * location; the binding's stride must equal the block's size), a built-in output of kind vertex_position (four f32), and
 * optionally user output location 0 = four f32, the colour the fragment stage interpolates (white without it).  A draw whose
 * state says vertex_program == KZN_VS_COMPILED runs the program bound to the queue as a dispatch of the PTX back end over the
 * vertices the draw can name (vertices first_vertex .. + vertex_count of a non-indexed draw, every whole vertex of the
 * binding for an indexed one) into the queue's vertex scratch, and rasterises from there; push constants and descriptor set
 * are the GRAPHICS bind point's.  kzn_queue_reserve_vertex_outputs sizes that scratch ahead of a recorded draw. */
typedef struct KznVertexProgram KznVertexProgram;
KZN_API KznResult kzn_vertex_program_create(KznDevice* device, const KznIrModule* module, KznVertexProgram** program);
KZN_API void kzn_vertex_program_destroy(KznVertexProgram* program);
KZN_API KznResult kzn_cmd_bind_vertex_program(KznQueue* queue, KznVertexProgram* program);
KZN_API uint64_t kzn_vertex_program_inputs_size(const KznVertexProgram* program);  /* = the vertex binding's stride */
KZN_API uint64_t kzn_vertex_program_outputs_size(const KznVertexProgram* program); /* bytes per vertex in the scratch */
KZN_API KznResult kzn_queue_reserve_vertex_outputs(KznQueue* queue, uint64_t max_vertices, const KznVertexProgram* program);

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
