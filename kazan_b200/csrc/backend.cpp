// backend.cpp — C ABI of `shader-compiler-backend-ptx` (include/kazan_b200_backend.h).
//
// Replaces / mirrors (paths in the kazan-3d/kazan tree):
//   shader-compiler-backend/src/lib.rs:15-257      the backend trait surface (Context, Module, Function, BasicBlock,
//                                                  Attached/DetachedBuilder, Compiler::run, CompiledCode::get)
//   shader-compiler-backend/src/types.rs:113-153   TypeBuilder
//   shader-compiler-backend-llvm-7/src/backend.rs  the LLVM-7 implementation of that surface (:444-459 name rule,
//                                                  :477 verify, :553-690 Compiler::run, :645-657 symbol lookup)
//   shader-compiler/src/lib.rs:753-821             ComputePipeline::new
//   vulkan-driver/src/api_impl.rs:4677-4684,:5079-5089   vkCmdDispatch / vkCmdDispatchBase (unimplemented!())
// Code is loaded with the CUDA *driver* API (cuModuleLoadDataEx / cuModuleGetFunction / cuLaunchKernel), resolved with
// dlopen at first use so that everything that only produces PTX text works without a GPU or a driver.
#include <dlfcn.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/kazan_b200_backend.h"
#include "kzn_internal.h"
#include "ptx_emit.h"

// ---- CUDA driver API, bound at run time -------------------------------------------------------------------------------
namespace {

typedef int CUresult_;
typedef void* CUmodule_;
typedef void* CUfunction_;
typedef void* CUcontext_;
enum { CU_JIT_INFO_LOG_BUFFER_ = 3, CU_JIT_INFO_LOG_BUFFER_SIZE_BYTES_ = 4, CU_JIT_ERROR_LOG_BUFFER_ = 5, CU_JIT_ERROR_LOG_BUFFER_SIZE_BYTES_ = 6 };

struct DriverApi {
    void* handle = nullptr;
    CUresult_ (*cuInit)(unsigned) = nullptr;
    CUresult_ (*cuCtxGetCurrent)(CUcontext_*) = nullptr;
    CUresult_ (*cuModuleLoadDataEx)(CUmodule_*, const void*, unsigned, int*, void**) = nullptr;
    CUresult_ (*cuModuleGetFunction)(CUfunction_*, CUmodule_, const char*) = nullptr;
    CUresult_ (*cuModuleUnload)(CUmodule_) = nullptr;
    CUresult_ (*cuLaunchKernel)(CUfunction_, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    CUresult_ (*cuLaunchKernelEx)(const void* config, CUfunction_, void**, void**) = nullptr;  // optional: programmatic dependent launch
    CUresult_ (*cuFuncSetAttribute)(CUfunction_, int, int) = nullptr;
    CUresult_ (*cuGetErrorString)(CUresult_, const char**) = nullptr;
    std::string error;
};

DriverApi* driver() {
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libcuda.so.1", "libcuda.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = "libcuda.so.1 cannot be loaded: no NVIDIA driver on this machine";
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.handle, n);
            if (!p && api.error.empty()) api.error = std::string("libcuda is missing ") + n;
            return p;
        };
        api.cuInit = (decltype(api.cuInit))sym("cuInit");
        api.cuCtxGetCurrent = (decltype(api.cuCtxGetCurrent))sym("cuCtxGetCurrent");
        api.cuModuleLoadDataEx = (decltype(api.cuModuleLoadDataEx))sym("cuModuleLoadDataEx");
        api.cuModuleGetFunction = (decltype(api.cuModuleGetFunction))sym("cuModuleGetFunction");
        api.cuModuleUnload = (decltype(api.cuModuleUnload))sym("cuModuleUnload");
        api.cuLaunchKernel = (decltype(api.cuLaunchKernel))sym("cuLaunchKernel");
        api.cuFuncSetAttribute = (decltype(api.cuFuncSetAttribute))sym("cuFuncSetAttribute");
        api.cuGetErrorString = (decltype(api.cuGetErrorString))sym("cuGetErrorString");
        api.cuLaunchKernelEx = (decltype(api.cuLaunchKernelEx))dlsym(api.handle, "cuLaunchKernelEx");
    });
    return &api;
}

KznResult driver_failed(CUresult_ r, const char* what) {
    const char* s = nullptr;
    DriverApi* d = driver();
    if (d->cuGetErrorString) d->cuGetErrorString(r, &s);
    kzn_set_error("CUDA driver error %d (%s) in %s", r, s ? s : "?", what);
    return r == 100 /* NO_DEVICE */ || r == 3 /* NOT_INITIALIZED */ ? KZN_ERROR_INITIALIZATION_FAILED : KZN_ERROR_DEVICE_LOST;
}

// PTX text -> module in the device's primary context
KznResult load_ptx(KznDevice* device, const std::string& ptx, CUmodule_* module) {
    KZN_VALID(device, "null device");
    DriverApi* d = driver();
    KZN_REQUIRE(d->error.empty(), KZN_ERROR_INITIALIZATION_FAILED, "%s", d->error.c_str());
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KZN_CUDA(cudaFree(0));
    CUcontext_ ctx = nullptr;
    CUresult_ r = d->cuCtxGetCurrent(&ctx);
    if (r != 0) return driver_failed(r, "cuCtxGetCurrent");
    KZN_REQUIRE(ctx, KZN_ERROR_INITIALIZATION_FAILED, "no current CUDA context");
    static char info[8192], err[8192];  // (one compile at a time: pipeline creation is externally synchronised)
    info[0] = err[0] = 0;
    int opts[4] = {CU_JIT_INFO_LOG_BUFFER_, CU_JIT_INFO_LOG_BUFFER_SIZE_BYTES_, CU_JIT_ERROR_LOG_BUFFER_, CU_JIT_ERROR_LOG_BUFFER_SIZE_BYTES_};
    void* vals[4] = {info, (void*)(uintptr_t)sizeof info, err, (void*)(uintptr_t)sizeof err};
    r = d->cuModuleLoadDataEx(module, ptx.c_str(), 4, opts, vals);
    if (r != 0) {
        const char* s = nullptr;
        d->cuGetErrorString(r, &s);
        kzn_set_error("cuModuleLoadDataEx failed: %d (%s): %.300s", r, s ? s : "?", err);
        return KZN_ERROR_DEVICE_LOST;
    }
    return KZN_SUCCESS;
}

char* dup_text(const std::string& s, size_t* length) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.c_str(), s.size() + 1);
    if (length) *length = s.size();
    return p;
}

// run `fn`, mapping the IR library's exceptions to result codes (text goes to kzn_last_error)
template <typename F>
KznResult guarded(F&& fn) {
    try {
        return fn();
    } catch (const kir::Unsupported& e) {
        kzn_set_error("%s", e.what());
        return KZN_ERROR_FEATURE_NOT_PRESENT;
    } catch (const kir::Error& e) {
        kzn_set_error("%s", e.what());
        return KZN_ERROR_VALIDATION_FAILED;
    } catch (const std::bad_alloc&) {
        kzn_set_error("out of host memory");
        return KZN_ERROR_OUT_OF_HOST_MEMORY;
    } catch (const std::exception& e) {
        kzn_set_error("%s", e.what());
        return KZN_ERROR_VALIDATION_FAILED;
    }
}

}  // namespace

struct KznIrModule {
    std::unique_ptr<kir::Module> module;
};

// A fragment shader built into the raster kernels (include/kazan_b200_backend.h).
struct KznFragmentProgram {
    KznDevice* device = nullptr;
    CUmodule_ module = nullptr;
    CUfunction_ kernel[2][3] = {};  // [raster_direct, raster][destination mode]
    uint32_t flags = 0;
    int call_sites = 0;
};

struct KznVertexProgram {
    struct KznPipeline* pipeline = nullptr;  // the shader as a kernel of the PTX back end, one invocation per vertex
    KznVertexLayout layout = {};
};

struct KznPipeline {
    KznDevice* device = nullptr;
    CUmodule_ module = nullptr;
    CUfunction_ function = nullptr;
    uint32_t local_size[3] = {1, 1, 1};
    uint64_t inputs_size = 0, outputs_size = 0;
    std::string ptx;
};

// vulkan-driver/src/descriptor_set.rs:230-267 `Descriptor` for the two kinds a kernel can read through a pointer
// (UniformBuffer / StorageBuffer hold BufferSlices, :460-527 `write`); the device-side form is a flat table of
// {address, size} pairs in binding order that the shader indexes (include/kazan_b200_backend.h)
struct KznDescriptorSet {
    KznDevice* device = nullptr;
    struct Binding {
        uint32_t binding, type, count, first_slot;
    };
    std::vector<Binding> bindings;
    std::vector<uint64_t> host;  // 2 words per slot
    void* table = nullptr;       // device copy
};

// ---- backend trait objects ----------------------------------------------------------------------------------------------
enum class BtKind { Void, Bool, Int, F32, F64, Pointer, Array, Vector, Struct, Function };

struct KznBackendType {
    BtKind kind;
    uint32_t bits = 0;
    bool is_signed = false;
    KznBackendType* element = nullptr;  // pointer target / array, vector element / function return (null = void)
    uint64_t count = 0;
    std::vector<KznBackendType*> members;  // struct members / function arguments
    uint64_t size = 0, align = 1;
};

enum class BvKind { Parameter, Alloca, Function, Block };
struct KznBackendValue {
    BvKind kind;
    KznBackendType* type = nullptr;
    KznBackendFunction* function = nullptr;
    uint32_t index = 0;  // parameter index / alloca slot
};

struct BackendInstr {
    bool is_return = false;
    KznBackendValue* value = nullptr;  // alloca result / returned value (may be null)
    KznBackendType* type = nullptr;    // alloca type
};

struct KznBackendBlock {
    KznBackendFunction* function = nullptr;
    std::string name;
    std::vector<BackendInstr> instrs;
    bool terminated = false;
    KznBackendValue* as_value = nullptr;
};

struct KznBackendFunction {
    KznBackendModule* module = nullptr;
    std::string name, symbol;
    KznBackendType* type = nullptr;
    std::vector<KznBackendValue*> parameters;
    std::vector<KznBackendBlock*> blocks;
    uint32_t allocas = 0;
    KznBackendValue* as_value = nullptr;
};

struct KznBackendModule {
    KznBackendContext* context = nullptr;
    std::string name, source_file_name;
    std::vector<KznBackendFunction*> functions;
};

struct KznBackendBuilder {
    KznBackendContext* context = nullptr;
    KznBackendBlock* block = nullptr;  // null = detached
};

struct KznBackendContext {  // owns everything created through it (the trait's 'a lifetime)
    std::vector<std::unique_ptr<KznBackendType>> types;
    std::vector<std::unique_ptr<KznBackendValue>> values;
    std::vector<std::unique_ptr<KznBackendBlock>> blocks;
    std::vector<std::unique_ptr<KznBackendFunction>> functions;
    std::vector<std::unique_ptr<KznBackendModule>> modules;
    std::vector<std::unique_ptr<KznBackendBuilder>> builders;

    KznBackendType* type(KznBackendType t) {
        types.push_back(std::make_unique<KznBackendType>(std::move(t)));
        return types.back().get();
    }
    KznBackendValue* value(KznBackendValue v) {
        values.push_back(std::make_unique<KznBackendValue>(v));
        return values.back().get();
    }
};

struct CompiledFunction {
    CUfunction_ function = nullptr;
    bool has_return_slot = false;
    size_t parameter_count = 0;
};

struct KznCompiledCode {
    KznDevice* device = nullptr;
    CUmodule_ module = nullptr;
    std::map<std::string, CompiledFunction> by_key;
};

namespace {

bool valid_function_name(const char* s) {  // backend.rs:444-459: [A-Za-z_.$-][A-Za-z0-9_.$-]*
    if (!s || !*s) return false;
    for (const char* p = s; *p; ++p) {
        unsigned char c = (unsigned char)*p;
        bool head = isalpha(c) || c == '_' || c == '.' || c == '$' || c == '-';
        if (!(head || (p != s && isdigit(c)))) return false;
    }
    return true;
}

std::string param_decl(const KznBackendType* t, const std::string& name) {
    std::ostringstream o;
    switch (t->kind) {
        case BtKind::Bool: o << ".param .u8 " << name; break;
        case BtKind::Int: o << ".param ." << (t->is_signed ? "s" : "u") << t->bits << " " << name; break;
        case BtKind::F32: o << ".param .f32 " << name; break;
        case BtKind::F64: o << ".param .f64 " << name; break;
        case BtKind::Pointer:
        case BtKind::Function: o << ".param .u64 " << name; break;
        default: o << ".param .align " << t->align << " .b8 " << name << "[" << (t->size ? t->size : 1) << "]"; break;
    }
    return o.str();
}

// PTX of a module built through the trait surface: every function is a kernel; a non-void function gets a trailing
// `.u64 kzn_return_slot` parameter where build_return stores its value
std::string backend_module_ptx(const KznBackendModule* m) {
    std::ostringstream o;
    o << "//\n// Generated by kazan_b200 shader-compiler-backend-ptx; module '" << m->name << "'";
    if (!m->source_file_name.empty()) o << ", source '" << m->source_file_name << "'";
    o << "\n//\n.version 8.7\n.target sm_100a\n.address_size 64\n";
    for (const KznBackendFunction* f : m->functions) {
        if (f->blocks.empty()) continue;  // a declaration
        const KznBackendType* ret = f->type->element;
        o << "\n// " << f->name << "\n.visible .entry " << f->symbol << "(";
        size_t np = f->type->members.size();
        for (size_t i = 0; i < np; ++i) o << (i ? ",\n    " : "\n    ") << param_decl(f->type->members[i], "kzn_arg" + std::to_string(i));
        if (ret) o << (np ? ",\n    " : "\n    ") << ".param .u64 kzn_return_slot";
        o << (np || ret ? "\n)\n{\n" : ")\n{\n");
        o << "    .reg .b16 %h<2>;\n    .reg .b32 %r<2>;\n    .reg .b64 %rd<4>;\n";
        uint32_t slot = 0;
        for (const KznBackendBlock* b : f->blocks)
            for (const BackendInstr& in : b->instrs)
                if (!in.is_return)
                    o << "    .local .align " << (in.type->align ? in.type->align : 1) << " .b8 kzn_alloca" << slot++ << "[" << (in.type->size ? in.type->size : 1) << "];\n";
        size_t bi = 0;
        for (const KznBackendBlock* b : f->blocks) {
            o << "$bb" << bi++ << "_" << kir::ptx_mangle(b->name) << ":\n";
            for (const BackendInstr& in : b->instrs) {
                if (!in.is_return) continue;
                if (in.value && ret) {
                    const KznBackendValue* v = in.value;
                    o << "    ld.param.u64 %rd0, [kzn_return_slot];\n";
                    if (v->kind == BvKind::Alloca) {
                        o << "    mov.u64 %rd1, kzn_alloca" << v->index << ";\n    cvta.local.u64 %rd1, %rd1;\n    st.u64 [%rd0], %rd1;\n";
                    } else if (v->kind == BvKind::Parameter) {
                        const KznBackendType* t = v->type;
                        std::string p = "[kzn_arg" + std::to_string(v->index) + "]";
                        uint64_t sz = t->size;
                        bool scalar = t->kind == BtKind::Bool || t->kind == BtKind::Int || t->kind == BtKind::F32 || t->kind == BtKind::F64 ||
                                      t->kind == BtKind::Pointer || t->kind == BtKind::Function;
                        if (scalar && sz == 8) o << "    ld.param.u64 %rd1, " << p << ";\n    st.u64 [%rd0], %rd1;\n";
                        else if (scalar && sz == 4) o << "    ld.param.u32 %r0, " << p << ";\n    st.u32 [%rd0], %r0;\n";
                        else if (scalar && sz == 2) o << "    ld.param.u16 %h0, " << p << ";\n    st.u16 [%rd0], %h0;\n";
                        else if (scalar && sz == 1) o << "    ld.param.u8 %h0, " << p << ";\n    st.u8 [%rd0], %h0;\n";
                        else
                            for (uint64_t k = 0; k < sz; ++k)
                                o << "    ld.param.u8 %h0, [kzn_arg" << v->index << "+" << k << "];\n    st.u8 [%rd0+" << k << "], %h0;\n";
                    }
                }
                o << "    ret;\n";
            }
        }
        o << "}\n";
    }
    return o.str();
}

KznResult verify_module(const KznBackendModule* m) {
    KZN_VALID(m, "null module");
    for (const KznBackendFunction* f : m->functions) {
        for (const KznBackendBlock* b : f->blocks) {
            KZN_VALID(b->terminated, "module verification failure: basic block '%s' of function '%s' does not have a terminator", b->name.c_str(),
                      f->name.c_str());
            for (const BackendInstr& in : b->instrs) {
                if (!in.is_return) continue;
                const KznBackendType* ret = f->type->element;
                KZN_VALID(!!ret == !!in.value, "module verification failure: function '%s' return value does not match its return type", f->name.c_str());
                if (in.value) {
                    KZN_VALID(in.value->kind == BvKind::Alloca || in.value->kind == BvKind::Parameter,
                              "module verification failure: function '%s' returns a value that is not first-class in this back end", f->name.c_str());
                    KZN_VALID(in.value->function == f, "module verification failure: function '%s' returns a value of another function", f->name.c_str());
                    KZN_VALID(in.value->type->size == ret->size, "module verification failure: function '%s' return value does not match its return type",
                              f->name.c_str());
                }
            }
        }
    }
    return KZN_SUCCESS;
}

KznResult callable_trampoline(void* self, KznQueue* queue, const KznLaunchDims* dims, void** arguments) {
    CompiledFunction* f = (CompiledFunction*)self;
    KZN_VALID(f && queue && dims, "KznCallable: null argument");
    KznResult r = kzn_queue_enter(queue);
    if (r != KZN_SUCCESS) return r;
    for (int i = 0; i < 3; ++i) KZN_VALID(dims->grid[i] > 0 && dims->block[i] > 0, "KznCallable: empty launch dimension");
    KZN_VALID(arguments || f->parameter_count + f->has_return_slot == 0, "KznCallable: null argument array");
    CUresult_ cr = driver()->cuLaunchKernel(f->function, dims->grid[0], dims->grid[1], dims->grid[2], dims->block[0], dims->block[1], dims->block[2], 0,
                                            (void*)queue->stream, arguments, nullptr);
    if (cr != 0) return driver_failed(cr, "cuLaunchKernel");
    queue->launches++;
    return KZN_SUCCESS;
}

}  // namespace

extern "C" {

// ---- IR text ------------------------------------------------------------------------------------------------------------

KznResult kzn_ir_parse_ex(const char* text, size_t length, uint32_t version, KznIrModule** module) {
    KZN_VALID(text && module, "kzn_ir_parse: null argument");
    *module = nullptr;
    KZN_VALID(version == KZN_IR_VERSION_1 || version == KZN_IR_VERSION_2, "kzn_ir_parse_ex: unknown IR version %u", version);
    return guarded([&] {
        auto m = kir::parse_module(std::string(text, length), (int)version);
        *module = new KznIrModule{std::move(m)};
        return KZN_SUCCESS;
    });
}
KznResult kzn_ir_parse(const char* text, size_t length, KznIrModule** module) { return kzn_ir_parse_ex(text, length, KZN_IR_VERSION_1, module); }

void kzn_ir_module_destroy(KznIrModule* module) { delete module; }

KznResult kzn_ir_to_text(const KznIrModule* module, char** text, size_t* length) {
    KZN_VALID(module && text, "kzn_ir_to_text: null argument");
    *text = nullptr;
    return guarded([&] {
        *text = dup_text(kir::module_to_text(*module->module), length);
        KZN_REQUIRE(*text, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

KznResult kzn_ir_round_trip(const char* kind, const char* text, size_t length, char** out, size_t* out_length) {
    KZN_VALID(kind && text && out, "kzn_ir_round_trip: null argument");
    *out = nullptr;
    return guarded([&] {
        *out = dup_text(kir::round_trip(kind, std::string(text, length)), out_length);
        KZN_REQUIRE(*out, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

void kzn_free(void* ptr) { free(ptr); }

// ---- SPIR-V -> IR -> PTX ---------------------------------------------------------------------------------------------------

KznResult kzn_translate_spirv_to_ir_ex(const uint32_t* words, size_t word_count, const char* entry_point, uint32_t execution_model,
                                       uint32_t version, KznIrModule** module, uint32_t local_size[3]) {
    KZN_VALID(words && entry_point && module, "kzn_translate_spirv_to_ir: null argument");
    *module = nullptr;
    KZN_VALID(version == KZN_IR_VERSION_1 || version == KZN_IR_VERSION_2, "kzn_translate_spirv_to_ir_ex: unknown IR version %u", version);
    return guarded([&] {
        auto m = version == KZN_IR_VERSION_2 ? kir::translate_spirv_v2(words, word_count, entry_point, execution_model, local_size)
                                             : kir::translate_spirv(words, word_count, entry_point, execution_model, local_size);
        *module = new KznIrModule{std::move(m)};
        return KZN_SUCCESS;
    });
}
KznResult kzn_translate_spirv_to_ir(const uint32_t* words, size_t word_count, const char* entry_point, uint32_t execution_model,
                                    KznIrModule** module, uint32_t local_size[3]) {
    return kzn_translate_spirv_to_ir_ex(words, word_count, entry_point, execution_model, KZN_IR_VERSION_1, module, local_size);
}

KznResult kzn_spirv_disassemble(const uint32_t* words, size_t word_count, char** text, size_t* length) {
    KZN_VALID(words && text, "kzn_spirv_disassemble: null argument");
    *text = nullptr;
    return guarded([&] {
        *text = dup_text(kir::disassemble_spirv(words, word_count), length);
        KZN_REQUIRE(*text, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

KznResult kzn_ptx_from_ir(const KznIrModule* module, char** ptx, size_t* length) {
    KZN_VALID(module && ptx, "kzn_ptx_from_ir: null argument");
    *ptx = nullptr;
    return guarded([&] {
        *ptx = dup_text(kir::emit_ptx(*module->module, nullptr), length);
        KZN_REQUIRE(*ptx, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

// ---- compute pipelines -------------------------------------------------------------------------------------------------------

KznResult kzn_compute_pipeline_create(KznDevice* device, const KznIrModule* module, const uint32_t local_size[3], KznPipeline** pipeline) {
    KZN_VALID(device && module && pipeline, "kzn_compute_pipeline_create: null argument");
    *pipeline = nullptr;
    auto p = std::make_unique<KznPipeline>();
    p->device = device;
    for (int i = 0; i < 3; ++i) {
        p->local_size[i] = local_size ? local_size[i] : 1;
        KZN_VALID(p->local_size[i] > 0, "LocalSize component %d is zero", i);
    }
    kir::EntryInfo info;
    KznResult r = guarded([&] {
        p->ptx = kir::emit_ptx(*module->module, &info);
        return KZN_SUCCESS;
    });
    if (r != KZN_SUCCESS) return r;
    p->inputs_size = info.inputs_size;
    p->outputs_size = info.outputs_size;
    r = load_ptx(device, p->ptx, &p->module);
    if (r != KZN_SUCCESS) return r;
    CUresult_ cr = driver()->cuModuleGetFunction(&p->function, p->module, info.entry_name.c_str());
    if (cr != 0) {
        driver()->cuModuleUnload(p->module);
        return driver_failed(cr, "cuModuleGetFunction");
    }
    *pipeline = p.release();
    return KZN_SUCCESS;
}

void kzn_pipeline_destroy(KznPipeline* pipeline) {
    if (!pipeline) return;
    if (pipeline->module) {
        cudaSetDevice(pipeline->device->ordinal);
        driver()->cuModuleUnload(pipeline->module);
    }
    delete pipeline;
}

uint64_t kzn_pipeline_inputs_size(const KznPipeline* pipeline) { return pipeline ? pipeline->inputs_size : 0; }
uint64_t kzn_pipeline_outputs_size(const KznPipeline* pipeline) { return pipeline ? pipeline->outputs_size : 0; }

KznResult kzn_cmd_dispatch(KznQueue* queue, KznPipeline* pipeline, const uint32_t base_group[3], const uint32_t group_count[3],
                           uint64_t user_inputs_address, uint64_t user_outputs_address, uint64_t first_invocation) {
    KZN_VALID(queue && pipeline && group_count, "kzn_cmd_dispatch: null argument");
    KznResult r = kzn_queue_enter(queue);
    if (r != KZN_SUCCESS) return r;
    KZN_VALID(queue->device->ordinal == pipeline->device->ordinal, "pipeline was created for another device");
    // invocation count = groups * LocalSize; the base group only renumbers WorkgroupId (api_impl.rs:5079-5089), which
    // the IR cannot observe (interface.rs:61-67 has no compute built-ins), so it does not move the interface slots
    unsigned __int128 total = 1;
    for (int i = 0; i < 3; ++i)
        if (group_count[i] == 0 || pipeline->local_size[i] == 0) return KZN_SUCCESS;  // a zero-sized dispatch is valid and does nothing
    for (int i = 0; i < 3; ++i) {
        if (base_group) KZN_VALID((uint64_t)base_group[i] + group_count[i] <= 0xffffffffull, "baseGroup + groupCount overflows in dimension %d", i);
        total *= group_count[i];
        total *= pipeline->local_size[i];
        // checked per dimension: six 32-bit factors can pass 2^128 and wrap around to something small
        KZN_VALID(total <= (unsigned __int128)1 << 62, "dispatch of more than 2^62 invocations");
    }
    uint64_t count = (uint64_t)total;
    KZN_VALID(pipeline->inputs_size == 0 || user_inputs_address, "pipeline reads user_inputs_block but no address was bound");
    KZN_VALID(pipeline->outputs_size == 0 || user_outputs_address, "pipeline writes user_outputs_block but no address was bound");
    // one invocation per thread, 256 per CTA whatever the LocalSize; CTAs folded into 2-D past the 2^31-1 x-limit
    const unsigned threads = 256;
    uint64_t ctas = (count + threads - 1) / threads;
    uint64_t gx = ctas, gy = 1;
    if (ctas > 0x7fffffffull) {
        gx = 1ull << 30;
        gy = (ctas + gx - 1) / gx;
        KZN_VALID(gy <= 65535, "dispatch too large for one launch");
    }
    // version 2 parameters: the global size (global_invocation_id is cut from the linear index, x fastest), the first
    // workgroup of vkCmdDispatchBase, and what vkCmdBindDescriptorSets / vkCmdPushConstants left on the queue
    uint64_t size_x = (uint64_t)group_count[0] * pipeline->local_size[0], size_y = (uint64_t)group_count[1] * pipeline->local_size[1];
    struct alignas(16) { uint32_t v[4]; } base = {{0, 0, 0, 0}};
    for (int i = 0; i < 3; ++i) base.v[i] = base_group ? base_group[i] * pipeline->local_size[i] : 0u;
    uint64_t table = queue->descriptor_table[KZN_PIPELINE_BIND_POINT_COMPUTE];
    void* args[9] = {&user_inputs_address, &user_outputs_address, &count, &first_invocation, &size_x, &size_y, &table, &base, queue->push_constants};
    CUresult_ cr = driver()->cuLaunchKernel(pipeline->function, (unsigned)gx, (unsigned)gy, 1, threads, 1, 1, 0, (void*)queue->stream, args, nullptr);
    if (cr != 0) return driver_failed(cr, "cuLaunchKernel");
    queue->launches++;
    return KZN_SUCCESS;
}

// ---- fragment programs --------------------------------------------------------------------------------------------------------------

namespace {

// The raster kernels' PTX (raster.cu compiled with KZN_FS_TEMPLATE, a string constant of this library) calls the fragment
// program as an external function.  nvcc writes every call as one delimited sequence
//     { // callseq N, 0   .param ...; st.param.f32 [param0], %f2; ... call.uni (retval0), kzn_fs_main, (...);
//       ld.param.v4.f32 {%f7, %f8, %f9, %f10}, [retval0];   } // callseq N
// which is replaced, call by call, by the shader's own PTX reading the stored operands and writing the loaded registers
// (kir::emit_fragment_inline): after that the fragment program is ordinary code of the kernel — registers allocated and
// instructions scheduled across it, unused inputs and outputs eliminated with the code that made them.
std::string build_fragment_module(const kir::Module& m, int* call_sites) {
    const std::string text = kzn_raster_template_ptx;
    std::string out;
    out.reserve(text.size() + (1u << 20));
    // the external declarations of the two entry points of the fragment program go: no reference to them is left
    size_t pos = 0;
    int declarations = 0;
    for (size_t d = text.find(".extern .func"); d != std::string::npos; d = text.find(".extern .func", d + 1)) {
        const size_t paren = text.find("\n(", d), end = text.find(';', d);
        if (paren == std::string::npos || end == std::string::npos) break;
        const std::string head = text.substr(d, paren - d);
        if (head.find(" kzn_fs_main") == std::string::npos) continue;
        out.append(text, pos, d - pos);
        pos = end + 1;
        ++declarations;
    }
    if (declarations != 2) throw kir::Error("internal: the raster template does not declare the fragment program the way this library expects");
    int sites = 0;
    for (;;) {
        const size_t b = text.find("{ // callseq ", pos);
        if (b == std::string::npos) break;
        // inside which definition?  A kernel entry keeps the push constants in its parameter space (DrawParams is its
        // only parameter); a helper function of the template sees them through a pointer
        std::string entry;
        {
            const size_t ent = text.rfind(".entry ", b), fn = text.rfind("\n.func", b);
            if (ent != std::string::npos && (fn == std::string::npos || fn < ent)) {
                const size_t n0 = ent + 7, n1 = text.find('(', n0);
                entry = text.substr(n0, n1 - n0);
            }
        }
        const size_t e0 = text.find("} // callseq ", b);
        if (e0 == std::string::npos) throw kir::Error("internal: unterminated call sequence in the raster template");
        const size_t e = text.find('\n', e0);
        const std::string seq = text.substr(b, e0 - b);
        if (seq.find("kzn_fs_main") == std::string::npos) {  // a call of another function of the template
            out.append(text, pos, e - pos);
            pos = e;
            continue;
        }
        kir::FragmentSite site;
        site.id = sites++;
        site.fragments = seq.find("kzn_fs_main2") != std::string::npos ? 2 : 1;
        site.in.assign(8 * site.fragments + 2, "0");
        site.out.assign(4 * site.fragments, "");
        if (!entry.empty()) site.push_constants_param = entry + "_param_0+" + std::to_string(kzn_draw_params_push_constants_offset);
        size_t at = 0;
        while (at < seq.size()) {
            size_t nl = seq.find('\n', at);
            if (nl == std::string::npos) nl = seq.size();
            std::string ln = seq.substr(at, nl - at);
            at = nl + 1;
            const size_t first = ln.find_first_not_of(" \t");
            if (first == std::string::npos) continue;
            ln = ln.substr(first);
            if (ln.rfind("st.param", 0) == 0) {
                const size_t pb = ln.find("[param"), pe = ln.find(']', pb), semi = ln.rfind(';');
                if (pb == std::string::npos || pe == std::string::npos || semi == std::string::npos) throw kir::Error("internal: unexpected st.param in a fragment call");
                const int k = atoi(ln.c_str() + pb + 6);
                size_t v = ln.find(',', pe);
                std::string operand = ln.substr(v + 1, semi - v - 1);
                operand.erase(0, operand.find_first_not_of(" \t"));
                if (k < 0 || k >= (int)site.in.size()) throw kir::Error("internal: fragment call with more parameters than its callee has");
                site.in[k] = operand;
            } else if (ln.rfind("ld.param", 0) == 0) {
                const size_t rb = ln.find("[retval0");
                if (rb == std::string::npos) throw kir::Error("internal: unexpected ld.param in a fragment call");
                int component = 0;
                if (ln[rb + 8] == '+') component = atoi(ln.c_str() + rb + 9) / 4;
                const size_t comma = ln.rfind(',', rb);
                std::string dst = ln.substr(ln.find_first_of(" \t"), comma - ln.find_first_of(" \t"));
                std::string regs;
                for (char ch : dst)
                    if (ch != '{' && ch != '}' && ch != ' ' && ch != '\t') regs += ch;
                size_t r0 = 0;
                while (r0 <= regs.size() && component < (int)site.out.size()) {
                    size_t r1 = regs.find(',', r0);
                    if (r1 == std::string::npos) r1 = regs.size();
                    if (r1 > r0) site.out[component++] = regs.substr(r0, r1 - r0);
                    r0 = r1 + 1;
                }
            }
        }
        out.append(text, pos, b - pos);
        out += "// fragment program, call site " + std::to_string(site.id) + "\n";
        out += kir::emit_fragment_inline(m, site, nullptr);
        pos = e;
    }
    out.append(text, pos, std::string::npos);
    if (sites == 0) throw kir::Error("internal: the raster template holds no call of the fragment program");
    if (call_sites) *call_sites = sites;
    return out;
}

struct LaunchAttribute_ {  // CUlaunchAttribute (cuda.h): id, 4 bytes of padding, 64-byte value
    int id;
    char pad[4];
    union {
        char bytes[64];
        int programmatic_stream_serialization_allowed;
    } value;
};
struct LaunchConfig_ {  // CUlaunchConfig
    unsigned grid_x, grid_y, grid_z, block_x, block_y, block_z, shared_bytes;
    void* stream;
    LaunchAttribute_* attrs;
    unsigned num_attrs;
};

}  // namespace

KznResult kzn_fragment_program_ptx(const KznIrModule* module, char** ptx, size_t* length) {
    KZN_VALID(module && ptx, "kzn_fragment_program_ptx: null argument");
    *ptx = nullptr;
    return guarded([&] {
        kir::fragment_interface(*module->module);
        kir::FragmentSite site;
        site.id = 0;
        site.in = {"%fx", "%fy", "%fz", "%fw", "%in0", "%in1", "%in2", "%in3", "%push_constants", "%descriptor_table"};
        site.out = {"%out0", "%out1", "%out2", "%out3"};
        *ptx = dup_text(kir::emit_fragment_inline(*module->module, site, nullptr), length);
        KZN_REQUIRE(*ptx, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

KznResult kzn_fragment_program_module_ptx(const KznIrModule* module, char** ptx, size_t* length) {
    KZN_VALID(module && ptx, "kzn_fragment_program_module_ptx: null argument");
    *ptx = nullptr;
    return guarded([&] {
        kir::fragment_interface(*module->module);
        *ptx = dup_text(build_fragment_module(*module->module, nullptr), length);
        KZN_REQUIRE(*ptx, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
        return KZN_SUCCESS;
    });
}

KznResult kzn_fragment_program_create(KznDevice* device, const KznIrModule* module, KznFragmentProgram** program) {
    KZN_VALID(device && module && program, "kzn_fragment_program_create: null argument");
    *program = nullptr;
    auto p = std::make_unique<KznFragmentProgram>();
    p->device = device;
    std::string ptx;
    KznResult r = guarded([&] {
        const kir::FragmentInterface fi = kir::fragment_interface(*module->module);
        p->flags = (fi.reads_inputs || fi.reads_frag_coord_zw) ? 1u : 0u;
        ptx = build_fragment_module(*module->module, &p->call_sites);
        return KZN_SUCCESS;
    });
    if (r != KZN_SUCCESS) return r;
    if (const char* dump = getenv("KZN_DUMP_FRAGMENT_PTX")) {
        if (FILE* f = fopen(dump, "w")) {
            fwrite(ptx.data(), 1, ptx.size(), f);
            fclose(f);
        }
    }
    r = load_ptx(device, ptx, &p->module);
    if (r != KZN_SUCCESS) return r;
    static const char* const names[2][3] = {{"kzn_fs_raster_direct_0", "kzn_fs_raster_direct_1", "kzn_fs_raster_direct_2"},
                                            {"kzn_fs_raster_0", "kzn_fs_raster_1", "kzn_fs_raster_2"}};
    for (int k = 0; k < 2; ++k)
        for (int mode = 0; mode < 3; ++mode) {
            CUresult_ cr = driver()->cuModuleGetFunction(&p->kernel[k][mode], p->module, names[k][mode]);
            if (cr != 0) {
                driver()->cuModuleUnload(p->module);
                return driver_failed(cr, "cuModuleGetFunction");
            }
        }
    *program = p.release();
    return KZN_SUCCESS;
}

void kzn_fragment_program_destroy(KznFragmentProgram* program) {
    if (!program) return;
    if (program->module) {
        cudaSetDevice(program->device->ordinal);
        driver()->cuModuleUnload(program->module);
    }
    delete program;
}

KznResult kzn_cmd_bind_fragment_program(KznQueue* queue, KznFragmentProgram* program) {
    KZN_VALID(queue, "vkCmdBindPipeline: null queue");
    KZN_VALID(!program || program->device->ordinal == queue->device->ordinal, "fragment program was created for another device");
    queue->fragment_program = program;
    return KZN_SUCCESS;
}

// ---- vertex programs ------------------------------------------------------------------------------------------------------------

KznResult kzn_vertex_program_create(KznDevice* device, const KznIrModule* module, KznVertexProgram** program) {
    KZN_VALID(device && module && program, "kzn_vertex_program_create: null argument");
    *program = nullptr;
    const kir::Module& m = *module->module;
    KZN_REQUIRE(m.version >= 2, KZN_ERROR_FEATURE_NOT_PRESENT, "a vertex program needs version 2 of the IR (stores, the vertex_position built-in)");
    // where the emitter puts the built-in outputs: after the user outputs, each at a multiple of 16 bytes (ptx_emit.cpp emit_ptx)
    uint64_t at = m.user_outputs.size;
    int64_t position = -1;
    for (const auto& v : m.built_in_outputs) {
        at = (at + 15) / 16 * 16;
        const bool vec4 = v.type && v.type->kind == kir::TypeKind::Vector && v.type->length == 4 && v.type->element &&
                          v.type->element->kind == kir::TypeKind::Float && v.type->element->float_kind == kir::FloatKind::F32;
        if (v.built_in_kind == "vertex_position") {
            KZN_VALID(vec4, "vertex program: vertex_position must be <4 x f32>");
            position = (int64_t)at;
        }
        at += (kir::type_size(v.type) + 15) / 16 * 16;
    }
    KZN_VALID(position >= 0, "vertex program: the module has no built-in output of kind vertex_position");
    KZN_VALID(m.user_inputs.size > 0 && m.user_inputs.size % 4 == 0 && m.user_inputs.size < (1u << 28), "vertex program: bad user_inputs_block size");
    auto vp = std::make_unique<KznVertexProgram>();
    const uint32_t one[3] = {1, 1, 1};
    KznResult r = kzn_compute_pipeline_create(device, module, one, &vp->pipeline);
    if (r != KZN_SUCCESS) return r;
    vp->layout.inputs_size = vp->pipeline->inputs_size;
    vp->layout.outputs_size = vp->pipeline->outputs_size;
    vp->layout.position_offset = (uint32_t)position;
    vp->layout.color_offset = m.user_outputs.size >= 16 ? 0u : ~0u;  // user output location 0
    if (vp->layout.outputs_size != at || vp->layout.outputs_size >= (1u << 28)) {
        kzn_pipeline_destroy(vp->pipeline);
        kzn_set_error("internal: vertex program output slot of %llu bytes, expected %llu", (unsigned long long)vp->layout.outputs_size, (unsigned long long)at);
        return KZN_ERROR_VALIDATION_FAILED;
    }
    *program = vp.release();
    return KZN_SUCCESS;
}

void kzn_vertex_program_destroy(KznVertexProgram* program) {
    if (!program) return;
    kzn_pipeline_destroy(program->pipeline);
    delete program;
}

KznResult kzn_cmd_bind_vertex_program(KznQueue* queue, KznVertexProgram* program) {
    KZN_VALID(queue, "vkCmdBindPipeline: null queue");
    KZN_VALID(!program || program->pipeline->device->ordinal == queue->device->ordinal, "vertex program was created for another device");
    queue->vertex_program = program;
    return KZN_SUCCESS;
}

uint64_t kzn_vertex_program_inputs_size(const KznVertexProgram* program) { return program ? program->layout.inputs_size : 0; }
uint64_t kzn_vertex_program_outputs_size(const KznVertexProgram* program) { return program ? program->layout.outputs_size : 0; }

KznResult kzn_queue_reserve_vertex_outputs(KznQueue* queue, uint64_t max_vertices, const KznVertexProgram* program) {
    KznResult r = kzn_queue_enter(queue);
    if (r) return r;
    KZN_VALID(program && max_vertices > 0 && max_vertices <= 0xffffffffull, "kzn_queue_reserve_vertex_outputs: bad argument");
    KZN_VALID(!queue->recording, "kzn_queue_reserve_vertex_outputs allocates: call it before vkBeginCommandBuffer");
    uint64_t address = 0;
    return kzn_raster_vertex_scratch(queue, max_vertices * program->layout.outputs_size, &address);
}

}  // extern "C"

// internal (kzn_internal.h): used by raster.cu
KznVertexLayout kzn_vertex_program_layout(const KznVertexProgram* program) { return program ? program->layout : KznVertexLayout{}; }

KznResult kzn_vertex_program_run(KznQueue* queue, KznVertexProgram* program, uint64_t inputs, uint64_t outputs, uint64_t first, uint64_t count) {
    KZN_VALID(program && count <= 0xffffffffull, "internal: bad vertex program run");
    if (count == 0) return KZN_SUCCESS;
    // one invocation per vertex: invocation i works on slot first + i of both buffers; the GRAPHICS bind point's descriptor
    // set and the push constants go into the launch like a dispatch's
    const uint64_t saved = queue->descriptor_table[KZN_PIPELINE_BIND_POINT_COMPUTE];
    queue->descriptor_table[KZN_PIPELINE_BIND_POINT_COMPUTE] = queue->descriptor_table[KZN_PIPELINE_BIND_POINT_GRAPHICS];
    const uint32_t groups[3] = {(uint32_t)count, 1, 1};
    const KznResult r = kzn_cmd_dispatch(queue, program->pipeline, nullptr, groups, inputs, outputs, first);
    queue->descriptor_table[KZN_PIPELINE_BIND_POINT_COMPUTE] = saved;
    return r;
}

uint32_t kzn_fragment_program_flags(const KznFragmentProgram* program) { return program ? program->flags : 0u; }

KznResult kzn_fragment_program_launch(KznFragmentProgram* program, int kernel, int mode, unsigned grid_x, unsigned threads, size_t dynamic_smem,
                                      cudaStream_t stream, const void* params, size_t) {
    KZN_VALID(program && kernel >= 0 && kernel < 2 && mode >= 0 && mode < 3, "internal: bad fragment program launch");
    DriverApi* d = driver();
    CUfunction_ fn = program->kernel[kernel][mode];
    if (dynamic_smem > 48 * 1024) {
        CUresult_ cr = d->cuFuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)dynamic_smem);
        if (cr != 0) return driver_failed(cr, "cuFuncSetAttribute");
    }
    void* args[1] = {const_cast<void*>(params)};
    CUresult_ cr;
    if (d->cuLaunchKernelEx && kzn_pdl_enabled()) {
        // programmatic dependent launch, as the built-in kernels are launched (kzn_internal.h kzn_launch_pdl)
        LaunchAttribute_ attr;
        memset(&attr, 0, sizeof attr);
        attr.id = 6;  // CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION
        attr.value.programmatic_stream_serialization_allowed = 1;
        LaunchConfig_ cfg = {grid_x, 1, 1, threads, 1, 1, (unsigned)dynamic_smem, (void*)stream, &attr, 1};
        cr = d->cuLaunchKernelEx(&cfg, fn, args, nullptr);
    } else {
        cr = d->cuLaunchKernel(fn, grid_x, 1, 1, threads, 1, 1, (unsigned)dynamic_smem, (void*)stream, args, nullptr);
    }
    if (cr != 0) return driver_failed(cr, "cuLaunchKernel (fragment program)");
    return KZN_SUCCESS;
}

extern "C" {

// ---- descriptor sets and push constants -----------------------------------------------------------------------
// vkCreateDescriptorSetLayout / vkAllocateDescriptorSets api_impl.rs:4218, vkUpdateDescriptorSets :4339,
// vkCmdBindDescriptorSets :4597, vkCmdPushConstants :4932; vulkan-driver/src/descriptor_set.rs

KznResult kzn_descriptor_set_create(KznDevice* device, uint32_t binding_count, const KznDescriptorSetLayoutBinding* bindings, KznDescriptorSet** set) {
    KZN_VALID(device && set && (bindings || binding_count == 0), "kzn_descriptor_set_create: null argument");
    *set = nullptr;
    auto s = std::make_unique<KznDescriptorSet>();
    s->device = device;
    uint64_t slots = 0;
    for (uint32_t i = 0; i < binding_count; ++i) {
        const KznDescriptorSetLayoutBinding& b = bindings[i];
        // the kinds descriptor_set.rs can write (:460-527): buffers are what a kernel reads through a pointer; images and
        // samplers need a sampling path this back end does not have; texel and dynamic buffers are unimplemented!() upstream
        KZN_REQUIRE(b.descriptor_type == KZN_DESCRIPTOR_TYPE_UNIFORM_BUFFER || b.descriptor_type == KZN_DESCRIPTOR_TYPE_STORAGE_BUFFER,
                    KZN_ERROR_FEATURE_NOT_PRESENT, "descriptor type %u: only UNIFORM_BUFFER and STORAGE_BUFFER descriptors are supported", b.descriptor_type);
        for (auto& o : s->bindings) KZN_VALID(o.binding != b.binding, "binding %u appears twice in the layout", b.binding);
        KZN_VALID(slots + b.descriptor_count <= (1u << 20), "descriptor set with more than 2^20 descriptors");
        s->bindings.push_back({b.binding, b.descriptor_type, b.descriptor_count, (uint32_t)slots});
        slots += b.descriptor_count;
    }
    s->host.assign(2 * (slots ? slots : 1), 0);
    KZN_CUDA(cudaSetDevice(device->ordinal));
    KZN_CUDA(cudaMalloc(&s->table, s->host.size() * sizeof(uint64_t)));
    KZN_CUDA(cudaMemset(s->table, 0, s->host.size() * sizeof(uint64_t)));
    *set = s.release();
    return KZN_SUCCESS;
}

void kzn_descriptor_set_destroy(KznDescriptorSet* set) {
    if (!set) return;
    cudaSetDevice(set->device->ordinal);
    cudaFree(set->table);
    delete set;
}

KznResult kzn_descriptor_set_write_buffers(KznDescriptorSet* set, uint32_t binding, uint32_t first_element, uint32_t count,
                                           const KznDescriptorBufferInfo* buffers) {
    KZN_VALID(set && (buffers || count == 0), "kzn_descriptor_set_write_buffers: null argument");
    for (auto& b : set->bindings) {
        if (b.binding != binding) continue;
        KZN_VALID((uint64_t)first_element + count <= b.count, "descriptor write past the end of binding %u", binding);
        for (uint32_t i = 0; i < count; ++i) {
            set->host[2 * (size_t)(b.first_slot + first_element + i)] = buffers[i].address + buffers[i].offset;
            set->host[2 * (size_t)(b.first_slot + first_element + i) + 1] = buffers[i].range;
        }
        if (count) {
            // vkUpdateDescriptorSets is a host operation that is complete when it returns
            KZN_CUDA(cudaSetDevice(set->device->ordinal));
            KZN_CUDA(cudaMemcpy((uint64_t*)set->table + 2 * (size_t)(b.first_slot + first_element), &set->host[2 * (size_t)(b.first_slot + first_element)],
                                (size_t)count * 16, cudaMemcpyHostToDevice));
        }
        return KZN_SUCCESS;
    }
    kzn_set_error("binding %u is not in the set's layout", binding);
    return KZN_ERROR_VALIDATION_FAILED;
}

uint64_t kzn_descriptor_set_table_address(const KznDescriptorSet* set) { return set ? (uint64_t)(uintptr_t)set->table : 0; }

int64_t kzn_descriptor_set_slot(const KznDescriptorSet* set, uint32_t binding) {
    if (set)
        for (auto& b : set->bindings)
            if (b.binding == binding) return b.first_slot;
    return -1;
}

KznResult kzn_cmd_bind_descriptor_set(KznQueue* queue, uint32_t pipeline_bind_point, const KznDescriptorSet* set) {
    KZN_VALID(queue, "vkCmdBindDescriptorSets: null queue");
    KZN_VALID(pipeline_bind_point <= KZN_PIPELINE_BIND_POINT_COMPUTE, "vkCmdBindDescriptorSets: bad pipeline bind point %u", pipeline_bind_point);
    KZN_VALID(!set || set->device->ordinal == queue->device->ordinal, "descriptor set belongs to another device");
    queue->descriptor_table[pipeline_bind_point] = set ? (uint64_t)(uintptr_t)set->table : 0;
    return KZN_SUCCESS;
}

KznResult kzn_cmd_push_constants(KznQueue* queue, uint32_t offset, uint32_t size, const void* values) {
    KZN_VALID(queue && values, "vkCmdPushConstants: null argument");
    KZN_VALID(size > 0 && offset % 4 == 0 && size % 4 == 0, "vkCmdPushConstants: offset and size must be multiples of 4 and size non-zero");
    KZN_VALID((uint64_t)offset + size <= KZN_MAX_PUSH_CONSTANTS_SIZE, "vkCmdPushConstants: range past maxPushConstantsSize (%u)", KZN_MAX_PUSH_CONSTANTS_SIZE);
    memcpy(queue->push_constants + offset, values, size);
    return KZN_SUCCESS;
}

// ---- the backend trait surface --------------------------------------------------------------------------------------------------

const char* kzn_backend_compiler_name(void) { return "PTX (sm_100a)"; }  // Compiler::name lib.rs:249; LLVM's is "LLVM 7"

KznResult kzn_backend_context_create(KznBackendContext** context) {
    KZN_VALID(context, "null out pointer");
    *context = new (std::nothrow) KznBackendContext();
    KZN_REQUIRE(*context, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    return KZN_SUCCESS;
}
void kzn_backend_context_destroy(KznBackendContext* context) { delete context; }

static KznBackendType* scalar_type(KznBackendContext* c, BtKind k, uint32_t bits, bool sgn) {
    if (!c) return nullptr;
    KznBackendType t{k};
    t.bits = bits;
    t.is_signed = sgn;
    t.size = t.align = bits / 8;
    return c->type(t);
}
KznBackendType* kzn_backend_type_bool(KznBackendContext* c) { return scalar_type(c, BtKind::Bool, 8, false); }
KznBackendType* kzn_backend_type_int(KznBackendContext* c, uint32_t bits, int32_t is_signed) {
    if (bits != 8 && bits != 16 && bits != 32 && bits != 64) {
        kzn_set_error("integer types are 8, 16, 32 or 64 bits wide (types.rs:117-132)");
        return nullptr;
    }
    return scalar_type(c, BtKind::Int, bits, is_signed != 0);
}
KznBackendType* kzn_backend_type_f32(KznBackendContext* c) { return scalar_type(c, BtKind::F32, 32, true); }
KznBackendType* kzn_backend_type_f64(KznBackendContext* c) { return scalar_type(c, BtKind::F64, 64, true); }
KznBackendType* kzn_backend_type_pointer(KznBackendContext* c, KznBackendType* target) {
    if (!c || !target) return nullptr;
    KznBackendType t{BtKind::Pointer};
    t.element = target;
    t.size = t.align = 8;
    return c->type(t);
}
// No type of this back end can be larger than what a kernel could ever address as one variable; the bound also keeps the
// size arithmetic below from wrapping around (an array of 2^62 i64 "had" size 0).
static const uint64_t kMaxTypeBytes = 1ull << 40;
static bool sized_ok(uint64_t element_size, uint64_t count) {
    if (element_size != 0 && count > kMaxTypeBytes / element_size) {
        kzn_set_error("type of %llu x %llu bytes is larger than the back end's limit of 2^40 bytes", (unsigned long long)count,
                      (unsigned long long)element_size);
        return false;
    }
    return true;
}

KznBackendType* kzn_backend_type_array(KznBackendContext* c, KznBackendType* element, uint64_t count) {
    if (!c || !element || !sized_ok(element->size, count)) return nullptr;
    KznBackendType t{BtKind::Array};
    t.element = element;
    t.count = count;
    t.size = element->size * count;
    t.align = element->align;
    return c->type(t);
}
KznBackendType* kzn_backend_type_vector(KznBackendContext* c, KznBackendType* element, uint64_t length) {
    if (!c || !element) return nullptr;
    if (length == 0) {
        kzn_set_error("a vector type has at least one element");
        return nullptr;
    }
    if (!sized_ok(element->size, length)) return nullptr;
    KznBackendType t{BtKind::Vector};
    t.element = element;
    t.count = length;
    t.size = element->size * length;
    uint64_t a = 1;  // LLVM aligns a vector to its size rounded up to a power of two
    while (a < t.size) a *= 2;
    t.align = a > 16 ? 16 : a;
    return c->type(t);
}
KznBackendType* kzn_backend_type_struct(KznBackendContext* c, KznBackendType* const* members, uint64_t count) {
    if (!c || (count && !members)) return nullptr;
    KznBackendType t{BtKind::Struct};
    uint64_t off = 0, align = 1;
    for (uint64_t i = 0; i < count; ++i) {
        if (!members[i]) return nullptr;
        t.members.push_back(members[i]);
        uint64_t a = members[i]->align ? members[i]->align : 1;
        off = (off + a - 1) / a * a;
        if (!sized_ok(1, off) || !sized_ok(1, members[i]->size)) return nullptr;
        off += members[i]->size;
        if (a > align) align = a;
    }
    t.size = (off + align - 1) / align * align;
    t.align = align;
    return c->type(t);
}
KznBackendType* kzn_backend_type_function(KznBackendContext* c, KznBackendType* return_type, KznBackendType* const* arguments, uint64_t count) {
    if (!c || (count && !arguments)) return nullptr;
    KznBackendType t{BtKind::Function};
    t.element = return_type;
    for (uint64_t i = 0; i < count; ++i) {
        if (!arguments[i]) return nullptr;
        t.members.push_back(arguments[i]);
    }
    t.size = t.align = 8;
    return c->type(t);
}
uint64_t kzn_backend_type_size(const KznBackendType* type) { return type ? type->size : 0; }

KznBackendModule* kzn_backend_create_module(KznBackendContext* c, const char* name) {
    if (!c || !name) return nullptr;
    c->modules.push_back(std::make_unique<KznBackendModule>());
    KznBackendModule* m = c->modules.back().get();
    m->context = c;
    m->name = name;
    return m;
}
KznResult kzn_backend_module_set_source_file_name(KznBackendModule* m, const char* source_file_name) {
    KZN_VALID(m && source_file_name, "null argument");
    m->source_file_name = source_file_name;
    return KZN_SUCCESS;
}
KznBackendFunction* kzn_backend_add_function(KznBackendModule* m, const char* name, KznBackendType* function_type) {
    if (!m || !function_type || function_type->kind != BtKind::Function) {
        kzn_set_error("add_function: the type is not a function type");
        return nullptr;
    }
    if (!valid_function_name(name)) {
        kzn_set_error("invalid function name '%s' (backend.rs:444-459)", name ? name : "");
        return nullptr;
    }
    for (auto* f : m->functions)
        if (f->name == name) {
            kzn_set_error("function '%s' is already defined in module '%s'", name, m->name.c_str());
            return nullptr;
        }
    KznBackendContext* c = m->context;
    c->functions.push_back(std::make_unique<KznBackendFunction>());
    KznBackendFunction* f = c->functions.back().get();
    f->module = m;
    f->name = name;
    f->symbol = kir::ptx_mangle(name);
    f->type = function_type;
    for (uint32_t i = 0; i < function_type->members.size(); ++i)
        f->parameters.push_back(c->value({BvKind::Parameter, function_type->members[i], f, i}));
    f->as_value = c->value({BvKind::Function, function_type, f, 0});
    m->functions.push_back(f);
    return f;
}
uint64_t kzn_backend_function_parameter_count(const KznBackendFunction* f) { return f ? f->parameters.size() : 0; }
KznBackendValue* kzn_backend_function_parameter(KznBackendFunction* f, uint64_t index) {
    return f && index < f->parameters.size() ? f->parameters[index] : nullptr;
}
KznBackendValue* kzn_backend_function_as_value(KznBackendFunction* f) { return f ? f->as_value : nullptr; }
KznBackendBlock* kzn_backend_append_new_basic_block(KznBackendFunction* f, const char* name) {
    if (!f) return nullptr;
    KznBackendContext* c = f->module->context;
    c->blocks.push_back(std::make_unique<KznBackendBlock>());
    KznBackendBlock* b = c->blocks.back().get();
    b->function = f;
    b->name = name ? name : "";
    b->as_value = c->value({BvKind::Block, nullptr, f, (uint32_t)f->blocks.size()});
    f->blocks.push_back(b);
    return b;
}
KznBackendValue* kzn_backend_block_as_value(KznBackendBlock* b) { return b ? b->as_value : nullptr; }
KznBackendBuilder* kzn_backend_create_builder(KznBackendContext* c) {
    if (!c) return nullptr;
    c->builders.push_back(std::make_unique<KznBackendBuilder>());
    c->builders.back()->context = c;
    return c->builders.back().get();
}
KznResult kzn_backend_builder_attach(KznBackendBuilder* b, KznBackendBlock* block) {
    KZN_VALID(b && block, "null argument");
    KZN_VALID(!b->block, "builder is already attached (DetachedBuilder::attach consumes a detached builder, lib.rs:33-41)");
    KZN_VALID(!block->terminated, "basic block already has a terminator");
    b->block = block;
    return KZN_SUCCESS;
}
KznBackendBlock* kzn_backend_builder_current_basic_block(KznBackendBuilder* b) { return b ? b->block : nullptr; }
KznBackendValue* kzn_backend_build_alloca(KznBackendBuilder* b, KznBackendType* variable_type) {
    if (!b || !b->block || !variable_type) {
        kzn_set_error("build_alloca needs an attached builder and a type");
        return nullptr;
    }
    if (variable_type->kind == BtKind::Function || variable_type->kind == BtKind::Void) {
        kzn_set_error("build_alloca: the type has no size");
        return nullptr;
    }
    KznBackendContext* c = b->context;
    KznBackendFunction* f = b->block->function;
    KznBackendType* ptr = kzn_backend_type_pointer(c, variable_type);
    KznBackendValue* v = c->value({BvKind::Alloca, ptr, f, f->allocas++});
    BackendInstr in;
    in.value = v;
    in.type = variable_type;
    b->block->instrs.push_back(in);
    return v;
}
KznResult kzn_backend_build_return(KznBackendBuilder* b, KznBackendValue* value) {
    KZN_VALID(b && b->block, "build_return needs an attached builder");
    BackendInstr in;
    in.is_return = true;
    in.value = value;
    b->block->instrs.push_back(in);
    b->block->terminated = true;
    b->block = nullptr;  // build_return hands back a DetachedBuilder (lib.rs:26-29)
    return KZN_SUCCESS;
}
KznResult kzn_backend_module_verify(KznBackendModule* m) { return verify_module(m); }

KznResult kzn_backend_module_to_ptx(KznBackendModule* m, char** ptx, size_t* length) {
    KZN_VALID(m && ptx, "null argument");
    *ptx = nullptr;
    KznResult r = verify_module(m);
    if (r != KZN_SUCCESS) return r;
    *ptx = dup_text(backend_module_ptx(m), length);
    KZN_REQUIRE(*ptx, KZN_ERROR_OUT_OF_HOST_MEMORY, "out of host memory");
    return KZN_SUCCESS;
}

KznResult kzn_backend_compile(KznDevice* device, KznBackendModule* m, const char* const* keys, KznBackendFunction* const* callable_functions,
                              uint64_t count, KznCompiledCode** code) {
    KZN_VALID(device && m && code && (count == 0 || (keys && callable_functions)), "kzn_backend_compile: null argument");
    *code = nullptr;
    KznResult r = verify_module(m);
    if (r != KZN_SUCCESS) return r;
    for (uint64_t i = 0; i < count; ++i) {
        KZN_VALID(keys[i] && callable_functions[i], "callable function %llu is null", (unsigned long long)i);
        KZN_VALID(callable_functions[i]->module == m, "callable function '%s' belongs to another module", callable_functions[i]->name.c_str());
        KZN_VALID(!callable_functions[i]->blocks.empty(), "callable function '%s' has no body", callable_functions[i]->name.c_str());
    }
    auto cc = std::make_unique<KznCompiledCode>();
    cc->device = device;
    r = load_ptx(device, backend_module_ptx(m), &cc->module);
    if (r != KZN_SUCCESS) return r;
    for (uint64_t i = 0; i < count; ++i) {
        CompiledFunction f;
        CUresult_ cr = driver()->cuModuleGetFunction(&f.function, cc->module, callable_functions[i]->symbol.c_str());  // backend.rs:645-657
        if (cr != 0) {
            driver()->cuModuleUnload(cc->module);
            return driver_failed(cr, "cuModuleGetFunction");
        }
        f.has_return_slot = callable_functions[i]->type->element != nullptr;
        f.parameter_count = callable_functions[i]->parameters.size();
        cc->by_key[keys[i]] = f;
    }
    *code = cc.release();
    return KZN_SUCCESS;
}

KznResult kzn_compiled_code_get(KznCompiledCode* code, const char* key, KznCallable* callable) {
    KZN_VALID(code && key && callable, "null argument");
    auto it = code->by_key.find(key);
    if (it == code->by_key.end()) {  // CompiledCode::get returns None (lib.rs:204)
        callable->call = nullptr;
        callable->self = nullptr;
        kzn_set_error("no callable function with key '%s'", key);
        return KZN_NOT_READY;
    }
    callable->call = callable_trampoline;
    callable->self = &it->second;
    return KZN_SUCCESS;
}

void kzn_compiled_code_destroy(KznCompiledCode* code) {
    if (!code) return;
    if (code->module) {
        cudaSetDevice(code->device->ordinal);
        driver()->cuModuleUnload(code->module);
    }
    delete code;
}

}  // extern "C"
