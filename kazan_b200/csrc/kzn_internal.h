// kzn_internal.h — shared by the translation units of libkazan_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kazan_b200.h"

struct KznDevice {
    int ordinal;
    int sm_count;
    int l2_bytes;
    int cc_major, cc_minor;
    int l2_persist_bytes;   // cudaLimitPersistingL2CacheSize as set at creation (0: no persisting lines)
    int l2_window_bytes;    // cudaDevAttrMaxAccessPolicyWindowSize
};

struct KznRasterScratch;  // raster.cu
struct KznCommandBuffer;  // api.cu

// Contiguous, 16-byte aligned fills and copies recorded back to back into a command buffer whose address ranges are
// disjoint run as ONE launch (merged_transfer_kernel, transfer.cu): Vulkan orders nothing between transfer commands
// that no barrier separates, and disjoint ranges make the merged result identical to in-order execution anyway.
#define KZN_MAX_MERGED 96
struct KznPendingTransfer {
    uint64_t dst, src, bytes;  // src == 0: fill with `value`; bytes = payload (row_bytes x rows for a pitched box)
    uint32_t value;
    // pitched boxes of LINEAR memory (padded bufferRowLength, sub-rectangles): rows of row_bytes at the two pitches;
    // row_bytes == 0: one contiguous run.
    uint32_t row_bytes, rows;
    uint32_t dst_pitch, src_pitch;
};

// vkCmdBeginRenderPass with LOAD_OP_CLEAR, not yet performed: the next qualifying draw starts its bricks from these values
// (raster.cu), anything else performs the clear first (kzn_flush_render_pass_clear, transfer.cu).
struct KznPendingClear {
    bool active;
    KznFramebuffer fb;
    int32_t x0, y0, x1, y1;  // inclusive pixel rectangle: the render area cut to the framebuffer and its band
    uint32_t load;           // bit 0: colour, bit 1: depth
    uint32_t color_word, depth_bits;
};

struct KznQueue {
    KznDevice* device;
    cudaStream_t stream;
    bool owns_stream;
    uint64_t launches;
    uint64_t stats_address;
    KznRasterScratch* raster;  // lazily created, grows to the largest draw seen
    // kzn_queue_set_kernel_timing: event pair around the dominant kernel of the last command
    bool timing;
    bool timing_armed;  // begin recorded, the next kzn_launched records the end
    cudaEvent_t time_begin, time_end;
    KznCommandBuffer* recording;  // non-null between kzn_command_buffer_begin and _end: kzn_cmd_* are captured, not run
    KznPendingTransfer pending[KZN_MAX_MERGED];  // recorded, not yet captured (only while recording)
    uint32_t n_pending;
    uint64_t merged_commands;  // commands that went into a merged launch (statistics)
    // vkCmdBindDescriptorSets / vkCmdPushConstants state, read by the next dispatch or draw (its values go into that
    // launch's kernel parameters, so a recorded command keeps what was bound when it was recorded)
    uint64_t descriptor_table[2];  // [VkPipelineBindPoint]: device address of the bound set's table, 0 = none
    alignas(16) unsigned char push_constants[128];
    KznPendingClear pending_clear;
    bool in_render_pass;
    struct KznFragmentProgram* fragment_program;  // kzn_cmd_bind_fragment_program
    struct KznVertexProgram* vertex_program;      // kzn_cmd_bind_vertex_program
};

// backend.cpp: a draw with a compiled vertex stage.  layout: what one vertex occupies in the vertex scratch and where its
// position (four f32) and colour (four f32; ~0u: none) are; run: shade `count` vertices starting at vertex `first` of the
// binding at `inputs` (stride = the program's input block) into `outputs`, on the queue's stream.
struct KznVertexLayout {
    uint64_t inputs_size, outputs_size;
    uint32_t position_offset, color_offset;
};
KznVertexLayout kzn_vertex_program_layout(const struct KznVertexProgram* program);
KznResult kzn_vertex_program_run(KznQueue* queue, struct KznVertexProgram* program, uint64_t inputs, uint64_t outputs, uint64_t first, uint64_t count);
// raster.cu: the queue's vertex scratch, grown to `bytes`
KznResult kzn_raster_vertex_scratch(KznQueue* queue, uint64_t bytes, uint64_t* address);

// backend.cpp: launch one of the raster kernels built around a compiled fragment program.  kernel: 0 = raster_direct,
// 1 = raster (binned); mode: kDstNone / kDstGeneric / kDstAlphaOver; `params` is the kernel's single by-value parameter.
KznResult kzn_fragment_program_launch(struct KznFragmentProgram* program, int kernel, int mode, unsigned grid_x, unsigned threads,
                                      size_t dynamic_smem, cudaStream_t stream, const void* params, size_t params_size);
// what the program reads: bit 0 = interpolated inputs or frag_coord z / w
uint32_t kzn_fragment_program_flags(const struct KznFragmentProgram* program);
// byte offset of the push constant block inside the raster kernels' parameter (raster.cu DrawParams)
extern "C" const uint32_t kzn_draw_params_push_constants_offset;
// the raster kernels' PTX with the fragment program as an external call (generated at build time from raster.cu)
extern "C" const char kzn_raster_template_ptx[];

struct KznMemory {
    KznDevice* device;
    uint32_t type;
    uint64_t size;  // rounded up to KZN_MIN_MEMORY_MAP_ALIGNMENT
    void* device_ptr;
    void* host_ptr;  // HOST_COHERENT only
    bool imported;   // opened from another process's export (kzn_memory_import): closed, not freed
};

struct KznFence {
    KznDevice* device;
    cudaEvent_t event;
    bool signaled_without_work;  // created signaled, or reset and never submitted
    bool pending;                // an event record has been enqueued
};

struct KznSemaphore {
    KznDevice* device;
    cudaEvent_t event;
    bool signaled;  // a signal has been enqueued and not yet consumed by a wait
};

// A recorded draw names the queue's raster scratch buffers in its CUDA graph: the command buffer pins them (raster.cu) until
// it is destroyed, and may only be submitted to the queue it was recorded on (whose scratch and work counters it uses).
void kzn_command_buffer_uses_raster_scratch(KznCommandBuffer* cb, KznRasterScratch* scratch);
void kzn_raster_scratch_pin(KznRasterScratch* s);
void kzn_raster_scratch_unpin(KznRasterScratch* s);

// error plumbing -----------------------------------------------------------------------------
void kzn_set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
KznResult kzn_cuda_failed(cudaError_t e, const char* what, const char* file, int line);

#define KZN_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return kzn_cuda_failed(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define KZN_REQUIRE(cond, code, ...)   \
    do {                               \
        if (!(cond)) {                 \
            kzn_set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

#define KZN_VALID(cond, ...) KZN_REQUIRE(cond, KZN_ERROR_VALIDATION_FAILED, __VA_ARGS__)

// make the queue's device current and check the queue is usable; captures any pending merged transfers first
KznResult kzn_queue_enter(KznQueue* q);
KznResult kzn_queue_enter_draw(KznQueue* q);  // as kzn_queue_enter, but a pending render-pass clear stays pending (the draw decides)
KznResult kzn_flush_render_pass_clear(KznQueue* q);  // transfer.cu
// the same for the transfer commands themselves, which may join the pending batch instead of flushing it
KznResult kzn_queue_enter_transfer(KznQueue* q);
// capture the pending batch now (transfer.cu); no-op when nothing is pending
KznResult kzn_flush_transfers(KznQueue* q);
// call after every kernel launch: counts it and maps launch errors
KznResult kzn_launched(KznQueue* q, const char* kernel);
// call right before launching the kernel that dominates a command (no-op unless kernel timing is enabled)
void kzn_timing_begin(KznQueue* q);
// host memory whose contents a recorded command reads at replay time: copied into storage owned by the command buffer
// being recorded (returns `data` itself when the queue is not recording)
const void* kzn_recorded_host_copy(KznQueue* q, const void* data, size_t size);

// layout helpers (image.rs:58-88 + SURVEY.md E.10) ----------------------------------------------
static inline uint32_t kzn_texel_size(uint32_t format) {
    return (format == KZN_FORMAT_R8G8B8A8_UNORM || format == KZN_FORMAT_D32_SFLOAT) ? 4u : 0u;
}
static inline uint64_t kzn_round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
static inline uint32_t kzn_tiles(uint32_t texels) { return (uint32_t)(((uint64_t)texels + KZN_TILE_DIM - 1) / KZN_TILE_DIM); }
static inline uint64_t kzn_row_pitch(const KznImageDesc* im) {  // LINEAR only
    return (uint64_t)kzn_texel_size(im->format) * im->width;
}
static inline uint64_t kzn_slice_pitch(const KznImageDesc* im) {
    uint64_t ts = kzn_texel_size(im->format);
    if (im->tiling == KZN_TILING_TILED)
        return (uint64_t)kzn_tiles(im->width) * kzn_tiles(im->height) * KZN_TILE_DIM * KZN_TILE_DIM * ts;
    return ts * im->width * im->height;
}
// false when the image's byte size does not fit 2^62: the reference computes it with checked_mul().unwrap()
// (image.rs:74-82) and panics; here the image is rejected before any address arithmetic can wrap around
static inline bool kzn_image_size_fits(const KznImageDesc* im) {
    // VkOffset3D is signed 32-bit: an extent past 2^31 - 1 could not be addressed by any region, and the kernels keep
    // coordinates in int32
    if (im->width > 0x7fffffffu || im->height > 0x7fffffffu || im->depth > 0x7fffffffu || im->array_layers > 0x7fffffffu) return false;
    unsigned long long w = im->width, h = im->height;
    if (im->tiling == KZN_TILING_TILED) {
        w = (unsigned long long)kzn_tiles(im->width) * KZN_TILE_DIM;
        h = (unsigned long long)kzn_tiles(im->height) * KZN_TILE_DIM;
    }
    unsigned long long n = 0;
    if (__builtin_mul_overflow(w, h, &n) || __builtin_mul_overflow(n, (unsigned long long)im->depth, &n) ||
        __builtin_mul_overflow(n, (unsigned long long)im->array_layers, &n) || __builtin_mul_overflow(n, 4ull, &n))
        return false;
    return n < (1ull << 62);
}
KznResult kzn_check_image(const KznImageDesc* im, const char* what);

// E.1 on the host (used for the one-off clear colour conversion)
uint32_t kzn_pack_rgba8_host(const float rgba[4]);

#ifdef __CUDACC__
// Programmatic dependent launch (used by the transfer and draw kernels): a kernel launched through kzn_launch_pdl may
// become resident before the preceding kernel of the stream has drained; it must call kzn_pdl_wait() before it touches
// memory.  The wait returns once that kernel has completed and its writes are visible (a no-op for an ordinary launch), so
// in-order semantics are unchanged and only the launch gap between consecutive kernels is hidden.  KZN_PDL=0 in the
// environment falls back to ordinary launches.
__device__ __forceinline__ void kzn_pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool kzn_pdl_enabled();
template <typename... KArgs, typename... Args>
inline void kzn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = kzn_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through kzn_launched (cudaGetLastError)
}
#endif
